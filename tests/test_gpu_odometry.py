"""GPU parity: scan-to-scan association + LM (laserOdometry.cpp:274-568, lidarFactor.hpp) vs the CPU oracle."""
import numpy as np
import pytest

from conftest import rot_angle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair(orc, synth, scans):
    """features of two consecutive HDL-64 scans (oracle, canonical)"""
    ns, _, mr = synth.SENSORS["HDL-64"][:3]
    f0 = orc.Features(scans("HDL-64", 0), ns, mr)
    f1 = orc.Features(scans("HDL-64", 1), ns, mr)
    return f0, f1


@pytest.fixture(scope="module")
def ctx(aloam):
    c = aloam.Aloam(n_scans=64, max_points=200000)
    yield c
    c.close()


def test_knn_exact(ctx, orc, pair):
    f0, f1 = pair
    ctx.odometry_set_last(f0.less_sharp, f0.less_flat)
    for which, cloud, q in [(0, f0.less_sharp, f1.sharp), (1, f0.less_flat, f1.flat), (1, f0.less_flat, f1.less_sharp[:3000])]:
        idx, sqd = ctx.knn(which, q, 1)
        ridx, rsqd = orc.bruteforce_knn(cloud, q, 1)
        assert np.array_equal(idx, ridx)
        assert np.array_equal(sqd, rsqd)  # float-exact distances
        kidx, ksqd = orc.KdTree(cloud).knn(q, 1)
        assert np.array_equal(kidx, ridx) and np.array_equal(ksqd, rsqd)


@pytest.mark.parametrize("pose", [([0, 0, 0, 1.0], [0, 0, 0.0]), ([0.001, -0.002, 0.008, 0.99996], [0.7, 0.02, -0.01])])
def test_association_exact(ctx, orc, pair, pose):
    f0, f1 = pair
    q = np.array(pose[0]); q /= np.linalg.norm(q)
    t = np.array(pose[1])
    od = orc.Odometry()
    od.set_last(f0.less_sharp, f0.less_flat)
    cc, pc, blocks = od.associate(f1.sharp, f1.flat, q, t)
    ctx.odometry_set_last(f0.less_sharp, f0.less_flat)
    gcc, gpc = ctx.odometry_associate(f1.sharp, f1.flat, q, t)
    got_c = [(i, a, b) for i, (a, b, v) in enumerate(gcc) if v]
    got_p = [(i, a, b, c) for i, (a, b, c, v) in enumerate(gpc) if v]
    assert got_c == [tuple(r) for r in cc]
    assert got_p == [tuple(r) for r in pc]
    assert len(got_c) > 500 and len(got_p) > 1000


def test_normal_equations(ctx, orc, pair):
    f0, f1 = pair
    od = orc.Odometry()
    od.set_last(f0.less_sharp, f0.less_flat)
    q = np.array([0.002, 0.001, 0.004, 1.0]); q /= np.linalg.norm(q)
    t = np.array([0.5, 0.05, 0.0])
    _, _, blocks = od.associate(f1.sharp, f1.flat, q, t)
    x = np.concatenate([q, t])
    for autodiff in (True, False):
        JtJ, Jtr, cost = orc.normal_equations(blocks, x, autodiff=autodiff)
        g_JtJ, g_Jtr, g_cost = ctx.normal_equations(blocks, x)
        assert abs(g_cost - cost) <= 1e-12 * cost
        assert np.abs(g_JtJ - JtJ).max() <= 1e-10 * np.abs(JtJ).max()
        assert np.abs(g_Jtr - Jtr).max() <= 1e-10 * np.abs(Jtr).max()


def test_solve_matches_ceres_restatement(ctx, orc, pair):
    f0, f1 = pair
    od = orc.Odometry()
    od.set_last(f0.less_sharp, f0.less_flat)
    x0 = np.array([0, 0, 0, 1.0, 0, 0, 0])
    _, _, blocks = od.associate(f1.sharp, f1.flat, x0[:4], x0[4:])
    xr, sr, tr = orc.solve(blocks, x0, max_iters=4)
    xg, sg, tg = ctx.solve(blocks, x0)
    assert sg["termination"] == sr["termination"]
    assert sg["num_iterations"] == sr["num_iterations"] and sg["num_successful"] == sr["num_successful"]
    assert np.abs(xg - xr).max() < 1e-9
    assert abs(sg["final_cost"] - sr["final_cost"]) < 1e-9 * sr["final_cost"]
    n = min(len(tr), len(tg))
    assert np.allclose(tg[:n, [0, 5]], tr[:n, [0, 5]], rtol=1e-8)  # cost and trust-region radius, step by step
    # empty problem: parameters untouched
    xe, se, _ = ctx.solve(np.zeros((0, 11)), x0)
    assert np.array_equal(xe, x0) and se["termination"] == 4


def test_register_pose(ctx, orc, pair):
    f0, f1 = pair
    od = orc.Odometry()
    od.set_last(f0.less_sharp, f0.less_flat)
    q, t, info = od.register(f1.sharp, f1.flat, [0, 0, 0, 1.0], [0, 0, 0.0])
    ctx.odometry_set_last(f0.less_sharp, f0.less_flat)
    gq, gt, st = ctx.odometry_register(f1.sharp, f1.flat, [0, 0, 0, 1.0], [0, 0, 0.0])
    assert np.abs(gt - t).max() < 1e-4 and rot_angle(gq, q) < 1e-4     # the north-star tolerance
    assert np.abs(gt - t).max() < 1e-8 and rot_angle(gq, q) < 1e-7     # what the design actually delivers
    assert st["n_corner_corr"] == info["corner_corr"] and st["n_plane_corr"] == info["plane_corr"]


@pytest.mark.parametrize("sensor,frames", [("HDL-64", 6), ("VLP-16", 5), ("HDL-32", 4)])
def test_fused_pipeline_poses(aloam, orc, synth, scans, sensor, frames):
    """raw scans -> world poses, fused device pipeline vs oracle pipeline (extract -> register -> integrate -> set_last)"""
    ns, _, mr = synth.SENSORS[sensor][:3]
    c = aloam.Aloam(n_scans=ns, max_points=200000)
    od = orc.Odometry()
    q = np.array([0, 0, 0, 1.0]); t = np.zeros(3)
    qw = np.array([0, 0, 0, 1.0]); tw = np.zeros(3)
    for k in range(frames):
        raw = scans(sensor, k)
        f = orc.Features(raw, ns, mr)
        if k > 0:
            q, t, _ = od.register(f.sharp, f.flat, q, t)
            qw, tw = orc.integrate_pose(qw, tw, q, t)
        od.set_last(f.less_sharp, f.less_flat)
        gq, gt, st = c.scan_to_pose(raw)
        if k == 0:
            assert st["flags"] & aloam.FLAG_INITIALISED_ONLY
        assert np.abs(gt - tw).max() < 1e-4 and rot_angle(gq, qw) < 1e-4, (k, gt, tw)
        assert np.abs(gt - tw).max() < 1e-7, (k, gt - tw)
    c.close()


def test_scan_stream_equals_per_scan_calls(aloam, synth, scans):
    """the multi-stream pipelined call returns exactly the poses of one aloam_scan_to_pose call per scan (host and device input)"""
    import torch
    raws = [scans("HDL-64", k) for k in range(7)]
    a = aloam.Aloam(n_scans=64, max_points=140000)
    ref = [np.concatenate(a.scan_to_pose(r)[:2]) for r in raws]
    n = max(len(r) for r in raws)
    host = torch.zeros((len(raws), n, 4), dtype=torch.float32).pin_memory()
    for i, r in enumerate(raws):
        host[i, :len(r)] = torch.from_numpy(r)
    dev = host.cuda()
    for buf, devres in [(host, False), (dev, True)]:
        a.reset_odometry()
        p1, _ = a.scan_stream([buf[i].data_ptr() for i in range(3)], [len(r) for r in raws[:3]], devres)     # split across two calls
        p2, _ = a.scan_stream([buf[i].data_ptr() for i in range(3, 7)], [len(r) for r in raws[3:]], devres)
        got = np.concatenate([p1, p2])
        assert np.array_equal(got, np.array(ref))
    a.close()


def test_scan_stream_mixes_with_per_scan_calls(aloam, synth, scans):
    """stream calls and synchronous calls share one trajectory state: any interleaving gives the per-scan poses, and a
    long stream (several turns of the feature-set ring, programmatic launches across scans) stays bit-identical"""
    import torch
    raws = [scans("VLP-16", k, n_az=900) for k in range(14)]
    a = aloam.Aloam(n_scans=16, max_points=20000)
    ref = np.array([np.concatenate(a.scan_to_pose(r)[:2]) for r in raws])
    n = max(len(r) for r in raws)
    dev = torch.zeros((len(raws), n, 4), dtype=torch.float32)
    for i, r in enumerate(raws):
        dev[i, :len(r)] = torch.from_numpy(r)
    dev = dev.cuda()
    cnt = [len(r) for r in raws]
    a.reset_odometry()
    got = [np.concatenate(a.scan_to_pose(raws[0])[:2])]
    p, _ = a.scan_stream([dev[i].data_ptr() for i in range(1, 6)], cnt[1:6], True); got += list(p)
    got.append(np.concatenate(a.scan_to_pose(raws[6])[:2]))
    p, _ = a.scan_stream([dev[i].data_ptr() for i in range(7, 14)], cnt[7:14], True); got += list(p)
    assert np.array_equal(np.array(got), ref)
    a.reset_odometry()
    p, _ = a.scan_stream([dev[i].data_ptr() for i in range(14)], cnt, True)
    assert np.array_equal(p, ref)
    a.close()


def test_replay_of_kitti_files_equals_direct_calls(aloam, synth, scans, tmp_path):
    """tools/replay_kitti.py: scans written as KITTI .bin files and replayed through aloam_scan_stream give the poses of
    direct aloam_scan_to_pose calls; the written KITTI pose lines parse back to the same lidar-frame poses"""
    import importlib, os, sys
    io = importlib.import_module("a-loam_b200.io")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import replay_kitti
    raws = [scans("VLP-16", k, n_az=900) for k in range(6)]
    paths = []
    for k, r in enumerate(raws):
        p = str(tmp_path / ("%06d.bin" % k)); io.write_kitti_bin(p, r); paths.append(p)
    got = replay_kitti.replay(paths, n_scans=16, chunk=4)
    a = aloam.Aloam(n_scans=16, max_points=20000)
    ref = np.array([np.concatenate(a.scan_to_pose(r)[:2]) for r in raws])
    a.close()
    assert np.array_equal(got, ref)
    T = io.lidar_pose_to_kitti(got[-1][:4], got[-1][4:])
    q, t = io.kitti_pose_to_lidar(T)
    assert np.abs(t - got[-1][4:]).max() < 1e-12 and rot_angle(q, got[-1][:4]) < 1e-7


def test_per_stage_calls_do_not_disturb_the_fused_pipeline(aloam, synth, scans):
    """aloam_extract_features / aloam_odometry_set_last / aloam_odometry_register / aloam_solve between aloam_scan_to_pose calls
    of the same context leave the fused pipeline's feature sets, search indices and warm start alone (the per-stage entry
    points own separate feature sets and a separate pose), also while the index of the previous scan is still being built
    on the index stream"""
    raws = [scans("VLP-16", k, n_az=900) for k in range(5)]
    clean = aloam.Aloam(n_scans=16, max_points=40000)
    ref = [np.concatenate(clean.scan_to_pose(r)[:2]) for r in raws]
    clean.close()
    c = aloam.Aloam(n_scans=16, max_points=40000)
    rng = np.random.default_rng(5)
    for k, r in enumerate(raws):
        q, t, _ = c.scan_to_pose(r)
        assert np.array_equal(np.concatenate([q, t]), ref[k]), k
        f = c.extract_features(raws[(k + 2) % 5])                      # a different scan through the per-stage path
        c.odometry_set_last(f["less_sharp"], f["less_flat"])
        c.odometry_register(f["sharp"], f["flat"], np.array([0, 0, 0, 1.0]), rng.normal(0, 0.1, 3))
        blocks = np.array([[2, 1, 2, 3, 0, 0, 1, 0, 0, 0, -2.5]], float)
        c.solve(blocks, np.array([0, 0, 0, 1.0, 0.3, 0.2, 0.1]))
    c.close()
