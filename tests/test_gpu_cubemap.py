"""GPU parity: the device-resident map cube store (cubemap.cu) vs the oracle restatement (oracle/cubemap.cc) of
laserMapping.cpp:309-550,736-801 -- shift, gather, stack filters, insertion, per-cube VoxelGrid, and the whole mapping
loop over a short trajectory."""
import numpy as np
import pytest

from conftest import rot_angle

pytestmark = pytest.mark.gpu

W, H = 21, 21


def test_cube_store_bit_exact_when_the_pose_is_the_odometry_pose(aloam, orc):
    """thin clouds (< 10 corner points in reach) => no optimisation, pose = odometry pose: every cube of the block must
    then equal the oracle's bit for bit along a path that scrolls the ring buffer in all six directions"""
    rng = np.random.default_rng(11)
    c = aloam.Aloam(n_scans=16, max_points=20000, max_map_points=200000, line_res=0.4, plane_res=0.8)
    c.mapper_reset()
    cm = orc.CubeMap()
    ident = np.array([0, 0, 0, 1.0])
    path = [(0, 0, 0), (60, -35, 12), (130, -80, 30), (260, -170, 75), (420, -290, 140), (300, -100, 60), (-90, 40, -30),
            (-400, 380, -160), (-700, 600, -260), (-640, 610, -250)]
    for t in path:
        t = np.array(t, float)
        corner = (rng.normal(size=(6, 4)) * [8, 8, 2, 0]).astype(np.float32)
        surf = (rng.normal(size=(300, 4)) * [30, 30, 3, 0]).astype(np.float32)
        pose, info = cm.step(corner, surf, ident, t, 0.4, 0.8)
        q, tt, st = c.mapper_step(corner, surf, ident, t)
        assert np.array_equal(np.concatenate([q, tt]), pose) and st["flags"] & aloam.FLAG_MAP_TOO_THIN
        so, sg = cm.state(), c.mapper_state()
        assert sg["centre"] == so["centre"] and sg["valid"] == so["valid"]
        assert (sg["total_corner"], sg["total_surf"]) == (so["total_corner"], so["total_surf"])
        for idx in so["valid"]:
            for which in (0, 1):
                assert np.array_equal(c.mapper_cube(which, idx), cm.cube(which, idx)), (t, idx, which)
    c.close()


def test_mapping_loop_over_a_trajectory(aloam, orc, synth, scans):
    """odometry poses from the oracle, map refinement on the device: refined poses within the north-star tolerance of
    the oracle's mapping loop, same optimisation decisions, submaps of (nearly) the same size"""
    ns, _, mr = synth.SENSORS["VLP-16"][:3]
    c = aloam.Aloam(n_scans=16, max_points=40000, max_map_points=400000)      # VLP-16 launch file: 0.2 / 0.4
    c.mapper_reset()
    cm = orc.CubeMap()
    od = orc.Odometry()
    q = np.array([0, 0, 0, 1.0]); t = np.zeros(3)
    qw = np.array([0, 0, 0, 1.0]); tw = np.zeros(3)
    for k in range(6):
        f = orc.Features(scans("VLP-16", k, n_az=900), ns, mr)
        if k > 0:
            q, t, _ = od.register(f.sharp, f.flat, q, t)
            qw, tw = orc.integrate_pose(qw, tw, q, t)
        od.set_last(f.less_sharp, f.less_flat)
        pose, info = cm.step(f.less_sharp, f.less_flat, qw, tw, 0.2, 0.4)
        gq, gt, st = c.mapper_step(f.less_sharp, f.less_flat, qw, tw)
        assert bool(info["optimised"]) == (not (st["flags"] & aloam.FLAG_MAP_TOO_THIN))
        assert np.abs(gt - pose[4:]).max() < 1e-4 and rot_angle(gq, pose[:4]) < 1e-4, (k, gt - pose[4:])
        so, sg = cm.state(), c.mapper_state()
        assert sg["centre"] == so["centre"] and sg["valid"] == so["valid"]
        # the poses agree to ~1e-8, so a transformed point can land on the other side of a voxel boundary: sizes may
        # differ by a handful of points, not more
        assert abs(sg["total_corner"] - so["total_corner"]) <= 5 and abs(sg["total_surf"] - so["total_surf"]) <= 20
        # cube CONTENTS after the optimisation: same voxels in the same order, centroids equal to float rounding of a 1e-8 pose
        # difference (a cube whose size differs had a point cross a voxel / cube boundary and is skipped)
        same = differ = 0
        for idx in so["valid"]:
            for which in (0, 1):
                a, b = c.mapper_cube(which, idx), cm.cube(which, idx)
                if a.shape != b.shape:
                    differ += 1
                    continue
                same += 1
                if len(a):
                    assert np.abs(a[:, :3] - b[:, :3]).max() < 2e-5 and np.abs(a[:, 3] - b[:, 3]).max() < 1e-4, (k, idx, which)
        assert same >= 1 and differ <= 4
    c.close()


def test_stream_with_mapping_equals_per_scan_calls_and_oracle(aloam, orc, synth, scans):
    """aloam_scan_stream_mapped (odometry -> scan-to-map hand-off on the device) == aloam_scan_to_pose +
    aloam_mapper_step per scan, and both follow the oracle's ascanRegistration | alaserOdometry | alaserMapping chain"""
    import torch
    ns, _, mr = synth.SENSORS["VLP-16"][:3]
    raws = [scans("VLP-16", k, n_az=900) for k in range(6)]
    maxn = max(r.shape[0] for r in raws)
    c = aloam.Aloam(n_scans=16, max_points=maxn + 1024, max_map_points=400000)
    dev = [torch.from_numpy(r).cuda() for r in raws]
    # split in two calls: the mapper state carries over
    o1, m1 = c.scan_stream_mapped([d.data_ptr() for d in dev[:4]], [r.shape[0] for r in raws[:4]], True)
    o2, m2 = c.scan_stream_mapped([r.ctypes.data for r in raws[4:]], [r.shape[0] for r in raws[4:]], False)
    odom, mapped = np.concatenate([o1, o2]), np.concatenate([m1, m2])
    c.close()
    # per-scan calls
    c2 = aloam.Aloam(n_scans=16, max_points=maxn + 1024, max_map_points=400000)
    c2.mapper_reset()
    cm = orc.CubeMap()
    od = orc.Odometry()
    q = np.array([0, 0, 0, 1.0]); t = np.zeros(3); qw = q.copy(); tw = t.copy()
    for k, raw in enumerate(raws):
        gq, gt, _ = c2.scan_to_pose(raw)
        assert np.array_equal(np.concatenate([gq, gt]), odom[k])
        f = c2.extract_features(raw)
        mq, mt, _ = c2.mapper_step(f["less_sharp"], f["less_flat"], gq, gt)
        assert np.array_equal(np.concatenate([mq, mt]), mapped[k]), k
        fo = orc.Features(raw, ns, mr)
        if k > 0:
            q, t, _ = od.register(fo.sharp, fo.flat, q, t)
            qw, tw = orc.integrate_pose(qw, tw, q, t)
        od.set_last(fo.less_sharp, fo.less_flat)
        pose, info = cm.step(fo.less_sharp, fo.less_flat, qw, tw, 0.2, 0.4)
        assert np.abs(mapped[k, 4:] - pose[4:]).max() < 1e-4 and rot_angle(mapped[k, :4], pose[:4]) < 1e-4
    c2.close()


def test_slab_overflow_is_flagged_not_fatal(aloam):
    """a cube that receives more points than a slab holds keeps running: the overflow is dropped and flagged (the
    reference's cubes grow without bound between re-filters)"""
    rng = np.random.default_rng(3)
    c = aloam.Aloam(n_scans=16, max_points=3000, max_map_points=20000, line_res=0.01, plane_res=0.01)   # slab capacity = 3000
    c.mapper_reset()
    ident = np.array([0, 0, 0, 1.0])
    flags = 0
    for k in range(3):
        corner = (rng.uniform(-10, 10, size=(5, 4))).astype(np.float32)
        surf = (rng.uniform(-20, 20, size=(2500, 4)) * [1, 1, 0.2, 0]).astype(np.float32)     # all inside the centre cube
        q, t, st = c.mapper_step(corner, surf, ident, np.zeros(3))
        flags |= st["flags"]
    assert flags & aloam.FLAG_CUBE_OVERFLOW
    assert c.mapper_state()["total_surf"] == 3000
    c.close()
