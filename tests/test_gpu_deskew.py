"""GPU parity of the DISTORTION == 1 build (laserOdometry.cpp:59,111-148,376-379,470-473 and the slerp inside
lidarFactor.hpp:27-33,79-85): per-point interpolation ratio s = (intensity - int(intensity)) / SCAN_PERIOD in
TransformToStart and in every residual block, analytic slerp Jacobian in the LM kernel -- against the CPU oracle with the
same flag."""
import numpy as np
import pytest

from conftest import rot_angle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(aloam):
    c = aloam.Aloam(n_scans=16, max_points=60000, distortion=1)
    yield c
    c.close()


def test_association_and_register_with_distortion(ctx, orc, scans):
    # features from the product's own extraction: the very same intensity bits (ring + 0.1 relTime) go to both sides
    f0 = ctx.extract_features(scans("VLP-16", 0))
    f1 = ctx.extract_features(scans("VLP-16", 1))
    frac = f1["sharp"][:, 3] - np.floor(f1["sharp"][:, 3])
    assert frac.max() > 0.05 and frac.min() >= 0.0          # the ratios really are spread over the sweep
    od = orc.Odometry(distortion=True)
    od.set_last(f0["less_sharp"], f0["less_flat"])
    ctx.odometry_set_last(f0["less_sharp"], f0["less_flat"])
    q0 = np.array([0.002, -0.001, 0.008, 1.0]); q0 /= np.linalg.norm(q0)
    t0 = np.array([0.7, 0.05, -0.01])
    cc, pc, blocks = od.associate(f1["sharp"], f1["flat"], q0, t0)
    gc, gp = ctx.odometry_associate(f1["sharp"], f1["flat"], q0, t0)
    ok = gc[:, 2] == 1
    assert np.array_equal(np.flatnonzero(ok), cc[:, 0]) and np.array_equal(gc[ok][:, :2], cc[:, 1:3])
    okp = gp[:, 3] == 1
    assert np.array_equal(np.flatnonzero(okp), pc[:, 0]) and np.array_equal(gp[okp][:, :3], pc[:, 1:4])
    assert len(cc) > 100 and len(pc) > 200
    assert 0.0 < blocks[:, 10].min() < 0.2 and blocks[:, 10].max() > 0.8        # per-block ratios, not 1.0
    # normal equations of exactly these blocks (s != 1: slerp + analytic Jacobian on the device vs Jet autodiff)
    x0 = np.concatenate([q0, t0])
    JtJ, Jtr, cost = ctx.normal_equations(blocks, x0)
    rJ, rr, rc = orc.normal_equations(blocks, x0, autodiff=True)
    sc = np.abs(rJ).max()
    assert np.abs(JtJ - rJ).max() / sc < 1e-10 and np.abs(Jtr - rr).max() / sc < 1e-10 and abs(cost - rc) < 1e-10 * rc
    # full 2 x (association + LM)
    qr, tr, info = od.register(f1["sharp"], f1["flat"], q0, t0)
    qg, tg, st = ctx.odometry_register(f1["sharp"], f1["flat"], q0, t0)
    assert np.abs(tg - tr).max() < 1e-7 and rot_angle(qg, qr) < 1e-7
    assert st["lm_iters"] == sum(int(s["num_iterations"]) for s in info["summaries"])
    assert (st["n_corner_corr"], st["n_plane_corr"]) == (info["corner_corr"], info["plane_corr"])


def test_fused_pipeline_with_distortion(aloam, orc, synth, scans):
    c = aloam.Aloam(n_scans=16, max_points=60000, distortion=1)
    ns, _, mr = synth.SENSORS["VLP-16"][:3]
    od = orc.Odometry(distortion=True)
    q = np.array([0, 0, 0, 1.0]); t = np.zeros(3); qw = q.copy(); tw = t.copy()
    for k in range(5):
        raw = scans("VLP-16", k)
        f = orc.Features(raw, ns, mr)
        if k > 0:
            q, t, _ = od.register(f.sharp, f.flat, q, t)
            qw, tw = orc.integrate_pose(qw, tw, q, t)
        od.set_last(f.less_sharp, f.less_flat)
        gq, gt, st = c.scan_to_pose(raw)
        # the relTime fraction differs by <= 1 ulp between atan2f implementations, so s differs by ~1e-6: far inside 1e-4
        assert np.abs(gt - tw).max() < 1e-4 and rot_angle(gq, qw) < 1e-4
    assert np.abs(gt - tw).max() < 1e-5
    c.close()


def test_transform_to_end(ctx, orc, scans):
    f = ctx.extract_features(scans("VLP-16", 2))
    q = np.array([0.004, -0.003, 0.02, 1.0]); q /= np.linalg.norm(q)
    t = np.array([0.8, -0.02, 0.01])
    for dist in (True, False):
        got = ctx.transform_to_end(f["less_flat"], q, t, dist)
        ref = orc.transform_to_end(f["less_flat"], q, t, dist)
        assert got.shape == ref.shape
        assert np.array_equal(got[:, 3], ref[:, 3]) and np.array_equal(got[:, 3], np.floor(f["less_flat"][:, 3]))
        assert np.abs(got[:, :3] - ref[:, :3]).max() < 2e-6      # acos / sin differ in the last bits between libm and CUDA
    assert ctx.transform_to_end(np.zeros((0, 4), np.float32), q, t).shape == (0, 4)
