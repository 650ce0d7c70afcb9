"""CPU: association + odometry + mapping oracle behaviour (laserOdometry.cpp:274-506, laserMapping.cpp:554-734)."""
import os

import numpy as np
import pytest

from conftest import rot_angle

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _pair(orc, synth, scans, sensor="VLP-16"):
    ns, az, mr = synth.SENSORS[sensor][:3]
    return orc.Features(scans(sensor, 0), ns, mr), orc.Features(scans(sensor, 1), ns, mr)


def test_association_rules(orc, synth, scans):
    f0, f1 = _pair(orc, synth, scans)
    od = orc.Odometry()
    od.set_last(f0.less_sharp, f0.less_flat)
    cc, pc, bl = od.associate(f1.sharp, f1.flat, [0, 0, 0, 1.0], [0, 0, 0.0])
    assert len(bl) == len(cc) + len(pc) and len(cc) > 100 and len(pc) > 200
    ring_c = f0.less_sharp[:, 3].astype(int)
    ring_s = f0.less_flat[:, 3].astype(int)
    for q, a, b in cc:      # second edge point: a different ring within +-2 (laserOdometry.cpp:312-361)
        assert ring_c[b] != ring_c[a] and abs(ring_c[b] - ring_c[a]) <= 2
    for q, a, b, c in pc:   # b: same ring as a ; c: another ring within +-2 (:402-455)
        assert ring_s[b] == ring_s[a] and b != a and ring_s[c] != ring_s[a] and abs(ring_s[c] - ring_s[a]) <= 2
    # nearest neighbour really is the nearest, in float arithmetic
    idx, sqd = orc.bruteforce_knn(f0.less_sharp, f1.sharp, 1)
    for q, a, b in cc:
        assert idx[q, 0] == a and sqd[q, 0] < 25


def test_odometry_recovers_motion(orc, synth, scans):
    sensor = "VLP-16"
    ns, az, mr = synth.SENSORS[sensor][:3]
    od = orc.Odometry()
    q = np.array([0, 0, 0, 1.0]); t = np.zeros(3)
    qw = np.array([0, 0, 0, 1.0]); tw = np.zeros(3)
    for k in range(6):
        f = orc.Features(scans(sensor, k), ns, mr)
        if k > 0:
            q, t, info = od.register(f.sharp, f.flat, q, t)
            qw, tw = orc.integrate_pose(qw, tw, q, t)
            assert all(s["final_cost"] <= s["initial_cost"] for s in info["summaries"])
        od.set_last(f.less_sharp, f.less_flat)
    qg, tg = synth.pose(5)
    assert np.abs(tw - tg)[1:].max() < 0.1 and abs(tw[0] - tg[0]) < 0.8   # the corridor axis is the weak direction
    assert rot_angle(qw, qg) < 0.02


def test_golden_odometry(orc):
    g = np.load(os.path.join(GOLD, "odometry_vlp16_az360.npz"))
    od = orc.Odometry()
    od.set_last(g["corner_last"], g["surf_last"])
    cc, pc, bl = od.associate(g["sharp"], g["flat"], g["q0"], g["t0"])
    assert np.array_equal(cc, g["corner_corr"]) and np.array_equal(pc, g["plane_corr"])
    q, t, _ = od.register(g["sharp"], g["flat"], g["q0"], g["t0"])
    assert np.allclose(q, g["q"], atol=1e-12) and np.allclose(t, g["t"], atol=1e-12)


def test_small_linear_algebra(orc):
    rng = np.random.default_rng(0)
    for _ in range(50):
        B = rng.normal(0, 1, (5, 3)) * rng.uniform(0.01, 5)
        A = (B - B.mean(0)).T @ (B - B.mean(0))
        ev, V = orc.eig3_sym(A)
        w, U = np.linalg.eigh(A)
        assert np.allclose(ev, w, rtol=1e-10, atol=1e-12 * abs(w).max())
        assert np.allclose(np.abs(np.sum(V * U, axis=0)), 1, atol=1e-7)
        P = rng.normal(0, 1, (5, 3)) + rng.uniform(-50, 50, 3)
        n = orc.lsq_5x3(P, -np.ones(5))
        assert np.allclose(n, np.linalg.lstsq(P, -np.ones(5), rcond=None)[0], rtol=1e-8, atol=1e-10)


def test_mapping_refines_pose(orc, synth, scans):
    """scan-to-map on a map built from neighbouring scans at their true poses: a perturbed initial guess is pulled back"""
    sensor = "VLP-16"
    ns, az, mr, lres, pres = synth.SENSORS[sensor]
    corner, surf = [], []
    for k in [0, 1, 3, 4]:
        f = orc.Features(scans(sensor, k), ns, mr)
        qk, tk = synth.pose(k)
        R = _rot(qk)
        for src, dst in [(f.less_sharp, corner), (f.less_flat, surf)]:
            w = src.copy(); w[:, :3] = (src[:, :3].astype(np.float64) @ R.T + tk).astype(np.float32); dst.append(w)
    cmap = orc.voxel_grid(np.concatenate(corner), lres)
    smap = orc.voxel_grid(np.concatenate(surf), pres)
    f2 = orc.Features(scans(sensor, 2), ns, mr)
    cs, ss = orc.voxel_grid(f2.less_sharp, lres), orc.voxel_grid(f2.less_flat, pres)
    q2, t2 = synth.pose(2)
    x0 = np.concatenate([q2, t2 + np.array([0.05, -0.04, 0.02])])
    m = orc.Mapping()
    m.set_map(cmap, smap)
    fits, bl = m.associate(cs, ss, x0)
    assert len(bl) == len(fits) and len(bl) > 300
    x, info = m.register(cs, ss, x0)
    assert info["optimised"]
    assert np.abs(x[4:] - t2).max() < 0.03 < np.abs(x0[4:] - t2).max()
    thin = orc.Mapping(); thin.set_map(cmap[:5], smap[:20])
    x2, info2 = thin.register(cs, ss, x0)
    assert not info2["optimised"] and np.array_equal(x2, x0)      # laserMapping.cpp:554,730-733


def _rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def test_transform_helpers(orc):
    import ctypes as C
    L = orc.lib()
    qm = np.array([0.01, 0.02, -0.03, 0.9993]); qm /= np.linalg.norm(qm)
    tm = np.array([1.0, 2, 3]); qo = np.array([0, 0, 0.1, 0.995]); qo /= np.linalg.norm(qo); to = np.array([4.0, 5, 6])
    x = np.zeros(7)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    L.orc_transform_associate_to_map(dp(qm), dp(tm), dp(qo), dp(to), dp(x))
    qm2, tm2 = np.zeros(4), np.zeros(3)
    L.orc_transform_update(dp(x), dp(qo), dp(to), dp(qm2), dp(tm2))
    assert np.allclose(qm2, qm, atol=1e-12) and np.allclose(tm2, tm, atol=1e-12)   # laserMapping.cpp:142-152 round trip


def test_distortion_mode_of_the_oracle(orc, synth, scans):
    """#define DISTORTION 1 (laserOdometry.cpp:59): per-point ratio s = (intensity - int(intensity)) / SCAN_PERIOD in
    TransformToStart and the residual blocks; TransformToEnd (:133-148) maps the sweep-start frame to the sweep-end frame"""
    ns, _, mr = synth.SENSORS["VLP-16"][:3]
    f0 = orc.Features(scans("VLP-16", 0), ns, mr)
    f1 = orc.Features(scans("VLP-16", 1), ns, mr)
    q = np.array([0.003, -0.002, 0.01, 1.0]); q /= np.linalg.norm(q)
    t = np.array([0.75, 0.02, -0.01])
    # s == 1 everywhere: TransformToEnd(p) = q^-1 (q p + t - t) = p up to float rounding, intensity -> scan id
    same = orc.transform_to_end(f1.less_flat, q, t, distortion=False)
    assert np.abs(same[:, :3] - f1.less_flat[:, :3]).max() < 2e-5
    assert np.array_equal(same[:, 3], np.floor(f1.less_flat[:, 3]))
    # with the ratio: a point at the sweep end (s -> 1) stays, a point at the sweep start (s -> 0) moves by about q^-1 (p - t) - p
    moved = orc.transform_to_end(f1.less_flat, q, t, distortion=True)
    s = (f1.less_flat[:, 3] - np.floor(f1.less_flat[:, 3])) / 0.1
    d = np.linalg.norm(moved[:, :3] - f1.less_flat[:, :3], axis=1)
    late, early = s > 0.95, s < 0.05
    assert late.any() and early.any() and d[late].max() < 0.08 and d[early].min() > 0.5
    # the blocks carry the ratio, and the association differs from the DISTORTION 0 one
    od0, od1 = orc.Odometry(), orc.Odometry(distortion=True)
    for od in (od0, od1):
        od.set_last(f0.less_sharp, f0.less_flat)
    _, _, b0 = od0.associate(f1.sharp, f1.flat, q, t)
    _, _, b1 = od1.associate(f1.sharp, f1.flat, q, t)
    assert np.all(b0[:, 10] == 1.0) and b1[:, 10].min() < 0.2 and b1[:, 10].max() <= 1.0
    # analytic evaluation falls back to autodiff for s != 1 inside the oracle: both flags give the same normal equations
    x = np.concatenate([q, t])
    Ja, ga, ca = orc.normal_equations(b1, x, autodiff=True)
    Jb, gb, cb = orc.normal_equations(b1, x, autodiff=False)
    assert np.array_equal(Ja, Jb) and ca == cb
