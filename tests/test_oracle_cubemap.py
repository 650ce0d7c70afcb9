"""CPU: the oracle's restatement of alaserMapping's map cube store (laserMapping.cpp:74-108,309-550,736-801 --
SURVEY.md 8 f-1, the next row of the scope table; there is no GPU counterpart yet).  The ring-buffer shift, gather,
insertion and per-cube re-filter are checked against an independent dictionary model, and the optimisation step
against orc.Mapping driven by hand."""
import numpy as np
import pytest

W, H, D = 21, 21, 11


def cube_of(v, centre):
    """int((v + 25.0) / 50.0) + centre, minus one below zero (laserMapping.cpp:316-325,741-750); v float32 -> double"""
    v = float(np.float32(v))
    c = int((v + 25.0) / 50.0) + centre
    if v + 25.0 < 0:
        c -= 1
    return c


def test_shift_gather_insert_refilter_against_dictionary_model(orc):
    rng = np.random.default_rng(11)
    cm = orc.CubeMap()
    model = {}                      # world cube (i - cen_w, j - cen_h, k - cen_d) -> [points]  per type
    models = [dict(), dict()]
    cen = [10, 10, 5]
    ident = np.array([0, 0, 0, 1.0])
    # a path that walks out of the initial 21 x 21 x 11 block in every direction (positive and negative)
    path = [(0, 0, 0), (60, -35, 12), (130, -80, 30), (260, -170, 75), (420, -290, 140), (300, -100, 60), (-90, 40, -30),
            (-400, 380, -160), (-700, 600, -260), (-640, 610, -250)]
    for step, t in enumerate(path):
        t = np.array(t, float)
        # fewer than 10 corner points in reach => the optimisation is skipped and the pose is the odometry pose
        corner = (rng.normal(size=(6, 4)) * [8, 8, 2, 0]).astype(np.float32)
        surf = (rng.normal(size=(300, 4)) * [30, 30, 3, 0]).astype(np.float32)
        pose, info = cm.step(corner, surf, ident, t, 0.4, 0.8)
        assert info["optimised"] == 0 and np.array_equal(pose, np.concatenate([ident, t]))
        # --- model: centre cube and shift (:314-509)
        c = [cube_of(t[a], cen[a]) for a in range(3)]
        dims = [W, H, D]
        for a in range(3):
            while c[a] < 3:
                c[a] += 1; cen[a] += 1
            while c[a] >= dims[a] - 3:
                c[a] -= 1; cen[a] -= 1
        for m in models:          # cubes that scrolled out of the block are gone
            for key in [k for k in m if not all(0 <= k[a] + cen[a] < dims[a] for a in range(3))]:
                del m[key]
        st = cm.state()
        assert st["centre"] == tuple(cen)
        valid = [(i, j, k) for i in range(c[0] - 2, c[0] + 3) for j in range(c[1] - 2, c[1] + 3) for k in range(c[2] - 1, c[2] + 2)
                 if 0 <= i < W and 0 <= j < H and 0 <= k < D]
        assert st["valid"] == [i + W * j + W * H * k for i, j, k in valid]
        # --- gather (:531-539) happens BEFORE insertion
        for which, m in enumerate(models):
            parts = [np.array(m[(i - cen[0], j - cen[1], k - cen[2])], np.float32).reshape(-1, 4) for i, j, k in valid
                     if (i - cen[0], j - cen[1], k - cen[2]) in m]
            exp = np.concatenate(parts) if parts else np.zeros((0, 4), np.float32)
            assert np.array_equal(cm.cloud(which), exp), (step, which)
        # --- stacks, insertion (:736-767), re-filter of the valid cubes (:770-788)
        for which, (last, leaf) in enumerate([(corner, 0.4), (surf, 0.8)]):
            stack = orc.voxel_grid(last, leaf, orc.SORT_CANONICAL)
            assert np.array_equal(cm.cloud(2 + which), stack)
            m = models[which]
            for p in stack:
                w = np.array([np.float32(float(p[0]) + t[0]), np.float32(float(p[1]) + t[1]), np.float32(float(p[2]) + t[2]), p[3]], np.float32)
                ijk = [cube_of(w[a], cen[a]) for a in range(3)]
                if all(0 <= ijk[a] < dims[a] for a in range(3)):
                    m.setdefault((ijk[0] - cen[0], ijk[1] - cen[1], ijk[2] - cen[2]), []).append(w)
            for i, j, k in valid:
                key = (i - cen[0], j - cen[1], k - cen[2])
                if key in m:
                    m[key] = list(orc.voxel_grid(np.array(m[key], np.float32), leaf, orc.SORT_CANONICAL))
        # --- every cube of the block agrees with the model
        for which, m in enumerate(models):
            total = 0
            for key, pts in m.items():
                idx = (key[0] + cen[0]) + W * (key[1] + cen[1]) + W * H * (key[2] + cen[2])
                got = cm.cube(which, idx)
                assert np.array_equal(got, np.array(pts, np.float32).reshape(-1, 4)), (step, which, key)
                total += len(pts)
            assert total == (st := cm.state())["total_corner" if which == 0 else "total_surf"]
    assert cen != [10, 10, 5]        # the path did force shifts


def test_mapping_loop_matches_hand_driven_mapping(orc, synth, scans):
    """a short VLP-16 trajectory: CubeMap.step == transformAssociateToMap -> Mapping on the gathered submap ->
    transformUpdate, done by hand with the pieces tested elsewhere; the map refinement keeps the pose near the truth"""
    ns, _, mr = synth.SENSORS["VLP-16"][:3]
    cm = orc.CubeMap()
    od = orc.Odometry()
    q = np.array([0, 0, 0, 1.0]); t = np.zeros(3)
    qw = np.array([0, 0, 0, 1.0]); tw = np.zeros(3)
    q_wmap_wodom = np.array([0, 0, 0, 1.0]); t_wmap_wodom = np.zeros(3)
    optimised = 0
    for k in range(5):
        f = orc.Features(scans("VLP-16", k, n_az=900), ns, mr)
        if k > 0:
            q, t, _ = od.register(f.sharp, f.flat, q, t)
            qw, tw = orc.integrate_pose(qw, tw, q, t)
        od.set_last(f.less_sharp, f.less_flat)
        # the hand-driven frame uses the submap the store is ABOUT to gather: read it from the cubes before stepping
        st_before = cm.state()
        pose, info = cm.step(f.less_sharp, f.less_flat, qw, tw, 0.2, 0.4)     # VLP-16 launch file resolutions
        x0 = orc.transform_associate_to_map(q_wmap_wodom, t_wmap_wodom, qw, tw)
        m = orc.Mapping(); m.set_map(cm.cloud(0), cm.cloud(1))
        x, minfo = m.register(cm.cloud(2), cm.cloud(3), x0)
        assert bool(info["optimised"]) == minfo["optimised"]
        assert np.array_equal(pose, x)
        q_wmap_wodom, t_wmap_wodom = orc.transform_update(x, qw, tw)
        st = cm.state()
        assert np.array_equal(st["q_wmap_wodom"], q_wmap_wodom) and np.array_equal(st["t_wmap_wodom"], t_wmap_wodom)
        optimised += info["optimised"]
        if k == 0:
            assert info["corner_from_map"] == 0 and not info["optimised"]     # empty map on the first frame (:554)
        else:
            assert info["corner_from_map"] > 10 and info["surf_from_map"] > 50
        qt, tt = synth.pose(k)
        assert np.abs(pose[4:] - tt).max() < 0.05
    assert optimised == 4
