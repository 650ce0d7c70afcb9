"""CPU: the oracle's feature extraction (restating scanRegistration.cpp:129-408) -- invariants the reference code implies,
LITERAL vs CANONICAL tie handling, and the committed golden vectors."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("sensor", ["VLP-16", "HDL-32", "HDL-64"])
def test_feature_invariants(sensor, orc, synth, scans):
    ns, az, mr = synth.SENSORS[sensor][:3]
    raw = scans(sensor, 2)
    f = orc.Features(raw, ns, mr)
    rings = f.full[:, 3].astype(int)
    assert np.all(np.diff(rings) >= 0), "ring-major ascending (scanRegistration.cpp:247-252)"
    assert np.all((f.full[:, 3] - rings >= 0) & (f.full[:, 3] - rings < 0.1001)), "intensity = ring + 0.1*relTime (:239)"
    r2 = np.einsum("ij,ij->i", f.full[:, :3], f.full[:, :3])
    assert r2.min() >= np.float32(mr) ** 2, "minimum_range filter (:99)"
    used = [r for r in range(ns) if f.scan_end[r] - f.scan_start[r] >= 6]
    assert len(f.sharp) <= 2 * 6 * len(used) and len(f.flat) <= 4 * 6 * len(used) and len(f.less_sharp) <= 20 * 6 * len(used)
    # every sharp point is also a less-sharp point (:303-305), label bookkeeping
    ls = {tuple(p) for p in f.less_sharp}
    assert all(tuple(p) in ls for p in f.sharp)
    assert (f.label == 2).sum() == len(f.sharp) and (f.label >= 1).sum() == len(f.less_sharp) and (f.label == -1).sum() == len(f.flat)
    assert np.all(f.curvature[f.label >= 1] > 0.1) and np.all(f.curvature[f.label == -1] < 0.1)
    # less-flat cloud is ring-major too and every voxel centroid stays on its ring
    assert np.all(np.diff(f.less_flat[:, 3].astype(int)) >= 0)
    # curvature formula spot check (:256-266)
    i = int(f.scan_start[used[0]]) + 7
    d = f.full[i - 5:i + 6, :3].astype(np.float32)
    s = np.zeros(3, np.float32)
    for k in list(range(0, 5)):
        s = s + d[k]
    s = s - np.float32(10) * d[5]
    for k in range(6, 11):
        s = s + d[k]
    assert np.float32(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]) == f.curvature[i]


@pytest.mark.parametrize("sensor", ["VLP-16", "HDL-64"])
def test_literal_vs_canonical(sensor, orc, synth, scans):
    """std::sort's tie order only matters for equal curvatures (measure zero with range noise) and inside voxels
    (<= a few ulp on the less-flat centroids): SURVEY.md 8a note 4"""
    ns, az, mr = synth.SENSORS[sensor][:3]
    raw = scans(sensor, 1)
    a = orc.Features(raw, ns, mr, orc.SORT_CANONICAL)
    b = orc.Features(raw, ns, mr, orc.SORT_LITERAL)
    for name in ["full", "sharp", "less_sharp", "flat"]:
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    assert a.less_flat.shape == b.less_flat.shape
    assert np.abs(a.less_flat - b.less_flat).max() < 5e-5
    # no exact curvature ties inside any sorted segment of the golden-style inputs
    for r in range(ns):
        s, e = a.scan_start[r], a.scan_end[r]
        if e - s < 6:
            continue
        for j in range(6):
            sp, ep = s + (e - s) * j // 6, s + (e - s) * (j + 1) // 6 - 1
            c = a.curvature[sp:ep + 1]
            assert len(np.unique(c)) == len(c)


def test_bad_inputs(orc):
    with pytest.raises(RuntimeError):
        orc.Features(np.zeros((10, 4), np.float32), 48, 0.3)          # unsupported scan count (:472-476)
    with pytest.raises(RuntimeError):
        orc.Features(np.full((10, 4), np.nan, np.float32), 16, 0.3)   # nothing survives
    f = orc.Features(np.array([[10, 0, 0, 0], [0, 10, 0.1, 0], [-10, 0, 0.2, 0]], np.float32), 16, 0.3)
    assert len(f.sharp) == 0 and len(f.less_flat) == 0 and len(f.full) == 3  # rings too short to process (:279)


def test_golden_features(orc):
    g = np.load(os.path.join(GOLD, "features_vlp16_az360.npz"))
    f = orc.Features(g["raw"], 16, 0.3, orc.SORT_CANONICAL)
    for name in ["full", "sharp", "less_sharp", "flat", "less_flat"]:
        assert np.array_equal(getattr(f, name)[:, :3], g[name][:, :3]), name
        assert np.allclose(getattr(f, name)[:, 3], g[name][:, 3], atol=1e-6), name
    assert np.array_equal(f.label, g["label"]) and np.array_equal(f.curvature, g["curvature"])
