"""GPU parity: feature extraction (scanRegistration.cpp:129-408) -- CUDA path vs the CPU oracle, through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SENSORS = ["VLP-16", "HDL-32", "HDL-64"]


@pytest.fixture(scope="module")
def ctxs(aloam, synth):
    out = {}
    for s in SENSORS:
        out[s] = aloam.Aloam(n_scans=synth.SENSORS[s][0], max_points=200000)
    yield out
    for c in out.values():
        c.close()


@pytest.mark.parametrize("sensor", SENSORS)
@pytest.mark.parametrize("index", [0, 3])
def test_features_match_oracle(sensor, index, ctxs, orc, synth, scans):
    ns, _, mr = synth.SENSORS[sensor][:3]
    raw = scans(sensor, index)
    ref = orc.Features(raw, ns, mr, orc.SORT_CANONICAL)
    got = ctxs[sensor].extract_features(raw)
    # ring-major cloud: coordinates bit-exact; intensity = ring + 0.1*relTime may differ in the last ulps because
    # atan2f is not correctly rounded on either side (SURVEY.md 8a note 5) -- the integer part (the ring) must agree
    assert got["full"].shape == ref.full.shape
    assert np.array_equal(got["full"][:, :3], ref.full[:, :3])
    assert np.array_equal(got["full"][:, 3].astype(np.int32), ref.full[:, 3].astype(np.int32))
    assert np.all(np.abs(got["full"][:, 3] - ref.full[:, 3]) <= 1e-6 + np.spacing(ref.full[:, 3]))
    curv, label, s, e = ctxs[sensor].debug_features(ref.full.shape[0])
    assert np.array_equal(s, ref.scan_start) and np.array_equal(e, ref.scan_end)
    for r in range(ns):  # curvature is only defined (and used) inside [scanStart, scanEnd]
        a, b = ref.scan_start[r], ref.scan_end[r]
        if b - a < 6:
            continue
        assert np.array_equal(curv[a:b + 1], ref.curvature[a:b + 1])
        assert np.array_equal(label[a:b], ref.label[a:b])
    for name in ["sharp", "less_sharp", "flat"]:
        g, r_ = got[name], getattr(ref, name)
        assert g.shape == r_.shape, name
        assert np.array_equal(g[:, :3], r_[:, :3]), name
        assert np.all(np.abs(g[:, 3] - r_[:, 3]) <= 1e-6 + np.spacing(r_[:, 3])), name
    g, r_ = got["less_flat"], ref.less_flat
    assert g.shape == r_.shape
    assert np.array_equal(g[:, :3], r_[:, :3])
    assert np.array_equal(g[:, 3].astype(np.int32), r_[:, 3].astype(np.int32))
    assert np.abs(g[:, 3] - r_[:, 3]).max() <= 1e-5


def test_stride8_and_nan_and_close_points(ctxs, orc, synth, scans):
    """PCL 32-byte points, NaN returns and returns inside minimum_range are handled like the reference (:136-137)"""
    raw = scans("VLP-16", 1).copy()
    rng = np.random.default_rng(5)
    bad = rng.choice(raw.shape[0], 200, replace=False)
    raw[bad[:100], 0] = np.nan
    raw[bad[100:], :3] *= 0.001  # inside 0.3 m
    raw[0, 1] = np.nan            # first return invalid => start azimuth comes from the next one
    raw8 = np.zeros((raw.shape[0], 8), np.float32)
    raw8[:, :4] = raw
    raw8[:, 4:] = 123.0
    ref = orc.Features(raw, 16, 0.3, orc.SORT_CANONICAL)
    got = ctxs["VLP-16"].extract_features(raw8)
    assert np.array_equal(got["full"][:, :3], ref.full[:, :3])
    for name in ["sharp", "less_sharp", "flat", "less_flat"]:
        assert np.array_equal(got[name][:, :3], getattr(ref, name)[:, :3]), name


def test_error_codes(aloam, ctxs):
    c = ctxs["VLP-16"]
    with pytest.raises(aloam.AloamError) as e:
        c.extract_features(np.full((100, 4), np.nan, np.float32))
    assert e.value.code == -3
    with pytest.raises(aloam.AloamError) as e:
        aloam.Aloam(n_scans=48)
    assert e.value.code == -2
    with pytest.raises(aloam.AloamError) as e:
        c.extract_features(np.zeros((200001, 4), np.float32))
    assert e.value.code == -4
