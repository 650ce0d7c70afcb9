"""GPU parity of the batched scan stream (BASELINE configs[4]; aloam_scan_stream_batch): several trajectories advance in
lockstep sharing every kernel launch.  Each trajectory must be bit-identical to its solo run through aloam_scan_stream and
within 1e-4 m / 1e-4 rad of the CPU oracle; the warm-start chain of laserOdometry.cpp:97-98 is kept per trajectory."""
import numpy as np
import pytest

from conftest import rot_angle

pytestmark = pytest.mark.gpu


def _trajectories(synth, sensor, n_traj, n_scans, n_az=None):
    return [[synth.scan(sensor, k, seed=synth.BASE_SEED + 50 + b, n_az=n_az) for k in range(n_scans)] for b in range(n_traj)]


def _oracle_poses(orc, synth, scans, sensor):
    ns, _, mr = synth.SENSORS[sensor][:3]
    od = orc.Odometry()
    q = np.array([0, 0, 0, 1.0]); t = np.zeros(3)
    qw = np.array([0, 0, 0, 1.0]); tw = np.zeros(3)
    out = []
    for k, raw in enumerate(scans):
        f = orc.Features(raw, ns, mr)
        if k > 0:
            q, t, _ = od.register(f.sharp, f.flat, q, t)
            qw, tw = orc.integrate_pose(qw, tw, q, t)
        od.set_last(f.less_sharp, f.less_flat)
        out.append(np.concatenate([qw, tw]))
    return np.array(out)


@pytest.mark.parametrize("sensor,n_traj,max_ring", [("HDL-32", 4, 2304), ("HDL-64", 3, 2048), ("VLP-16", 16, 4096)])
def test_batch_equals_solo_and_oracle(aloam, orc, synth, sensor, n_traj, max_ring):
    import torch
    n_scans = 5
    ns = synth.SENSORS[sensor][0]
    trajs = _trajectories(synth, sensor, n_traj, n_scans, n_az=900 if sensor == "VLP-16" else None)
    maxn = max(s.shape[0] for tr in trajs for s in tr)
    dev = [[torch.from_numpy(s).cuda() for s in tr] for tr in trajs]
    ptrs = np.array([[dev[b][k].data_ptr() for b in range(n_traj)] for k in range(n_scans)], np.uint64)
    counts = np.array([[trajs[b][k].shape[0] for b in range(n_traj)] for k in range(n_scans)])
    ctx = aloam.Aloam(n_scans=ns, max_points=maxn + 1024, max_batch=n_traj, max_ring_points=max_ring)
    # two calls (3 + 2 scans): the batch state carries over between calls like the single-trajectory stream
    poses = np.concatenate([ctx.scan_stream_batch(ptrs[:3], counts[:3], True), ctx.scan_stream_batch(ptrs[3:], counts[3:], True)])
    assert poses.shape == (n_scans, n_traj, 7)
    # host-buffer path gives the same bits
    ctx.reset_odometry()
    hptrs = np.array([[trajs[b][k].ctypes.data for b in range(n_traj)] for k in range(n_scans)], np.uint64)
    poses_h = ctx.scan_stream_batch(hptrs, counts, False)
    assert np.array_equal(poses, poses_h)
    ctx.close()
    solo = aloam.Aloam(n_scans=ns, max_points=maxn + 1024, max_ring_points=max_ring)
    for b in range(n_traj if n_traj <= 4 else 3):
        solo.reset_odometry()
        sp, _ = solo.scan_stream([dev[b][k].data_ptr() for k in range(n_scans)], counts[:, b], True)
        assert np.array_equal(sp, poses[:, b]), "trajectory %d differs from its solo run" % b
        ref = _oracle_poses(orc, synth, trajs[b], sensor)
        assert np.abs(poses[:, b, 4:] - ref[:, 4:]).max() < 1e-4
        assert max(rot_angle(poses[k, b, :4], ref[k, :4]) for k in range(n_scans)) < 1e-4
    solo.close()


def test_batch_argument_checks(aloam, synth):
    import torch
    raw = synth.scan("VLP-16", 0, n_az=900)
    d = torch.from_numpy(raw).cuda()
    ctx = aloam.Aloam(n_scans=16, max_points=raw.shape[0] + 1024, max_batch=2)
    with pytest.raises(aloam.AloamError):   # more trajectories than the context was created for
        ctx.scan_stream_batch(np.full((1, 3), d.data_ptr(), np.uint64), np.full((1, 3), raw.shape[0]), True)
    with pytest.raises(aloam.AloamError):   # an empty scan
        ctx.scan_stream_batch(np.full((1, 2), d.data_ptr(), np.uint64), np.array([[raw.shape[0], 0]]), True)
    far = raw.copy(); far[:, :3] *= 1e-3    # everything closer than minimum_range: nothing survives (scanRegistration.cpp:137)
    dfar = torch.from_numpy(far).cuda()
    with pytest.raises(aloam.AloamError) as e:
        ctx.scan_stream_batch(np.array([[d.data_ptr(), dfar.data_ptr()]], np.uint64), np.full((1, 2), raw.shape[0]), True)
    assert e.value.code == -3
    # the context is usable afterwards (state was reset)
    p = ctx.scan_stream_batch(np.full((2, 2), d.data_ptr(), np.uint64), np.full((2, 2), raw.shape[0]), True)
    assert np.isfinite(p).all()
    ctx.close()
    with pytest.raises(aloam.AloamError):
        aloam.Aloam(n_scans=16, max_points=1000, max_batch=17)
    with pytest.raises(aloam.AloamError):
        aloam.Aloam(n_scans=16, max_points=1000, max_ring_points=100)
