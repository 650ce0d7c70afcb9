"""CPU: the format helpers of include/aloam_io.h against independent numpy restatements of
kittiHelper.cpp:25-35 (scan files), :78-80,97-113 (ground-truth poses through float, camera -> lidar frame) and of
the pcl::PointXYZI PointCloud2 payload."""
import importlib
import struct

import numpy as np
import pytest

io = importlib.import_module("a-loam_b200.io")

R_T = np.array([[0, 0, 1], [-1, 0, 0], [0, -1, 0]], float)   # kittiHelper.cpp:78-79


def quat_to_mat(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def test_kitti_bin_round_trip(tmp_path, synth):
    scan = synth.scan("HDL-64", 3)
    scan[:, 3] = np.linspace(0, 1, len(scan), dtype=np.float32)
    p = tmp_path / "000003.bin"
    io.write_kitti_bin(p, scan)
    got = io.read_kitti_bin(p)
    assert got.dtype == np.float32 and np.array_equal(got, scan)
    # a trailing partial point is dropped exactly like `for (i = 0; i < size; i += 4)` would stop short of it ... the
    # reference would in fact read past the buffer; we stop at whole points
    with open(p, "ab") as f:
        f.write(struct.pack("<2f", 1.0, 2.0))
    assert len(io.read_kitti_bin(p)) == len(scan)
    with pytest.raises(OSError):
        io.read_kitti_bin(tmp_path / "missing.bin")


def test_kitti_pose_goes_through_float_and_axis_swap():
    rng = np.random.default_rng(7)
    for _ in range(50):
        # a random rotation (all four branches of the matrix -> quaternion conversion get exercised) and translation
        a = rng.normal(size=4); a /= np.linalg.norm(a)
        R = quat_to_mat(a)
        t = rng.normal(size=3) * 100
        T = np.hstack([R, t[:, None]])
        line = " ".join("%.9e" % v for v in T.reshape(-1))
        Tp = io.parse_kitti_pose(line)
        assert np.array_equal(Tp, T.astype(np.float32).astype(np.float64))      # stof(): float precision, kittiHelper.cpp:106
        q, tl = io.kitti_pose_to_lidar(Tp)
        assert abs(np.linalg.norm(q) - 1) < 1e-12
        assert np.abs(quat_to_mat(q) - R_T @ Tp[:, :3]).max() < 1e-6            # float-parsed R is only orthonormal to 1e-7
        assert np.abs(tl - R_T @ Tp[:, 3]).max() < 1e-9
        back = io.lidar_pose_to_kitti(q, tl)
        assert np.abs(back[:, :3] - Tp[:, :3]).max() < 1e-6 and np.abs(back[:, 3] - Tp[:, 3]).max() < 1e-9
    # identity ground truth -> the pure axis swap: q_transform = Quaterniond(R_transform) = (0.5, -0.5, 0.5, -0.5)
    q, tl = io.kitti_pose_to_lidar(np.hstack([np.eye(3), np.zeros((3, 1))]))
    assert np.allclose(q, [0.5, -0.5, 0.5, -0.5]) and np.allclose(tl, 0)
    with pytest.raises(ValueError):
        io.parse_kitti_pose("1 0 0 0 0 1 0")


def test_pointcloud2_pointxyzi_layout():
    rng = np.random.default_rng(3)
    pts = rng.normal(size=(1000, 4)).astype(np.float32)
    data = io.pack_pointxyzi(pts)
    assert data.shape == (32000,)
    rec = data.view(np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("pad0", "<f4"), ("i", "<f4"), ("pad1", "<f4", 3)]))
    assert np.array_equal(rec["x"], pts[:, 0]) and np.array_equal(rec["z"], pts[:, 2]) and np.array_equal(rec["i"], pts[:, 3])
    assert not rec["pad0"].any() and not rec["pad1"].any()
    assert np.array_equal(io.unpack_points(data, 1000), pts)
    # the same payload is a stride-8 aloam_cloud_view: no repacking between a ROS message and the C ABI
    assert np.array_equal(data.view(np.float32).reshape(-1, 8)[:, [0, 1, 2, 4]], pts)
    # a velodyne-driver style layout: x, y, z, intensity, ring(u16), point_step 22 ; and a cloud without intensity
    raw = np.zeros((1000, 22), np.uint8)
    raw[:, :16] = pts.view(np.uint8).reshape(1000, 16)
    assert np.array_equal(io.unpack_points(raw.reshape(-1), 1000, point_step=22, off_intensity=12), pts)
    noi = io.unpack_points(raw.reshape(-1), 1000, point_step=22, off_intensity=-1)
    assert np.array_equal(noi[:, :3], pts[:, :3]) and not noi[:, 3].any()
    with pytest.raises(ValueError):
        io.unpack_points(raw.reshape(-1), 1000, point_step=22, off_intensity=20)
