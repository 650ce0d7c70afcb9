"""Seeded synthetic lidar scans (VLP-16 / HDL-32 / HDL-64 shaped) -- ctypes binding of synth.cc."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SENSORS = {  # name -> (n_scans, azimuth steps, minimum_range [launch/*.launch], line_res, plane_res)
    "VLP-16": (16, 1800, 0.3, 0.2, 0.4),
    "HDL-32": (32, 2200, 0.3, 0.2, 0.4),
    "HDL-64": (64, 2000, 5.0, 0.4, 0.8),
}
BASE_SEED = 20240901


def build(force=False):
    so = os.path.join(_HERE, "libaloam_synth.so")
    src = os.path.join(_HERE, "synth.cc")
    if force or not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
        subprocess.check_call(["g++", "-std=c++17", "-O3", "-fPIC", "-shared", "-o", so, src, "-lpthread"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libaloam_synth.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.synth_scan.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                 C.POINTER(C.c_float), C.c_int]
        L.synth_pose.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _LIB = L
    return _LIB


def scan(sensor, index, seed=BASE_SEED, n_az=None, noise=0.02, max_range=120.0):
    """Raw scan `index` of the trajectory: (n, 4) float32 x,y,z,0 in the sensor frame, firing order."""
    n_scans, az = SENSORS[sensor][0], SENSORS[sensor][1]
    if n_az is not None:
        az = n_az
    out = np.zeros((n_scans * az, 4), np.float32)
    n = _lib().synth_scan(seed, n_scans, az, index, noise, max_range, out.ctypes.data_as(C.POINTER(C.c_float)), out.shape[0])
    if n < 0:
        raise RuntimeError("synth_scan failed: %d" % n)
    return out[:n].copy()


def pose(index):
    """Ground-truth world pose of scan `index`: (q xyzw, t)."""
    q = np.zeros(4)
    t = np.zeros(3)
    _lib().synth_pose(index, q.ctypes.data_as(C.POINTER(C.c_double)), t.ctypes.data_as(C.POINTER(C.c_double)))
    return q, t


def voxel_downsample(cloud, leaf):
    """Voxel-centroid downsample used to BUILD synthetic maps (numpy; same grid rule as pcl::VoxelGrid, float64 sums --
    it only has to produce a plausible map, the parity-relevant VoxelGrid lives in the oracle / CUDA path)."""
    cloud = np.ascontiguousarray(cloud, np.float32)
    if len(cloud) == 0:
        return cloud.copy()
    ijk = np.floor(cloud[:, :3] / np.float32(leaf)).astype(np.int64)
    ijk -= ijk.min(0)
    dims = ijk.max(0) + 1
    key = ijk[:, 0] + dims[0] * (ijk[:, 1] + dims[1] * ijk[:, 2])
    order = np.argsort(key, kind="stable")
    ks = key[order]
    starts = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1]])
    sums = np.add.reduceat(cloud[order].astype(np.float64), starts, axis=0)
    counts = np.diff(np.r_[starts, len(ks)])[:, None]
    return (sums / counts).astype(np.float32)


def rotation_matrix(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


MAP_SCANS = list(range(0, 24))        # scans whose features make the base map
MAP_QUERY_SCANS = list(range(24, 40))  # scans registered against it (not part of the map)


def build_map(features, total_points, sensor="HDL-64", seed=BASE_SEED + 3, line_res=0.4, plane_res=0.8):
    """Synthetic voxel map for the scan-to-map configs: union of the less-sharp / less-flat features of
    scans MAP_SCANS at their true poses, voxel-filtered at line_res / plane_res, then replicated with 2 cm jitter on a
    lattice of parallel streets (y pitch 30 m, 7 columns) and stacked levels (z pitch 8 m, 18 levels) INSIDE the
    250 x 250 x 150 m volume of the reference's 5 x 5 x 3-cube submap (laserMapping.cpp:512-529) to exactly
    total_points // 5 corner + the rest surf points.  The tile at offset 0 (the one the query scans see) is a clean
    filtered map; if one lattice of tiles is not enough the other tiles are added again with an x shift and fresh jitter.
    `features(raw) -> (less_sharp, less_flat)` is the extraction to use (the product's or the oracle's)."""
    corner, surf = [], []
    for k in MAP_SCANS:
        ls, lf = features(scan(sensor, k))
        qk, tk = pose(k)
        R = rotation_matrix(qk)
        for src, dst in ((ls, corner), (lf, surf)):
            w = src.copy()
            w[:, :3] = (src[:, :3].astype(np.float64) @ R.T + tk).astype(np.float32)
            dst.append(w)
    cbase = voxel_downsample(np.concatenate(corner), line_res)
    sbase = voxel_downsample(np.concatenate(surf), plane_res)
    rng = np.random.default_rng(seed)
    offsets = [(0.0, 0.0, 0.0)]
    for lvl in sorted(range(-9, 9), key=abs):
        for col in sorted(range(-3, 4), key=abs):
            if lvl or col:
                offsets.append((0.0, 30.0 * col, 8.0 * lvl))

    def tile(base, target):
        out, n, rounds = [base], len(base), 0
        while n < target:
            for ox, oy, oz in offsets[1:]:
                c = base.copy()
                c[:, :3] += np.array([ox + 0.37 * rounds, oy, oz], np.float32) + rng.normal(0, 0.02, (len(base), 3)).astype(np.float32)
                out.append(c)
                n += len(c)
                if n >= target:
                    break
            rounds += 1
        return np.ascontiguousarray(np.concatenate(out)[:target])

    return tile(cbase, total_points // 5), tile(sbase, total_points - total_points // 5)
