#!/usr/bin/env python
"""Phase time stamps (SM clock cycles) of k_ring_features per ring: load, curvature + reach, greedy picks, voxel keys, sort, centroids."""
import ctypes as C
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("a-loam_b200")
synth = importlib.import_module("a-loam_b200.synth")

sensor = sys.argv[1] if len(sys.argv) > 1 else "HDL-64"
ns = synth.SENSORS[sensor][0]
ctx = pkg.Aloam(n_scans=ns, max_points=140000, max_ring_points=int(sys.argv[2]) if len(sys.argv) > 2 else 4096)
for k in range(3):
    ctx.extract_features(synth.scan(sensor, k))
out = (C.c_longlong * (65 * 8))()
pkg.lib().aloam_debug_feature_cycles.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
pkg.lib().aloam_debug_feature_cycles(ctx._h, out)
a = np.array(out[:64 * 8]).reshape(64, 8)[:min(ns, 51)]
names = ["load+curv+reach", "(unused)", "picks", "bbox+keys", "sort", "centroids"]
d = np.diff(a[:, :7], axis=1)
print("phase mean cycles over rings:", {n: int(v) for n, v in zip(names, d.mean(0))}, "total", int((a[:, 6] - a[:, 0]).mean()), "max", int((a[:, 6] - a[:, 0]).max()))
ctx.close()
