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
pkg.lib().aloam_debug_feature_cycles.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
pkg.lib().aloam_debug_feature_cycles(ctx._h, None)   # arm the stamps
for k in range(3):
    ctx.extract_features(synth.scan(sensor, k))
out = (C.c_longlong * (65 * 8 + 16))()
pkg.lib().aloam_debug_feature_cycles(ctx._h, out)
a = np.array(out[:64 * 8]).reshape(64, 8)[:min(ns, 51)]
names = ["load+curv+reach", "(unused)", "picks", "bbox+keys", "sort", "centroids"]
d = np.diff(a[:, :7], axis=1)
pk = np.array(out[64 * 8:65 * 8])
w = np.array(out[65 * 8:65 * 8 + 12]).reshape(6, 2)
print("ring 8 per-warp segment pass: start offset", list(w[:, 0] - a[8, 2]), "duration", list(w[:, 1] - w[:, 0]))
print("ring 8 segment 2 picks: load", pk[1] - pk[0], "sharp walk", pk[2] - pk[1], "picks", pk[4], "-> per selection", (pk[2] - pk[1]) / max(pk[4] + 1, 1), "; flat walk", pk[3] - pk[2], "picks", pk[5])
print("ring 8: speculative pass of the six segments", pk[6] - a[8, 2], "re-run loop", pk[7] - pk[6], "labels + outputs", a[8, 3] - pk[7])
print("phase mean cycles over rings:", {n: int(v) for n, v in zip(names, d.mean(0))}, "total", int((a[:, 6] - a[:, 0]).mean()), "max", int((a[:, 6] - a[:, 0]).max()))
ctx.close()
