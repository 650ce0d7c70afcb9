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
pk = np.array(out[64 * 8:65 * 8])
w = np.array(out[65 * 8:65 * 8 + 12]).reshape(6, 2)
print("ring 8 per-warp segment pass: start offset", list(w[:, 0] - a[8, 2]), "duration", list(w[:, 1] - w[:, 0]))
print("ring 8 segment 2 picks: load", pk[1] - pk[0], "sharp walk", pk[2] - pk[1], "picks", pk[4], "-> per selection", (pk[2] - pk[1]) / max(pk[4] + 1, 1), "; flat walk", pk[3] - pk[2], "picks", pk[5])
print("ring 8: speculative pass of the six segments", pk[6] - a[8, 2], "re-run loop", pk[7] - pk[6], "labels + outputs", a[8, 3] - pk[7])
# A ring is a pair of CTAs on two SMs (clock64 is per SM, so the two rows are not comparable with each other):
#   picks CTA: [0] start, [1] load + curvature + reach done, [3] picks + labels done
#   voxel CTA: [7] start, [4] load + bounding box done, [5] sort done, [6] labels received + centroids written
picks = {"load+curv+reach": a[:, 1] - a[:, 0], "picks+labels": a[:, 3] - a[:, 1], "total": a[:, 3] - a[:, 0]}
voxel = {"load+bbox": a[:, 4] - a[:, 7], "sort": a[:, 5] - a[:, 4], "wait for labels + centroids": a[:, 6] - a[:, 5], "total": a[:, 6] - a[:, 7]}
print("picks CTA mean cycles over rings:", {k: int(v.mean()) for k, v in picks.items()}, "max total", int(picks["total"].max()))
print("voxel CTA mean cycles over rings:", {k: int(v.mean()) for k, v in voxel.items()}, "max total", int(voxel["total"].max()))
ctx.close()
