#!/usr/bin/env python
"""SM-clock cycle split of the scan-to-scan LM solves (k_lm_solve) of the last aloam_scan_to_pose call:
whole solve / evaluation passes (residual blocks, reductions, cluster barriers) / trust-region step on thread 0.
usage: python tools/lm_phases.py [sensor]"""
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
ctx = pkg.Aloam(n_scans=ns, max_points=140000)
lib = pkg.lib()
lib.aloam_debug_lm_cycles.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.c_int]
rows = []
for k in range(6):
    q, t, st = ctx.scan_to_pose(synth.scan(sensor, k))
    if k == 0:
        continue
    out = (C.c_longlong * 20)()
    lib.aloam_debug_lm_cycles(ctx._h, out, 2)
    for it in range(2):
        tot, ev, tr, blocks, barrier = [int(out[5 * it + j]) for j in range(5)]
        rows.append((tot, ev, tr, blocks, barrier, st["lm_iters"]))
        print("scan %d solve %d: total %6d  eval passes %6d (residual blocks %6d, cluster barriers %6d, reductions %6d)  trust-region steps %6d  rest %5d   [lm_iters of the scan %s]"
              % (k, it, tot, ev, blocks, barrier, ev - blocks - barrier, tr, tot - ev - tr, st["lm_iters"]))
a = np.array([r[:5] for r in rows], float).mean(0)
print("mean: total %.0f = eval %.0f (blocks %.0f + barriers %.0f + reductions %.0f) + trust region %.0f + rest %.0f"
      % (a[0], a[1], a[3], a[4], a[1] - a[3] - a[4], a[2], a[0] - a[1] - a[2]))
ctx.close()
