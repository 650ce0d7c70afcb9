import sys, importlib, ctypes as C
sys.path.insert(0, "/root/repo")
import numpy as np
pkg = importlib.import_module("a-loam_b200"); synth = importlib.import_module("a-loam_b200.synth")
ctx = pkg.Aloam(n_scans=64, max_points=140000)
L = pkg.lib(); L.aloam_debug_lm_cycles.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.c_int]
L.aloam_debug_feature_cycles.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
for k in range(4):
    q, t, st = ctx.scan_to_pose(synth.scan("HDL-64", k))
    out = (C.c_longlong * 20)(); L.aloam_debug_lm_cycles(ctx._h, out, 2)
    print(k, st["lm_iters"], st["accepted_steps"], "cycles [total, eval, chol, plus, grad] x2:", list(out)[:10])
f = (C.c_longlong * 512)(); L.aloam_debug_feature_cycles(ctx._h, f)
a = np.array(list(f)).reshape(64, 8)
d = np.diff(a[:51, :7], axis=1)
print("feature phases (cycles) load+curv, sort1, greedy, bbox, sort2, heads+sums: median", np.median(d, 0).astype(int), "max", d.max(0))
print("total per ring median/max", np.median(a[:51, 6] - a[:51, 0]), (a[:51, 6] - a[:51, 0]).max())
