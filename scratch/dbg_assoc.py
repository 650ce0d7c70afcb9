import sys, importlib, ctypes as C
sys.path.insert(0, "/root/repo")
import numpy as np
pkg = importlib.import_module("a-loam_b200"); synth = importlib.import_module("a-loam_b200.synth")
ctx = pkg.Aloam(n_scans=64, max_points=140000)
L = pkg.lib(); L.aloam_debug_assoc_cycles.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
for k in range(12):
    ctx.scan_to_pose(synth.scan("HDL-64", k))
buf = (C.c_longlong * (4096 * 4))(); L.aloam_debug_assoc_cycles(ctx._h, buf)
a = np.array(list(buf)).reshape(4096, 4)[:2304]
ok = a[:, 0] > 0
a = a[ok]
t0 = a[:, 0].min()
print("queries recorded", len(a), " kernel span (first start -> last end) cycles", a[:, 2].max() - t0)
nn = a[:, 1] - a[:, 0]; cl = a[:, 2] - a[:, 1]
print("start offsets pct 0/50/90/100:", np.percentile(a[:, 0] - t0, [0, 50, 90, 100]).astype(int))
print("NN cycles    pct 50/90/99/100:", np.percentile(nn, [50, 90, 99, 100]).astype(int))
print("class cycles pct 50/90/99/100:", np.percentile(cl, [50, 90, 99, 100]).astype(int))
print("end offsets  pct 50/90/99/100:", np.percentile(a[:, 2] - t0, [50, 90, 99, 100]).astype(int))
sm = a[:, 3]; print("warps per SM max", np.bincount(sm.astype(int)).max(), "SMs used", len(np.unique(sm)))
