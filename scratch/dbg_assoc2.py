import importlib, ctypes as C, sys
sys.path.insert(0, "/root/repo")
import numpy as np
pkg = importlib.import_module("a-loam_b200"); synth = importlib.import_module("a-loam_b200.synth")
ctx = pkg.Aloam(n_scans=64, max_points=140000)
L = pkg.lib(); L.aloam_debug_assoc.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
for k in range(4):
    q, t, st = ctx.scan_to_pose(synth.scan("HDL-64", k))
    out = (C.c_int * (2304 * 4))(); L.aloam_debug_assoc(ctx._h, out)
    a = np.array(list(out)).reshape(2304, 4)
    for name, sl in (("corner", slice(0, 768)), ("surf", slice(768, 2304))):
        b = a[sl]; v = b[b[:, 0] > 0]
        if len(v) == 0: continue
        print(k, name, "n", len(v), "nn cycles p50/p90/p99/max", np.percentile(v[:, 0], [50, 90, 99, 100]).astype(int),
              "2nd p50/p90/p99/max", np.percentile(v[:, 1], [50, 90, 99, 100]).astype(int), "total p50/max", np.percentile(v[:, 2], [50, 100]).astype(int))
    v = a[a[:, 0] > 0]
    kf = v[:, 3] // 1000; rho = (v[:, 3] % 1000) / 10.0
    print("  k histogram", np.bincount(kf)[:12])
    for lo, hi in ((0, 6), (6, 8), (8, 12), (12, 20), (20, 100)):
        m = (rho >= lo) & (rho < hi)
        if m.sum(): print("  rho [%d,%d): n %d, k hist %s, nn cycles median %d" % (lo, hi, m.sum(), np.bincount(kf[m])[:10], np.median(v[m, 0])))
