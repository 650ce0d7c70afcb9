import importlib, ctypes as C, sys
sys.path.insert(0, "/root/repo")
import numpy as np
pkg = importlib.import_module("a-loam_b200"); synth = importlib.import_module("a-loam_b200.synth")
ctx = pkg.Aloam(n_scans=64, max_points=140000)
L = pkg.lib(); L.aloam_debug_assoc.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
for k in range(4):
    q, t, st = ctx.scan_to_pose(synth.scan("HDL-64", k))
    out = (C.c_int * (2304 * 8))(); L.aloam_debug_assoc(ctx._h, out)
    a = np.array(list(out)).reshape(2304, 8)
    for name, sl in (("corner", slice(0, 768)), ("surf", slice(768, 2304))):
        b = a[sl]; v = b[b[:, 6] > 0]
        if len(v) == 0: continue
        print(k, name, "n", len(v), "median [setup, bucket_of, delims, points, argmin/check, second, total]", np.median(v[:, :7], 0).astype(int), "max", v[:, :7].max(0))
