import importlib, ctypes as C, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
pkg = importlib.import_module("a-loam_b200"); synth = importlib.import_module("a-loam_b200.synth")
ctx = pkg.Aloam(n_scans=64, max_points=140000)
L = pkg.lib()
N = 40
scans = [synth.scan("HDL-64", k) for k in range(N)]
dev = [torch.from_numpy(s).cuda() for s in scans]
def read(fn):
    out = (C.c_ulonglong * (16 * 1024))(); n = (C.c_int * 16)()
    getattr(L, fn)(out, n)
    a = np.array(list(out), dtype=np.uint64).reshape(16, 1024)
    return a, list(n)
# warm
ctx.scan_stream([d.data_ptr() for d in dev[:8]], [d.shape[0] for d in dev[:8]], True)
for fn in ("aloam_tl_features", "aloam_tl_odometry", "aloam_tl_lm"): read(fn)
ctx.reset_odometry()
poses, st = ctx.scan_stream([d.data_ptr() for d in dev], [d.shape[0] for d in dev], True)
F, nf = read("aloam_tl_features"); O, no = read("aloam_tl_odometry"); M, nm = read("aloam_tl_lm")
print("counts", nf[:5], nf[8:11], no[:4], no[8:11], nm[:2], nm[8])
ev = []
names_f = {0: "classify", 1: "ring_scan", 2: "scatter", 3: "ring_features", 4: "compact", 8: "ring_scan.res", 9: "scatter.res", 10: "compact.res"}
names_o = {0: "rab_count", 1: "rab_scan", 2: "rab_fill", 3: "assoc", 8: "rab_scan.res", 9: "rab_fill.res", 10: "assoc.res"}
names_m = {0: "lm", 1: "lm.end", 8: "lm.res"}
for arr, cnt, names in ((F, nf, names_f), (O, no, names_o), (M, nm, names_m)):
    for kid, nm_ in names.items():
        for i in range(min(cnt[kid], 1024)):
            ev.append((int(arr[kid, i]), nm_, i))
ev.sort()
t0 = ev[0][0]
# print window around scan 20: find the 20th ring_features start
rf = [e for e in ev if e[1] == "ring_features"]
lo = rf[20][0] - 5000; hi = rf[23][0]
for t, nm_, i in ev:
    if lo <= t <= hi and not nm_.endswith(".res"):
        print("%9.1f us  %-14s #%d" % ((t - lo) / 1000.0, nm_, i))
# periods
for nm_ in ("classify", "ring_features", "rab_count", "assoc", "lm"):
    ts = np.array([e[0] for e in ev if e[1] == nm_], dtype=np.float64)
    d = np.diff(ts) / 1000.0
    print(nm_, "period median %.1f us" % np.median(d[len(d)//2:]))
lm_s = np.array([e[0] for e in ev if e[1] == "lm"], dtype=np.float64); lm_e = np.array([e[0] for e in ev if e[1] == "lm.end"], dtype=np.float64)
n = min(len(lm_s), len(lm_e)); print("lm start->finish median %.1f us" % np.median((lm_e[:n] - lm_s[:n]) / 1000.0))
as_ = np.array([e[0] for e in ev if e[1] == "assoc"], dtype=np.float64)
print("assoc start -> lm start median %.1f us" % np.median((lm_s[:len(as_)] - as_[:len(lm_s)]) / 1000.0))
