import sys, importlib
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import pyoracle as orc
pkg = importlib.import_module("a-loam_b200"); synth = importlib.import_module("a-loam_b200.synth")
from test_gpu_mapping import _rot
sensor = "HDL-64"; ns, az, mr, lres, pres = synth.SENSORS[sensor]
corner, surf = [], []
for k in [0, 1, 3, 4]:
    f = orc.Features(synth.scan(sensor, k), ns, mr); qk, tk = synth.pose(k); R = _rot(qk)
    for src, dst in [(f.less_sharp, corner), (f.less_flat, surf)]:
        w = src.copy(); w[:, :3] = (src[:, :3].astype(np.float64) @ R.T + tk).astype(np.float32); dst.append(w)
cmap = orc.voxel_grid(np.concatenate(corner), lres); smap = orc.voxel_grid(np.concatenate(surf), pres)
f2 = orc.Features(synth.scan(sensor, 2), ns, mr)
cs, ss = orc.voxel_grid(f2.less_sharp, lres), orc.voxel_grid(f2.less_flat, pres)
q2, t2 = synth.pose(2); x0 = np.concatenate([q2, t2 + np.array([0.05, -0.04, 0.02])])
c = pkg.Aloam(n_scans=64, max_points=200000, max_map_points=400000); c.map_upload(cmap, smap)
m = orc.Mapping(); m.set_map(cmap, smap)
fits, blocks = m.associate(cs, ss, x0)
got = c.mapping_associate(cs, ss, x0); acc = got[got[:, 1] >= 0]
e = np.where(fits[:, 1] == 0)[0][:3]
np.set_printoptions(precision=6, suppress=True, linewidth=200)
for i in e:
    print("oracle", fits[i, :9]); print("gpu   ", acc[i, :9])
    P = cmap[fits[i, 9:].astype(int), :3].astype(np.float64); cc = P.mean(0); ev, V = np.linalg.eigh((P - cc).T @ (P - cc)); print("numpy centre", cc, "dir", V[:, 2], "ev", ev)
blocks_e = blocks[blocks[:, 0] == 0][:2]
print("oracle blocks", blocks_e)
