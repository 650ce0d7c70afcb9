"""Generates the committed golden vectors from the CPU oracle (oracle/).  The reference ships none (no tests, no
fixtures) and cannot be run here, so these pin the ORACLE against regressions -- they are not reference outputs.
Run from the repo root:  python tests/golden/make_golden.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle as orc  # noqa
synth = importlib.import_module("a-loam_b200.synth")
OUT = os.path.dirname(os.path.abspath(__file__))

raw0 = synth.scan("VLP-16", 0, n_az=360)
raw1 = synth.scan("VLP-16", 1, n_az=360)
f0 = orc.Features(raw0, 16, 0.3, orc.SORT_CANONICAL)
f1 = orc.Features(raw1, 16, 0.3, orc.SORT_CANONICAL)
np.savez_compressed(os.path.join(OUT, "features_vlp16_az360.npz"), raw=raw0, full=f0.full, sharp=f0.sharp, less_sharp=f0.less_sharp,
                    flat=f0.flat, less_flat=f0.less_flat, label=f0.label, curvature=f0.curvature)
od = orc.Odometry()
od.set_last(f0.less_sharp, f0.less_flat)
q0 = np.array([0, 0, 0, 1.0]); t0 = np.zeros(3)
cc, pc, bl = od.associate(f1.sharp, f1.flat, q0, t0)
q, t, _ = od.register(f1.sharp, f1.flat, q0, t0)
np.savez_compressed(os.path.join(OUT, "odometry_vlp16_az360.npz"), corner_last=f0.less_sharp, surf_last=f0.less_flat, sharp=f1.sharp,
                    flat=f1.flat, q0=q0, t0=t0, corner_corr=cc, plane_corr=pc, q=q, t=t)
# LM golden on an HDL-64 pair's first association
g0 = orc.Features(synth.scan("HDL-64", 0, n_az=500), 64, 5.0)
g1 = orc.Features(synth.scan("HDL-64", 1, n_az=500), 64, 5.0)
od2 = orc.Odometry(); od2.set_last(g0.less_sharp, g0.less_flat)
_, _, blocks = od2.associate(g1.sharp, g1.flat, q0, t0)
x0 = np.concatenate([q0, t0])
x, s, trace = orc.solve(blocks, x0, max_iters=4)
JtJ, Jtr, cost = orc.normal_equations(blocks, x0)
np.savez_compressed(os.path.join(OUT, "solve_hdl64_pair.npz"), blocks=blocks, x0=x0, x=x, trace=trace, JtJ=JtJ, Jtr=Jtr, cost=cost)
print("golden written:", [f for f in os.listdir(OUT) if f.endswith(".npz")])
