#!/usr/bin/env python
"""A few frames of the full pipeline through the synchronous API (aloam_scan_to_pose + aloam_mapper_step) for ncu launch lists.
usage: python tools/prof_mapped.py [frames]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("a-loam_b200")
synth = importlib.import_module("a-loam_b200.synth")


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    raws = [synth.scan("HDL-64", k) for k in range(frames)]
    ctx = pkg.Aloam(n_scans=64, max_points=max(r.shape[0] for r in raws) + 1024, max_map_points=600000)
    ctx.mapper_reset()
    for k, raw in enumerate(raws):
        q, t, _ = ctx.scan_to_pose(raw)
        f = ctx.extract_features(raw)
        mq, mt, st = ctx.mapper_step(f["less_sharp"], f["less_flat"], q, t)
        print(k, mt, st["lm_iters"], st["flags"], ctx.mapper_state()["total_surf"])
    ctx.close()


if __name__ == "__main__":
    main()
