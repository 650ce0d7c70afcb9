#!/usr/bin/env python
"""Scan-to-map steps only (map upload = index build, then register) for ncu captures of the mapping kernels.
usage: python tools/prof_mapping.py [total_map_points] [steps]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

pkg = importlib.import_module("a-loam_b200")
synth = importlib.import_module("a-loam_b200.synth")


def main():
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    fctx = pkg.Aloam(n_scans=64, max_points=200000)

    def feats(raw):
        f = fctx.extract_features(raw)
        return f["less_sharp"], f["less_flat"]
    cmap, smap = synth.build_map(feats, total)
    ctx = pkg.Aloam(n_scans=64, max_points=200000, max_map_points=max(len(cmap), len(smap)) + 1024)
    dc, ds = torch.from_numpy(cmap).cuda(), torch.from_numpy(smap).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for i in range(steps):
        k = synth.MAP_QUERY_SCANS[i % len(synth.MAP_QUERY_SCANS)]
        f = fctx.extract_features(synth.scan("HDL-64", k))
        cs, ss = ctx.voxel_filter(f["less_sharp"], 0.4), ctx.voxel_filter(f["less_flat"], 0.8)
        q, t = synth.pose(k)
        flush.fill_(i)
        torch.cuda.synchronize()
        ctx.map_upload_ptr(dc.data_ptr(), dc.shape[0], ds.data_ptr(), ds.shape[0])
        x, st = ctx.mapping_register(cs, ss, np.concatenate([q, t + np.array([0.05, -0.04, 0.02])]))
        print("step", i, "queries", len(cs) + len(ss), "pose err", np.abs(x[4:] - t).max(), st["lm_iters"])
    ctx.close(); fctx.close()


if __name__ == "__main__":
    main()
