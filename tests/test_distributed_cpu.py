"""CPU, gloo, world_size 2: the host-side logic of the multi-GPU scan-to-map path.

* the slab + halo split of a-loam_b200/shard.py: a rank's shard reproduces the global 5-NN of every query it owns whenever
  the reference would accept the neighbourhood (5th distance < 1 m);
* the per-rank normal equations, summed with ONE all-reduce of 28 doubles, equal the unsharded normal equations;
* bench.py's replica bookkeeping (max over ranks, summed scans)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _scene():
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import pyoracle as orc
    synth = importlib.import_module("a-loam_b200.synth")
    ns, az, mr, lres, pres = synth.SENSORS["VLP-16"]
    corner, surf = [], []
    for k in [0, 1, 3, 4]:
        f = orc.Features(synth.scan("VLP-16", k), ns, mr)
        qk, tk = synth.pose(k); R = _rot(qk)
        for src, dst in [(f.less_sharp, corner), (f.less_flat, surf)]:
            w = src.copy(); w[:, :3] = (src[:, :3].astype(np.float64) @ R.T + tk).astype(np.float32); dst.append(w)
    cmap = orc.voxel_grid(np.concatenate(corner), lres); smap = orc.voxel_grid(np.concatenate(surf), pres)
    f2 = orc.Features(synth.scan("VLP-16", 2), ns, mr)
    cs, ss = orc.voxel_grid(f2.less_sharp, lres), orc.voxel_grid(f2.less_flat, pres)
    q2, t2 = synth.pose(2)
    return orc, cmap, smap, cs, ss, np.concatenate([q2, t2 + np.array([0.04, -0.03, 0.02])])


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc, cmap, smap, cs, ss, x0 = _scene()
    shard = importlib.import_module("a-loam_b200.shard")
    R = _rot(x0[:4])

    def owned(stack):
        w = (stack[:, :3].astype(np.float64) @ R.T + x0[4:]).astype(np.float32)   # pointAssociateToMap, float store
        return shard.owner_of_cell(shard.cell_x(w[:, 0]), world) == rank
    m = orc.Mapping()
    m.set_map(shard.shard_cloud(cmap, rank, world), shard.shard_cloud(smap, rank, world))
    fits, blocks = m.associate(cs[owned(cs)], ss[owned(ss)], x0)
    JtJ, Jtr, cost = orc.normal_equations(blocks, x0) if len(blocks) else (np.zeros((6, 6)), np.zeros(6), 0.0)
    v = torch.tensor(np.concatenate([JtJ[np.triu_indices(6)], Jtr, [cost], [len(blocks)]]))   # 21 + 6 + 1 (+ count)
    dist.all_reduce(v)
    # bench bookkeeping: max over ranks of the elapsed time, sum of the scans
    t = torch.tensor([1.0 + rank]); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        np.save(out, np.concatenate([v.numpy(), t.numpy()]))
    dist.destroy_process_group()


def test_sharded_normal_equations_equal_unsharded(tmp_path):
    out = str(tmp_path / "reduced.npy")
    mp.spawn(_worker, args=(2, 29733, out), nprocs=2, join=True)
    got = np.load(out)
    orc, cmap, smap, cs, ss, x0 = _scene()
    m = orc.Mapping(); m.set_map(cmap, smap)
    fits, blocks = m.associate(cs, ss, x0)
    JtJ, Jtr, cost = orc.normal_equations(blocks, x0)
    ref = np.concatenate([JtJ[np.triu_indices(6)], Jtr, [cost]])
    assert got[28] == len(blocks) and len(blocks) > 300                   # every accepted query was owned by exactly one rank
    assert np.abs(got[:28] - ref).max() <= 1e-12 * np.abs(ref).max()      # identical to ~1e-12 relative
    assert got[29] == 2.0                                                 # max over ranks


def test_shard_halo_covers_accepted_neighbourhoods():
    orc, cmap, smap, cs, ss, x0 = _scene()
    shard = importlib.import_module("a-loam_b200.shard")
    R = _rot(x0[:4])
    w = ss.copy(); w[:, :3] = (ss[:, :3].astype(np.float64) @ R.T + x0[4:]).astype(np.float32)
    gidx, gsqd = orc.bruteforce_knn(smap, w, 5)
    for world in (2, 3, 8):
        total = 0
        for rank in range(world):
            mask = shard.shard_mask(smap, rank, world)
            local = np.flatnonzero(mask)
            mine = shard.owner_of_cell(shard.cell_x(w[:, 0]), world) == rank
            lidx, lsqd = orc.bruteforce_knn(smap[mask], w[mine], 5)
            acc = gsqd[mine][:, 4] < 1.0
            assert np.array_equal(local[lidx[acc]], gidx[mine][acc]) and np.array_equal(lsqd[acc], gsqd[mine][acc])
            assert np.all(lsqd[~acc][:, 4] >= 1.0)                       # rejected globally => rejected locally too
            total += mine.sum()
        assert total == len(w)
