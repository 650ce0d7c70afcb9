"""Launched by tests/test_gpu_multi.py under torchrun (one rank per GPU): sharded scan-to-map registration must give the
pose of the unsharded run.  Each rank uploads only its shard of the map (owned slabs + one-cell halo), fits only the
stack points it owns, and the ranks meet in one ncclAllReduce of the normal equations per evaluation."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa
import torch.distributed as dist  # noqa

pkg = importlib.import_module("a-loam_b200")
synth = importlib.import_module("a-loam_b200.synth")
shard = importlib.import_module("a-loam_b200.shard")


def rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = pkg.Aloam(n_scans=64, device=local, max_points=200000, max_map_points=600000)
    # map from the product's own feature extraction of scans 0,1,3,4 at their true poses (no oracle on this path)
    corner, surf = [], []
    for k in [0, 1, 3, 4]:
        f = ctx.extract_features(synth.scan("HDL-64", k))
        qk, tk = synth.pose(k); R = rot(qk)
        for src, dst in [(f["less_sharp"], corner), (f["less_flat"], surf)]:
            w = src.copy(); w[:, :3] = (src[:, :3].astype(np.float64) @ R.T + tk).astype(np.float32); dst.append(w)
    cmap = synth.voxel_downsample(np.concatenate(corner), 0.4)
    smap = synth.voxel_downsample(np.concatenate(surf), 0.8)
    f2 = ctx.extract_features(synth.scan("HDL-64", 2))
    cs, ss = synth.voxel_downsample(f2["less_sharp"], 0.4), synth.voxel_downsample(f2["less_flat"], 0.8)
    q2, t2 = synth.pose(2)
    x0 = np.concatenate([q2, t2 + np.array([0.05, -0.04, 0.02])])
    # unsharded reference on every rank
    ctx.map_upload(cmap, smap)
    x_ref, st_ref = ctx.mapping_register(cs, ss, x0)
    # sharded
    idb = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idb = torch.tensor(list(pkg.Aloam.comm_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(idb, 0)
    ctx.comm_init(rank, world, bytes(idb.cpu().tolist()))
    path = "NVLink peer memory inside the LM kernel" if ctx.comm_uses_peer_memory() else "ncclAllReduce between per-evaluation kernels"
    ctx.map_upload(shard.shard_cloud(cmap, rank, world), shard.shard_cloud(smap, rank, world))
    x_sh, st_sh = ctx.mapping_register(cs, ss, x0)
    # the same with the split done on the device from the whole map (aloam_map_upload_sharded), host and device views
    ctx.map_upload_sharded(cmap, smap)
    x_dv, _ = ctx.mapping_register(cs, ss, x0)
    dc, dsf = torch.from_numpy(np.ascontiguousarray(cmap)).cuda(), torch.from_numpy(np.ascontiguousarray(smap)).cuda()
    ctx.map_upload_sharded_ptr(dc.data_ptr(), dc.shape[0], dsf.data_ptr(), dsf.shape[0])
    x_dd, _ = ctx.mapping_register(cs, ss, x0)
    split_ok = bool(np.array_equal(x_dv, x_sh) and np.array_equal(x_dd, x_sh))
    dt = float(np.abs(x_sh[4:] - x_ref[4:]).max()); dq = float(abs(abs(float(x_sh[:4] @ x_ref[:4])) - 1.0))
    counts = torch.tensor([st_sh["n_corner_corr"], st_sh["n_plane_corr"]], device="cuda")
    if rank == 0:
        print("exchange path: %s; device-side split equals host split: %s" % (path, split_ok), flush=True)
    print("rank %d/%d: shard map %d+%d of %d+%d pts, blocks total %s (unsharded %d+%d), |dt| %.3e, 1-|dq| %.3e, lm_iters %d vs %d"
          % (rank, world, int(shard.shard_mask(cmap, rank, world).sum()), int(shard.shard_mask(smap, rank, world).sum()), len(cmap), len(smap),
             counts.tolist(), st_ref["n_corner_corr"], st_ref["n_plane_corr"], dt, dq, st_sh["lm_iters"], st_ref["lm_iters"]), flush=True)
    ok = dt < 1e-9 and dq < 1e-12 and st_sh["lm_iters"] == st_ref["lm_iters"] and counts.tolist() == [st_ref["n_corner_corr"], st_ref["n_plane_corr"]] and split_ok
    # ---- the whole mapping loop on a rank of a sharded job: cube store replicated, submap index + 5-NN + normal equations sharded
    # (aloam_scan_stream_mapped = odometry -> aloam_mapper_step per scan on the device) against a context without a communicator
    raws = [synth.scan("HDL-64", k) for k in range(5)]
    single = pkg.Aloam(n_scans=64, device=local, max_points=200000, max_map_points=600000)
    o_ref, m_ref = single.scan_stream_mapped([r.ctypes.data for r in raws], [r.shape[0] for r in raws], False)
    ms_ref = single.mapper_state()
    single.close()
    ctx.reset_odometry()
    ctx.mapper_reset()
    o_sh, m_sh = ctx.scan_stream_mapped([r.ctypes.data for r in raws], [r.shape[0] for r in raws], False)
    ms_sh = ctx.mapper_state()
    loop_dt = float(np.abs(m_sh[:, 4:] - m_ref[:, 4:]).max())
    loop_dq = float(np.abs(np.abs(np.sum(m_sh[:, :4] * m_ref[:, :4], axis=1)) - 1.0).max())
    moved = float(np.abs(m_ref[:, 4:] - o_ref[:, 4:]).max())   # the refinement did something
    loop_ok = (np.array_equal(o_sh, o_ref) and loop_dt < 1e-9 and loop_dq < 1e-12 and ms_sh["centre"] == ms_ref["centre"] and ms_sh["valid"] == ms_ref["valid"]
               and abs(ms_sh["total_corner"] - ms_ref["total_corner"]) <= 5 and abs(ms_sh["total_surf"] - ms_ref["total_surf"]) <= 20 and moved > 0)
    print("rank %d/%d: mapping loop over %d scans, sharded vs single GPU: |dt| %.3e, 1-|dq| %.3e, cube store %d+%d vs %d+%d pts, refinement moved the pose by %.3e m"
          % (rank, world, len(raws), loop_dt, loop_dq, ms_sh["total_corner"], ms_sh["total_surf"], ms_ref["total_corner"], ms_ref["total_surf"], moved), flush=True)
    ok = ok and loop_ok
    ml = torch.tensor(m_sh.reshape(-1), device="cuda"); mlo = ml.clone(); mhi = ml.clone()
    dist.all_reduce(mlo, op=dist.ReduceOp.MIN); dist.all_reduce(mhi, op=dist.ReduceOp.MAX)
    ok = ok and bool(torch.equal(mlo, mhi))   # every rank refined to the identical poses
    # all ranks must hold the identical pose
    xs = torch.tensor(x_sh, device="cuda"); lo = xs.clone(); hi = xs.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    ok = ok and bool(torch.equal(lo, hi))
    flag = torch.tensor([1 if ok else 0], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ctx.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_MAPPING_OK" if int(flag[0]) == 1 else "MULTI_GPU_MAPPING_FAILED", flush=True)
    sys.exit(0 if int(flag[0]) == 1 else 1)


if __name__ == "__main__":
    main()
