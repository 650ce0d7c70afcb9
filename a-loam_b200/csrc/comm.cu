// Multi-GPU scan-to-map: one process per GPU, the submap is split into spatial slabs (+ one-cell halo),
// every rank fits only the stack points whose cell it owns and the ranks meet in ONE collective per evaluation:
// ncclAllReduce(sum) of 32 doubles = [J^T J upper 21 | J^T r 6 | cost | #edge | #plane | pad] over NVLink / NVSwitch.
// Every rank then takes the identical trust-region step (k_lm_tr_shard) -- no broadcast, no host round trip.
//
// NCCL is loaded with dlopen at communicator creation, so libaloam_b200.so itself has no NCCL dependency and loads on a
// CPU-only box; inside a torch process the already-loaded torch-bundled libnccl.so.2 is picked up by soname.
#include <dlfcn.h>
#include <nccl.h>
#include "ctx.h"

using namespace aloam;

namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

bool load_nccl() {
  if (g_nccl.handle) return true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { fprintf(stderr, "[aloam_b200] cannot load libnccl: %s\n", dlerror()); return false; }
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
  g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(h, "ncclAllGather");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) return false;
  g_nccl.handle = h;
  return true;
}
}  // namespace

// peer-memory exchange state of a context (see PeerX, kernels.h)
struct PeerState {
  PeerX px = {};
  void* own = nullptr;                 // this rank's mailbox allocation (boxes, flags, counter, totals, error word)
  void* opened[ALOAM_MAX_RANKS] = {};  // peers' mailboxes mapped into this process (cudaIpcOpenMemHandle)
};

void launch_lm_sharded(aloam_ctx* c, const BlockRec* blocks, const int* d_n, double* pose, const LmParams& lp, LmSummary* summary) {
  if (c->peer) {
    // ONE launch per solve: the cluster kernel of the single-GPU path with the all-reduce of every evaluation done inside it
    // over NVLink peer memory; early termination ends the solve on every rank at the same evaluation
    PeerState* ps = static_cast<PeerState*>(c->peer);
    Batch<LmArgs> b = {};
    b.a[0] = LmArgs{blocks, d_n, 0, pose, summary, nullptr, nullptr};
    launch_lm_batch(c, true, b, 1, lp, 0, 0, true, &ps->px);
    return;
  }
  const int evals = 1 + lp.max_iters;
  double* local = c->d_lm_tot + 32;
  for (int e = 0; e < evals; ++e) {
    const int first = e == 0, last = e == evals - 1;
    prof_begin(c, KID_LM_SHARD);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kLmCluster); cfg.blockDim = dim3(ALOAM_LM_THREADS); cfg.dynamicSmemBytes = lm_dynamic_smem_bytes(); cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kLmCluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k_lm_eval_shard, blocks, d_n, (const double*)pose, c->d_lm_state, first, lp.huber_a, local);
    prof_end(c);
    ncclResult_t r = g_nccl.AllReduce(local, c->d_lm_tot, 32, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream);
    if (r != ncclSuccess) fprintf(stderr, "[aloam_b200] ncclAllReduce failed: %s\n", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    LAUNCH(c, KID_LM_SHARD, k_lm_tr_shard, 1, 32, 0, c->d_lm_state, (const double*)c->d_lm_tot, pose, first, last, lp, summary);
  }
}

// Maps every rank's mailbox into every other rank (CUDA IPC handles exchanged with one ncclAllGather).  Returns false (and leaves
// the NCCL path in place) when peer access is not available or ALOAM_NO_PEER is set.
static bool peer_setup(aloam_ctx* c, int rank, int world) {
  if (getenv("ALOAM_NO_PEER") || world > ALOAM_MAX_RANKS || !g_nccl.AllGather) return false;
  PeerState* ps = new (std::nothrow) PeerState();
  if (!ps) return false;
  const size_t box_bytes = (size_t)2 * world * 32 * 2 * sizeof(unsigned long long), flag_bytes = (size_t)2 * world * sizeof(unsigned);
  const size_t total = box_bytes + flag_bytes + 64 + 2 * 32 * sizeof(double) + 64;
  bool ok = cudaMalloc(&ps->own, total) == cudaSuccess && cudaMemset(ps->own, 0, total) == cudaSuccess;
  cudaIpcMemHandle_t mine;
  ok = ok && cudaIpcGetMemHandle(&mine, ps->own) == cudaSuccess;
  cudaIpcMemHandle_t* d_all = nullptr;
  std::vector<cudaIpcMemHandle_t> all(world);
  ok = ok && cudaMalloc((void**)&d_all, sizeof(mine) * (world + 1)) == cudaSuccess;
  // every rank must take part in the gather even if its own setup failed: a zeroed handle marks the failure
  if (!ok) std::memset(&mine, 0, sizeof(mine));
  bool gathered = false;
  if (d_all) {
    cudaMemcpyAsync(d_all + world, &mine, sizeof(mine), cudaMemcpyHostToDevice, c->stream);
    if (g_nccl.AllGather(d_all + world, d_all, sizeof(mine), ncclChar, (ncclComm_t)c->comm, c->stream) == ncclSuccess &&
        cudaMemcpyAsync(all.data(), d_all, sizeof(mine) * world, cudaMemcpyDeviceToHost, c->stream) == cudaSuccess &&
        cudaStreamSynchronize(c->stream) == cudaSuccess)
      gathered = true;
    cudaFree(d_all);
  }
  ok = ok && gathered;
  const cudaIpcMemHandle_t zero = {};
  for (int r = 0; r < world && ok; ++r) {
    void* base = ps->own;
    if (r != rank) {
      if (!std::memcmp(&all[r], &zero, sizeof(zero)) || cudaIpcOpenMemHandle(&ps->opened[r], all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; break; }
      base = ps->opened[r];
    }
    ps->px.box[r] = reinterpret_cast<double*>(base);
    ps->px.flag[r] = reinterpret_cast<unsigned*>(static_cast<char*>(base) + box_bytes);
  }
  if (!ok) {
    cudaGetLastError();
    for (void* p : ps->opened) if (p) cudaIpcCloseMemHandle(p);
    if (ps->own) cudaFree(ps->own);
    delete ps;
    fprintf(stderr, "[aloam_b200] peer-memory exchange unavailable on rank %d: the sharded LM uses ncclAllReduce\n", rank);
    return false;
  }
  char* own = static_cast<char*>(ps->own);
  ps->px.seq = reinterpret_cast<unsigned long long*>(own + box_bytes + flag_bytes);
  ps->px.gtot = reinterpret_cast<double*>(own + box_bytes + flag_bytes + 64);
  ps->px.err = reinterpret_cast<int*>(own + box_bytes + flag_bytes + 64 + 2 * 32 * sizeof(double));
  ps->px.rank = rank; ps->px.world = world;
  c->peer = ps;
  return true;
}

// sum of two ints over the ranks, in place (global submap sizes for the thin-map test, mapping.cu)
int comm_allreduce_int2(aloam_ctx* c, int* d_two) {
  if (c->shard_count <= 1) return ALOAM_OK;
  ncclResult_t r = g_nccl.AllReduce(d_two, d_two, 2, ncclInt32, ncclSum, (ncclComm_t)c->comm, c->stream);
  if (r != ncclSuccess) { fprintf(stderr, "[aloam_b200] ncclAllReduce failed: %s\n", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); return ALOAM_ERR_COMM; }
  return ALOAM_OK;
}

extern "C" {

void aloam_comm_free_impl(aloam_ctx* c) {
  if (c->comm && g_nccl.CommDestroy) { g_nccl.CommDestroy((ncclComm_t)c->comm); c->comm = nullptr; }
  if (c->peer) {
    PeerState* ps = static_cast<PeerState*>(c->peer);
    for (void* p : ps->opened) if (p) cudaIpcCloseMemHandle(p);
    if (ps->own) cudaFree(ps->own);
    delete ps;
    c->peer = nullptr;
  }
  if (c->d_lm_tot) cudaFree(c->d_lm_tot);
  if (c->d_lm_state) cudaFree(c->d_lm_state);
  c->d_lm_tot = nullptr; c->d_lm_state = nullptr;
}

// rank 0 creates the 128-byte NCCL id; the caller ships it to the other ranks (e.g. torch.distributed.broadcast)
int aloam_comm_unique_id(char out128[128]) {
  if (!out128) return ALOAM_ERR_INVALID_ARG;
  if (!load_nccl()) return ALOAM_ERR_COMM;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return ALOAM_ERR_COMM;
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  std::memcpy(out128, &id, 128);
  return ALOAM_OK;
}

// joins this context to a world of `world` ranks: from now on aloam_mapping_register fits only the stack points whose
// map cell this rank owns and all-reduces the normal equations.  The uploaded map must be this rank's shard
// (owned slabs + one-cell halo; see a-loam_b200/shard.py for the host-side split).
int aloam_comm_init(aloam_ctx* c, int rank, int world, const char id128[128]) {
  if (!c || !id128 || world < 1 || rank < 0 || rank >= world) return ALOAM_ERR_INVALID_ARG;
  CUDA_CHECK_RET(cudaSetDevice(c->cfg.device));
  if (world == 1) { c->shard_rank = 0; c->shard_count = 1; return ALOAM_OK; }
  if (!load_nccl()) return ALOAM_ERR_COMM;
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  ncclComm_t comm = nullptr;
  ncclResult_t r = g_nccl.CommInitRank(&comm, world, id, rank);
  if (r != ncclSuccess) { fprintf(stderr, "[aloam_b200] ncclCommInitRank failed: %s\n", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); return ALOAM_ERR_COMM; }
  c->comm = comm;
  if (!c->d_lm_tot) CUDA_CHECK_RET(cudaMalloc((void**)&c->d_lm_tot, 64 * sizeof(double)));
  if (!c->d_lm_state) CUDA_CHECK_RET(cudaMalloc(&c->d_lm_state, lm_state_bytes()));
  c->shard_rank = rank; c->shard_count = world;
  // the 256-byte exchange of every LM evaluation goes through NVLink peer memory inside the solve kernel when the ranks can map
  // each other's memory (all ranks decide alike: the outcome of the handle exchange is the same everywhere ... see below)
  const bool mine_ok = peer_setup(c, rank, world);
  // agree on the path: one rank without peer access forces everybody onto NCCL
  int* d_flag = nullptr;
  CUDA_CHECK_RET(cudaMalloc((void**)&d_flag, 2 * sizeof(int)));
  int h[2] = {mine_ok ? 0 : 1, 0};
  CUDA_CHECK_RET(cudaMemcpyAsync(d_flag, h, 8, cudaMemcpyHostToDevice, c->stream));
  if (g_nccl.AllReduce(d_flag, d_flag, 2, ncclInt32, ncclSum, (ncclComm_t)c->comm, c->stream) != ncclSuccess) { cudaFree(d_flag); return ALOAM_ERR_COMM; }
  CUDA_CHECK_RET(cudaMemcpyAsync(h, d_flag, 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK_RET(cudaStreamSynchronize(c->stream));
  cudaFree(d_flag);
  if (h[0] != 0 && c->peer) {   // somebody failed: drop the peer path here too
    PeerState* ps = static_cast<PeerState*>(c->peer);
    for (void* p : ps->opened) if (p) cudaIpcCloseMemHandle(p);
    if (ps->own) cudaFree(ps->own);
    delete ps;
    c->peer = nullptr;
  }
  return ALOAM_OK;
}

int aloam_comm_uses_peer_memory(aloam_ctx* c) { return c && c->peer ? 1 : 0; }

int aloam_shard_slab_cells(void) { return 8; }   // ownership: slab = floor((cell_x + 2^20) / 8), owner = slab mod world

}  // extern "C"
