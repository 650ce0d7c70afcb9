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
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) return false;
  g_nccl.handle = h;
  return true;
}
}  // namespace

void launch_lm_sharded(aloam_ctx* c, const BlockRec* blocks, const int* d_n, double* pose, const LmParams& lp, LmSummary* summary) {
  const int evals = 1 + lp.max_iters;
  double* local = c->d_lm_tot + 32;
  for (int e = 0; e < evals; ++e) {
    const int first = e == 0, last = e == evals - 1;
    prof_begin(c, KID_LM_SOLVE);
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
    LAUNCH(c, KID_LM_SOLVE, k_lm_tr_shard, 1, 32, 0, c->d_lm_state, (const double*)c->d_lm_tot, pose, first, last, lp, summary);
  }
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
  return ALOAM_OK;
}

int aloam_shard_slab_cells(void) { return 8; }   // ownership: slab = floor((cell_x + 2^20) / 8), owner = slab mod world

}  // extern "C"
