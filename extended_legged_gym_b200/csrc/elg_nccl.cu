// elg_nccl.cu -- the two genuine reductions of the path on the COMPUTE STREAM, graph-capturable, no host sync
// (the reference env is single-GPU and has no collective of its own):
//
//   elg_episode_stats_allreduce   sum over ranks of the (per-term sums, count[, extra words]) accumulated by elg_reset_envs:
//                                 extras["episode"] means identical to one GPU owning every env (legged_robot.py:200-206)
//   elg_mppi_update               the whole cost-weighted update with the rollout dimension sharded over ranks: local costs ->
//                                 ncclAllGather (4 B per sample) -> local weights and partial sums -> ncclAllReduce of
//                                 [sum_e, sum_e * sample] per main env -> mean trajectories (cmp_mppi_wbfo.py:216-233)
//
// NCCL is resolved at run time with dlopen: a process that has imported torch already has the torch-bundled libnccl.so.2
// mapped, and the same copy is picked up (RTLD_NOLOAD first); the library itself has no link-time NCCL dependency, so the
// single-GPU path and the CPU-side ABI tests load it on a box without NCCL.  One communicator per rank, created from a
// ncclUniqueId that the host side distributes (utils/distributed.py broadcasts it over torch.distributed).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "elg_common.cuh"

// the handful of NCCL declarations used here (ABI-stable across NCCL 2.x; values as in nccl.h)
extern "C" {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;     // ncclSuccess == 0
typedef int ncclDataType_t;   // ncclFloat32 == 7, ncclFloat64 == 8
typedef int ncclRedOp_t;      // ncclSum == 0
}

namespace {

struct Nccl {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};
Nccl g_nccl;

int nccl_load() {
  if (g_nccl.handle) return ELG_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // the copy torch has mapped, if any
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return elg::set_error(ELG_ERR_UNSUPPORTED, "libnccl.so.2 not found (multi-GPU reductions need NCCL)");
#define ELG_SYM(field, name)                                                         \
  *(void**)(&g_nccl.field) = dlsym(h, name);                                         \
  if (!g_nccl.field) return elg::set_error(ELG_ERR_UNSUPPORTED, "libnccl: missing symbol " name);
  ELG_SYM(GetUniqueId, "ncclGetUniqueId")
  ELG_SYM(CommInitRank, "ncclCommInitRank")
  ELG_SYM(CommDestroy, "ncclCommDestroy")
  ELG_SYM(AllReduce, "ncclAllReduce")
  ELG_SYM(AllGather, "ncclAllGather")
  ELG_SYM(GetErrorString, "ncclGetErrorString")
  ELG_SYM(GetVersion, "ncclGetVersion")
#undef ELG_SYM
  g_nccl.handle = h;
  return ELG_OK;
}

int nccl_check(ncclResult_t r, const char* what) {
  if (r == 0) return ELG_OK;
  static thread_local char msg[256];
  snprintf(msg, sizeof(msg), "%s: NCCL error %d (%s)", what, r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  return elg::set_error(ELG_ERR_CUDA, msg);
}

}  // namespace

struct ElgComm {
  ncclComm_t comm;
  int rank, world;
};

extern "C" {

int elg_comm_unique_id(void* out128) {
  if (!out128) return elg::set_error(ELG_ERR_NULL_POINTER, "elg_comm_unique_id: out is NULL");
  if (int rc = nccl_load()) return rc;
  ncclUniqueId id;
  if (int rc = nccl_check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId")) return rc;
  memcpy(out128, &id, sizeof(id));
  return ELG_OK;
}

int elg_comm_init(const void* unique_id128, int rank, int world, ElgComm** out) {
  if (!unique_id128 || !out) return elg::set_error(ELG_ERR_NULL_POINTER, "elg_comm_init: id/out is NULL");
  if (world < 1 || rank < 0 || rank >= world) return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "elg_comm_init: bad rank / world size");
  if (int rc = nccl_load()) return rc;
  ncclUniqueId id;
  memcpy(&id, unique_id128, sizeof(id));
  ncclComm_t c = nullptr;
  if (int rc = nccl_check(g_nccl.CommInitRank(&c, world, id, rank), "ncclCommInitRank")) return rc;
  *out = new ElgComm{c, rank, world};
  return ELG_OK;
}

int elg_comm_destroy(ElgComm* comm) {
  if (!comm) return ELG_OK;
  int rc = ELG_OK;
  if (g_nccl.CommDestroy) rc = nccl_check(g_nccl.CommDestroy(comm->comm), "ncclCommDestroy");
  delete comm;
  return rc;
}

int elg_comm_info(const ElgComm* comm, int* rank, int* world, int* nccl_version) {
  if (!comm) return elg::set_error(ELG_ERR_NULL_POINTER, "elg_comm_info: comm is NULL");
  if (rank) *rank = comm->rank;
  if (world) *world = comm->world;
  if (nccl_version && g_nccl.GetVersion) g_nccl.GetVersion(nccl_version);
  return ELG_OK;
}

// one all-gather of `n` floats per rank (+ one float all-reduce) so that NCCL's lazy peer set-up for the collectives of
// elg_mppi_update happens here, outside any CUDA-graph capture; scratch holds >= 8 * world bytes * n
int elg_comm_warmup(ElgComm* comm, void* scratch, int32_t n, void* stream) {
  if (!comm || comm->world == 1) return ELG_OK;
  if (!scratch || n < 1) return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "elg_comm_warmup: scratch is NULL or n < 1");
  float* f = static_cast<float*>(scratch);
  if (int rc = nccl_check(g_nccl.AllGather(f + (size_t)comm->rank * n, f, (size_t)n, /*ncclFloat32*/ 7, comm->comm, (cudaStream_t)stream), "ncclAllGather(warm-up)"))
    return rc;
  return nccl_check(g_nccl.AllReduce(f, f, (size_t)n, 7, 0, comm->comm, (cudaStream_t)stream), "ncclAllReduce(warm-up)");
}

int elg_episode_stats_allreduce(double* stats, int32_t n, ElgComm* comm, void* stream) {
  if (n < 0) return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "elg_episode_stats_allreduce: negative length");
  if (n == 0 || !comm || comm->world == 1) return ELG_OK;   // one rank: the local sums are the global ones
  if (!stats) return elg::set_error(ELG_ERR_NULL_POINTER, "elg_episode_stats_allreduce: stats is NULL");
  return nccl_check(g_nccl.AllReduce(stats, stats, (size_t)n, /*ncclFloat64*/ 8, /*ncclSum*/ 0, comm->comm, (cudaStream_t)stream),
                    "ncclAllReduce(episode stats)");
}

// the local stages (elg_mppi.cu)
int elg_mppi_costs(const float* rewards, int64_t num_main, int64_t num_samples, int32_t horizon, float* costs, void* stream);
int elg_mppi_finish(const float* partial, int64_t num_main, int32_t traj_size, float* mean_traj, void* stream);
int elg_mppi_partials_ranked(const float* costs_ranked, int64_t num_main, int32_t num_ranks, int32_t rank, int32_t samples_local,
                             const float* samples, int32_t traj_size, float temperature, float* partial, void* stream);

int elg_mppi_update(const float* rewards, const float* samples, int64_t num_main, int32_t samples_local, int32_t horizon, int32_t traj_size,
                    float temperature, float* costs_ranked, float* partial, float* mean_traj, ElgComm* comm, void* stream) {
  const int world = comm ? comm->world : 1, rank = comm ? comm->rank : 0;
  if (num_main < 0 || samples_local < 1 || horizon < 0 || traj_size < 1) return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "elg_mppi_update: bad sizes");
  if (num_main == 0) return ELG_OK;
  if (!rewards || !samples || !costs_ranked || !partial || !mean_traj) return elg::set_error(ELG_ERR_NULL_POINTER, "elg_mppi_update: a buffer is NULL");
  const size_t block = (size_t)num_main * samples_local;   // costs_ranked is [world][num_main][samples_local]
  if (int rc = elg_mppi_costs(rewards, num_main, samples_local, horizon, costs_ranked + (size_t)rank * block, stream)) return rc;
  if (world > 1)
    if (int rc = nccl_check(g_nccl.AllGather(costs_ranked + (size_t)rank * block, costs_ranked, block, /*ncclFloat32*/ 7, comm->comm, (cudaStream_t)stream),
                            "ncclAllGather(costs)"))
      return rc;
  if (int rc = elg_mppi_partials_ranked(costs_ranked, num_main, world, rank, samples_local, samples, traj_size, temperature, partial, stream)) return rc;
  if (world > 1)
    if (int rc = nccl_check(g_nccl.AllReduce(partial, partial, (size_t)num_main * (1 + traj_size), 7, 0, comm->comm, (cudaStream_t)stream),
                            "ncclAllReduce(partials)"))
      return rc;
  return elg_mppi_finish(partial, num_main, traj_size, mean_traj, stream);
}

}  // extern "C"
