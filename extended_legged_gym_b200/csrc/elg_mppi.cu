// elg_mppi.cu -- cost-weighted control update of sampling-based MPC (MPPI) for sm_100a.
//
// The production optimiser of the reference lives in the external `traj_sampling` package (PegasusFlow, absent from
// the tree); the only in-tree statement of the update is legged_gym/tests/score_sampling/cmp_mppi_wbfo.py:216-233:
//     costs = sum_t step_rewards;  n = (costs - mean) / (std + 1e-6);  w = softmax(n / temp);  traj = sum_s w_s sample_s
// Here it is batched over the main envs and split so that the rollout dimension can be sharded across GPUs
// (SURVEY section 8e): every rank reduces its own samples, the collectives in between carry a few KB.
//   elg_mppi_costs        rewards [M, S, T] -> costs [M, S]                                     (local)
//   elg_mppi_partials     all costs [M, S_total] + local samples [M, S_local, K*D] -> per main
//                         [sum_e, sum_e * sample (K*D)] with e = exp((n_s - n_max) / temp)      (local; all-reduce next)
//   elg_mppi_finish       partial sums -> mean trajectory [M, K*D]
#include <cuda_runtime.h>
#include <stdint.h>

#include "elg_common.cuh"

namespace elg {

constexpr int kMppiThreads = 256;

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0f;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (l == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}
__device__ __forceinline__ float block_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : -3.402823466e38f;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
    if (l == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

__global__ void __launch_bounds__(kMppiThreads)
elg_mppi_costs_kernel(const float* __restrict__ rewards, const long long rows, const int T, float* __restrict__ costs) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  const float* r = rewards + i * T;
  float s = 0.0f;
  for (int t = 0; t < T; ++t) s += r[t];
  costs[i] = s;
}

// one CTA per main env.  costs_all is [num_ranks][M][S_rank] (what ncclAllGather leaves: rank-major blocks); the single-tensor form
// [M, S_total] is num_ranks == 1.  The local samples are columns [s_first, s_first + S_local) of rank block `rank`.
__global__ void __launch_bounds__(kMppiThreads)
elg_mppi_partials_kernel(const float* __restrict__ costs_all, const int M, const int num_ranks, const int S_rank, const int rank,
                         const int s_first, const int S_local, const float* __restrict__ samples, const int KD, const float temp,
                         float* __restrict__ partial) {
  __shared__ float red[32];
  extern __shared__ float s_e[];   // [S_total] this main env's costs, then the weights of the local samples in place
  const int m = blockIdx.x;
  const int S_total = num_ranks * S_rank;
  float* const c = s_e;
  for (int s = threadIdx.x; s < S_total; s += blockDim.x) {
    const int r = s / S_rank;
    c[s] = costs_all[((size_t)r * M + m) * S_rank + (s - r * S_rank)];
  }
  __syncthreads();
  const int first = rank * S_rank + s_first;
  float acc = 0.0f;
  for (int s = threadIdx.x; s < S_total; s += blockDim.x) acc += c[s];
  const float mean = block_sum(acc, red) / (float)S_total;
  acc = 0.0f;
  for (int s = threadIdx.x; s < S_total; s += blockDim.x) {
    const float d = c[s] - mean;
    acc += d * d;
  }
  const float var = block_sum(acc, red) / (float)(S_total > 1 ? S_total - 1 : 1);   // torch.std: unbiased
  const float denom = sqrtf(var) + 1e-6f;
  float mx = -3.402823466e38f;
  for (int s = threadIdx.x; s < S_total; s += blockDim.x) mx = fmaxf(mx, (c[s] - mean) / denom);
  mx = block_max(mx, red);
  acc = 0.0f;
  float* const w_e = s_e + S_total;   // [S_local]
  for (int s = threadIdx.x; s < S_local; s += blockDim.x) {
    const float n = (c[first + s] - mean) / denom;
    const float e = expf((n - mx) / temp);
    w_e[s] = e;
    acc += e;
  }
  const float sum_e = block_sum(acc, red);
  float* out = partial + (size_t)m * (1 + KD);
  if (threadIdx.x == 0) out[0] = sum_e;
  // sum_s w_s * sample[s][j]: the block is 4 sample groups x 64 columns -- consecutive lanes read consecutive floats of one sample
  // row, every thread runs a quarter of the samples, the four partial sums meet in shared memory (fixed order: reproducible)
  __shared__ float s_acc[4][64];
  const float* smp = samples + (size_t)m * S_local * KD;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  for (int j0 = 0; j0 < KD; j0 += 64) {
    const int j = j0 + tx;
    float a = 0.0f;
    if (j < KD)
      for (int s = ty; s < S_local; s += 4) a = fmaf(w_e[s], smp[(size_t)s * KD + j], a);
    __syncthreads();
    s_acc[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && j < KD) out[1 + j] = (s_acc[0][tx] + s_acc[1][tx]) + (s_acc[2][tx] + s_acc[3][tx]);
  }
}

__global__ void __launch_bounds__(kMppiThreads)
elg_mppi_finish_kernel(const float* __restrict__ partial, const long long M, const int KD, float* __restrict__ mean_traj) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * KD) return;
  const long long m = i / KD;
  const int j = (int)(i - m * KD);
  mean_traj[i] = partial[m * (1 + KD) + 1 + j] / partial[m * (1 + KD)];
}

}  // namespace elg

namespace {
int pfail(int code, const char* msg) { return elg::set_error(code, msg); }
}  // namespace

extern "C" {

int elg_mppi_costs(const float* rewards, int64_t num_main, int64_t num_samples, int32_t horizon, float* costs, void* stream) {
  if (num_main < 0 || num_samples < 0 || horizon < 0) return pfail(ELG_ERR_INVALID_ARGUMENT, "negative size");
  const long long rows = (long long)num_main * num_samples;
  if (rows == 0) return ELG_OK;
  if (!rewards || !costs) return pfail(ELG_ERR_NULL_POINTER, "rewards/costs is NULL");
  elg::elg_mppi_costs_kernel<<<(unsigned)((rows + elg::kMppiThreads - 1) / elg::kMppiThreads), elg::kMppiThreads, 0, (cudaStream_t)stream>>>(
      rewards, rows, horizon, costs);
  return elg::check_launch("elg_mppi_costs");
}

static int launch_partials(const float* costs_all, int64_t num_main, int32_t num_ranks, int32_t samples_rank, int32_t rank, int32_t first_local_sample,
                           int32_t samples_local, const float* samples, int32_t traj_size, float temperature, float* partial, void* stream) {
  const long long samples_total = (long long)num_ranks * samples_rank;
  if (num_main < 0 || samples_total < 1 || samples_local < 0 || traj_size < 1) return pfail(ELG_ERR_INVALID_ARGUMENT, "bad MPPI sizes");
  if (first_local_sample < 0 || first_local_sample + samples_local > samples_rank || rank < 0 || rank >= num_ranks)
    return pfail(ELG_ERR_INVALID_ARGUMENT, "local sample range outside [0, samples_total)");
  if (!(temperature > 0.0f)) return pfail(ELG_ERR_INVALID_ARGUMENT, "temperature must be > 0");
  if (num_main == 0) return ELG_OK;
  if (!costs_all || !partial || (samples_local > 0 && !samples)) return pfail(ELG_ERR_NULL_POINTER, "an MPPI buffer is NULL");
  const size_t smem = 4 * (size_t)(samples_total + (samples_local > 0 ? samples_local : 1));
  if (smem > 200 * 1024) return pfail(ELG_ERR_UNSUPPORTED, "more than 51200 samples (all ranks + local) per main env");
  static elg::SmemCache smem_cache = {};
  size_t& smem_set = elg::smem_slot(smem_cache);
  if (smem > 48 * 1024 && smem > smem_set) {
    if (cudaFuncSetAttribute(elg::elg_mppi_partials_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return pfail(ELG_ERR_CUDA, "cannot reserve shared memory for elg_mppi_partials_kernel");
    smem_set = smem;
  }
  elg::elg_mppi_partials_kernel<<<(unsigned)num_main, elg::kMppiThreads, smem, (cudaStream_t)stream>>>(
      costs_all, (int)num_main, num_ranks, samples_rank, rank, first_local_sample, samples_local, samples, traj_size, temperature, partial);
  return elg::check_launch("elg_mppi_partials");
}

int elg_mppi_partials(const float* costs_all, int64_t num_main, int32_t samples_total, int32_t first_local_sample, int32_t samples_local,
                      const float* samples, int32_t traj_size, float temperature, float* partial, void* stream) {
  return launch_partials(costs_all, num_main, 1, samples_total, 0, first_local_sample, samples_local, samples, traj_size, temperature, partial, stream);
}

int elg_mppi_partials_ranked(const float* costs_ranked, int64_t num_main, int32_t num_ranks, int32_t rank, int32_t samples_local,
                             const float* samples, int32_t traj_size, float temperature, float* partial, void* stream) {
  return launch_partials(costs_ranked, num_main, num_ranks, samples_local, rank, 0, samples_local, samples, traj_size, temperature, partial, stream);
}

int elg_mppi_finish(const float* partial, int64_t num_main, int32_t traj_size, float* mean_traj, void* stream) {
  if (num_main < 0 || traj_size < 1) return pfail(ELG_ERR_INVALID_ARGUMENT, "bad MPPI sizes");
  if (num_main == 0) return ELG_OK;
  if (!partial || !mean_traj) return pfail(ELG_ERR_NULL_POINTER, "partial/mean_traj is NULL");
  const long long n = (long long)num_main * traj_size;
  elg::elg_mppi_finish_kernel<<<(unsigned)((n + elg::kMppiThreads - 1) / elg::kMppiThreads), elg::kMppiThreads, 0, (cudaStream_t)stream>>>(
      partial, num_main, traj_size, mean_traj);
  return elg::check_launch("elg_mppi_finish");
}

}  // extern "C"
