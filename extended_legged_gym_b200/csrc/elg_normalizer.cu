// elg_normalizer.cu -- caller-side fusion of the step outputs for sm_100a:
// rsl_rl EmpiricalNormalization.forward in training mode (rsl_rl/modules/normalizer.py:43-75: running mean / variance update
// over the batch of envs, then (x - mean) / (std + eps)) with the normalised rows written straight into their destination
// (the rollout-storage slot, rsl_rl/storage/rollout_storage.py:95-100) together with the reward / done columns, so the
// observations go from the step kernel's output to the policy's input buffer in one pass.
//
// Two launches, chained by programmatic dependent launch, no atomics, bit-reproducible:
//   elg_norm_stats_kernel  per row block and column the (count, mean, M2) triple -> scratch; CTA 0 snapshots the old state
//   elg_norm_apply_kernel  every CTA merges the (<= 128) triples in block order (Chan et al.) -- the same arithmetic in every CTA,
//                          so all of them hold the identical new mean / std -- and normalises its rows; CTA 0 stores the state.
// [N, O] fp32 is read twice (second time from L2) and written once: 8 N O bytes of HBM traffic for the pair.
#include <cuda_runtime.h>
#include <stdint.h>

#include "elg_common.cuh"
#include "elg_async.cuh"

namespace elg {

constexpr int kNormParts = 128;      // row blocks of the statistics pass (upper bound)
constexpr int kStatsThreads = 256;
constexpr int kApplyThreads = 1024;
constexpr int kChunk = 32;           // rows held in registers at a time

struct NormGeom {
  int64_t rows;
  int cols;
  int cpt;              // columns per tile (power of two, 32..256)
  int parts;            // row blocks actually used
  int64_t rows_per_part;
};

__host__ __device__ inline int norm_cpt(int cols) { return cols > 128 ? 256 : cols > 64 ? 128 : cols > 32 ? 64 : 32; }

// scratch layout (floats after a 16-byte header holding the old count): old_mean[O], old_var[O], then [parts][3][O]
__device__ __forceinline__ float* scr_old_mean(void* s) { return reinterpret_cast<float*>(reinterpret_cast<char*>(s) + 16); }
__device__ __forceinline__ float* scr_part(void* s, int cols, int part, int which) {
  return scr_old_mean(s) + (size_t)cols * (2 + 3 * (size_t)part + which);
}

struct Moments { float n, mean, m2; };
__device__ __forceinline__ Moments merge(const Moments a, const Moments b) {     // Chan, Golub, LeVeque pairwise update
  if (b.n == 0.0f) return a;
  if (a.n == 0.0f) return b;
  const float n = a.n + b.n, d = b.mean - a.mean;
  return Moments{n, a.mean + d * (b.n / n), a.m2 + b.m2 + d * d * (a.n * (b.n / n))};
}

__global__ void __launch_bounds__(kStatsThreads)
elg_norm_stats_kernel(const __grid_constant__ NormGeom g, const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ var,
                      const int64_t* __restrict__ count, const int64_t until, void* __restrict__ scratch) {
  __shared__ float sm[3 * kStatsThreads];
  const int c = threadIdx.x % g.cpt, rs = threadIdx.x / g.cpt, rsub = kStatsThreads / g.cpt;
  const int col = blockIdx.y * g.cpt + c;
  const bool live = col < g.cols;
  pdl_launch_dependents();
  pdl_wait();
  const int64_t old_count = *count;
  if (blockIdx.x == 0) {      // snapshot of the state the apply pass updates from (its CTA 0 overwrites the live tensors)
    if (threadIdx.x == 0 && blockIdx.y == 0) *reinterpret_cast<int64_t*>(scratch) = old_count;
    if (rs == 0 && live) { scr_old_mean(scratch)[col] = mean[col]; scr_old_mean(scratch)[g.cols + col] = var[col]; }
  }
  if (until >= 0 && old_count >= until) return;          // learning has stopped (normalizer.py:62-63)
  const int64_t r0 = (int64_t)blockIdx.x * g.rows_per_part;
  const int64_t r1 = min(r0 + g.rows_per_part, g.rows);
  Moments acc{0.0f, 0.0f, 0.0f};
  for (int64_t base = r0 + rs; base < r1; base += (int64_t)kChunk * rsub) {
    float v[kChunk];
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < kChunk; ++k) {
      const int64_t r = base + (int64_t)k * rsub;
      const bool ok = live && r < r1;
      v[k] = ok ? x[r * g.cols + col] : 0.0f;
      cnt += ok;
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < kChunk; ++k) s += v[k];
    const float m = cnt ? s / (float)cnt : 0.0f;
    float m2 = 0.0f;
#pragma unroll
    for (int k = 0; k < kChunk; ++k) { const float d = v[k] - m; m2 += (k < cnt) ? d * d : 0.0f; }
    acc = merge(acc, Moments{(float)cnt, m, m2});
  }
  sm[threadIdx.x] = acc.n; sm[kStatsThreads + threadIdx.x] = acc.mean; sm[2 * kStatsThreads + threadIdx.x] = acc.m2;
  __syncthreads();
  if (rs == 0 && live) {
    for (int j = 1; j < rsub; ++j) {
      const int t = j * g.cpt + c;
      acc = merge(acc, Moments{sm[t], sm[kStatsThreads + t], sm[2 * kStatsThreads + t]});
    }
    scr_part(scratch, g.cols, blockIdx.x, 0)[col] = acc.n;
    scr_part(scratch, g.cols, blockIdx.x, 1)[col] = acc.mean;
    scr_part(scratch, g.cols, blockIdx.x, 2)[col] = acc.m2;
  }
}

__global__ void __launch_bounds__(kApplyThreads)
elg_norm_apply_kernel(const __grid_constant__ NormGeom g, const float* __restrict__ x, float* __restrict__ mean, float* __restrict__ var,
                      float* __restrict__ stdv, int64_t* __restrict__ count, const float eps, const int64_t until, const int training,
                      float* __restrict__ out, void* __restrict__ scratch, const float* __restrict__ rew, float* __restrict__ rew_out,
                      const uint8_t* __restrict__ dones, uint8_t* __restrict__ dones_out) {
  __shared__ float sm[3 * kApplyThreads];
  __shared__ float s_mean[256], s_den[256];
  const int c = threadIdx.x % g.cpt, grp = threadIdx.x / g.cpt, groups = kApplyThreads / g.cpt;
  const int col = blockIdx.y * g.cpt + c;
  const bool live = col < g.cols;
  pdl_launch_dependents();
  pdl_wait();
  const int64_t old_count = training ? *reinterpret_cast<const int64_t*>(scratch) : 0;
  const bool learn = training && !(until >= 0 && old_count >= until);
  if (learn) {
    Moments acc{0.0f, 0.0f, 0.0f};
    if (live)
      for (int p = grp; p < g.parts; p += groups)
        acc = merge(acc, Moments{scr_part(scratch, g.cols, p, 0)[col], scr_part(scratch, g.cols, p, 1)[col], scr_part(scratch, g.cols, p, 2)[col]});
    sm[threadIdx.x] = acc.n; sm[kApplyThreads + threadIdx.x] = acc.mean; sm[2 * kApplyThreads + threadIdx.x] = acc.m2;
    __syncthreads();
    if (grp == 0 && live) {
      for (int j = 1; j < groups; ++j) {
        const int t = j * g.cpt + c;
        acc = merge(acc, Moments{sm[t], sm[kApplyThreads + t], sm[2 * kApplyThreads + t]});
      }
      // normalizer.py:65-75
      const int64_t new_count = old_count + g.rows;
      const float rate = (float)g.rows / (float)new_count;
      const float mean_x = acc.mean, var_x = acc.m2 / (float)g.rows;
      const float m_old = scr_old_mean(scratch)[col], v_old = scr_old_mean(scratch)[g.cols + col];
      const float delta = mean_x - m_old;
      const float m_new = m_old + rate * delta;
      const float v_new = v_old + rate * (var_x - v_old + delta * (mean_x - m_new));
      const float s_new = __fsqrt_rn(v_new);
      s_mean[c] = m_new;
      s_den[c] = s_new + eps;
      if (blockIdx.x == 0) {
        mean[col] = m_new; var[col] = v_new; stdv[col] = s_new;
        if (col == 0) *count = new_count;
      }
    }
  } else if (grp == 0 && live) {
    s_mean[c] = mean[col];
    s_den[c] = stdv[col] + eps;
  }
  __syncthreads();
  const int64_t rpb = (g.rows + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * rpb, r1 = min(r0 + rpb, g.rows);
  if (live && out) {
    const float m = s_mean[c], den = s_den[c];
    for (int64_t r = r0 + grp; r < r1; r += groups) out[r * g.cols + col] = (x[r * g.cols + col] - m) / den;
  }
  if (blockIdx.y == 0)
    for (int64_t r = r0 + threadIdx.x; r < r1; r += kApplyThreads) {
      if (rew_out) rew_out[r] = rew[r];
      if (dones_out) dones_out[r] = dones[r];
    }
}

static NormGeom make_geom(int64_t rows, int cols) {
  NormGeom g{};
  g.rows = rows;
  g.cols = cols;
  g.cpt = norm_cpt(cols);
  g.rows_per_part = (rows + kNormParts - 1) / kNormParts;
  if (g.rows_per_part < kChunk) g.rows_per_part = kChunk;
  g.parts = (int)((rows + g.rows_per_part - 1) / g.rows_per_part);
  return g;
}

template <typename... Args>
static void launch_pdl(void (*k)(Args...), dim3 grid, int threads, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(threads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, k, args...);
}

}  // namespace elg

extern "C" {

int64_t elg_normalizer_scratch_bytes(int64_t num_rows, int32_t num_cols) {
  if (num_rows < 0 || num_cols < 1) return 0;
  const elg::NormGeom g = elg::make_geom(num_rows > 0 ? num_rows : 1, num_cols);
  return 16 + (int64_t)sizeof(float) * num_cols * (2 + 3 * (int64_t)g.parts);
}

int elg_normalize_observations(int64_t num_rows, int32_t num_cols, const float* x, float* mean, float* var, float* std, int64_t* count, float eps,
                               int64_t until, int32_t training, float* out, void* scratch, const float* rew, float* rew_out,
                               const uint8_t* dones, uint8_t* dones_out, void* stream) {
  if (num_rows < 0 || num_cols < 1) return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "normalizer: num_rows < 0 or num_cols < 1");
  if (!x || !mean || !var || !std || !count) return elg::set_error(ELG_ERR_NULL_POINTER, "normalizer: x / mean / var / std / count is NULL");
  if (training && !scratch) return elg::set_error(ELG_ERR_NULL_POINTER, "normalizer: training mode needs the scratch buffer");
  if (scratch && (reinterpret_cast<uintptr_t>(scratch) & 15u) != 0) return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "normalizer: scratch must be 16-byte aligned");
  if ((rew_out && !rew) || (dones_out && !dones)) return elg::set_error(ELG_ERR_NULL_POINTER, "normalizer: reward / done destination without a source");
  if (training && num_rows == 0) return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "normalizer: an empty batch cannot update the statistics");
  if (num_rows == 0) return ELG_OK;
  const elg::NormGeom g = elg::make_geom(num_rows, num_cols);
  const unsigned tiles = (unsigned)((num_cols + g.cpt - 1) / g.cpt);
  cudaStream_t s = (cudaStream_t)stream;
  if (training)
    elg::launch_pdl(elg::elg_norm_stats_kernel, dim3((unsigned)g.parts, tiles), elg::kStatsThreads, s, g, x, (const float*)mean, (const float*)var,
                    (const int64_t*)count, until, scratch);
  const int64_t want = (num_rows + 31) / 32;
  const unsigned blocks = (unsigned)(want < 148 ? want : 148);
  elg::launch_pdl(elg::elg_norm_apply_kernel, dim3(blocks, tiles), elg::kApplyThreads, s, g, x, mean, var, std, count, eps, until, (int)training, out, scratch,
                  rew, rew_out, dones, dones_out);
  return elg::check_launch("elg_normalize_observations");
}

}  // extern "C"
