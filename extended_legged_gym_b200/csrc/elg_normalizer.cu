// elg_normalizer.cu -- caller-side fusion of the step outputs for sm_100a:
// rsl_rl EmpiricalNormalization.forward in training mode (rsl_rl/modules/normalizer.py:43-75: running mean / variance update
// over the batch of envs, then (x - mean) / (std + eps)) with the normalised rows written straight into their destination
// (the rollout-storage slot, rsl_rl/storage/rollout_storage.py:95-100) together with the reward / done columns, so the
// observations go from the step kernel's output to the policy's input buffer in one pass.
//
// Batches of <= 65 536 rows take ONE launch (elg_norm_cols_kernel): a CTA owns four columns, its rows stay in registers from the load
// to the normalised store, and the rows of a column group are split over a thread-block cluster of 1 / 2 / 4 / 8 CTAs whose partial
// sums cross through distributed shared memory.  [N, O] is read ONCE (4 N O bytes in, 4 N O out).  Kept as an A/B form
// (elg_set_normalizer_tuning(2)): elg_norm_fused_kernel, a row-parallel single launch with a grid-wide hand-over through global
// memory -- measured slower than the pair below.  Larger batches (and elg_set_normalizer_tuning(1)) take
// two launches, chained by programmatic dependent launch, bit-reproducible:
//   elg_norm_stats_kernel  (<= 32 row blocks) x (column tiles of 64): a thread owns one column of a few rows, forms their
//                          (count, mean, M2) exactly in registers, the CTA tree-merges its row groups (Chan et al.) and
//                          stores one triple per column; the LAST CTA of a column tile to finish (one atomic ticket -- it
//                          decides who merges, never the order) merges the tile's triples in block order -> batch mean / M2.
//   elg_norm_apply_kernel  applies the reference's update rule (every CTA computes the identical new mean / std from the batch
//                          moments; CTA 0 stores the state) and normalises its rows.
// [N, O] fp32 is read twice (second time from L2) and written once: 8 N O bytes of HBM traffic for the pair.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "elg_common.cuh"
#include "elg_async.cuh"

namespace elg {

constexpr int kNormParts = 32;       // row blocks of the statistics pass (upper bound)
constexpr int kNormThreads = 1024;
constexpr int kNormHeader = 256;     // bytes: old count (int64) at 0, one ticket (uint32) per column tile from byte 16
constexpr int kNormMaxTiles = (kNormHeader - 16) / 4;

struct NormGeom {
  int64_t rows;
  int cols;
  int cpt;              // columns per tile of the apply pass (power of two, 32..256)
  int cpt_s;            // columns per tile of the statistics pass (32 or 64: many narrow CTAs -- the pass is issue / latency bound)
  int parts;            // row blocks actually used
  int rows_per_part;
};

__host__ __device__ inline int norm_cpt(int cols) { return cols > 128 ? 256 : cols > 64 ? 128 : cols > 32 ? 64 : 32; }

// scratch floats after the header: old_mean[O], old_var[O], batch_mean[O], batch_m2[O], then the triples [parts][3][O]
__device__ __forceinline__ float* scr_vec(void* s, int cols, int which) {
  return reinterpret_cast<float*>(reinterpret_cast<char*>(s) + kNormHeader) + (size_t)cols * which;
}
__device__ __forceinline__ float* scr_part(void* s, int cols, int part, int which) { return scr_vec(s, cols, 4 + 3 * part + which); }

struct Moments { float n, mean, m2; };
__device__ __forceinline__ Moments merge(const Moments a, const Moments b) {     // Chan, Golub, LeVeque pairwise update
  if (b.n == 0.0f) return a;
  if (a.n == 0.0f) return b;
  const float n = a.n + b.n, d = b.mean - a.mean, w = b.n / n;
  return Moments{n, a.mean + d * w, a.m2 + b.m2 + d * d * (a.n * w)};
}

// tree merge over the row groups of a CTA (group rs, column slot c): group 0 ends up with groups 0..rsub-1 in a fixed order
__device__ __forceinline__ Moments merge_groups(Moments acc, float* sm, const int c, const int rs, const int rsub, const int cpt) {
  for (int half = rsub >> 1; half >= 1; half >>= 1) {
    __syncthreads();
    if (rs >= half && rs < 2 * half) {
      sm[threadIdx.x] = acc.n; sm[kNormThreads + threadIdx.x] = acc.mean; sm[2 * kNormThreads + threadIdx.x] = acc.m2;
    }
    __syncthreads();
    if (rs < half) {
      const int t = (rs + half) * cpt + c;
      acc = merge(acc, Moments{sm[t], sm[kNormThreads + t], sm[2 * kNormThreads + t]});
    }
  }
  return acc;
}

template <int kC>      // rows a thread holds in registers at a time
__global__ void __launch_bounds__(kNormThreads)
elg_norm_stats_kernel(const __grid_constant__ NormGeom g, const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ var,
                      const int64_t* __restrict__ count, const int64_t until, void* __restrict__ scratch) {
  __shared__ float sm[3 * kNormThreads];
  __shared__ int s_last;
  const int cpt = g.cpt_s, c = threadIdx.x & (cpt - 1), rs = threadIdx.x / cpt, rsub = kNormThreads / cpt;
  const int col = blockIdx.y * cpt + c;
  const bool live = col < g.cols;
  pdl_launch_dependents();
  pdl_wait();
  const int64_t old_count = *count;
  if (blockIdx.x == 0) {      // snapshot of the state the apply pass updates from (its CTA 0 overwrites the live tensors)
    if (threadIdx.x == 0 && blockIdx.y == 0) *reinterpret_cast<int64_t*>(scratch) = old_count;
    if (rs == 0 && live) { scr_vec(scratch, g.cols, 0)[col] = mean[col]; scr_vec(scratch, g.cols, 1)[col] = var[col]; }
  }
  if (until >= 0 && old_count >= until) return;          // learning has stopped (normalizer.py:62-63)
  const int64_t r0 = (int64_t)blockIdx.x * g.rows_per_part;
  const int nrows = (int)min((int64_t)g.rows_per_part, g.rows - r0);      // rows of this block; offsets inside it fit 32 bits
  const float* xb = x + r0 * g.cols + col;
  Moments acc{0.0f, 0.0f, 0.0f};
  for (int base = rs; base < nrows; base += kC * rsub) {
    float v[kC];
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < kC; ++k) {
      const int r = base + k * rsub;
      const bool ok = live && r < nrows;
      v[k] = ok ? xb[(uint32_t)(r * g.cols)] : 0.0f;
      cnt += ok;
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < kC; ++k) s += v[k];
    const float m = cnt ? s / (float)cnt : 0.0f;
    float m2 = 0.0f;
#pragma unroll
    for (int k = 0; k < kC; ++k) { const float d = v[k] - m; m2 += (k < cnt) ? d * d : 0.0f; }
    acc = merge(acc, Moments{(float)cnt, m, m2});
  }
  acc = merge_groups(acc, sm, c, rs, rsub, cpt);
  if (rs == 0 && live) {
    scr_part(scratch, g.cols, blockIdx.x, 0)[col] = acc.n;
    scr_part(scratch, g.cols, blockIdx.x, 1)[col] = acc.mean;
    scr_part(scratch, g.cols, blockIdx.x, 2)[col] = acc.m2;
  }
  // the last CTA of this column tile merges the tile's triples in block order
  unsigned* ticket = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(scratch) + 16) + blockIdx.y;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  Moments tot{0.0f, 0.0f, 0.0f};
  if (live)
    for (int p = rs; p < g.parts; p += rsub)
      tot = merge(tot, Moments{__ldcg(scr_part(scratch, g.cols, p, 0) + col), __ldcg(scr_part(scratch, g.cols, p, 1) + col),
                               __ldcg(scr_part(scratch, g.cols, p, 2) + col)});
  tot = merge_groups(tot, sm, c, rs, rsub, cpt);
  if (rs == 0 && live) { scr_vec(scratch, g.cols, 2)[col] = tot.mean; scr_vec(scratch, g.cols, 3)[col] = tot.m2; }
  if (threadIdx.x == 0) *ticket = 0u;      // ready for the next call
}

__global__ void __launch_bounds__(kNormThreads)
elg_norm_apply_kernel(const __grid_constant__ NormGeom g, const float* __restrict__ x, float* __restrict__ mean, float* __restrict__ var,
                      float* __restrict__ stdv, int64_t* __restrict__ count, const float eps, const int64_t until, const int training,
                      float* __restrict__ out, void* __restrict__ scratch, const float* __restrict__ rew, float* __restrict__ rew_out,
                      const uint8_t* __restrict__ dones, uint8_t* __restrict__ dones_out) {
  __shared__ float s_mean[256], s_den[256];
  const int c = threadIdx.x & (g.cpt - 1), grp = threadIdx.x / g.cpt, groups = kNormThreads / g.cpt;
  const int col = blockIdx.y * g.cpt + c;
  const bool live = col < g.cols;
  pdl_launch_dependents();
  pdl_wait();
  if (grp == 0 && live) {
    const int64_t old_count = training ? *reinterpret_cast<const int64_t*>(scratch) : 0;
    if (training && !(until >= 0 && old_count >= until)) {      // normalizer.py:65-75
      const int64_t new_count = old_count + g.rows;
      const float rate = (float)g.rows / (float)new_count;
      const float mean_x = scr_vec(scratch, g.cols, 2)[col], var_x = scr_vec(scratch, g.cols, 3)[col] / (float)g.rows;
      const float m_old = scr_vec(scratch, g.cols, 0)[col], v_old = scr_vec(scratch, g.cols, 1)[col];
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
    } else {
      s_mean[c] = mean[col];
      s_den[c] = stdv[col] + eps;
    }
  }
  __syncthreads();
  const int64_t rpb = (g.rows + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * rpb;
  const int nrows = (int)max((int64_t)0, min(rpb, g.rows - r0));
  if (live && out) {
    const float m = s_mean[c], den = s_den[c];
    const float* xb = x + r0 * g.cols + col;
    float* ob = out + r0 * g.cols + col;
    for (int base = grp; base < nrows; base += 8 * groups) {      // 8 rows in flight per thread
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int r = base + k * groups;
        v[k] = r < nrows ? xb[(uint32_t)(r * g.cols)] : 0.0f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int r = base + k * groups;
        if (r < nrows) ob[(uint32_t)(r * g.cols)] = (v[k] - m) / den;
      }
    }
  }
  if (blockIdx.y == 0)
    for (int r = threadIdx.x; r < nrows; r += kNormThreads) {
      if (rew_out) rew_out[r0 + r] = rew[r0 + r];
      if (dones_out) dones_out[r0 + r] = dones[r0 + r];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// single-launch form.  Header words (uint32, from byte 16; all zero between calls): [0] arrivals, [1] published flag, [2] consumers.
// All CTAs of the grid must be resident together (they wait for each other): the host launches it only with <= one CTA per SM.
// ---------------------------------------------------------------------------------------------------------------
struct FusedGeom {
  int64_t rows;
  int cols;
  int cslots;          // power of two >= cols (<= 1024): column slots of a CTA
  int rows_per_cta;
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// per-CTA sums of the single-launch form: doubles [cta][2][cols] (sum of x - pivot, sum of its square) behind the four float vectors.
// Sums run in double and in a fixed order -- float partial means of a few rows each would carry half an ulp of the MEAN into every
// (mean_i - mean)^2, which is large against a small variance (seen: 300 x 1024, 1.6e-6 on a std of 0.019).
__device__ __forceinline__ double* scr_dpart(void* s, int cols, int cta, int which) {
  return reinterpret_cast<double*>(reinterpret_cast<char*>(s) + kNormHeader + (size_t)cols * 16) + ((size_t)cta * 2 + which) * cols;
}

template <int kC>      // rows a thread keeps in registers
__global__ void __launch_bounds__(kNormThreads, 1)
elg_norm_fused_kernel(const __grid_constant__ FusedGeom g, const float* __restrict__ x, float* __restrict__ mean, float* __restrict__ var,
                      float* __restrict__ stdv, int64_t* __restrict__ count, const float eps, const int64_t until, float* __restrict__ out,
                      void* __restrict__ scratch, const float* __restrict__ rew, float* __restrict__ rew_out, const uint8_t* __restrict__ dones,
                      uint8_t* __restrict__ dones_out, const int debug) {
  __shared__ double sd[kNormThreads];
  __shared__ int s_last;
  const int cpt = g.cslots, c = threadIdx.x & (cpt - 1), rs = threadIdx.x / cpt, rsub = kNormThreads / cpt;
  const bool live = c < g.cols;
  const int P = gridDim.x;
  pdl_launch_dependents();
  pdl_wait();
  const int64_t old_count = *count;
  const bool learn = !(until >= 0 && old_count >= until) && !(debug & 8);          // (normalizer.py:62-63)
  const int64_t r0 = (int64_t)blockIdx.x * g.rows_per_cta;
  const int nrows = (int)max((int64_t)0, min((int64_t)g.rows_per_cta, g.rows - r0));
  const float* xb = x + r0 * g.cols + c;
  float v[kC];
#pragma unroll
  for (int k = 0; k < kC; ++k) {
    const int r = rs + k * rsub;
    v[k] = (live && r < nrows) ? xb[(uint32_t)(r * g.cols)] : 0.0f;
  }
  // reward / done columns of this CTA's rows ride along
  for (int r = threadIdx.x; r < nrows; r += kNormThreads) {
    if (rew_out) rew_out[r0 + r] = rew[r0 + r];
    if (dones_out) dones_out[r0 + r] = dones[r0 + r];
  }
  // sum over the row groups of column slot c, in group order (every thread of the column forms the identical sum)
  auto column_sum = [&](double mine) {
    __syncthreads();
    sd[threadIdx.x] = mine;
    __syncthreads();
    double t = 0.0;
    for (int q = 0; q < rsub; ++q) t += sd[q * cpt + c];
    return t;
  };
  float m_use, den;
  if (learn) {
    // CTA sums of (x - pivot) and (x - pivot)^2 in double, pivot = the column's entry in row 0 of the batch: a value of the data's
    // own scale, so the final  M2 = S2 - S1^2 / N  cancels a few bits of 53, whatever the mean / std ratio of the column
    const double pivot = live ? (double)__ldg(x + c) : 0.0;
    double s = 0.0, q2 = 0.0;
#pragma unroll
    for (int k = 0; k < kC; ++k) {
      const double d = (double)v[k] - pivot;
      const bool ok = rs + k * rsub < nrows;
      s += ok ? d : 0.0;
      q2 += ok ? d * d : 0.0;
    }
    const double s_cta = column_sum(s), q_cta = column_sum(q2);
    if (rs == 0 && live) {
      scr_dpart(scratch, g.cols, blockIdx.x, 0)[c] = s_cta;
      scr_dpart(scratch, g.cols, blockIdx.x, 1)[c] = q_cta;
    }
    unsigned* words = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(scratch) + 16);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(words, 1u) == (unsigned)(P - 1);
    __syncthreads();
    if (s_last) {
      // the last CTA to arrive: every pair is in L2.  Plain sums in block order -- the loads of a batch go out together (a loop that
      // consumes each value before it asks for the next pays the L2 latency once per CTA: measured 24 us at 128 CTAs) --, then the
      // update rule (normalizer.py:65-75), the state, and mean / std + eps for everybody
      __threadfence();
      constexpr int kB = 8;
      double a1 = 0.0, a2 = 0.0;
      if (live) {
        for (int p0 = rs; p0 < P; p0 += kB * rsub) {
          double sv[kB], qv[kB];
#pragma unroll
          for (int u = 0; u < kB; ++u) {
            const int p = p0 + u * rsub;
            sv[u] = p < P ? __ldcg(scr_dpart(scratch, g.cols, p, 0) + c) : 0.0;
            qv[u] = p < P ? __ldcg(scr_dpart(scratch, g.cols, p, 1) + c) : 0.0;
          }
#pragma unroll
          for (int u = 0; u < kB; ++u) { a1 += sv[u]; a2 += qv[u]; }
        }
      }
      const double S1 = column_sum(a1), S2 = column_sum(a2);
      const double mean_x_d = pivot + S1 / (double)g.rows;
      const double m2_x_d = fmax(S2 - S1 * S1 / (double)g.rows, 0.0);
      if (rs == 0 && live) {
        const int64_t new_count = old_count + g.rows;
        const float rate = (float)g.rows / (float)new_count;
        const float mean_x = (float)mean_x_d, var_x = (float)(m2_x_d / (double)g.rows);
        const float m_old = mean[c], v_old = var[c];
        const float delta = mean_x - m_old;
        const float m_new = m_old + rate * delta;
        const float v_new = v_old + rate * (var_x - v_old + delta * (mean_x - m_new));
        const float s_new = __fsqrt_rn(v_new);
        mean[c] = m_new; var[c] = v_new; stdv[c] = s_new;
        if (c == 0) *count = new_count;
        scr_vec(scratch, g.cols, 2)[c] = m_new;
        scr_vec(scratch, g.cols, 3)[c] = s_new + eps;
      }
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) {
        if (P > 1) st_release_u32(words + 1, 1u);
        else words[0] = 0u;                                   // nobody waits: leave the header zeroed
      }
    } else {
      if (threadIdx.x == 0) {
        while (ld_acquire_u32(words + 1) == 0u && !(debug & 4)) {
          if (!(debug & 16)) __nanosleep(20);
        }
        // the consumer that brings the count to P - 1 is the last reader of the flag: it zeroes the header for the next call
        if (atomicAdd(words + 2, 1u) == (unsigned)(P - 2)) { words[0] = 0u; words[2] = 0u; st_release_u32(words + 1, 0u); }
      }
      __syncthreads();
    }
    m_use = live ? __ldcg(scr_vec(scratch, g.cols, 2) + c) : 0.0f;
    den = live ? __ldcg(scr_vec(scratch, g.cols, 3) + c) : 1.0f;
  } else {
    m_use = live ? mean[c] : 0.0f;
    den = live ? stdv[c] + eps : 1.0f;
  }
  if (live && out) {
    float* ob = out + r0 * g.cols + c;
#pragma unroll
    for (int k = 0; k < kC; ++k) {
      const int r = rs + k * rsub;
      if (r < nrows) ob[(uint32_t)(r * g.cols)] = (v[k] - m_use) / den;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// column-parallel single launch: a CTA owns FOUR columns and every row of them (<= 32 values per thread: up to 8192 rows), so the
// batch statistics of its columns need no other CTA -- no grid-wide hand-over, no second launch, the rows stay in registers from the
// load to the normalised store.  Four consecutive lanes read the 16 contiguous bytes of a row (half a sector: the other half is
// the neighbour CTA's and comes out of L2).  Sums of (x - pivot) and (x - pivot)^2 in double (pivot: the column's entry in row 0),
// reduced in a fixed order: shuffles inside a warp, warp totals in shared memory, one thread per column adds them in warp order.
// The CTA that finishes last (one ticket) stores the new count -- every CTA has read the old one by then.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kColsPerCta = 4;
// The rows of a column group may be split over the CTAs of a THREAD-BLOCK CLUSTER (1, 2, 4 or 8 CTAs: launch attribute): CTA `rank`
// takes the row passes p = k * cluster_size + rank, forms its partial sums, and the partials cross through DISTRIBUTED SHARED MEMORY
// (one cluster barrier; every CTA reads all partials in rank order and arrives at the identical totals).  4096 x 235 then runs on
// 118 or 236 CTAs with 8 or 4 values per thread instead of 59 CTAs with 16.
template <int kC>
__global__ void __launch_bounds__(kNormThreads, 1)
elg_norm_cols_kernel(const int64_t rows, const int cols, const float* __restrict__ x, float* __restrict__ mean, float* __restrict__ var,
                     float* __restrict__ stdv, int64_t* __restrict__ count, const float eps, const int64_t until, float* __restrict__ out,
                     void* __restrict__ scratch, const float* __restrict__ rew, float* __restrict__ rew_out, const uint8_t* __restrict__ dones,
                     uint8_t* __restrict__ dones_out) {
  namespace cg = cooperative_groups;
  constexpr int kRowsPerPass = kNormThreads / kColsPerCta;   // 256
  __shared__ double s_w[2][kNormThreads / 32][kColsPerCta];
  __shared__ double s_part[2][kColsPerCta];                  // this CTA's partial sums: read by the other CTAs of the cluster
  __shared__ float s_mean[kColsPerCta], s_den[kColsPerCta];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned cs = cluster.num_blocks(), rank = cluster.block_rank();
  const int c4 = threadIdx.x & (kColsPerCta - 1), rb = threadIdx.x / kColsPerCta;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = (blockIdx.x / cs) * kColsPerCta + c4;
  const bool live = col < cols;
  pdl_launch_dependents();
  pdl_wait();
  const int64_t old_count = *count;
  const bool learn = !(until >= 0 && old_count >= until);          // (normalizer.py:62-63)
  const float* xb = x + col;
  float v[kC];
#pragma unroll
  for (int k = 0; k < kC; ++k) {
    const int64_t r = rb + (int64_t)(k * cs + rank) * kRowsPerPass;
    v[k] = (live && r < rows) ? xb[r * cols] : 0.0f;
  }
  {      // reward / done columns: this CTA's share of the rows
    const int64_t per = (rows + gridDim.x - 1) / gridDim.x, lo = (int64_t)blockIdx.x * per, hi = min(rows, lo + per);
    for (int64_t r = lo + threadIdx.x; r < hi; r += kNormThreads) {
      if (rew_out) rew_out[r] = rew[r];
      if (dones_out) dones_out[r] = dones[r];
    }
  }
  if (learn) {
    const double pivot = live ? (double)__ldg(xb) : 0.0;
    double s = 0.0, q = 0.0;
#pragma unroll
    for (int k = 0; k < kC; ++k) {
      const double d = (double)v[k] - pivot;
      const bool ok = rb + (int64_t)(k * cs + rank) * kRowsPerPass < rows;
      s += ok ? d : 0.0;
      q += ok ? d * d : 0.0;
    }
#pragma unroll
    for (int o = kColsPerCta; o < 32; o <<= 1) {      // lanes of one column: lane & 3
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane < kColsPerCta) { s_w[0][warp][lane] = s; s_w[1][warp][lane] = q; }
    __syncthreads();
    if (threadIdx.x < kColsPerCta) {
      double p1 = 0.0, p2 = 0.0;
      for (int w = 0; w < kNormThreads / 32; ++w) { p1 += s_w[0][w][c4]; p2 += s_w[1][w][c4]; }
      s_part[0][c4] = p1;
      s_part[1][c4] = p2;
    }
    cluster.sync();      // every CTA's partials are in its shared memory (a plain CTA barrier when the cluster is one CTA)
    if (threadIdx.x < kColsPerCta && live) {
      double S1 = 0.0, S2 = 0.0;
      for (unsigned r = 0; r < cs; ++r) {      // rank order: every CTA of the cluster forms the identical totals
        const double* remote = cluster.map_shared_rank(&s_part[0][0], r);
        S1 += remote[c4];
        S2 += remote[kColsPerCta + c4];
      }
      const double mean_x_d = pivot + S1 / (double)rows;
      const double m2_x_d = fmax(S2 - S1 * S1 / (double)rows, 0.0);
      const int64_t new_count = old_count + rows;      // (normalizer.py:65-75)
      const float rate = (float)rows / (float)new_count;
      const float mean_x = (float)mean_x_d, var_x = (float)(m2_x_d / (double)rows);
      const float m_old = mean[col], v_old = var[col];
      const float delta = mean_x - m_old;
      const float m_new = m_old + rate * delta;
      const float v_new = v_old + rate * (var_x - v_old + delta * (mean_x - m_new));
      const float s_new = __fsqrt_rn(v_new);
      s_mean[c4] = m_new;
      s_den[c4] = s_new + eps;
      // (the CTAs of a cluster read the old state here and rank 0 writes the new one after the second barrier)
      if (rank == 0) { s_w[0][0][c4] = (double)v_new; s_w[1][0][c4] = (double)s_new; }
    }
    cluster.sync();      // the partials have been read (a CTA may leave), s_mean / s_den are visible
    if (rank == 0 && threadIdx.x < kColsPerCta && live) {
      mean[col] = s_mean[c4]; var[col] = (float)s_w[0][0][c4]; stdv[col] = (float)s_w[1][0][c4];
    }
    if (threadIdx.x == 0) {      // the last CTA to get here stores the new count: every CTA has read the old one
      unsigned* ticket = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(scratch) + 16);
      if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
        *count = old_count + rows;
        *ticket = 0u;
      }
    }
  } else {
    if (threadIdx.x < kColsPerCta && live) { s_mean[c4] = mean[col]; s_den[c4] = stdv[col] + eps; }
    __syncthreads();
  }
  if (live && out) {
    const float m = s_mean[c4], den = s_den[c4];
    float* ob = out + col;
#pragma unroll
    for (int k = 0; k < kC; ++k) {
      const int64_t r = rb + (int64_t)(k * cs + rank) * kRowsPerPass;
      if (r < rows) ob[r * cols] = (v[k] - m) / den;
    }
  }
}

template <typename K, typename... Args>
static void launch_cols(K k, unsigned col_groups, unsigned cluster, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(col_groups * cluster);
  cfg.blockDim = dim3(kNormThreads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = cluster;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cluster > 1 ? 2 : 1;
  cudaLaunchKernelEx(&cfg, k, args...);
}

// geometry of the single-launch form, or cslots == 0 when the batch does not fit one wave of register-resident rows
static FusedGeom make_fused_geom(int64_t rows, int cols, int sms) {
  FusedGeom g{};
  g.rows = rows;
  g.cols = cols;
  if (cols > kNormThreads || sms < 1) return g;
  int cs = 32;
  while (cs < cols) cs <<= 1;
  const int rsub = kNormThreads / cs;
  int ctas = sms > kNormParts * 2 ? kNormParts * 2 : sms;      // <= 64 CTAs: the last one sums one (S1, S2) pair per CTA and column
  // power-of-two CTA counts keep the per-thread row count a small power of two as well (4096 rows -> 64 CTAs x 64 rows)
  int p2 = 1;
  while (p2 * 2 <= ctas) p2 <<= 1;
  int64_t rpc = (rows + p2 - 1) / p2;
  if (rpc < rsub) rpc = rsub;
  const int64_t per_thread = (rpc + rsub - 1) / rsub;
  if (per_thread > 32) return g;
  if (rpc * cols >= ((int64_t)1 << 31)) return g;
  g.cslots = cs;
  g.rows_per_cta = (int)rpc;
  return g;
}

static NormGeom make_geom(int64_t rows, int cols) {
  NormGeom g{};
  g.rows = rows;
  g.cols = cols;
  g.cpt = norm_cpt(cols);
  g.cpt_s = cols > 32 ? 64 : 32;
  int64_t rpp = (rows + kNormParts - 1) / kNormParts;
  if (rpp < 32) rpp = 32;
  g.rows_per_part = (int)rpp;
  g.parts = (int)((rows + rpp - 1) / rpp);
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

// bits 0-1: 0 (default) = the column-parallel single launch for batches of <= 65 536 rows, the statistics + apply pair otherwise;
// 1 = always the pair; 2 = the row-parallel single launch with a grid-wide hand-over (measured slower than the pair: 14.8 vs 11.0 us
// at 4096 x 235; kept for A/B runs).  Bits 2-4: force the cluster size of form 0 (1 -> 1 CTA, 2 -> 2, 3 -> 4, 4 -> 8).  Bits 5+:
// measurement switches of form 2 (results invalid): 32 = consumers do not wait, 64 = no statistics, 128 = spin without nanosleep
int g_norm_mode = 0;

template <typename K, typename... Args>
static void launch_fused(K k, dim3 grid, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kNormThreads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, k, args..., (int)((g_norm_mode >> 5) << 2));
}

}  // namespace elg

extern "C" {

int elg_set_normalizer_tuning(int mode) {
  if (mode < 0 || mode > 255) return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "normalizer tuning mode must be 0 or 1 (+ measurement bits)");
  elg::g_norm_mode = mode;
  return ELG_OK;
}

int64_t elg_normalizer_scratch_bytes(int64_t num_rows, int32_t num_cols) {
  if (num_rows < 0 || num_cols < 1) return 0;
  const elg::NormGeom g = elg::make_geom(num_rows > 0 ? num_rows : 1, num_cols);
  // the single-launch form stores (mean, M2) as doubles for each of its <= 128 CTAs in the same region
  int64_t part_bytes = (int64_t)12 * num_cols * g.parts;
  if (part_bytes < (int64_t)16 * num_cols * 4 * elg::kNormParts) part_bytes = (int64_t)16 * num_cols * 4 * elg::kNormParts;
  return elg::kNormHeader + (int64_t)16 * num_cols + part_bytes;
}

int elg_normalize_observations(int64_t num_rows, int32_t num_cols, const float* x, float* mean, float* var, float* std, int64_t* count, float eps,
                               int64_t until, int32_t training, float* out, void* scratch, const float* rew, float* rew_out,
                               const uint8_t* dones, uint8_t* dones_out, void* stream) {
  if (num_rows < 0 || num_cols < 1) return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "normalizer: num_rows < 0 or num_cols < 1");
  if (num_cols > 32 * elg::kNormMaxTiles) return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "normalizer: more than 1920 columns");
  if (num_rows * (int64_t)num_cols >= ((int64_t)1 << 40)) return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "normalizer: batch too large");
  // (an empty batch first: torch hands out NULL for the storage of an empty tensor)
  if (num_rows == 0) return training ? elg::set_error(ELG_ERR_INVALID_ARGUMENT, "normalizer: an empty batch cannot update the statistics") : ELG_OK;
  if (!x || !mean || !var || !std || !count) return elg::set_error(ELG_ERR_NULL_POINTER, "normalizer: x / mean / var / std / count is NULL");
  if (training && !scratch) return elg::set_error(ELG_ERR_NULL_POINTER, "normalizer: training mode needs the scratch buffer");
  if (scratch && (reinterpret_cast<uintptr_t>(scratch) & 15u) != 0) return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "normalizer: scratch must be 16-byte aligned");
  if ((rew_out && !rew) || (dones_out && !dones)) return elg::set_error(ELG_ERR_NULL_POINTER, "normalizer: reward / done destination without a source");
  const elg::NormGeom g = elg::make_geom(num_rows, num_cols);
  // 32-bit row offsets inside one row block / one apply block
  if ((int64_t)g.rows_per_part * num_cols >= ((int64_t)1 << 31)) return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "normalizer: batch too large");
  cudaStream_t s = (cudaStream_t)stream;
  const int norm_form = elg::g_norm_mode & 3;
  if (training && (norm_form == 0 || norm_form == 3) && num_rows <= 8 * 32 * (elg::kNormThreads / elg::kColsPerCta)) {
    const int passes = (int)((num_rows + elg::kNormThreads / elg::kColsPerCta - 1) / (elg::kNormThreads / elg::kColsPerCta));
    // thread-block cluster over the rows: the largest of 1 / 2 / 4 / 8 CTAs that keeps the grid within one CTA per SM (measured at
    // 4096 x 235, 59 column groups: 9.4 / 7.1 / 10.3 us with 1 / 2 / 4 CTAs per cluster -- 236 CTAs no longer fit one per SM; at
    // 4096 x 48, 12 groups: 8.1 / 6.4 / 5.1 us) and leaves every CTA at least one row pass; bits 2-4 of the tuning word force a size
    const unsigned groups = (unsigned)((num_cols + elg::kColsPerCta - 1) / elg::kColsPerCta);
    const int sms = elg::sm_count();
    unsigned cluster = 1u;
    while (cluster < 8u && groups * cluster * 2u <= (unsigned)(sms > 0 ? sms : 1) && (int)(cluster * 2u) <= passes) cluster *= 2u;
    const int forced = (elg::g_norm_mode >> 2) & 7;
    if (forced >= 1 && forced <= 4) cluster = 1u << (forced - 1);
    while (passes > 32 * (int)cluster && cluster < 8u) cluster *= 2u;
    // more clusters than fit one CTA per SM (e.g. 32 832 x 235: 59 groups x 8) lose to the pair below: 76.9 vs 26.8 us
    const bool fits = forced || groups * cluster <= (unsigned)(sms > 0 ? sms : 1);
    const int per_thread = (passes + (int)cluster - 1) / (int)cluster;
    if (!fits || per_thread > 32) goto two_launches;
    if (per_thread <= 4)
      elg::launch_cols(elg::elg_norm_cols_kernel<4>, groups, cluster, s, num_rows, (int)num_cols, x, mean, var, std, count, eps, until, out, scratch, rew, rew_out, dones, dones_out);
    else if (per_thread <= 8)
      elg::launch_cols(elg::elg_norm_cols_kernel<8>, groups, cluster, s, num_rows, (int)num_cols, x, mean, var, std, count, eps, until, out, scratch, rew, rew_out, dones, dones_out);
    else if (per_thread <= 16)
      elg::launch_cols(elg::elg_norm_cols_kernel<16>, groups, cluster, s, num_rows, (int)num_cols, x, mean, var, std, count, eps, until, out, scratch, rew, rew_out, dones, dones_out);
    else
      elg::launch_cols(elg::elg_norm_cols_kernel<32>, groups, cluster, s, num_rows, (int)num_cols, x, mean, var, std, count, eps, until, out, scratch, rew, rew_out, dones, dones_out);
    return elg::check_launch("elg_normalize_observations");
  }
two_launches:
  if (training && norm_form == 2) {      // the grid-wide hand-over form: measured slower than the pair below (A/B only)
    const elg::FusedGeom fg = elg::make_fused_geom(num_rows, num_cols, elg::sm_count());
    if (fg.cslots > 0) {
      const int rsub = elg::kNormThreads / fg.cslots;
      const int per_thread = (fg.rows_per_cta + rsub - 1) / rsub;
      const dim3 grid((unsigned)((num_rows + fg.rows_per_cta - 1) / fg.rows_per_cta));
      if (per_thread <= 8)
        elg::launch_fused(elg::elg_norm_fused_kernel<8>, grid, s, fg, x, mean, var, std, count, eps, until, out, scratch, rew, rew_out, dones, dones_out);
      else if (per_thread <= 16)
        elg::launch_fused(elg::elg_norm_fused_kernel<16>, grid, s, fg, x, mean, var, std, count, eps, until, out, scratch, rew, rew_out, dones, dones_out);
      else
        elg::launch_fused(elg::elg_norm_fused_kernel<32>, grid, s, fg, x, mean, var, std, count, eps, until, out, scratch, rew, rew_out, dones, dones_out);
      return elg::check_launch("elg_normalize_observations");
    }
  }
  if (training) {
    const dim3 grid((unsigned)g.parts, (unsigned)((num_cols + g.cpt_s - 1) / g.cpt_s));
    const int per_thread = (g.rows_per_part + (elg::kNormThreads / g.cpt_s) - 1) / (elg::kNormThreads / g.cpt_s);
    if (per_thread <= 8)
      elg::launch_pdl(elg::elg_norm_stats_kernel<8>, grid, elg::kNormThreads, s, g, x, (const float*)mean, (const float*)var, (const int64_t*)count, until, scratch);
    else
      elg::launch_pdl(elg::elg_norm_stats_kernel<32>, grid, elg::kNormThreads, s, g, x, (const float*)mean, (const float*)var, (const int64_t*)count, until, scratch);
  }
  const int64_t want = (num_rows + 31) / 32;
  const unsigned blocks = (unsigned)(want < 148 ? want : 148);
  elg::launch_pdl(elg::elg_norm_apply_kernel, dim3(blocks, (unsigned)((num_cols + g.cpt - 1) / g.cpt)), elg::kNormThreads, s, g, x, mean, var, std, count, eps,
                  until, (int)training, out, scratch, rew, rew_out, dones, dones_out);
  return elg::check_launch("elg_normalize_observations");
}

}  // extern "C"
