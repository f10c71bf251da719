// elg_clone.cu -- main -> rollout state clone, main-row cache and restore for sm_100a.
//
// Replaces RobotBatchRollout._sync_main_to_rollout / _cache_main_env_states / _restore_main_env_states
// (envs/batch_rollout/robot_batch_rollout.py:1447-1535, :1537-1583, :1585-1640 in the reference): a Python loop
// over the main envs that rebuilds an index tensor, then 14 gather -> scatter pairs, every call.
//
// Layout facts the kernel relies on (robot_batch_rollout.py:119-164): main env k sits at row k (1 + R), its R
// rollout envs directly behind it.  For every field the destination of main k is therefore ONE contiguous span of
// R * row_bytes bytes, and the source is the row in front of it, repeated.  One launch handles every field:
//   grid = (slices, num_main); a CTA stages the main row of every field in shared memory (a few hundred bytes),
//   then streams its slice of each span with 16-byte stores (scalar head / tail up to the first / after the last
//   16-byte boundary), the source word being (word index mod row words) of the staged row.
// The optional position drift (domain_rand.rollout_envs_sync_pos_drift) is applied to words 0..2 of the drift field
// (root_states: base_pos is a view of it) as pos + (u - 0.5) * drift with individually rounded ops, u either a
// caller tensor (= torch.rand_like, parity mode) or in-kernel Philox4x32-10.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "elg_common.cuh"

namespace elg {

constexpr int kCloneThreads = 256;
constexpr int kCloneMaxRowWords = 2048;   // staged words per main over all fields

__global__ void __launch_bounds__(kCloneThreads)
elg_clone_sync_kernel(const __grid_constant__ ElgCloneTable tb, const float drift, const float* __restrict__ drift_u,
                      const uint64_t seed, const uint64_t offset) {
  __shared__ __align__(16) uint32_t s_row[kCloneMaxRowWords];
  const int k = blockIdx.y;                       // main env
  const int R = tb.rollouts_per_main;
  const size_t main_row = (size_t)k * (1 + R);
  const int tid = threadIdx.x;
  // stage the main row of every field
  int w0 = 0;
  for (int f = 0; f < tb.num_fields; ++f) {
    const int rw = tb.fields[f].row_bytes >> 2;
    const uint32_t* src = static_cast<const uint32_t*>(tb.fields[f].base) + main_row * rw;
    for (int i = tid; i < rw; i += kCloneThreads) s_row[w0 + i] = src[i];
    w0 += rw;
  }
  __syncthreads();
  w0 = 0;
  for (int f = 0; f < tb.num_fields; ++f) {
    const int rw = tb.fields[f].row_bytes >> 2;
    const uint32_t* row = s_row + w0;
    w0 += rw;
    uint32_t* dst = static_cast<uint32_t*>(tb.fields[f].base) + (main_row + 1) * rw;   // span of R * rw words
    const int span = R * rw;                                                            // < 2^31 (checked by the host)
    const bool drifting = (f == tb.drift_field) && drift > 0.0f;
    // words up to the first 16-byte boundary, vectors, tail
    int headw = (int)(((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15) >> 2);
    if (headw > span) headw = span;
    const int nvec = (span - headw) >> 2;
    const int tailw = span - headw - 4 * nvec;
    auto drifted = [&](int w, int m, uint32_t v) -> uint32_t {   // word w of the span, m == w % rw < 3: base position + drift
      const long long gi = ((long long)k * R + w / rw) * 3 + m;
      float u;
      if (drift_u) {
        u = drift_u[gi];
      } else {
        const uint4 b = philox4x32_10(make_uint4((uint32_t)gi, (uint32_t)(gi >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)),
                                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        u = u01(b.x);
      }
      return __float_as_uint(add_r(__uint_as_float(v), mul_r(sub_r(u, 0.5f), drift)));
    };
    // this CTA's slice of the vectors; the source index advances modulo the row length without divisions
    const int per = (nvec + gridDim.x - 1) / gridDim.x;
    const int v_lo = blockIdx.x * per, v_hi = min(nvec, v_lo + per);
    uint4* dv = reinterpret_cast<uint4*>(dst + headw);
    int v = v_lo + tid;
    int m = (headw + 4 * v) % rw;
    const int step = (4 * kCloneThreads) % rw;
    for (; v < v_hi; v += kCloneThreads) {
      int m1 = m + 1; m1 = m1 >= rw ? m1 - rw : m1;
      int m2 = m1 + 1; m2 = m2 >= rw ? m2 - rw : m2;
      int m3 = m2 + 1; m3 = m3 >= rw ? m3 - rw : m3;
      uint4 o = make_uint4(row[m], row[m1], row[m2], row[m3]);
      if (drifting) {
        const int w = headw + 4 * v;
        if (m < 3) o.x = drifted(w, m, o.x);
        if (m1 < 3) o.y = drifted(w + 1, m1, o.y);
        if (m2 < 3) o.z = drifted(w + 2, m2, o.z);
        if (m3 < 3) o.w = drifted(w + 3, m3, o.w);
      }
      dv[v] = o;
      m += step;
      m = m >= rw ? m - rw : m;
    }
    if (blockIdx.x == 0) {
      if (tid < headw) {
        const int mm = tid % rw;
        const uint32_t val = row[mm];
        dst[tid] = (drifting && mm < 3) ? drifted(tid, mm, val) : val;
      }
      if (tid < tailw) {
        const int w = headw + 4 * nvec + tid, mm = w % rw;
        const uint32_t val = row[mm];
        dst[w] = (drifting && mm < 3) ? drifted(w, mm, val) : val;
      }
    }
  }
}

// rows whose size is not a multiple of 4 bytes (e.g. last_contacts with F % 4 != 0): byte-wise fallback, all fields
__global__ void __launch_bounds__(kCloneThreads)
elg_clone_sync_bytes_kernel(const __grid_constant__ ElgCloneTable tb, const float drift, const float* __restrict__ drift_u,
                            const uint64_t seed, const uint64_t offset) {
  const int k = blockIdx.y;
  const int R = tb.rollouts_per_main;
  const size_t main_row = (size_t)k * (1 + R);
  for (int f = 0; f < tb.num_fields; ++f) {
    const int rb = tb.fields[f].row_bytes;
    uint8_t* base = static_cast<uint8_t*>(tb.fields[f].base);
    const uint8_t* src = base + main_row * rb;
    uint8_t* dst = base + (main_row + 1) * rb;
    const long long span = (long long)R * rb;
    const bool drifting = (f == tb.drift_field) && drift > 0.0f;
    for (long long b = (long long)blockIdx.x * kCloneThreads + threadIdx.x; b < span; b += (long long)gridDim.x * kCloneThreads) {
      const int m = (int)(b % rb);
      if (drifting && m < 12) continue;   // drifted words are written below
      dst[b] = src[m];
    }
    if (drifting)
      for (long long i = (long long)blockIdx.x * kCloneThreads + threadIdx.x; i < (long long)R * 3; i += (long long)gridDim.x * kCloneThreads) {
        const long long r = i / 3;
        const int m = (int)(i - 3 * r);
        const long long gi = ((long long)k * R + r) * 3 + m;
        float u;
        if (drift_u) {
          u = drift_u[gi];
        } else {
          const uint4 bl = philox4x32_10(make_uint4((uint32_t)gi, (uint32_t)(gi >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)),
                                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
          u = u01(bl.x);
        }
        const float p = reinterpret_cast<const float*>(src)[m];
        reinterpret_cast<float*>(dst + r * rb)[m] = add_r(p, mul_r(sub_r(u, 0.5f), drift));
      }
  }
}

// cache (main rows -> cache tensors) and restore (cache -> main rows): num_main * sum(row_bytes) bytes, tiny
__global__ void __launch_bounds__(kCloneThreads)
elg_clone_cache_kernel(const __grid_constant__ ElgCloneTable tb, const int restore) {
  const int R = tb.rollouts_per_main;
  for (int f = 0; f < tb.num_fields; ++f) {
    const int rb = tb.fields[f].row_bytes;
    uint8_t* env = static_cast<uint8_t*>(tb.fields[f].base);
    uint8_t* cache = static_cast<uint8_t*>(tb.fields[f].cache);
    if (!cache) continue;
    const long long total = (long long)tb.num_main * rb;
    if ((rb & 3) == 0) {
      const int rw = rb >> 2;
      for (long long i = (long long)blockIdx.x * kCloneThreads + threadIdx.x; i < (total >> 2); i += (long long)gridDim.x * kCloneThreads) {
        const long long k = i / rw;
        const int m = (int)(i - k * rw);
        uint32_t* e = reinterpret_cast<uint32_t*>(env) + (size_t)k * (1 + R) * rw + m;
        uint32_t* c = reinterpret_cast<uint32_t*>(cache) + i;
        if (restore) *e = *c;
        else *c = *e;
      }
    } else {
      for (long long i = (long long)blockIdx.x * kCloneThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kCloneThreads) {
        const long long k = i / rb;
        const int m = (int)(i - k * rb);
        uint8_t* e = env + (size_t)k * (1 + R) * rb + m;
        if (restore) *e = cache[i];
        else cache[i] = *e;
      }
    }
  }
}

}  // namespace elg

namespace {
int cfail(int code, const char* msg) { return elg::set_error(code, msg); }
}  // namespace

extern "C" {

int elg_sizeof_clone_table(void) { return (int)sizeof(ElgCloneTable); }

int elg_clone_rows(const ElgCloneTable* table, int mode, float drift, const float* drift_u, uint64_t seed, uint64_t offset, void* stream) {
  if (!table) return cfail(ELG_ERR_NULL_POINTER, "clone table is NULL");
  if (table->num_fields < 0 || table->num_fields > ELG_MAX_CLONE_FIELDS) return cfail(ELG_ERR_INVALID_ARGUMENT, "num_fields outside [0, ELG_MAX_CLONE_FIELDS]");
  if (table->num_main < 0 || table->rollouts_per_main < 0) return cfail(ELG_ERR_INVALID_ARGUMENT, "negative env counts");
  if (mode < ELG_CLONE_SYNC || mode > ELG_CLONE_RESTORE) return cfail(ELG_ERR_INVALID_ARGUMENT, "mode must be ELG_CLONE_SYNC, _CACHE or _RESTORE");
  if (table->drift_field >= table->num_fields) return cfail(ELG_ERR_INVALID_ARGUMENT, "drift_field out of range");
  bool words = true;
  int staged = 0;
  for (int f = 0; f < table->num_fields; ++f) {
    const ElgCloneField& fd = table->fields[f];
    if (!fd.base) return cfail(ELG_ERR_NULL_POINTER, "a clone field has a NULL base pointer");
    if (fd.row_bytes <= 0) return cfail(ELG_ERR_INVALID_ARGUMENT, "a clone field has row_bytes <= 0");
    if (mode != ELG_CLONE_SYNC && !fd.cache) continue;
    words = words && (fd.row_bytes % 4 == 0) && (reinterpret_cast<uintptr_t>(fd.base) % 4 == 0);
    staged += fd.row_bytes / 4;
  }
  if (table->drift_field >= 0 && table->fields[table->drift_field].row_bytes < 12)
    return cfail(ELG_ERR_INVALID_ARGUMENT, "the drift field needs at least 3 floats per row");
  if (table->num_main == 0 || table->num_fields == 0) return ELG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == ELG_CLONE_SYNC) {
    if (table->rollouts_per_main == 0) return ELG_OK;   // no rollout envs (robot_batch_rollout.py:1452-1453)
    const int sms = elg::sm_count();
    if (sms <= 0) return cfail(ELG_ERR_CUDA, "cannot query the SM count");
    // enough CTAs for 2 per SM, but no more slices than 4 KB pieces of the largest span
    long long slices = (2LL * sms + table->num_main - 1) / table->num_main;
    long long biggest = 0;
    for (int f = 0; f < table->num_fields; ++f) biggest = biggest > table->fields[f].row_bytes ? biggest : table->fields[f].row_bytes;
    const long long cap = ((long long)table->rollouts_per_main * biggest + 4095) / 4096;
    if (slices > cap) slices = cap;
    if (slices < 1) slices = 1;
    if (table->num_main > 65535) return cfail(ELG_ERR_UNSUPPORTED, "more than 65535 main envs");
    if ((long long)table->rollouts_per_main * biggest / 4 > 0x7fffffffLL / 8) return cfail(ELG_ERR_UNSUPPORTED, "rollouts_per_main * row size too large");
    const dim3 grid((unsigned)slices, (unsigned)table->num_main);
    if (words && staged <= elg::kCloneMaxRowWords)
      elg::elg_clone_sync_kernel<<<grid, elg::kCloneThreads, 0, st>>>(*table, drift, drift_u, seed, offset);
    else
      elg::elg_clone_sync_bytes_kernel<<<grid, elg::kCloneThreads, 0, st>>>(*table, drift, drift_u, seed, offset);
  } else {
    long long bytes = 0;
    for (int f = 0; f < table->num_fields; ++f) bytes += (long long)table->num_main * table->fields[f].row_bytes;
    long long blocks = (bytes / 4 + elg::kCloneThreads - 1) / elg::kCloneThreads;
    if (blocks > 1024) blocks = 1024;
    if (blocks < 1) blocks = 1;
    elg::elg_clone_cache_kernel<<<(unsigned)blocks, elg::kCloneThreads, 0, st>>>(*table, mode == ELG_CLONE_RESTORE ? 1 : 0);
  }
  return elg::check_launch("elg_clone_rows");
}

}  // extern "C"
