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
#include "elg_async.cuh"

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

// ---------------------------------------------------------------------------------------------------------------
// The same clone with the TMA doing the writing.  For every field the CTA replicates the main row tile_rows times in
// shared memory -- rotated so that the tile starts at the first 16-byte boundary of the destination span -- and then
// ONE cp.async.bulk (shared -> global) per field and chunk moves tile_rows * row_bytes bytes: a few instructions per
// kilobyte instead of one 16-byte store per thread.  grid = (chunk slices, num_main); chunk c of a field covers span
// words [head + c * tile_words, head + (c + 1) * tile_words).  The <= 3 words in front of the first boundary and the
// <= 3 behind the last full vector are plain stores.  The drift field (root_states when the position drift is on) gets
// its tile rebuilt per chunk, the three position words of every row drifted; all other tiles are built once per CTA.
// No integer division anywhere: a modulo by a run-time row length costs more than a whole tile.
// ---------------------------------------------------------------------------------------------------------------
struct CloneBulkPlan {
  int tile_rows;                        // multiple of 4, <= rollouts_per_main
  int tile_off[ELG_MAX_CLONE_FIELDS];   // byte offset of the field's tile in dynamic shared memory (16-byte aligned)
  int row_off[ELG_MAX_CLONE_FIELDS];    // word offset of the field's staged main row
  int tiles_bytes;                      // staged rows live behind the tiles
  int row_words;                        // staged words per main over all fields
  int debug;                            // measurement switches: 2 = build only (no stores), 4 = stores only (no build)
};

__global__ void __launch_bounds__(kCloneThreads)
elg_clone_bulk_kernel(const __grid_constant__ ElgCloneTable tb, const __grid_constant__ CloneBulkPlan pl, const float drift,
                      const float* __restrict__ drift_u, const uint64_t seed, const uint64_t offset) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint32_t* const s_row = reinterpret_cast<uint32_t*>(smem_raw + pl.tiles_bytes);
  constexpr int kWarps = kCloneThreads / 32;
  const int k = blockIdx.y;                       // main env
  const int R = tb.rollouts_per_main, TR = pl.tile_rows;
  const size_t main_row = (size_t)k * (1 + R);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nf = tb.num_fields;
  const bool drift_on = tb.drift_field >= 0 && drift > 0.0f;
  pdl_launch_dependents();
  pdl_wait();
  auto head_words = [&](int f) {   // words of the span in front of its first 16-byte boundary
    const int rw = tb.fields[f].row_bytes >> 2;
    const uintptr_t dst = reinterpret_cast<uintptr_t>(tb.fields[f].base) + (main_row + 1) * rw * 4;
    const int h = (int)(((16 - (dst & 15)) & 15) >> 2);
    return h < R * rw ? h : R * rw;
  };
  auto drift_word = [&](int rr, int j, uint32_t v) -> uint32_t {   // position component j of rollout rr of this main
    const long long gi = ((long long)k * R + rr) * 3 + j;
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
  // Stage the main rows and build the replicated, rotated tiles (tile[i] = row[(head + i) % rw]) in one pass without a
  // barrier in between: staged word x of the row set belongs to ONE thread pair, which fetches it from global memory
  // once (all loads of the CTA in flight together), keeps it in a register and writes every copy of it -- a stream of
  // independent shared-memory stores, copy r to tile word r * rw + j - head.  Thread x + 128 h takes the copies r = h, h + 2, ...
  for (int x0 = 0; x0 < pl.row_words; x0 += kCloneThreads / 2) {
    const int x = x0 + (tid & (kCloneThreads / 2 - 1)), half = tid / (kCloneThreads / 2);
    if (x < pl.row_words) {
      int f = 0;
      while (f + 1 < nf && pl.row_off[f + 1] <= x) ++f;
      const int rw = tb.fields[f].row_bytes >> 2, j = x - pl.row_off[f];
      const uint32_t v = (static_cast<const uint32_t*>(tb.fields[f].base) + main_row * rw)[j];
      if (half == 0) s_row[x] = v;
      if (!(pl.debug & 4)) {
        uint32_t* tile = reinterpret_cast<uint32_t*>(smem_raw + pl.tile_off[f]);
        const int tw = TR * rw;
        int i = half * rw + j - head_words(f);
#pragma unroll 4
        for (int r = half; r <= TR + 3; r += 2, i += 2 * rw)   // head <= 3 words can span up to 3 one-word rows
          if (i >= 0 && i < tw) tile[i] = v;
      }
    }
  }
  fence_async_smem();
  __syncthreads();
  // chunks of this CTA: lane 0 of warp w issues the bulk stores of its fields
  const int nchunk = (R + TR - 1) / TR;
  if (lane == 0 && !(pl.debug & 2)) {
    for (int f = warp; f < nf; f += kWarps) {
      if (drift_on && f == tb.drift_field) continue;
      const int rw = tb.fields[f].row_bytes >> 2;
      const int head = head_words(f), span = R * rw, tw = TR * rw;
      uint32_t* dst = static_cast<uint32_t*>(tb.fields[f].base) + (main_row + 1) * rw;
      for (int c = blockIdx.x; c < nchunk; c += gridDim.x) {
        const int w0 = head + c * tw;
        int words = span - w0;
        words = words < tw ? words : tw;
        words &= ~3;
        if (words > 0) bulk_s2g(dst + w0, smem_raw + pl.tile_off[f], (uint32_t)words * 4u);
      }
    }
    bulk_commit();
  }
  // the drifting field: its tile was built like the others; per chunk only the three position words of every row change.
  // One thread per (row of the chunk, component): dense Philox / one load each, written over the tile in place.
  if (drift_on) {
    const int f = tb.drift_field;
    const int rw = tb.fields[f].row_bytes >> 2;
    const uint32_t* row = s_row + pl.row_off[f];
    uint32_t* tile = reinterpret_cast<uint32_t*>(smem_raw + pl.tile_off[f]);
    const int head = head_words(f), span = R * rw, tw = TR * rw;
    uint32_t* dst = static_cast<uint32_t*>(tb.fields[f].base) + (main_row + 1) * rw;
    bool pending = false;
    for (int c = blockIdx.x; c < nchunk; c += gridDim.x) {
      if (pending) bulk_wait_read_all();   // (thread 0) the previous chunk has left the tile
      __syncthreads();
      const int w0 = head + c * tw;
      int words = span - w0;
      words = words < tw ? words : tw;
      words &= ~3;
      // the chunk covers span words [w0, w0 + words): rows c * TR .. c * TR + TR + (head words' worth)
      for (int q = tid; q < 3 * (TR + 4); q += kCloneThreads) {
        const int rl = q / 3, j = q - 3 * rl;      // (constant divisor)
        const int rr = c * TR + rl, i = rr * rw + j - w0;
        if (rr < R && i >= 0 && i < words) tile[i] = drift_word(rr, j, row[j]);
      }
      fence_async_smem();
      __syncthreads();
      if (tid == 0 && words > 0) {
        bulk_s2g(dst + w0, tile, (uint32_t)words * 4u);
        bulk_commit();
        pending = true;
      }
    }
  }
  // head and tail words of every span (slice 0): 8 threads per field, thread t < head writes head word t, thread 4 + t tail word t
  if (blockIdx.x == 0 && tid < 8 * nf) {
    const int f = tid >> 3, t = tid & 7;
    const int rw = tb.fields[f].row_bytes >> 2;
    const uint32_t* row = s_row + pl.row_off[f];
    uint32_t* dst = static_cast<uint32_t*>(tb.fields[f].base) + (main_row + 1) * rw;
    const int head = head_words(f), span = R * rw;
    const int tailw = (span - head) & 3;
    int w = -1, mm = 0, rr = 0;   // span word, its column, its rollout row
    if (t < head) {
      w = t;
      mm = t;
      while (mm >= rw) { mm -= rw; ++rr; }
    } else if (t >= 4 && t < 4 + tailw) {
      w = span - tailw + (t - 4);
      mm = (t - 4) - tailw;        // == w (mod rw) because span is a multiple of rw
      rr = R;
      while (mm < 0) { mm += rw; --rr; }
    }
    if (w >= 0) {
      uint32_t v = row[mm];
      if (drift_on && f == tb.drift_field && mm < 3) v = drift_word(rr, mm, v);
      dst[w] = v;
    }
  }
  if (lane == 0) bulk_wait_read_all();   // shared memory must outlive the reads
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
int g_clone_tune = 0;   // elg_set_clone_tuning: 1 = per-thread 16-byte stores instead of TMA bulk stores; 2 / 4 = measurement switches
}  // namespace

extern "C" {

int elg_sizeof_clone_table(void) { return (int)sizeof(ElgCloneTable); }

int elg_set_clone_tuning(int disable_bulk) {
  g_clone_tune = disable_bulk;
  return ELG_OK;
}

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
    // TMA path: word-sized rows, at least 4 rollouts per main; tile of up to 32 rows per field, <= 160 KB of shared memory
    if (words && staged > 0 && staged <= elg::kCloneMaxRowWords && table->rollouts_per_main >= 4 && (g_clone_tune & 1) == 0) {
      int tr_cap = 32;      // (scripts/clone_ab.py --tiles, 64 mains x 512 rollouts: tiles of 32 / 64 / 128 / 256 rows -> 6.1-6.3 / 6.6 / 7.7 / 7.9 us)
      if (((g_clone_tune >> 16) & 0xfff) > 0) tr_cap = (g_clone_tune >> 16) & 0xfff;   // measurement override: rows per tile
      int tr = table->rollouts_per_main < tr_cap ? table->rollouts_per_main : tr_cap;
      const int fit = (160 * 1024 - staged * 4) / (staged * 4);
      if (tr > fit) tr = fit;
      tr &= ~3;
      if (tr >= 4) {
        elg::CloneBulkPlan pl{};
        pl.tile_rows = tr;
        pl.debug = g_clone_tune & 0xff;
        int toff = 0, roff = 0;
        for (int f = 0; f < table->num_fields; ++f) {
          pl.tile_off[f] = toff;
          pl.row_off[f] = roff;
          toff += tr * table->fields[f].row_bytes;   // multiple of 16: tr % 4 == 0, row_bytes % 4 == 0
          roff += table->fields[f].row_bytes / 4;
        }
        pl.tiles_bytes = toff;
        pl.row_words = roff;
        const size_t smem = (size_t)toff + (size_t)roff * 4;
        static elg::SmemCache smem_cache = {};
  size_t& smem_set = elg::smem_slot(smem_cache);
        if (smem > smem_set) {
          if (cudaFuncSetAttribute(elg::elg_clone_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return cfail(ELG_ERR_CUDA, "cannot reserve dynamic shared memory for elg_clone_bulk_kernel");
          smem_set = smem;
        }
        const long long nchunk = ((long long)table->rollouts_per_main + tr - 1) / tr;
        // one chunk per CTA and field while that keeps the grid below ~8 CTAs per SM (measured, scripts/clone_ab.py: at
        // 64 mains x 512 rollouts 8 slices take 6.6 us, 5 slices 8.5 us, 1 slice 7.9 us)
        long long sl = (8LL * sms + table->num_main - 1) / table->num_main;
        if (((g_clone_tune >> 8) & 0xff) > 0) sl = (g_clone_tune >> 8) & 0xff;   // measurement override
        if (sl > nchunk) sl = nchunk;
        if (sl < 1) sl = 1;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)sl, (unsigned)table->num_main);
        cfg.blockDim = dim3(elg::kCloneThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, elg::elg_clone_bulk_kernel, *table, pl, drift, drift_u, seed, offset);
        return elg::check_launch("elg_clone_rows");
      }
    }
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
