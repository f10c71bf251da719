// elg_actuator.cu -- actuator-network torques for sm_100a (Anymal._compute_torques, envs/anymal_c/anymal.py:93-105).
//
// The reference pushes a [N*12, 1, 2] batch through a TorchScript LSTM every physics sub-step (decimation x per env
// step): ~15 ATen launches and a [N*12, 32] gate tensor per layer.  Here one thread owns one (env, dof) row end to end:
// 2 inputs, 2 x (8 hidden + 8 cell) state floats in, the same out, one torque -- 276 bytes per row, streamed with
// 16-byte accesses (thread t touches bytes [32 t, 32 t + 32) of each state plane: perfectly coalesced).  The 973
// weights sit in shared memory and are read as warp-uniform broadcasts.  Roofline: HBM (13.6 MB per call at 4096 envs),
// with ~1000 FMAs + 80 transcendentals per row next to it.
#include <cuda_runtime.h>
#include <stdint.h>

#include "elg_common.cuh"
#include "elg_async.cuh"

namespace elg {

constexpr int kActThreads = 64;

// Gate activations.  The row's ~3000 instructions used to be 55 % libdevice expf / tanhf / __frcp_rn (range checks, an out-of-line
// special-operand path per reciprocal, a divergent two-branch tanhf): every one of them a scheduling barrier between the eight
// independent hidden units.  These forms are branch-free and stay inside the north-star tolerance with margin (float32 emulation
// with every MUFU result perturbed by its documented error, tests/golden/actuator_net.npz: <= 0.45 of the 1e-5 / 1e-6 bar, the
// libdevice forms 0.30 -- evaluation-order noise of the float32 graph itself):
//   * MUFU.EX2 on x * log2(e) (2 ulp + the rounding of the product; the sensitivity of both functions to it is t / (1 + t)^k < 1),
//   * 1 / y as MUFU.RCP + one Newton step (<= 1 ulp for y in [1, 1e38]),
//   * tanh below 0.55 as the odd minimax polynomial x + x^3 P(x^2) (< 0.9 ulp), above it 1 - 2 / (e^{2|x|} + 1).
// NaN inputs come out as NaN (the selects below are written so that an unordered compare keeps the NaN operand).
__device__ __forceinline__ float ex2_approx(float a) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ float rcp_newton(float y) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y));
  return fmaf(r, fmaf(-y, r, 1.0f), r);
}
__device__ __forceinline__ float sigmoid_f(float x) {
  float y = 1.0f + ex2_approx(x * -1.4426950408889634f);
  y = y > 1e38f ? 1e38f : y;   // e^-x overflows below x = -88.7: keep the Newton step finite (result 0)
  return rcp_newton(y);
}
__device__ __forceinline__ float tanh_f(float x) {
  const float ax = fabsf(x), x2 = x * x;
  float p = -0.006715521216392517f;
  p = fmaf(p, x2, 0.02136712521314621f);
  p = fmaf(p, x2, -0.05391916632652283f);
  p = fmaf(p, x2, 0.13333165645599365f);
  p = fmaf(p, x2, -0.3333333134651184f);
  const float small = fmaf(x * x2, p, x);
  const float t = ex2_approx((ax > 10.0f ? 10.0f : ax) * 2.8853900817779268f);   // tanh(10) rounds to 1
  const float big = copysignf(fmaf(-2.0f, rcp_newton(t + 1.0f), 1.0f), x);
  return ax >= 0.55f ? big : small;   // NaN: the compare is false, the polynomial propagates it
}

// The 973 parameters of the bound network (elg_actuator_net_bind) in the constant bank: every weight is a warp-uniform operand
// with a compile-time offset, i.e. an immediate constant-bank operand of the FFMA itself -- no shared-memory load, no register.
__constant__ float c_actw[ELG_ACTNET_WORDS];

// weight word i: constant bank (kConst) or the CTA's shared-memory copy
template <bool kConst>
__device__ __forceinline__ float actw(const float* __restrict__ s_w, int i) { return kConst ? c_actw[i] : s_w[i]; }

// one LSTM cell step for one row; torch gate order i, f, g, o (rows u, 8 + u, 16 + u, 24 + u); offsets are words of the blob
template <int kIn, bool kConst, int kWih, int kWhh, int kBih, int kBhh>
__device__ __forceinline__ void lstm_cell(const float* __restrict__ s_w, const float (&x)[kIn], float (&h)[8], float (&c)[8]) {
  float hn[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    float g[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = 8 * q + u;
      float a = actw<kConst>(s_w, kBih + r);
#pragma unroll
      for (int k = 0; k < kIn; ++k) a = fmaf(actw<kConst>(s_w, kWih + r * kIn + k), x[k], a);
      float b = actw<kConst>(s_w, kBhh + r);
#pragma unroll
      for (int k = 0; k < 8; ++k) b = fmaf(actw<kConst>(s_w, kWhh + r * 8 + k), h[k], b);
      g[q] = a + b;
    }
    const float ig = sigmoid_f(g[0]), fg = sigmoid_f(g[1]), gg = tanh_f(g[2]), og = sigmoid_f(g[3]);
    c[u] = fg * c[u] + ig * gg;
    hn[u] = og * tanh_f(c[u]);
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) h[u] = hn[u];
}

// kEarly (shared-memory weights of the BOUND blob, whose contents the caller has declared frozen -- elg_actuator_net_bind): the
// weights are staged before griddepcontrol.wait, i.e. under the tail of the preceding kernel of the stream.  The state loads are
// requested before the barrier that publishes the weights, so the two round trips overlap.
template <bool kConst, bool kEarly>
__global__ void __launch_bounds__(kActThreads)
elg_actuator_kernel(const int64_t rows, const int D, const float action_scale, const float* __restrict__ weights,
                    const float* __restrict__ actions, const float* __restrict__ dof_state, const float* __restrict__ default_dof_pos,
                    float* __restrict__ hidden, float* __restrict__ cell, float* __restrict__ torques) {
  __shared__ __align__(16) float s_w[kConst ? 4 : ELG_ACTNET_WORDS];
  pdl_launch_dependents();
  if (!kEarly) pdl_wait();
  if (!kConst) {
    for (int i = threadIdx.x; i < ELG_ACTNET_WORDS / 4; i += kActThreads)
      reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(weights) + i);
  }
  if (kEarly) pdl_wait();
  const int64_t r = (int64_t)blockIdx.x * kActThreads + threadIdx.x;
  const bool live = r < rows;
  const int64_t rr = live ? r : rows - 1;   // surplus threads of the last CTA shadow the last row (they reach the barrier, store nothing)
  const int j = (int)(rr % D);
  const float2 pv = *reinterpret_cast<const float2*>(dof_state + 2 * rr);
  const float act = actions[rr], q0 = __ldg(default_dof_pos + j);
  float h0[8], c0[8], h1[8], c1[8];
  auto load8 = [&](const float* base, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(base), b = *reinterpret_cast<const float4*>(base + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  };
  auto store8 = [&](float* base, const float (&v)[8]) {
    *reinterpret_cast<float4*>(base) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(base + 4) = make_float4(v[4], v[5], v[6], v[7]);
  };
  const int64_t plane = rows * 8;   // layer stride of the [2, rows, 8] state tensors
  load8(hidden + rr * 8, h0);
  load8(cell + rr * 8, c0);
  load8(hidden + plane + rr * 8, h1);
  load8(cell + plane + rr * 8, c1);
  if (!kConst) __syncthreads();   // the weights
  float x[2];
  x[0] = (act * action_scale + q0 - pv.x) * actw<kConst>(s_w, ELG_ACTNET_IN_SCALE);
  x[1] = pv.y * actw<kConst>(s_w, ELG_ACTNET_IN_SCALE + 1);
  lstm_cell<2, kConst, ELG_ACTNET_W_IH0, ELG_ACTNET_W_HH0, ELG_ACTNET_B_IH0, ELG_ACTNET_B_HH0>(s_w, x, h0, c0);
  lstm_cell<8, kConst, ELG_ACTNET_W_IH1, ELG_ACTNET_W_HH1, ELG_ACTNET_B_IH1, ELG_ACTNET_B_HH1>(s_w, h0, h1, c1);
  float y = actw<kConst>(s_w, ELG_ACTNET_B_LIN);
#pragma unroll
  for (int k = 0; k < 8; ++k) y = fmaf(actw<kConst>(s_w, ELG_ACTNET_W_LIN + k), h1[k], y);
  if (!live) return;
  torques[r] = actw<kConst>(s_w, ELG_ACTNET_OUT_SCALE) * y;
  store8(hidden + r * 8, h0);
  store8(cell + r * 8, c0);
  store8(hidden + plane + r * 8, h1);
  store8(cell + plane + r * 8, c1);
}

// ---------------------------------------------------------------------------------------------------------------
// unit-parallel form: EIGHT lanes per (env, dof) row, lane u owns hidden unit u of both layers.  At 4096 envs the row-per-thread
// kernel above puts 49 152 threads on the chip -- 10 warps per SM, each running ~1000 dependent FMAs: it is latency-bound at 15 %
// occupancy whatever its register count.  Here every lane evaluates the four gates of ONE unit (40 + 64 FMAs, 5 + 5 transcendentals)
// with exactly the expressions of lstm_cell -- the results are bit-identical -- and the 8-vectors every unit needs (previous hidden
// state, layer-0 output) are exchanged through shared memory (one 4-byte store, two 16-byte broadcast loads).  State loads / stores
// are one float per lane: 32 consecutive floats per warp.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kAct8Threads = 256;   // 32 rows per CTA

template <int kIn>
__device__ __forceinline__ void lstm_unit(const float* __restrict__ w_ih, const float* __restrict__ w_hh, const float* __restrict__ b_ih,
                                          const float* __restrict__ b_hh, const int u, const float (&x)[kIn], const float (&h)[8], float& c_u,
                                          float& h_u) {
  float g[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int r = 8 * q + u;
    float a = b_ih[r];
#pragma unroll
    for (int k = 0; k < kIn; ++k) a = fmaf(w_ih[r * kIn + k], x[k], a);
    float b = b_hh[r];
#pragma unroll
    for (int k = 0; k < 8; ++k) b = fmaf(w_hh[r * 8 + k], h[k], b);
    g[q] = a + b;
  }
  const float ig = sigmoid_f(g[0]), fg = sigmoid_f(g[1]), gg = tanh_f(g[2]), og = sigmoid_f(g[3]);
  c_u = fg * c_u + ig * gg;
  h_u = og * tanh_f(c_u);
}

__global__ void __launch_bounds__(kAct8Threads)
elg_actuator_unit_kernel(const int64_t rows, const int D, const float action_scale, const float* __restrict__ weights,
                         const float* __restrict__ actions, const float* __restrict__ dof_state, const float* __restrict__ default_dof_pos,
                         float* __restrict__ hidden, float* __restrict__ cell, float* __restrict__ torques) {
  __shared__ __align__(16) float s_w[ELG_ACTNET_WORDS];
  __shared__ __align__(16) float s_x[3][kAct8Threads];   // exchange buffers: old h0 | new h0 / new h1 | old h1
  pdl_launch_dependents();
  pdl_wait();
  for (int i = threadIdx.x; i < ELG_ACTNET_WORDS / 4; i += kAct8Threads)
    reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(weights) + i);
  const int64_t t = (int64_t)blockIdx.x * kAct8Threads + threadIdx.x;
  const int64_t r = t >> 3;
  const int u = threadIdx.x & 7;
  const bool live = r < rows;
  const int64_t rr = live ? r : rows - 1;
  const int64_t plane = rows * 8;   // layer stride of the [2, rows, 8] state tensors
  float h0 = hidden[rr * 8 + u], c0 = cell[rr * 8 + u], h1 = hidden[plane + rr * 8 + u], c1 = cell[plane + rr * 8 + u];
  const int j = (int)(rr % D);
  const float2 pv = *reinterpret_cast<const float2*>(dof_state + 2 * rr);
  const float act = actions[rr];
  s_x[0][threadIdx.x] = h0;
  s_x[2][threadIdx.x] = h1;
  __syncthreads();   // weights + the old hidden vectors
  float x[2];
  x[0] = (act * action_scale + __ldg(default_dof_pos + j) - pv.x) * s_w[ELG_ACTNET_IN_SCALE];
  x[1] = pv.y * s_w[ELG_ACTNET_IN_SCALE + 1];
  auto vec8 = [&](const float* base, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(base), b = *reinterpret_cast<const float4*>(base + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  };
  const int row0 = threadIdx.x & ~7;
  float H[8], X1[8];
  vec8(&s_x[0][row0], H);
  lstm_unit<2>(s_w + ELG_ACTNET_W_IH0, s_w + ELG_ACTNET_W_HH0, s_w + ELG_ACTNET_B_IH0, s_w + ELG_ACTNET_B_HH0, u, x, H, c0, h0);
  s_x[1][threadIdx.x] = h0;
  __syncwarp();   // a row's eight lanes sit in one warp
  vec8(&s_x[1][row0], X1);
  vec8(&s_x[2][row0], H);
  lstm_unit<8>(s_w + ELG_ACTNET_W_IH1, s_w + ELG_ACTNET_W_HH1, s_w + ELG_ACTNET_B_IH1, s_w + ELG_ACTNET_B_HH1, u, X1, H, c1, h1);
  __syncwarp();   // everybody has read the layer-0 outputs
  s_x[1][threadIdx.x] = h1;
  __syncwarp();
  if (live) {
    hidden[r * 8 + u] = h0;
    cell[r * 8 + u] = c0;
    hidden[plane + r * 8 + u] = h1;
    cell[plane + r * 8 + u] = c1;
    if (u == 0) {
      vec8(&s_x[1][row0], H);
      float y = s_w[ELG_ACTNET_B_LIN];
#pragma unroll
      for (int k = 0; k < 8; ++k) y = fmaf(s_w[ELG_ACTNET_W_LIN + k], H[k], y);
      torques[r] = s_w[ELG_ACTNET_OUT_SCALE] * y;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// unit-split form (measured, not the default): FOUR WARPS per 32 rows, warp p owns hidden units 2p, 2p + 1 of both layers of the rows
// lane = row.  Every weight a warp touches is still warp-uniform (constant bank through the uniform datapath, or a shared-memory
// broadcast), nothing is computed twice, and the chip holds 4x the warps of the row-per-thread form -- which sits at 10 warps per SM
// with a ~3000-instruction dependent chain each.  On B200 at 4096 envs x 12 dofs it takes 11.6 us against 10.5-10.8 us for the
// row-per-thread form with the same constant-bank weights (profiles/README.md r2x): the extra warps buy nothing, the redundant state
// loads, two CTA barriers and four code copies cost a little.  The 8-vectors a unit needs from the other warps (layer-0 output, layer-1 output) cross through shared
// memory ([unit][row]: conflict-free) at two CTA barriers.  Same expressions per unit as lstm_cell: bit-identical results.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kSplitParts = 4, kSplitUnits = 8 / kSplitParts;

template <int kIn, bool kConst, int kU0, int kWih, int kWhh, int kBih, int kBhh>
__device__ __forceinline__ void lstm_units(const float* __restrict__ s_w, const float (&x)[kIn], const float (&h)[8], float (&c)[kSplitUnits],
                                           float (&hn)[kSplitUnits]) {
  constexpr int wih = kWih, whh = kWhh, bih = kBih, bhh = kBhh;
#pragma unroll
  for (int uu = 0; uu < kSplitUnits; ++uu) {
    constexpr int u0 = kU0;
    const int u = u0 + uu;      // compile-time after unrolling: every weight offset is an immediate
    float g[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = 8 * q + u;
      float a = actw<kConst>(s_w, bih + r);
#pragma unroll
      for (int k = 0; k < kIn; ++k) a = fmaf(actw<kConst>(s_w, wih + r * kIn + k), x[k], a);
      float b = actw<kConst>(s_w, bhh + r);
#pragma unroll
      for (int k = 0; k < 8; ++k) b = fmaf(actw<kConst>(s_w, whh + r * 8 + k), h[k], b);
      g[q] = a + b;
    }
    const float ig = sigmoid_f(g[0]), fg = sigmoid_f(g[1]), gg = tanh_f(g[2]), og = sigmoid_f(g[3]);
    c[uu] = fg * c[uu] + ig * gg;
    hn[uu] = og * tanh_f(c[uu]);
  }
}

// everything a warp does once its part is known at compile time (the kernel dispatches on the warp index: four instantiations, each
// warp runs one; the CTA barriers inside are the same two in every instantiation)
template <bool kConst, int kPart>
__device__ __forceinline__ void split_body(const float* __restrict__ s_w, float (*s_h)[8][32], const int lane, const int64_t rows, const int D,
                                           const float action_scale, const float* __restrict__ actions, const float* __restrict__ dof_state,
                                           const float* __restrict__ default_dof_pos, float* __restrict__ hidden, float* __restrict__ cell,
                                           float* __restrict__ torques) {
  constexpr int u0 = kPart * kSplitUnits;
  const int64_t r = (int64_t)blockIdx.x * 32 + lane;
  const bool live = r < rows;
  const int64_t rr = live ? r : rows - 1;      // surplus lanes shadow the last row (loads stay in range, stores are guarded)
  const int64_t plane = rows * 8;              // layer stride of the [2, rows, 8] state tensors
  auto load8 = [&](const float* base, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(base), b = *reinterpret_cast<const float4*>(base + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  };
  float h0[8], h1[8], c0[kSplitUnits], c1[kSplitUnits];
  load8(hidden + rr * 8, h0);
  load8(hidden + plane + rr * 8, h1);
  {
    const float2 a = *reinterpret_cast<const float2*>(cell + rr * 8 + u0), b = *reinterpret_cast<const float2*>(cell + plane + rr * 8 + u0);
    c0[0] = a.x; c0[1] = a.y; c1[0] = b.x; c1[1] = b.y;
  }
  const int j = (int)(rr % D);
  const float2 pv = *reinterpret_cast<const float2*>(dof_state + 2 * rr);
  const float act = actions[rr], q0 = __ldg(default_dof_pos + j);
  if (!kConst) __syncthreads();      // the weights
  float x[2];
  x[0] = (act * action_scale + q0 - pv.x) * actw<kConst>(s_w, ELG_ACTNET_IN_SCALE);
  x[1] = pv.y * actw<kConst>(s_w, ELG_ACTNET_IN_SCALE + 1);
  float hn[kSplitUnits];
  lstm_units<2, kConst, u0, ELG_ACTNET_W_IH0, ELG_ACTNET_W_HH0, ELG_ACTNET_B_IH0, ELG_ACTNET_B_HH0>(s_w, x, h0, c0, hn);
  s_h[0][u0][lane] = hn[0];
  s_h[0][u0 + 1][lane] = hn[1];
  if (live) {
    *reinterpret_cast<float2*>(hidden + r * 8 + u0) = make_float2(hn[0], hn[1]);
    *reinterpret_cast<float2*>(cell + r * 8 + u0) = make_float2(c0[0], c0[1]);
  }
  __syncthreads();
  float x1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) x1[k] = s_h[0][k][lane];
  lstm_units<8, kConst, u0, ELG_ACTNET_W_IH1, ELG_ACTNET_W_HH1, ELG_ACTNET_B_IH1, ELG_ACTNET_B_HH1>(s_w, x1, h1, c1, hn);
  s_h[1][u0][lane] = hn[0];
  s_h[1][u0 + 1][lane] = hn[1];
  if (live) {
    *reinterpret_cast<float2*>(hidden + plane + r * 8 + u0) = make_float2(hn[0], hn[1]);
    *reinterpret_cast<float2*>(cell + plane + r * 8 + u0) = make_float2(c1[0], c1[1]);
  }
  __syncthreads();
  if (kPart == 0 && live) {
    float y = actw<kConst>(s_w, ELG_ACTNET_B_LIN);
#pragma unroll
    for (int k = 0; k < 8; ++k) y = fmaf(actw<kConst>(s_w, ELG_ACTNET_W_LIN + k), s_h[1][k][lane], y);
    torques[r] = actw<kConst>(s_w, ELG_ACTNET_OUT_SCALE) * y;
  }
}

template <bool kConst>
__global__ void __launch_bounds__(32 * kSplitParts)
elg_actuator_split_kernel(const int64_t rows, const int D, const float action_scale, const float* __restrict__ weights,
                          const float* __restrict__ actions, const float* __restrict__ dof_state, const float* __restrict__ default_dof_pos,
                          float* __restrict__ hidden, float* __restrict__ cell, float* __restrict__ torques) {
  __shared__ __align__(16) float s_w[kConst ? 4 : ELG_ACTNET_WORDS];
  __shared__ float s_h[2][8][32];      // new layer-0 / layer-1 outputs, [unit][row]
  pdl_launch_dependents();
  pdl_wait();
  if (!kConst)
    for (int i = threadIdx.x; i < ELG_ACTNET_WORDS / 4; i += 32 * kSplitParts)
      reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(weights) + i);
  const int lane = threadIdx.x & 31;
  static_assert(kSplitParts == 4, "the dispatch below lists four parts");
  switch (threadIdx.x >> 5) {
    case 0: split_body<kConst, 0>(s_w, s_h, lane, rows, D, action_scale, actions, dof_state, default_dof_pos, hidden, cell, torques); break;
    case 1: split_body<kConst, 1>(s_w, s_h, lane, rows, D, action_scale, actions, dof_state, default_dof_pos, hidden, cell, torques); break;
    case 2: split_body<kConst, 2>(s_w, s_h, lane, rows, D, action_scale, actions, dof_state, default_dof_pos, hidden, cell, torques); break;
    default: split_body<kConst, 3>(s_w, s_h, lane, rows, D, action_scale, actions, dof_state, default_dof_pos, hidden, cell, torques); break;
  }
}

// 0 (default): one thread per row, shared-memory weights -- staged before the grid-dependency wait when the blob is the bound one;
// 2: the same, always staged after the wait; 5: the weights as constant-bank operands (needs a bound blob); 3 / 4: the unit-split
// form (constant bank / shared memory); 1: eight lanes per row.  Measured on B200 at 4096 envs x 12 dofs (profiles/README.md r4): with
// libdevice expf / tanhf / __frcp_rn the row forms took 10.6 (constant bank) / 11.0 us and forms 3 / 1 12.4 / 13.5 us; with the
// branch-free activations (3312 -> 2296 instructions per row) and the state loads requested ahead of the weight barrier forms
// 0 / 2 / 5 / 3 take 8.6 / 9.3 / 8.6 / 9.0 us.  A form with the gate sums as packed FFMA2 on gate-interleaved weights (1912
// instructions, bit-identical) took 10.9 us -- FFMA2 occupies the FP32 pipe for both halves, the kernel is bound by that pipe and by
// dependent-issue latency, not by issue slots -- and was removed.
int g_act_mode = 0;
const float* g_act_bound[kMaxDevices] = {};   // the device blob whose contents sit in this device's constant bank

}  // namespace elg

extern "C" {

int elg_actuator_net_words(void) { return ELG_ACTNET_WORDS; }

int elg_set_actuator_tuning(int mode) {
  if (mode < 0 || mode > 5) return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "actuator tuning mode must be 0 ... 5");
  elg::g_act_mode = mode;
  return ELG_OK;
}

int elg_actuator_net_bind(const float* weights, void* stream) {
  if (!weights) {      // unbind: the next calls read the weights through shared memory again
    elg::g_act_bound[elg::device_index()] = nullptr;
    return ELG_OK;
  }
  if ((reinterpret_cast<uintptr_t>(weights) & 15u) != 0) return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "actuator net: weights must be 16-byte aligned");
  if (cudaMemcpyToSymbolAsync(elg::c_actw, weights, sizeof(float) * ELG_ACTNET_WORDS, 0, cudaMemcpyDeviceToDevice, (cudaStream_t)stream) != cudaSuccess)
    return elg::check_launch("elg_actuator_net_bind");
  elg::g_act_bound[elg::device_index()] = weights;
  return ELG_OK;
}

int elg_actuator_net_torques(const ElgDims* dims, const float* weights, float action_scale, const float* actions, const float* dof_state,
                             const float* default_dof_pos, float* hidden, float* cell, float* torques, void* stream) {
  if (!dims) return elg::set_error(ELG_ERR_NULL_POINTER, "dims is NULL");
  if (dims->num_envs < 0 || dims->num_dof < 1 || dims->num_dof > ELG_MAX_DOF) return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "num_envs / num_dof out of range");
  if (!weights || !actions || !dof_state || !default_dof_pos || !hidden || !cell || !torques)
    return elg::set_error(ELG_ERR_NULL_POINTER, "actuator net: a pointer is NULL");
  auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  if (!a16(weights) || !a16(hidden) || !a16(cell)) return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "actuator net: weights / hidden / cell must be 16-byte aligned");
  if ((reinterpret_cast<uintptr_t>(dof_state) & 7u) != 0) return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "actuator net: dof_state must be 8-byte aligned");
  const int64_t rows = (int64_t)dims->num_envs * dims->num_dof;
  if (rows == 0) return ELG_OK;
  const bool unit = elg::g_act_mode == 1;
  const bool split = elg::g_act_mode == 3 || elg::g_act_mode == 4;
  const int threads = unit ? elg::kAct8Threads : split ? 32 * elg::kSplitParts : elg::kActThreads;
  const int64_t nthreads = unit ? rows * 8 : split ? ((rows + 31) / 32) * 32 * elg::kSplitParts : rows;
  const bool bound = elg::g_act_bound[elg::device_index()] == weights;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((nthreads + threads - 1) / threads));
  cfg.blockDim = dim3(threads);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (unit)
    cudaLaunchKernelEx(&cfg, elg::elg_actuator_unit_kernel, rows, (int)dims->num_dof, action_scale, weights, actions, dof_state, default_dof_pos,
                       hidden, cell, torques);
  else if (split && bound && elg::g_act_mode != 4)
    cudaLaunchKernelEx(&cfg, elg::elg_actuator_split_kernel<true>, rows, (int)dims->num_dof, action_scale, weights, actions, dof_state, default_dof_pos,
                       hidden, cell, torques);
  else if (split)
    cudaLaunchKernelEx(&cfg, elg::elg_actuator_split_kernel<false>, rows, (int)dims->num_dof, action_scale, weights, actions, dof_state, default_dof_pos,
                       hidden, cell, torques);
  else if (bound && elg::g_act_mode == 5)
    cudaLaunchKernelEx(&cfg, elg::elg_actuator_kernel<true, false>, rows, (int)dims->num_dof, action_scale, weights, actions, dof_state, default_dof_pos,
                       hidden, cell, torques);
  else if (bound && elg::g_act_mode == 0)
    cudaLaunchKernelEx(&cfg, elg::elg_actuator_kernel<false, true>, rows, (int)dims->num_dof, action_scale, weights, actions, dof_state, default_dof_pos,
                       hidden, cell, torques);
  else
    cudaLaunchKernelEx(&cfg, elg::elg_actuator_kernel<false, false>, rows, (int)dims->num_dof, action_scale, weights, actions, dof_state, default_dof_pos,
                       hidden, cell, torques);
  return elg::check_launch("elg_actuator_net_torques");
}

}  // extern "C"
