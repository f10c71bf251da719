// elg_common.cuh -- device helpers shared by the sm_100a kernels of the per-step hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/elg_b200.h"

namespace elg {

constexpr int kWarp = 32;

// host-side error plumbing shared by the translation units (defined in elg_step.cu)
int set_error(int code, const char* msg);   // records msg for elg_last_error(), returns code
int check_launch(const char* what);         // cudaGetLastError() -> ELG_OK / ELG_ERR_CUDA
int sm_count();                             // SMs of the current device (0 on failure)
// Host-side caches are kept PER DEVICE: a process may drive several GPUs (one env object each), and the dynamic shared-memory
// opt-in of a kernel (cudaFuncSetAttribute) as well as the SM count belong to the device that is current at the call.
constexpr int kMaxDevices = 32;
int device_index();                         // current device, clamped to [0, kMaxDevices)
struct SmemCache { size_t bytes[kMaxDevices]; };
inline size_t& smem_slot(SmemCache& c) { return c.bytes[device_index()]; }

// ---------------------------------------------------------------------------------------------
// Individually rounded fp32 ops.  torch evaluates the reference expression one ATen op at a
// time, so every intermediate is rounded to fp32; where an integer / boolean output depends on
// the value (terrain cell index, contact / termination masks) the kernels must not let the
// compiler contract a*b+c into an FMA.  __fmul_rn / __fadd_rn are never contracted.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float mul_r(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_r(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_r(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_r(float a, float b) { return __fdiv_rn(a, b); }

// torch.norm(x, dim=-1) on CPU for a reduced extent of 2 or 3 accumulates acc = fma(x_k, x_k, acc)
// and takes an IEEE sqrt (measured against torch 2.11 CPU, see tests/test_oracle_pinned.py and
// DESIGN.md "rounding model"); the masks `norm > threshold` are bit-exact only with this chain.
__device__ __forceinline__ float norm2_t(float x, float y) { return __fsqrt_rn(__fmaf_rn(y, y, __fmul_rn(x, x))); }
__device__ __forceinline__ float norm3_t(float x, float y, float z) {
  return __fsqrt_rn(__fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x))));
}

// IEEE sqrt for operands that are exactly zero most of the time (contact forces of bodies in the air): the special-operand
// path of sqrt.rn (zero / denormal / inf / NaN) is an out-of-line routine at the far end of the kernel, and one zero lane
// sends the whole warp there.  Same value for every s >= 0 and NaN.
__device__ __forceinline__ float sqrt_rn_z(float s) {
  const float r = __fsqrt_rn(s > 0.0f ? s : 1.0f);
  return s > 0.0f ? r : s;
}
__device__ __forceinline__ float norm2_tz(float x, float y) { return sqrt_rn_z(__fmaf_rn(y, y, __fmul_rn(x, x))); }
__device__ __forceinline__ float norm3_tz(float x, float y, float z) { return sqrt_rn_z(__fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)))); }
__device__ __forceinline__ float sumsq3_t(float x, float y, float z) { return __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x))); }

// Correctly rounded x / c for a loop-invariant divisor: q0 = x*r, then two residual corrections.
// This is the tail of the IEEE division sequence the hardware path uses (reciprocal, quotient,
// two FMA corrections) with r = RN(1/c) supplied by the caller; valid for finite, normal-range
// operands.  Only used where the host has verified the divisor (see DivConst::exact).
struct DivConst {
  float c;      // divisor
  float r;      // RN(1/c)
  int exact;    // 1: fast path allowed
};
__device__ __forceinline__ float div_const(float x, const DivConst& d) {
  if (d.exact) {
    float q = __fmul_rn(x, d.r);
    float e = __fmaf_rn(-d.c, q, x);
    q = __fmaf_rn(e, d.r, q);
    e = __fmaf_rn(-d.c, q, x);
    return __fmaf_rn(e, d.r, q);
  }
  return __fdiv_rn(x, d.c);
}

// ---------------------------------------------------------------------------------------------
// quaternion helpers (xyzw), following isaacgym.torch_utils as restated in oracle/torch_utils.py
// ---------------------------------------------------------------------------------------------
struct Quat { float x, y, z, w; };
struct Vec3 { float x, y, z; };

// quat_rotate_inverse(q, v) = v*(2w^2-1) - 2w*(q_xyz x v) + 2*q_xyz*(q_xyz . v)
__device__ __forceinline__ Vec3 quat_rotate_inverse(const Quat& q, const Vec3& v) {
  const float s = 2.0f * q.w * q.w - 1.0f;
  const float cx = q.y * v.z - q.z * v.y;
  const float cy = q.z * v.x - q.x * v.z;
  const float cz = q.x * v.y - q.y * v.x;
  const float d = q.x * v.x + q.y * v.y + q.z * v.z;
  Vec3 r;
  r.x = v.x * s - cx * q.w * 2.0f + q.x * d * 2.0f;
  r.y = v.y * s - cy * q.w * 2.0f + q.y * d * 2.0f;
  r.z = v.z * s - cz * q.w * 2.0f + q.z * d * 2.0f;
  return r;
}

// wrap_to_pi (utils/math_utils.py:55-58): python-modulo by fp32(2*pi), then -2*pi where > pi
__device__ __forceinline__ float wrap_to_pi(float a) {
  const float two_pi = 6.283185307179586f;
  float m = fmodf(a, two_pi);
  if (m != 0.0f && m < 0.0f) m += two_pi;
  if (m > 3.141592653589793f) m -= two_pi;
  return m;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter-based: no generator state lives in HBM.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
// the same function for code that a FEW warps run once per launch (reset path): one out-of-line copy with the rounds rolled, so a
// kernel that needs several blocks fetches ~30 instructions once instead of ~120 per use (single-warp straight-line code is bound
// by instruction fetch, not by issue)
static __device__ __noinline__ uint4 philox4x32_10_cold(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll 1
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// In-kernel observation noise: lane `lane` of the warp that owns env draws 128-bit blocks
// Philox4x32-10(counter = (env, lane | block << 5, offset_lo, offset_hi), key = seed) and cuts each into eight 16-bit
// uniforms (sample s = half (s & 1) of word s >> 1).  Height point p = lane + 32 j uses sample j % 8 of block 1 + j / 8;
// head entry k = lane + 32 m uses sample (nj % 8) + m of block 1 + nj / 8 (nj = ceil(H / 32)) when
// nj % 8 + ceil(head / 32) <= 8, else sample m % 8 of block m / 8.  obs += (2 s / 65536 - 1) * noise_scale.
__device__ __forceinline__ uint4 noise_block(uint64_t seed, uint64_t offset, uint32_t env, uint32_t lane, uint32_t chunk) {
  return philox4x32_10(make_uint4(env, lane | (chunk << 5), (uint32_t)offset, (uint32_t)(offset >> 32)),
                       make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
__device__ __forceinline__ uint32_t pick(const uint4& r, int m) { return m == 0 ? r.x : m == 1 ? r.y : m == 2 ? r.z : r.w; }

// 16-bit uniform sample s (0..7) of a 128-bit Philox block, as a float in [0, 65535]
__device__ __forceinline__ float sample16(const uint4& r, int s) {   // s known at compile time after unrolling
  const uint32_t w = (s >> 1) == 0 ? r.x : (s >> 1) == 1 ? r.y : (s >> 1) == 2 ? r.z : r.w;
  return (float)((s & 1) ? (w >> 16) : (w & 0xffffu));
}
__device__ __forceinline__ float sample16_dyn(const uint4& r, int s) {
  const int i = s >> 1;
  const uint32_t w = i == 0 ? r.x : i == 1 ? r.y : i == 2 ? r.z : r.w;
  return (float)((w >> (16 * (s & 1))) & 0xffffu);
}

// 2 u - 1 for the 16-bit sample s of a Philox block, u = s16 / 65536: the same value as fmaf((float)s16, 1 / 32768, -1) without
// the integer -> float conversion (a quarter-rate pipe): one byte permute drops s16 into the mantissa of 2^23, so the float is
// 2^23 + s16 exactly; times 2^-15 it is 256 + s16 / 32768 (24 significant bits, exact), minus 257 exact again.
__device__ __forceinline__ float sym16(const uint4& r, int s) {
  const int i = s >> 1;
  const uint32_t w = i == 0 ? r.x : i == 1 ? r.y : i == 2 ? r.z : r.w;
  const uint32_t bits = __byte_perm(w, 0x4B000000u, (s & 1) ? 0x7632u : 0x7610u);
  return fmaf(__uint_as_float(bits), 1.0f / 32768.0f, -257.0f);
}

// lean fused step for the common quadruped layout (elg_step_fast.cu): returns 1 when it took the launch (*rc = status)
int launch_step_fast(const ElgDims* dims, const ElgStepParams* prm, const ElgStepBuffers* buf, uint32_t phase, int cap_override,
                     int flags, long long* dbg, void* stream, int* rc);

}  // namespace elg
