// elg_step.cu -- fused post-physics step, PD torques and terrain height scan for sm_100a.
//
// One launch of elg_step_kernel replaces the ~100 ATen launches of
// LeggedRobot.post_physics_step (envs/base/legged_robot.py:113-150 in the reference).
//
// Design (v2, see DESIGN.md "step kernel"):
//   * The environments are cut into CHUNKS of whole quads (4 envs), balanced so that every SM gets
//     the same number of quads to within one: 4096 envs -> 148 chunks of 24..28 envs, one
//     1024-thread CTA per SM.  Large N: 256-thread CTAs, 4 per SM, each looping over 16-env chunks.
//   * TMA in, TMA out.  Because the envs of a chunk are consecutive, every per-env array is ONE
//     contiguous global range per chunk: thread 0 issues one cp.async.bulk (global -> shared,
//     mbarrier completion) per input array and, at the end, one cp.async.bulk (shared -> global)
//     per output array -- including the whole [nenv, 235] observation block and the [nenv, 187]
//     height block.  The compute code only touches shared memory.
//   * "state warps" (one per 8 envs) run the O(D + F + B)-per-env work as FLAT (env, item) loops --
//     (env, rotation), (env, dof), (env, foot), (env, body) -- so every phase uses all 32 lanes;
//     per-env reductions go through a per-warp shared scratch.  "row warps" (one env at a time)
//     run the 187-point terrain scan and assemble the observation row; the x/y halves of the
//     terrain-cell chain are evaluated with the packed FMUL2/FADD2/FFMA2 instructions of sm_100a,
//     every op individually IEEE-rounded so the cell index stays bit-exact with torch.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "elg_common.cuh"

namespace elg {

constexpr int kGroup = 8;         // environments per state warp
constexpr int kMaxCap = 32;       // environments per chunk (multiple of 4)
constexpr int kMaxStepThreads = 1024;
constexpr int kFeetQ = 10;        // per-foot partial quantities
constexpr int kDofQ = 8;          // per-dof partial quantities

__device__ __forceinline__ bool term_on(const ElgStepParams& pr, int t) { return (pr.reward_mask >> t) & 1u; }

// ---------------------------------------------------------------------------------------------
// height scan helpers (LeggedRobot._get_heights, legged_robot.py:900-938) -- scalar, exact chain
// ---------------------------------------------------------------------------------------------
struct YawFrame {
  float zz, ww;   // normalised yaw quaternion (0,0,zz,ww)
  float X, Y, Z;  // base position
};

__device__ __forceinline__ YawFrame make_yaw_frame(const float* __restrict__ rs) {
  YawFrame f;
  f.X = rs[0];
  f.Y = rs[1];
  f.Z = rs[2];
  const float qz = rs[5], qw = rs[6];
  // normalize((0,0,qz,qw)): torch's 4-wide norm is the plain sequential sum (no FMA), clamp(min=1e-9)
  float n = __fsqrt_rn(add_r(mul_r(qz, qz), mul_r(qw, qw)));
  n = fmaxf(n, 1e-9f);
  f.zz = div_r(qz, n);
  f.ww = div_r(qw, n);
  return f;
}

// terrain cell (clipped) of grid point (bx, by): every torch op of quat_apply / += / div rounded on its own
__device__ __forceinline__ void terrain_cell(const YawFrame& f, float bx, float by, const ElgStepParams& pr, int& ix, int& iy) {
  const float cx = -mul_r(f.zz, by);        // (q_xyz x b).x with q_xyz = (0,0,zz)
  const float cy = mul_r(f.zz, bx);
  const float tx = mul_r(cx, 2.0f), ty = mul_r(cy, 2.0f);
  float px = add_r(bx, mul_r(f.ww, tx));
  float py = add_r(by, mul_r(f.ww, ty));
  px = add_r(px, -mul_r(f.zz, ty));         // + (q_xyz x t)
  py = add_r(py, mul_r(f.zz, tx));
  px = add_r(add_r(px, f.X), pr.border_size);
  py = add_r(add_r(py, f.Y), pr.border_size);
  px = div_r(px, pr.horizontal_scale);
  py = div_r(py, pr.horizontal_scale);
  // .long() truncates toward zero; NaN/out-of-range behave like the clip below after saturation
  ix = __float2int_rz(px);
  iy = __float2int_rz(py);
  ix = min(max(ix, 0), pr.hf_rows - 2);
  iy = min(max(iy, 0), pr.hf_cols - 2);
}

__device__ __forceinline__ float cell_height(const int16_t* __restrict__ hs, int ix, int iy, const ElgStepParams& pr) {
  const int16_t* p = hs + (size_t)ix * pr.hf_cols + iy;
  int h = min((int)__ldg(p), (int)__ldg(p + pr.hf_cols));
  h = min(h, (int)__ldg(p + 1));
  return mul_r((float)h, pr.vertical_scale);
}

// ---------------------------------------------------------------------------------------------
// packed fp32 pairs (sm_100a): one instruction, two individually IEEE-rounded results
// ---------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float x, float y) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// ---------------------------------------------------------------------------------------------
// observation post-processing shared by head and height entries (legged_robot.py:250-252, :107-108)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float finish_obs(float v, float u, float ns, const ElgStepParams& pr) {
  if (pr.noise_mode != ELG_NOISE_OFF) v = v + (2.0f * u - 1.0f) * ns;
  if (pr.clip_observations > 0.0f) v = fminf(fmaxf(v, -pr.clip_observations), pr.clip_observations);
  return v;
}

// ---------------------------------------------------------------------------------------------
// shared-memory plan of one chunk (word offsets; every region starts 16-byte aligned)
// ---------------------------------------------------------------------------------------------
struct StepPlan {
  // staged per-env arrays, [slot][per-env]
  int root, dof, act, lact, ldv, tq, cf, lrv, vec5, cmd, air, con, lc, ep, gidx, gprev, fpos, fvel;
  int head, hsum, sums, accs, rew, mh, obs;
  // CTA-constant tables
  int q0, plim, vlim, tlim, ns, grid;
  int scratch, scratch_words;   // per state warp
  int part_d, part_f;           // pitches inside the scratch
  int words;
  // launch geometry
  int cap, nchunks, nstate, use_bulk, obs_smem, nterms, head_pitch;
  int8_t term_ids[ELG_NUM_REWARD_TERMS];
};

inline int up4(int w) { return (w + 3) & ~3; }

inline StepPlan make_plan(const ElgDims& d, const ElgStepParams& pr, int cap, int nstate, bool obs_smem) {
  StepPlan L{};
  const int D = d.num_dof, F = d.num_feet, H = d.num_height_points, O = d.num_obs, B = d.num_bodies, C = d.num_commands;
  L.cap = cap;
  L.nstate = nstate;
  L.obs_smem = obs_smem ? 1 : 0;
  int nt = 0;
  for (int t = 0; t < ELG_NUM_REWARD_TERMS; ++t)
    if ((pr.reward_mask >> t) & 1u) L.term_ids[nt++] = (int8_t)t;
  L.nterms = nt;
  const int head = 12 + 3 * D;
  L.head_pitch = head | 1;   // odd pitch: per-env lanes write columns without bank conflicts
  int o = 0;
  auto take = [&](int words) { const int at = o; o += up4(words); return at; };
  L.root = take(cap * 13);
  L.dof = take(cap * 2 * D);
  L.act = take(cap * D);
  L.lact = take(cap * D);
  L.ldv = take(cap * D);
  L.tq = take(cap * D);
  L.cf = take(cap * B * 3);
  L.lrv = take(cap * 6);
  L.vec5 = take(5 * cap * 3);           // base_lin_vel, base_ang_vel, projected_gravity, base_lin_acc, base_ang_acc
  L.cmd = take(cap * C);
  L.air = take(cap * F);
  L.con = take(cap * F);
  L.lc = take((cap * F + 3) / 4);       // bytes
  L.ep = take(cap * 2);                 // int64
  L.gidx = take(cap);
  L.gprev = take(cap * F);
  L.fpos = take(cap * F * 3);
  L.fvel = take(cap * F * 3);
  L.head = take(cap * L.head_pitch);
  L.hsum = take(cap);
  L.sums = take((nt > 0 ? nt : 1) * cap);
  L.accs = take(ELG_NUM_REWARD_TERMS * cap);
  L.rew = take(cap);
  L.mh = take(cap * H);
  L.obs = take(obs_smem ? cap * O : 0);
  L.q0 = take(D);
  L.plim = take(2 * D);
  L.vlim = take(D);
  L.tlim = take(D);
  L.ns = take(O);
  L.grid = take(4 * H);
  L.part_d = kGroup * D + 1;
  L.part_f = kGroup * F + 1;
  int part = kDofQ * L.part_d;
  if (kFeetQ * L.part_f > part) part = kFeetQ * L.part_f;
  if (kGroup * (d.num_penalised + d.num_termination) > part) part = kGroup * (d.num_penalised + d.num_termination);
  L.scratch_words = up4(part + kFeetQ * kGroup + ELG_NUM_REWARD_TERMS * kGroup);
  L.scratch = take(nstate * L.scratch_words);
  L.words = o;
  return L;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA bulk copy global -> this CTA's shared memory; src, dst and bytes must be multiples of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// TMA bulk copy shared -> global (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// cooperative fallback copies (ragged tail chunk or unaligned caller tensors)
__device__ __forceinline__ void coop_copy(void* dst, const void* src, uint32_t bytes, int tid, int nthreads) {
  if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | bytes) & 3u) == 0) {
    const uint32_t* s = static_cast<const uint32_t*>(src);
    uint32_t* d = static_cast<uint32_t*>(dst);
    for (uint32_t i = tid; i < (bytes >> 2); i += nthreads) d[i] = s[i];
  } else {
    const uint8_t* s = static_cast<const uint8_t*>(src);
    uint8_t* d = static_cast<uint8_t*>(dst);
    for (uint32_t i = tid; i < bytes; i += nthreads) d[i] = s[i];
  }
}

// floor(i / d) for 0 <= i < 1024, 1 <= d <= 64
struct FastDiv {
  uint32_t m;
  __device__ explicit FastDiv(int d) : m((65536u + (uint32_t)d - 1u) / (uint32_t)(d > 0 ? d : 1)) {}
  __device__ __forceinline__ int div(int i) const { return (int)(((uint32_t)i * m) >> 16); }
};

// ---------------------------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMaxStepThreads, 1)
elg_step_kernel(const __grid_constant__ ElgDims dm, const __grid_constant__ ElgStepParams pr,
                const __grid_constant__ ElgStepBuffers bf, const __grid_constant__ StepPlan L, const uint32_t phase) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t s_bar;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nthreads = blockDim.x, nwarps = nthreads >> 5;
  const int nstate = L.nstate, nrow = nwarps - nstate;
  const int N = dm.num_envs, D = dm.num_dof, B = dm.num_bodies, F = dm.num_feet, H = dm.num_height_points;
  const int O = dm.num_obs, C = dm.num_commands, P = dm.num_penalised, T = dm.num_termination;
  const int cap = L.cap;
  const int head = 12 + 3 * D, headp = L.head_pitch;
  const bool do_derive = phase & ELG_PHASE_DERIVE, do_term = phase & ELG_PHASE_TERMINATION;
  const bool do_reward = phase & ELG_PHASE_REWARD, do_obs = phase & ELG_PHASE_OBS, do_hist = phase & ELG_PHASE_HISTORY;
  const bool need_hsum = do_reward && term_on(pr, ELG_REW_BASE_HEIGHT) && H > 0;   // CTA-uniform
  const bool gait = bf.gait_idx != nullptr && bf.gait_prev_foot_z != nullptr;
  const bool lim_terms = term_on(pr, ELG_REW_DOF_POS_LIMITS) | term_on(pr, ELG_REW_DOF_VEL_LIMITS) | term_on(pr, ELG_REW_TORQUE_LIMITS);
  const bool heights_live = H > 0 && do_derive && !pr.terrain_is_plane;
  const bool shared_grid = pr.height_points_env_stride == 0;

  float* s_root = smem + L.root;   float* s_dof = smem + L.dof;    float* s_act = smem + L.act;
  float* s_lact = smem + L.lact;   float* s_ldv = smem + L.ldv;    float* s_tq = smem + L.tq;
  float* s_cf = smem + L.cf;       float* s_lrv = smem + L.lrv;    float* s_vec5 = smem + L.vec5;
  float* s_cmd = smem + L.cmd;     float* s_air = smem + L.air;    float* s_con = smem + L.con;
  uint8_t* s_lc = reinterpret_cast<uint8_t*>(smem + L.lc);
  int64_t* s_ep = reinterpret_cast<int64_t*>(smem + L.ep);
  float* s_gidx = smem + L.gidx;   float* s_gprev = smem + L.gprev;
  float* s_fpos = smem + L.fpos;   float* s_fvel = smem + L.fvel;
  float* s_head = smem + L.head;   float* s_hsum = smem + L.hsum;
  float* s_sums = smem + L.sums;   float* s_accs = smem + L.accs;  float* s_rew = smem + L.rew;
  float* s_mh = smem + L.mh;       float* s_obs = smem + L.obs;
  float* s_q0 = smem + L.q0;       float* s_plim = smem + L.plim;  float* s_vlim = smem + L.vlim;
  float* s_tlim = smem + L.tlim;   float* s_ns = smem + L.ns;
  float4* s_grid = reinterpret_cast<float4*>(smem + L.grid);

  // ------------------------------- CTA-constant tables -------------------------------
  for (int j = tid; j < D; j += nthreads) {
    s_q0[j] = __ldg(bf.default_dof_pos + j);
    if (lim_terms) {
      s_plim[2 * j] = __ldg(bf.dof_pos_limits + 2 * j);
      s_plim[2 * j + 1] = __ldg(bf.dof_pos_limits + 2 * j + 1);
      s_vlim[j] = __ldg(bf.dof_vel_limits + j) * pr.soft_dof_vel_limit;
      s_tlim[j] = __ldg(bf.torque_limits + j) * pr.soft_torque_limit;
    }
  }
  for (int k = tid; k < O; k += nthreads) s_ns[k] = (bf.noise_scale_vec && pr.noise_mode != ELG_NOISE_OFF) ? __ldg(bf.noise_scale_vec + k) : 0.0f;
  if (H > 0 && shared_grid && bf.height_points)
    for (int p = tid; p < H; p += nthreads) {
      const float bx = __ldg(bf.height_points + 3 * p), by = __ldg(bf.height_points + 3 * p + 1);
      s_grid[p] = make_float4(bx, by, by, bx);
    }
  if (tid == 0) mbar_init(&s_bar, 1);
  __syncthreads();

  const long long Q = ((long long)N + 3) >> 2;
  uint32_t bar_parity = 0;
  bool stores_pending = false;

  for (int chunk = blockIdx.x; chunk < L.nchunks; chunk += gridDim.x) {
    const int q_lo = (int)((long long)chunk * Q / L.nchunks), q_hi = (int)((long long)(chunk + 1) * Q / L.nchunks);
    const int env0 = q_lo * 4;
    const int nenv = min(N, q_hi * 4) - env0;
    if (nenv <= 0) continue;
    const bool bulk = L.use_bulk && (nenv & 3) == 0;
    const size_t e0 = (size_t)env0;

    if (stores_pending) {   // the previous chunk's TMA stores must have read shared memory before it is overwritten
      if (tid == 0) bulk_wait_read_all();
      stores_pending = false;
    }
    __syncthreads();

    // ------------------------------- stage inputs -------------------------------
    auto for_each_input = [&](auto&& f) {
      f(s_root, bf.root_states + e0 * 13, 4u * nenv * 13);
      f(s_dof, bf.dof_state + e0 * 2 * D, 4u * nenv * 2 * D);
      f(s_act, bf.actions + e0 * D, 4u * nenv * D);
      f(s_lact, bf.last_actions + e0 * D, 4u * nenv * D);
      f(s_ldv, bf.last_dof_vel + e0 * D, 4u * nenv * D);
      f(s_tq, bf.torques + e0 * D, 4u * nenv * D);
      f(s_cf, bf.contact_forces + e0 * B * 3, 4u * nenv * B * 3);
      f(s_lrv, bf.last_root_vel + e0 * 6, 4u * nenv * 6);
      f(s_vec5 + 3 * cap * 3, bf.base_lin_acc + e0 * 3, 4u * nenv * 3);
      f(s_vec5 + 4 * cap * 3, bf.base_ang_acc + e0 * 3, 4u * nenv * 3);
      if (!do_derive) {
        f(s_vec5 + 0 * cap * 3, bf.base_lin_vel + e0 * 3, 4u * nenv * 3);
        f(s_vec5 + 1 * cap * 3, bf.base_ang_vel + e0 * 3, 4u * nenv * 3);
        f(s_vec5 + 2 * cap * 3, bf.projected_gravity + e0 * 3, 4u * nenv * 3);
        if (H > 0 && (do_obs || need_hsum)) f(s_mh, bf.measured_heights + e0 * H, 4u * nenv * H);
      }
      f(s_cmd, bf.commands + e0 * C, 4u * nenv * C);
      if (F > 0) {
        f(s_air, bf.feet_air_time + e0 * F, 4u * nenv * F);
        f(s_con, bf.feet_contact_time + e0 * F, 4u * nenv * F);
        f(s_lc, bf.last_contacts + e0 * F, (uint32_t)(nenv * F));
      }
      f(s_ep, bf.episode_length_buf + e0, 8u * nenv);
      if (gait) {
        f(s_gidx, bf.gait_idx + e0, 4u * nenv);
        f(s_gprev, bf.gait_prev_foot_z + e0 * F, 4u * nenv * F);
      }
      if (do_reward)
        for (int ti = 0; ti < L.nterms; ++ti) f(s_sums + ti * cap, bf.episode_sums + (size_t)L.term_ids[ti] * N + e0, 4u * nenv);
    };
    if (bulk) {
      if (tid == 0) {
        uint32_t total = 0;
        for_each_input([&](void*, const void*, uint32_t bytes) { total += bytes; });
        mbar_expect_tx(&s_bar, total);
        for_each_input([&](void* s, const void* g, uint32_t bytes) { bulk_g2s(s, g, bytes, &s_bar); });
      }
    } else {
      for_each_input([&](void* s, const void* g, uint32_t bytes) { coop_copy(s, g, bytes, tid, nthreads); });
    }
    // feet rows of rigid_body_state are a strided gather (52-byte rows, 6 useful floats per row): plain loads
    for (int r = tid; r < nenv * F * 6; r += nthreads) {
      const int c = r % 6, ef = r / 6;
      const int e = ef / F, f = ef - e * F;
      const float v = __ldg(bf.rigid_body_state + ((size_t)(env0 + e) * B + dm.feet_idx[f]) * 13 + (c < 3 ? c : c + 4));
      (c < 3 ? s_fpos : s_fvel)[ef * 3 + (c < 3 ? c : c - 3)] = v;
    }
    if (bulk) {
      mbar_wait(&s_bar, bar_parity);
      bar_parity ^= 1u;
    }
    __syncthreads();

    uint4 blk0 = make_uint4(0, 0, 0, 0);   // Philox block 0 of this warp's (single) env, reused by the head pass
    int blk0_env = -1;

    if (warp >= nstate) {
      // =========================== row warps: terrain scan + height part of the observation ===========================
      if (H > 0 && (do_derive || do_obs || need_hsum)) {
        const int16_t* __restrict__ hs = bf.height_samples;
        const int m_lo = head >> 5, m_hi = (head + H - 1) >> 5;
        const float r_h = __frcp_rn(pr.horizontal_scale);
        const f32x2 rr = pack2(r_h, r_h), nc = pack2(-pr.horizontal_scale, -pr.horizontal_scale);
        const f32x2 bord = pack2(pr.border_size, pr.border_size);
        const int cols = pr.hf_cols, rmax = pr.hf_rows - 2, cmax = pr.hf_cols - 2;
        const bool philox = do_obs && pr.noise_mode == ELG_NOISE_PHILOX;
        for (int slot = warp - nstate; slot < nenv; slot += nrow) {
          const int env = env0 + slot;
          const float* rs = s_root + slot * 13;
          const float rootz = rs[2];
          const float zc = sub_r(rootz, 0.5f);
          float zz = 0.0f, ww = 1.0f;
          if (heights_live) {
            float n = __fsqrt_rn(add_r(mul_r(rs[5], rs[5]), mul_r(rs[6], rs[6])));
            n = fmaxf(n, 1e-9f);
            zz = div_r(rs[5], n);
            ww = div_r(rs[6], n);
          }
          // t = 2 (q x b) = (-2 zz by, 2 zz bx); 2*RN(x) == RN(2x), so the doubling is folded into the multiplier
          const f32x2 c_t = pack2(-mul_r(zz, 2.0f), mul_r(zz, 2.0f));    // times (by, bx) -> (tx, ty)
          const f32x2 c_ts = pack2(mul_r(zz, 2.0f), -mul_r(zz, 2.0f));   // times (bx, by) -> (ty, tx)
          const f32x2 c_w = pack2(ww, ww), c_u = pack2(-zz, zz), xy = pack2(rs[0], rs[1]);
          const float* hp_env = shared_grid ? nullptr : bf.height_points + (size_t)env * pr.height_points_env_stride;
          float hsum = 0.0f;
          uint4 rnd = make_uint4(0, 0, 0, 0);
          for (int m = m_lo; m <= m_hi; ++m) {
            const int k = lane + 32 * m;
            const int p = k - head;
            if (philox && ((m & 3) == 0 || m == m_lo)) {   // warp-uniform: every lane owns word m&3 of its block
              rnd = noise_block(pr.noise_seed, pr.noise_offset, env, lane, m >> 2);
              if ((m >> 2) == 0) { blk0 = rnd; blk0_env = env; }
            }
            const bool valid = p >= 0 && p < H;
            float h = 0.0f;
            if (heights_live) {
              if (valid) {
                f32x2 b, bs;
                if (shared_grid) {
                  const float4 g = s_grid[p];
                  b = pack2(g.x, g.y);
                  bs = pack2(g.z, g.w);
                } else {
                  const float bx = __ldg(hp_env + 3 * p), by = __ldg(hp_env + 3 * p + 1);
                  b = pack2(bx, by);
                  bs = pack2(by, bx);
                }
                const f32x2 t = mul2(c_t, bs);          // (tx, ty)
                const f32x2 ts = mul2(c_ts, b);         // (ty, tx)
                f32x2 pt = add2(b, mul2(c_w, t));       // b + w t
                pt = add2(pt, mul2(c_u, ts));           // + q_xyz x t = (-zz ty, zz tx)
                pt = add2(add2(pt, xy), bord);          // + base xy, + border_size
                // correctly rounded pt / horizontal_scale: q0 = x r, two FMA residual corrections (Markstein)
                f32x2 q = mul2(pt, rr);
                f32x2 e = fma2(nc, q, pt);
                q = fma2(e, rr, q);
                e = fma2(nc, q, pt);
                q = fma2(e, rr, q);
                float qx, qy;
                unpack2(q, qx, qy);
                int ix = __float2int_rz(qx), iy = __float2int_rz(qy);
                ix = min(max(ix, 0), rmax);
                iy = min(max(iy, 0), cmax);
                const int16_t* cell = hs + (size_t)ix * cols + iy;
                const int a0 = __ldg(cell), a1 = __ldg(cell + cols), a2 = __ldg(cell + 1);
                h = mul_r((float)min(min(a0, a1), a2), pr.vertical_scale);
              }
            } else if (!do_derive && valid) {
              h = s_mh[slot * H + p];
            }
            if (valid) {
              if (do_derive) s_mh[slot * H + p] = h;
              hsum += sub_r(rootz, h);
              if (do_obs) {
                float v = mul_r(fminf(fmaxf(sub_r(zc, h), -1.0f), 1.0f), pr.obs_scale_height);
                float u = 0.0f;
                if (pr.noise_mode == ELG_NOISE_TENSOR) u = __ldg(bf.noise_u + (size_t)env * O + k);
                if (pr.noise_mode == ELG_NOISE_PHILOX) u = u01(pick(rnd, m & 3));
                v = finish_obs(v, u, s_ns[k], pr);
                if (L.obs_smem) s_obs[slot * O + k] = v;
                else bf.obs_buf[(size_t)env * O + k] = v;
              }
            }
          }
          if (need_hsum) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) hsum += __shfl_xor_sync(0xffffffffu, hsum, o);
            if (lane == 0) s_hsum[slot] = hsum;
          }
        }
      }
      if (need_hsum) __syncthreads();   // (A) heights -> state warps
    } else {
      // =========================== state warps: flat (env, item) loops over one group of 8 envs ===========================
      const int s0 = warp * kGroup;
      const int ne = min(kGroup, nenv - s0);   // <= 0: this warp only takes part in the barriers
      float* scratch = smem + L.scratch + warp * L.scratch_words;
      float* part = scratch;
      const int part_words = L.scratch_words - kFeetQ * kGroup - ELG_NUM_REWARD_TERMS * kGroup;
      float* fred = scratch + part_words;                       // [kFeetQ][kGroup]
      float* rterm = fred + kFeetQ * kGroup;                    // [nterms][kGroup]
      const FastDiv fdD(D), fdF(F > 0 ? F : 1), fdPT(P + T > 0 ? P + T : 1);
#define ACC(t, slot) s_accs[(t) * cap + (slot)]

      if (ne > 0) {
        // ---- episode counter (legged_robot.py:122)
        if (do_derive && lane < ne) s_ep[s0 + lane] += 1;

        // ---- phase R: (env, rotation) -- base-frame velocities, gravity, acceleration EMAs (:128-134)
        if (do_derive) {
          for (int i = lane; i < 5 * kGroup; i += 32) {
            const int e = i & (kGroup - 1), r = i >> 3;
            if (e < ne) {
              const int slot = s0 + e;
              const float* rs = s_root + slot * 13;
              const Quat q = {rs[3], rs[4], rs[5], rs[6]};
              Vec3 v;
              if (r == 2) {
                v = Vec3{pr.gravity_vec[0], pr.gravity_vec[1], pr.gravity_vec[2]};
              } else {
                const int k = (r == 1 || r == 4) ? 10 : 7;
                v = Vec3{rs[k], rs[k + 1], rs[k + 2]};
                if (r >= 3) {
                  const float* lrv = s_lrv + slot * 6 + (r - 3) * 3;
                  v.x -= lrv[0]; v.y -= lrv[1]; v.z -= lrv[2];
                }
              }
              Vec3 o = quat_rotate_inverse(q, v);
              float* dst = s_vec5 + (r * cap + slot) * 3;
              if (r >= 3) {
                const float ema = pr.acc_ema, w1 = pr.acc_ema_c;
                o.x = dst[0] * ema + (w1 * o.x) / pr.dt;
                o.y = dst[1] * ema + (w1 * o.y) / pr.dt;
                o.z = dst[2] * ema + (w1 * o.z) / pr.dt;
              }
              dst[0] = o.x; dst[1] = o.y; dst[2] = o.z;
            }
          }
        }

        // ---- phase D: (env, dof) -- per-dof reward partials, observation entries, history (:84-114, :237-244, :148-149)
        const int nD = ne * D;
        for (int i = lane; i < nD; i += 32) {
          const int e = fdD.div(i), j = i - e * D;
          const int fi = s0 * D + i;
          const float2 pv = *reinterpret_cast<const float2*>(s_dof + 2 * fi);
          const float pos = pv.x, vel = pv.y;
          const float a = s_act[fi], q0 = s_q0[j];
          if (do_reward) {
            const float la = s_lact[fi], lv = s_ldv[fi], tq = s_tq[fi];
            const float da = la - a;
            part[0 * L.part_d + i] = da * da;
            const float dv = (lv - vel) / pr.dt;
            part[1 * L.part_d + i] = dv * dv;
            part[2 * L.part_d + i] = vel * vel;
            part[3 * L.part_d + i] = tq * tq;
            part[4 * L.part_d + i] = fabsf(pos - q0);
            if (lim_terms) {
              part[5 * L.part_d + i] = -fminf(pos - s_plim[2 * j], 0.0f) + fmaxf(pos - s_plim[2 * j + 1], 0.0f);
              part[6 * L.part_d + i] = fminf(fmaxf(fabsf(vel) - s_vlim[j], 0.0f), 1.0f);
              part[7 * L.part_d + i] = fmaxf(fabsf(tq) - s_tlim[j], 0.0f);
            }
          }
          if (do_obs) {
            float* hrow = s_head + (s0 + e) * headp;
            hrow[12 + j] = (pos - q0) * pr.obs_scale_dof_pos;
            hrow[12 + D + j] = vel * pr.obs_scale_dof_vel;
            hrow[12 + 2 * D + j] = a;
          }
          if (do_hist) {
            s_lact[fi] = a;
            s_ldv[fi] = vel;
          }
        }
        __syncwarp();
        if (do_reward) {
          // registry ids of the 8 per-dof quantities, one byte each
          const unsigned long long qterm = (unsigned long long)ELG_REW_ACTION_RATE | ((unsigned long long)ELG_REW_DOF_ACC << 8) |
                                           ((unsigned long long)ELG_REW_DOF_VEL << 16) | ((unsigned long long)ELG_REW_TORQUES << 24) |
                                           ((unsigned long long)ELG_REW_STAND_STILL << 32) | ((unsigned long long)ELG_REW_DOF_POS_LIMITS << 40) |
                                           ((unsigned long long)ELG_REW_DOF_VEL_LIMITS << 48) | ((unsigned long long)ELG_REW_TORQUE_LIMITS << 56);
          const int nq = lim_terms ? kDofQ : 5;
          for (int i = lane; i < nq * kGroup; i += 32) {
            const int e = i & (kGroup - 1), qn = i >> 3;
            if (e < ne) {
              const float* src = part + qn * L.part_d + e * D;
              float sum = 0.0f;
              for (int j = 0; j < D; ++j) sum += src[j];
              ACC((int)((qterm >> (8 * qn)) & 0xffu), s0 + e) = sum;
            }
          }
          if (!lim_terms && lane < ne) {
            ACC(ELG_REW_DOF_POS_LIMITS, s0 + lane) = 0.0f;
            ACC(ELG_REW_DOF_VEL_LIMITS, s0 + lane) = 0.0f;
            ACC(ELG_REW_TORQUE_LIMITS, s0 + lane) = 0.0f;
          }
          __syncwarp();

          // ---- phase F: (env, foot) (legged_robot_rew_mixin.py:58-81, :121-212; gait_scheduler.py:74-81)
          // Terms that sort before feet_air_time read the OLD timers, terms after it the updated ones
          // and the rebound last_contacts.
          const bool air_on = term_on(pr, ELG_REW_FEET_AIR_TIME);
          const bool gs_on = term_on(pr, ELG_REW_GAIT_SCHEDULER) && gait;
          const int nF = ne * F;
          for (int i = lane; i < nF; i += 32) {
            const int e = fdF.div(i), f = i - e * F;
            const int slot = s0 + e, fi = s0 * F + i;
            const float* cf = s_cf + (slot * B + dm.feet_idx[f]) * 3;
            const float fxx = cf[0], fyy = cf[1], fz = cf[2];
            const float pz = s_fpos[fi * 3 + 2];
            const float vx = s_fvel[fi * 3], vy = s_fvel[fi * 3 + 1], vz = s_fvel[fi * 3 + 2];
            float air = s_air[fi], con = s_con[fi];
            const bool last_c = s_lc[fi] != 0;
            const bool contact = fz > 1.0f;
            const bool touching = con > 1e-3f;   // base_foot_height: nanmean over touching feet (old timers)
            part[0 * L.part_f + i] = touching ? pz : 0.0f;
            part[1 * L.part_f + i] = touching ? 1.0f : 0.0f;
            bool lc_after = last_c;
            float r_air = 0.0f;
            if (air_on) {
              const bool filt = contact | last_c;
              const bool first = (air > 0.0f) && filt;
              air += pr.dt;
              con += pr.dt;
              r_air = (air - 0.5f) * (first ? 1.0f : 0.0f);
              air *= filt ? 0.0f : 1.0f;
              con *= filt ? 1.0f : 0.0f;
              s_air[fi] = air;
              s_con[fi] = con;
              s_lc[fi] = contact ? 1 : 0;
              lc_after = contact;
            }
            const bool filt2 = contact | lc_after;
            part[2 * L.part_f + i] = r_air;
            part[3 * L.part_f + i] = fmaxf(norm3_t(fxx, fyy, fz) - pr.max_contact_force, 0.0f);
            const float vn = norm2_t(vx, vy);
            part[4 * L.part_f + i] = (filt2 ? 1.0f : 0.0f) * (vn * vn);
            const bool stumble = norm2_t(fxx, fyy) > mul_r(5.0f, fabsf(fz));
            part[5 * L.part_f + i] = (stumble ? 1.0f : 0.0f) * vz;
            part[6 * L.part_f + i] = (filt2 ? 0.0f : 1.0f) * (air - 0.5f);
            part[7 * L.part_f + i] = stumble ? 1.0f : 0.0f;
            part[8 * L.part_f + i] = fz < 1.0f ? 0.0f : 1.0f;   // number of feet that are NOT up
            float r_gs = 0.0f;
            if (gs_on) {
              float ph = s_gidx[slot] + pr.gait_foot_phases[f];
              ph = ph - floorf(ph);                                 // torch.remainder(x, 1.0)
              const float target = ph < 0.5f ? pr.gait_swing_height * sinf(6.283185307179586f * ph) : 0.0f;
              const float dz = target - s_gprev[fi];
              r_gs = dz * dz;
            }
            part[9 * L.part_f + i] = r_gs;
            if (gait) s_gprev[fi] = pz;   // GaitScheduler.step keeps this step's feet
          }
          __syncwarp();
          for (int i = lane; i < kFeetQ * kGroup; i += 32) {
            const int e = i & (kGroup - 1), qn = i >> 3;
            if (e < ne) {
              const float* src = part + qn * L.part_f + e * F;
              float sum = 0.0f;
              for (int f = 0; f < F; ++f) sum += src[f];
              fred[qn * kGroup + e] = sum;
            }
          }
          __syncwarp();
        }

        // ---- phase P: (env, body) -- collision count and termination contacts (:117-119, legged_robot.py:155-160)
        const int PT = P + T;
        if (do_term || do_reward) {
          for (int i = lane; i < ne * PT; i += 32) {
            const int e = fdPT.div(i), p = i - e * PT;
            const int body = p < P ? dm.penalised_idx[p] : dm.termination_idx[p - P];
            const float* f = s_cf + ((s0 + e) * B + body) * 3;
            part[i] = norm3_t(f[0], f[1], f[2]) > (p < P ? 0.1f : 1.0f) ? 1.0f : 0.0f;
          }
          __syncwarp();
        }

        // ---- phase E: one lane per env -- commands, termination, scalar reward terms, head, root-velocity history
        if (lane < ne) {
          const int slot = s0 + lane, env = env0 + slot;
          const float* rs = s_root + slot * 13;
          const float* blv = s_vec5 + (0 * cap + slot) * 3;
          const float* bav = s_vec5 + (1 * cap + slot) * 3;
          const float* pg = s_vec5 + (2 * cap + slot) * 3;
          float* cmd = s_cmd + slot * C;
          if (do_derive && pr.heading_command) {   // (legged_robot.py:394-398); forward = quat_apply(q, (1,0,0))
            const Quat q = {rs[3], rs[4], rs[5], rs[6]};
            const float fx = 1.0f + (q.y * (-2.0f * q.y) - q.z * (2.0f * q.z));
            const float fy = q.w * (2.0f * q.z) + (q.z * 0.0f - q.x * (-2.0f * q.y));
            const float heading = atan2f(fy, fx);
            cmd[2] = fminf(fmaxf(0.5f * wrap_to_pi(cmd[3] - heading), -1.0f), 1.0f);
          }
          const float cmd0 = cmd[0], cmd1 = cmd[1], cmd2 = cmd[2], cmd3 = C > 3 ? cmd[3] : 0.0f;
          bool reset = false, time_out = false;
          if (do_term) {
            bool contact_term = false;
            for (int t = 0; t < T; ++t) contact_term |= part[lane * PT + P + t] != 0.0f;
            time_out = s_ep[slot] > pr.max_episode_length;
            reset = contact_term | time_out;
            bf.reset_buf[env] = reset ? 1 : 0;
            bf.time_out_buf[env] = time_out ? 1 : 0;
          } else if (do_reward) {
            reset = bf.reset_buf[env] != 0;
            time_out = bf.time_out_buf[env] != 0;
          }
          if (do_reward) {
            const float rootz = rs[2];
            const float cmd_xy = norm2_t(cmd0, cmd1);
            ACC(ELG_REW_STAND_STILL, slot) *= (cmd_xy < pr.stand_still_threshold ? 1.0f : 0.0f);
            ACC(ELG_REW_LIN_VEL_Z, slot) = blv[2] * blv[2];
            ACC(ELG_REW_ANG_VEL_XY, slot) = bav[0] * bav[0] + bav[1] * bav[1];
            ACC(ELG_REW_ORIENTATION, slot) = pg[0] * pg[0] + pg[1] * pg[1];
            {
              const float ex = cmd0 - blv[0], ey = cmd1 - blv[1], ez = cmd2 - bav[2];
              ACC(ELG_REW_TRACKING_LIN_VEL, slot) = expf(-(ex * ex + ey * ey) / pr.tracking_sigma);
              ACC(ELG_REW_TRACKING_ANG_VEL, slot) = expf(-(ez * ez) / pr.tracking_sigma);
            }
            ACC(ELG_REW_TERMINATION, slot) = (reset && !time_out) ? 1.0f : 0.0f;
            {
              float n = 0.0f;
              for (int p = 0; p < P; ++p) n += part[lane * PT + p];
              ACC(ELG_REW_COLLISION, slot) = n;
            }
            const float* fr = fred + lane;   // fred[q * kGroup + lane]
            {
              const float cnt = fr[1 * kGroup];
              const float ground = cnt > 0.0f ? fr[0 * kGroup] / cnt : rootz - pr.base_height_target;
              const float rel = rootz - ground - pr.base_height_target;
              ACC(ELG_REW_BASE_FOOT_HEIGHT, slot) = rel * rel;
            }
            ACC(ELG_REW_FEET_AIR_TIME, slot) = fr[2 * kGroup] * (cmd_xy > 0.1f ? 1.0f : 0.0f);
            ACC(ELG_REW_FEET_CONTACT_FORCES, slot) = fr[3 * kGroup];
            ACC(ELG_REW_FEET_SLIP, slot) = fr[4 * kGroup];
            ACC(ELG_REW_FEET_STUMBLE_LIFTUP, slot) = fr[5 * kGroup];
            ACC(ELG_REW_JUMP_AIR, slot) = fmaxf(fr[6 * kGroup] - (float)F / 2.0f, 0.0f);
            ACC(ELG_REW_FEET_STUMBLE, slot) = fr[7 * kGroup] > 0.0f ? 1.0f : 0.0f;
            ACC(ELG_REW_FOUR_FOOTUP, slot) = fr[8 * kGroup] == 0.0f ? 0.1f : 0.0f;
            ACC(ELG_REW_GAIT_SCHEDULER, slot) = fr[9 * kGroup];
            {
              // gait_2_step (legged_robot_rew_mixin.py:170-206): FL/RR and FR/RL in phase, the rest anti-phase
              const float* ar = s_air + slot * F;
              const float* cn = s_con + slot * F;
              const float a0 = F > 0 ? ar[0] : 0.0f, a1 = F > 1 ? ar[1] : 0.0f, a2 = F > 2 ? ar[2] : 0.0f, a3 = F > 3 ? ar[3] : 0.0f;
              const float c0 = F > 0 ? cn[0] : 0.0f, c1 = F > 1 ? cn[1] : 0.0f, c2 = F > 2 ? cn[2] : 0.0f, c3 = F > 3 ? cn[3] : 0.0f;
              auto sq4 = [](float a, float b) { const float d = a - b; return fminf(d * d, 4.0f); };
              const float s = ((sq4(a0, a3) + sq4(c0, c3)) + (sq4(a1, a2) + sq4(c1, c2))) / 2.0f;
              const float a = ((sq4(a0, c1) + sq4(c0, a1)) + (sq4(a0, c2) + sq4(c0, a2)) + (sq4(a3, c2) + sq4(c3, a2)) +
                               (sq4(a3, c1) + sq4(c3, a1))) / 4.0f;
              const float yawish = pr.heading_command ? cmd3 : cmd2;
              const bool moving = (cmd_xy > pr.speed_min) | (fabsf(yawish) >= pr.speed_min / 2.0f);
              ACC(ELG_REW_GAIT_2_STEP, slot) = (s + a) * (moving ? 1.0f : 0.0f);
            }
            ACC(ELG_REW_BASE_HEIGHT, slot) = 0.0f;
            if (gait) {   // GaitScheduler.step (gait_scheduler.py:63-72) runs after the env step
              const float g = s_gidx[slot] + pr.gait_increment;
              s_gidx[slot] = g - floorf(g);
            }
          }
          if (do_obs) {
            float* hrow = s_head + slot * headp;
            hrow[0] = blv[0] * pr.obs_scale_lin_vel; hrow[1] = blv[1] * pr.obs_scale_lin_vel; hrow[2] = blv[2] * pr.obs_scale_lin_vel;
            hrow[3] = bav[0] * pr.obs_scale_ang_vel; hrow[4] = bav[1] * pr.obs_scale_ang_vel; hrow[5] = bav[2] * pr.obs_scale_ang_vel;
            hrow[6] = pg[0]; hrow[7] = pg[1]; hrow[8] = pg[2];
            hrow[9] = cmd0 * pr.commands_scale[0]; hrow[10] = cmd1 * pr.commands_scale[1]; hrow[11] = cmd2 * pr.commands_scale[2];
          }
          if (do_hist) {
            float* lrv = s_lrv + slot * 6;
#pragma unroll
            for (int k = 0; k < 6; ++k) lrv[k] = rs[7 + k];
          }
        }
        __syncwarp();
      }

      if (need_hsum) __syncthreads();   // (A)
      if (ne > 0 && do_reward) {
        if (need_hsum && lane < ne) {
          const float d = s_hsum[s0 + lane] / (float)H - pr.base_height_target;
          ACC(ELG_REW_BASE_HEIGHT, s0 + lane) = d * d;
        }
        __syncwarp();
        // ---- phase W: (term, env) scaled terms + episode sums, then the ordered fp32 sum (legged_robot.py:220-232)
        for (int i = lane; i < L.nterms * kGroup; i += 32) {
          const int e = i & (kGroup - 1), ti = i >> 3;
          if (e < ne) {
            const int t = L.term_ids[ti];
            const float r = ACC(t, s0 + e) * pr.reward_scales[t];
            rterm[ti * kGroup + e] = r;
            s_sums[ti * cap + s0 + e] += r;
          }
        }
        __syncwarp();
        if (lane < ne) {
          const int slot = s0 + lane;
          float total = 0.0f, r_term = 0.0f;
          for (int ti = 0; ti < L.nterms; ++ti) {
            const float r = rterm[ti * kGroup + lane];
            if (L.term_ids[ti] == ELG_REW_TERMINATION) r_term = r;   // added after the clip
            else total += r;
          }
          if (bf.extra_reward) total += bf.extra_reward[env0 + slot];
          if (pr.only_positive_rewards) total = fmaxf(total, 0.0f);
          if (term_on(pr, ELG_REW_TERMINATION)) total += r_term;
          s_rew[slot] = total;
        }
      }
#undef ACC
    }

    // ------------------------------- observation heads -------------------------------
    if (do_obs) {
      __syncthreads();   // (B) observation heads are in shared memory
      if (warp >= nstate) {
        for (int slot = warp - nstate; slot < nenv; slot += nrow) {
          const int env = env0 + slot;
          uint4 rnd = blk0;
          for (int k = lane; k < head; k += 32) {
            float u = 0.0f;
            if (pr.noise_mode == ELG_NOISE_TENSOR) u = __ldg(bf.noise_u + (size_t)env * O + k);
            if (pr.noise_mode == ELG_NOISE_PHILOX) {
              const int m = k >> 5;
              if (((m & 3) == 0 && (m > 0 || blk0_env != env)) ) rnd = noise_block(pr.noise_seed, pr.noise_offset, env, lane, m >> 2);
              u = u01(pick(rnd, m & 3));
            }
            const float v = finish_obs(s_head[slot * headp + k], u, s_ns[k], pr);
            if (L.obs_smem) s_obs[slot * O + k] = v;
            else bf.obs_buf[(size_t)env * O + k] = v;
          }
        }
      }
    }

    // ------------------------------- write back -------------------------------
    auto for_each_output = [&](auto&& f) {
      if (do_derive) {
        f(bf.base_lin_vel + e0 * 3, s_vec5 + 0 * cap * 3, 4u * nenv * 3);
        f(bf.base_ang_vel + e0 * 3, s_vec5 + 1 * cap * 3, 4u * nenv * 3);
        f(bf.projected_gravity + e0 * 3, s_vec5 + 2 * cap * 3, 4u * nenv * 3);
        f(bf.base_lin_acc + e0 * 3, s_vec5 + 3 * cap * 3, 4u * nenv * 3);
        f(bf.base_ang_acc + e0 * 3, s_vec5 + 4 * cap * 3, 4u * nenv * 3);
        if (F > 0) {
          f(bf.foot_positions + e0 * F * 3, s_fpos, 4u * nenv * F * 3);
          f(bf.foot_velocities + e0 * F * 3, s_fvel, 4u * nenv * F * 3);
        }
        if (pr.heading_command) f(bf.commands + e0 * C, s_cmd, 4u * nenv * C);
        f(bf.episode_length_buf + e0, s_ep, 8u * nenv);
        if (H > 0) f(bf.measured_heights + e0 * H, s_mh, 4u * nenv * H);
      }
      if (do_reward) {
        if (term_on(pr, ELG_REW_FEET_AIR_TIME) && F > 0) {
          f(bf.feet_air_time + e0 * F, s_air, 4u * nenv * F);
          f(bf.feet_contact_time + e0 * F, s_con, 4u * nenv * F);
          f(bf.last_contacts + e0 * F, s_lc, (uint32_t)(nenv * F));
        }
        for (int ti = 0; ti < L.nterms; ++ti) f(bf.episode_sums + (size_t)L.term_ids[ti] * N + e0, s_sums + ti * cap, 4u * nenv);
        f(bf.rew_buf + e0, s_rew, 4u * nenv);
        if (gait) {
          f(bf.gait_idx + e0, s_gidx, 4u * nenv);
          f(bf.gait_prev_foot_z + e0 * F, s_gprev, 4u * nenv * F);
        }
      }
      if (do_obs && L.obs_smem) f(bf.obs_buf + e0 * O, s_obs, 4u * nenv * O);
      if (do_hist) {
        f(bf.last_actions + e0 * D, s_lact, 4u * nenv * D);
        f(bf.last_dof_vel + e0 * D, s_ldv, 4u * nenv * D);
        f(bf.last_root_vel + e0 * 6, s_lrv, 4u * nenv * 6);
      }
    };
    if (bulk) {
      fence_async_smem();   // generic-proxy writes of every thread -> visible to the async (TMA) proxy
      __syncthreads();
      if (tid == 0) {
        for_each_output([&](void* g, const void* s, uint32_t bytes) { bulk_s2g(g, s, bytes); });
        bulk_commit();
      }
      stores_pending = true;
    } else {
      __syncthreads();
      for_each_output([&](void* g, const void* s, uint32_t bytes) { coop_copy(g, s, bytes, tid, nthreads); });
    }
  }
  if (stores_pending && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// PD torques (legged_robot.py:425-448): one thread per (env, dof)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
elg_torques_kernel(const int64_t n_rows, const int D, const int control_type, const float action_scale, const float sim_dt,
                   const float* __restrict__ actions, const float* __restrict__ dof_state,
                   const float* __restrict__ last_dof_vel, const float* __restrict__ p_gains,
                   const float* __restrict__ d_gains, const float* __restrict__ torque_limits,
                   const float* __restrict__ default_dof_pos, float* __restrict__ torques,
                   const int64_t* __restrict__ env_ids) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * D) return;
  const int64_t r = i / D;
  const int j = (int)(i - r * D);
  const int64_t env = env_ids ? env_ids[r] : r;
  const int64_t e = env * D + j;
  const float a = mul_r(actions[e], action_scale);
  float tq;
  if (control_type == ELG_CONTROL_P) {
    const float2 pv = *reinterpret_cast<const float2*>(dof_state + 2 * e);
    tq = sub_r(mul_r(__ldg(p_gains + j), sub_r(add_r(a, __ldg(default_dof_pos + j)), pv.x)), mul_r(__ldg(d_gains + j), pv.y));
  } else if (control_type == ELG_CONTROL_V) {
    const float2 pv = *reinterpret_cast<const float2*>(dof_state + 2 * e);
    tq = sub_r(mul_r(__ldg(p_gains + j), sub_r(a, pv.y)),
               div_r(mul_r(__ldg(d_gains + j), sub_r(pv.y, last_dof_vel[e])), sim_dt));
  } else {
    tq = a;
  }
  const float lim = __ldg(torque_limits + j);
  torques[e] = fminf(fmaxf(tq, -lim), lim);
}

// ---------------------------------------------------------------------------------------------
// standalone height scan (LeggedRobot._get_heights), one warp per environment
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
elg_heights_kernel(const __grid_constant__ ElgDims dm, const __grid_constant__ ElgStepParams pr,
                   const float* __restrict__ root_states, const int16_t* __restrict__ hs,
                   const float* __restrict__ height_points, float* __restrict__ out, int32_t* __restrict__ cells) {
  const int env = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (env >= dm.num_envs) return;
  const int H = dm.num_height_points;
  const YawFrame fr = make_yaw_frame(root_states + (size_t)env * 13);
  const float* hp = height_points + (size_t)env * pr.height_points_env_stride;
  for (int p = lane; p < H; p += kWarp) {
    float h = 0.0f;
    int ix = 0, iy = 0;
    if (!pr.terrain_is_plane) {
      terrain_cell(fr, __ldg(hp + 3 * p), __ldg(hp + 3 * p + 1), pr, ix, iy);
      h = cell_height(hs, ix, iy, pr);
    }
    out[(size_t)env * H + p] = h;
    if (cells) {
      cells[((size_t)env * H + p) * 2] = ix;
      cells[((size_t)env * H + p) * 2 + 1] = iy;
    }
  }
}

}  // namespace elg

// =================================================================================================
// C ABI
// =================================================================================================
namespace {
struct StepTune { int cap, threads, ctas_per_sm, no_bulk; };
StepTune g_tune = {0, 0, 0, 0};
thread_local char g_err[256] = "";
int fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return ELG_ERR_CUDA;
  }
  return ELG_OK;
}
const char* kTermNames[ELG_NUM_REWARD_TERMS] = {
    "action_rate", "ang_vel_xy", "base_foot_height", "base_height", "collision", "dof_acc", "dof_pos_limits", "dof_vel",
    "dof_vel_limits", "feet_air_time", "feet_contact_forces", "feet_slip", "feet_stumble", "feet_stumble_liftup",
    "four_footup", "gait_2_step", "gait_scheduler", "jump_air", "lin_vel_z", "orientation", "stand_still", "termination",
    "torque_limits", "torques", "tracking_ang_vel", "tracking_lin_vel"};

int validate_dims(const ElgDims* d) {
  if (!d) return fail(ELG_ERR_NULL_POINTER, "dims is NULL");
  if (d->num_envs < 0) return fail(ELG_ERR_INVALID_ARGUMENT, "num_envs < 0");
  if (d->num_dof < 1 || d->num_dof > ELG_MAX_DOF) return fail(ELG_ERR_INVALID_ARGUMENT, "num_dof outside [1, ELG_MAX_DOF]");
  if (d->num_feet < 0 || d->num_feet > ELG_MAX_FEET) return fail(ELG_ERR_INVALID_ARGUMENT, "num_feet outside [0, ELG_MAX_FEET]");
  if (d->num_penalised < 0 || d->num_penalised > ELG_MAX_PENALISED) return fail(ELG_ERR_INVALID_ARGUMENT, "num_penalised out of range");
  if (d->num_termination < 0 || d->num_termination > ELG_MAX_TERMINATION) return fail(ELG_ERR_INVALID_ARGUMENT, "num_termination out of range");
  for (int i = 0; i < d->num_feet; ++i)
    if (d->feet_idx[i] < 0 || d->feet_idx[i] >= d->num_bodies) return fail(ELG_ERR_INVALID_ARGUMENT, "feet_idx out of range");
  for (int i = 0; i < d->num_penalised; ++i)
    if (d->penalised_idx[i] < 0 || d->penalised_idx[i] >= d->num_bodies) return fail(ELG_ERR_INVALID_ARGUMENT, "penalised_idx out of range");
  for (int i = 0; i < d->num_termination; ++i)
    if (d->termination_idx[i] < 0 || d->termination_idx[i] >= d->num_bodies) return fail(ELG_ERR_INVALID_ARGUMENT, "termination_idx out of range");
  return ELG_OK;
}
}  // namespace

extern "C" {

int elg_abi_version(void) { return ELG_ABI_VERSION; }
int elg_sizeof_dims(void) { return (int)sizeof(ElgDims); }
int elg_sizeof_step_params(void) { return (int)sizeof(ElgStepParams); }
int elg_sizeof_step_buffers(void) { return (int)sizeof(ElgStepBuffers); }
const char* elg_last_error(void) { return g_err; }
const char* elg_reward_term_name(int term) { return (term >= 0 && term < ELG_NUM_REWARD_TERMS) ? kTermNames[term] : nullptr; }

int elg_compute_torques(const ElgDims* dims, const ElgStepParams* prm, const float* actions, const float* dof_state,
                        const float* last_dof_vel, const float* p_gains, const float* d_gains, const float* torque_limits,
                        const float* default_dof_pos, float* torques, const int64_t* env_ids, int64_t num_ids, void* stream) {
  if (int rc = validate_dims(dims)) return rc;
  if (!prm) return fail(ELG_ERR_NULL_POINTER, "params is NULL");
  if (!actions || !torques || !torque_limits) return fail(ELG_ERR_NULL_POINTER, "actions/torques/torque_limits is NULL");
  if (prm->control_type < ELG_CONTROL_P || prm->control_type > ELG_CONTROL_T)
    return fail(ELG_ERR_INVALID_ARGUMENT, "Unknown controller type");
  if (prm->control_type != ELG_CONTROL_T && (!dof_state || !p_gains || !d_gains || !default_dof_pos))
    return fail(ELG_ERR_NULL_POINTER, "P/V control needs dof_state, gains and default_dof_pos");
  if (prm->control_type == ELG_CONTROL_V && !last_dof_vel) return fail(ELG_ERR_NULL_POINTER, "V control needs last_dof_vel");
  const int64_t rows = env_ids ? num_ids : dims->num_envs;
  if (rows <= 0) return ELG_OK;
  const int64_t total = rows * dims->num_dof;
  const int threads = 256;
  const unsigned blocks = (unsigned)((total + threads - 1) / threads);
  elg::elg_torques_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(
      rows, dims->num_dof, prm->control_type, prm->action_scale, prm->sim_dt, actions, dof_state, last_dof_vel, p_gains,
      d_gains, torque_limits, default_dof_pos, torques, env_ids);
  return check_launch("elg_compute_torques");
}

int elg_post_physics_step(const ElgDims* dims, const ElgStepParams* prm, const ElgStepBuffers* buf, uint32_t phase, void* stream) {
  if (int rc = validate_dims(dims)) return rc;
  if (!prm || !buf) return fail(ELG_ERR_NULL_POINTER, "params/buffers is NULL");
  if (phase == 0 || phase > ELG_PHASE_FUSED) return fail(ELG_ERR_INVALID_ARGUMENT, "phase must be a non-empty OR of ELG_PHASE_* bits");
  const int head = 12 + 3 * dims->num_dof;
  if (dims->num_obs < head + dims->num_height_points) return fail(ELG_ERR_INVALID_ARGUMENT, "num_obs < 12 + 3*num_dof + num_height_points");
  if (dims->num_commands < 3) return fail(ELG_ERR_INVALID_ARGUMENT, "num_commands < 3");
  if (prm->heading_command && dims->num_commands < 4) return fail(ELG_ERR_INVALID_ARGUMENT, "heading_command needs 4 commands");
  if (!buf->root_states || !buf->dof_state || !buf->contact_forces || !buf->rigid_body_state || !buf->actions || !buf->torques)
    return fail(ELG_ERR_NULL_POINTER, "a PhysX state / action / torque pointer is NULL");
  if (!buf->default_dof_pos || !buf->commands || !buf->last_actions || !buf->last_dof_vel || !buf->last_root_vel)
    return fail(ELG_ERR_NULL_POINTER, "an env-owned state pointer is NULL");
  if (!buf->base_lin_vel || !buf->base_ang_vel || !buf->projected_gravity || !buf->base_lin_acc || !buf->base_ang_acc ||
      !buf->foot_positions || !buf->foot_velocities || !buf->feet_air_time || !buf->feet_contact_time || !buf->last_contacts ||
      !buf->episode_length_buf || !buf->episode_sums || !buf->reset_buf || !buf->time_out_buf || !buf->rew_buf || !buf->obs_buf)
    return fail(ELG_ERR_NULL_POINTER, "an output pointer is NULL");
  if (dims->num_height_points > 0) {
    if (!buf->measured_heights || !buf->height_points) return fail(ELG_ERR_NULL_POINTER, "height scan needs measured_heights and height_points");
    if (!prm->terrain_is_plane && (!buf->height_samples || prm->hf_rows < 2 || prm->hf_cols < 2))
      return fail(ELG_ERR_INVALID_ARGUMENT, "height scan needs height_samples with rows, cols >= 2");
  }
  if ((prm->reward_mask >> ELG_REW_BASE_HEIGHT) & 1u)
    if (dims->num_height_points <= 0) return fail(ELG_ERR_UNSUPPORTED, "_reward_base_height needs measured heights");
  const uint32_t lim = (1u << ELG_REW_DOF_POS_LIMITS) | (1u << ELG_REW_DOF_VEL_LIMITS) | (1u << ELG_REW_TORQUE_LIMITS);
  if ((prm->reward_mask & lim) && (!buf->dof_pos_limits || !buf->dof_vel_limits || !buf->torque_limits))
    return fail(ELG_ERR_NULL_POINTER, "limit reward terms need dof_pos_limits, dof_vel_limits and torque_limits");
  if (prm->noise_mode == ELG_NOISE_TENSOR && (!buf->noise_u || !buf->noise_scale_vec)) return fail(ELG_ERR_NULL_POINTER, "ELG_NOISE_TENSOR needs noise_u and noise_scale_vec");
  if (prm->noise_mode == ELG_NOISE_PHILOX && !buf->noise_scale_vec) return fail(ELG_ERR_NULL_POINTER, "ELG_NOISE_PHILOX needs noise_scale_vec");
  if (prm->noise_mode < ELG_NOISE_OFF || prm->noise_mode > ELG_NOISE_PHILOX) return fail(ELG_ERR_INVALID_ARGUMENT, "bad noise_mode");
  const int N = dims->num_envs;
  if (N == 0) return ELG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // TMA bulk staging needs 16-byte aligned array bases; chunk offsets and sizes are multiples of 16 by construction
  // (whole quads of envs) when N and the foot count are multiples of 4.  Otherwise: cooperative element-wise staging.
  auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  const void* ptrs[] = {buf->root_states, buf->dof_state, buf->actions, buf->last_actions, buf->last_dof_vel, buf->torques,
                        buf->contact_forces, buf->last_root_vel, buf->base_lin_acc, buf->base_ang_acc, buf->commands,
                        buf->feet_air_time, buf->feet_contact_time, buf->last_contacts, buf->episode_length_buf, buf->gait_idx,
                        buf->gait_prev_foot_z, buf->base_lin_vel, buf->base_ang_vel, buf->projected_gravity, buf->foot_positions,
                        buf->foot_velocities, buf->measured_heights, buf->episode_sums, buf->rew_buf, buf->obs_buf};
  int use_bulk = (N % 4 == 0) && (dims->num_feet % 4 == 0) && g_tune.no_bulk == 0;
  for (const void* p : ptrs) use_bulk = use_bulk && a16(p);

  static int sms = 0;
  if (sms == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
      return fail(ELG_ERR_CUDA, "cannot query the SM count");
    sms = v;
  }
  // chunking: whole quads of envs, balanced over the SMs (see the header comment of this file)
  const long long Q = ((long long)N + 3) / 4;
  const bool obs_smem = dims->num_obs == head + dims->num_height_points;
  const int kSmemLimit = 227 * 1024;
  int cap = 0, threads = 0, grid = 0, nchunks = 0, nstate = 0;
  elg::StepPlan plan{};
  auto try_plan = [&](int c, int th, int g, int nch) -> bool {
    const int ns = (c + elg::kGroup - 1) / elg::kGroup;
    if (th / 32 - ns < 1) return false;
    elg::StepPlan pl = elg::make_plan(*dims, *prm, c, ns, obs_smem);
    if ((size_t)pl.words * 4 + 1024 > (size_t)kSmemLimit) return false;
    cap = c; threads = th; grid = g; nchunks = nch; nstate = ns; plan = pl;
    return true;
  };
  bool ok = false;
  if (g_tune.cap > 0) {   // explicit tuning (elg_set_step_tuning): cap envs per chunk, threads per CTA, CTAs per SM
    const int c = g_tune.cap;
    long long nch = (Q + c / 4 - 1) / (c / 4);
    const long long g = (long long)sms * g_tune.ctas_per_sm;
    if (nch > g) nch = (nch + g - 1) / g * g;
    if (nch > Q) nch = Q;
    ok = try_plan(c, g_tune.threads, (int)(nch < g ? nch : g), (int)nch);
  }
  if (!ok && (Q + 7) / 8 <= sms) {   // at most one chunk per SM: one wide CTA per SM, every chunk <= 32 envs
    const int nch = (int)(Q < sms ? Q : sms);
    const int c = 4 * (int)((Q + nch - 1) / nch);
    const int ns = (c + elg::kGroup - 1) / elg::kGroup;
    int nrow = c < 32 - ns ? c : 32 - ns;
    ok = try_plan(c, 32 * (ns + nrow), nch, nch);
  }
  for (int c = 16; !ok && c >= 4; c -= 4) {   // many chunks: narrow CTAs, 4 per SM, each looping over its chunks
    const long long g = (long long)sms * 4;
    long long nch = (Q + c / 4 - 1) / (c / 4);
    if (nch > g) nch = (nch + g - 1) / g * g;
    if (nch > Q) nch = Q;
    const int ns = (c + elg::kGroup - 1) / elg::kGroup;
    ok = try_plan(c, 32 * (ns + 6), (int)(nch < g ? nch : g), (int)nch);
  }
  if (!ok) return fail(ELG_ERR_UNSUPPORTED, "robot dimensions do not fit the shared-memory plan of elg_step_kernel");
  plan.nchunks = nchunks;
  plan.use_bulk = use_bulk;
  const size_t smem = (size_t)plan.words * 4;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    if (cudaFuncSetAttribute(elg::elg_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return fail(ELG_ERR_CUDA, "cannot reserve dynamic shared memory for elg_step_kernel");
    smem_set = smem;
  }
  elg::elg_step_kernel<<<grid, threads, smem, st>>>(*dims, *prm, *buf, plan, phase);
  return check_launch("elg_post_physics_step");
}

int elg_set_step_tuning(int envs_per_chunk, int threads_per_cta, int ctas_per_sm, int disable_bulk) {
  if (envs_per_chunk == 0) {
    g_tune = StepTune{0, 0, 0, disable_bulk};
    return ELG_OK;
  }
  if (envs_per_chunk < 4 || envs_per_chunk > elg::kMaxCap || envs_per_chunk % 4 != 0)
    return fail(ELG_ERR_INVALID_ARGUMENT, "envs_per_chunk must be a multiple of 4 in [4, 32]");
  if (threads_per_cta < 64 || threads_per_cta > elg::kMaxStepThreads || threads_per_cta % 32 != 0)
    return fail(ELG_ERR_INVALID_ARGUMENT, "threads_per_cta must be a multiple of 32 in [64, 1024]");
  if (ctas_per_sm < 1 || ctas_per_sm > 16) return fail(ELG_ERR_INVALID_ARGUMENT, "ctas_per_sm must be in [1, 16]");
  g_tune = StepTune{envs_per_chunk, threads_per_cta, ctas_per_sm, disable_bulk};
  return ELG_OK;
}

int elg_get_heights(const ElgDims* dims, const ElgStepParams* prm, const float* root_states, const int16_t* height_samples,
                    const float* height_points, float* measured_heights, int32_t* cells_out, void* stream) {
  if (int rc = validate_dims(dims)) return rc;
  if (!prm) return fail(ELG_ERR_NULL_POINTER, "params is NULL");
  if (!root_states || !height_points || !measured_heights) return fail(ELG_ERR_NULL_POINTER, "root_states/height_points/measured_heights is NULL");
  if (!prm->terrain_is_plane && (!height_samples || prm->hf_rows < 2 || prm->hf_cols < 2))
    return fail(ELG_ERR_INVALID_ARGUMENT, "height scan needs height_samples with rows, cols >= 2");
  if (dims->num_envs == 0 || dims->num_height_points == 0) return ELG_OK;
  const int warps = 8;
  elg::elg_heights_kernel<<<(dims->num_envs + warps - 1) / warps, warps * 32, 0, (cudaStream_t)stream>>>(
      *dims, *prm, root_states, height_samples, height_points, measured_heights, cells_out);
  return check_launch("elg_get_heights");
}

}  // extern "C"
