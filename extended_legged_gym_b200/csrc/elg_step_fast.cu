// elg_step_fast.cu -- the fused post-physics step for the common quadruped layout (v5; round 2 measured a v6 whose terrain scan
// started on root_states alone and whose reward registry was split over four warps -- 0.2-0.5 us SLOWER per launch on the same box,
// profiles/README.md r2 -- and kept this structure).
//
// Same contract as elg_step_kernel (elg_step.cu) for ELG_PHASE_FUSED on a 12-DOF / 4-foot robot with the
// shared [H,3] height grid and the min-of-3 terrain table; everything else keeps going through the generic
// kernel.  What the v4 profile and the in-kernel timelines (profiles/README.md) say about this step at
// 28 envs per SM: it is bound by ISSUE SLOTS and by the LENGTH OF SINGLE-WARP CODE PATHS (1557
// warp-instructions per env; straight-line code that only one warp runs is fetched cold and crawls at
// > 10 cycles per instruction), not by HBM.  Hence:
//
//   * one chunk of n <= 28 consecutive envs per CTA; TMA in / TMA out (host-built copy tables, whole
//     quads of envs, one mbarrier); programmatic dependent launch so that barrier set-up and geometry
//     overlap the tail of the previous kernel in the stream.
//   * phase A, all warps: warp == ITEM, lane == env.  19 short tasks (4 feet, 5 base-frame rotations,
//     1 yaw frame + commands, 6 DOF pairs, 3 contact-body groups) run side by side, the warps of one kind
//     share their code, nobody reduces across lanes.  Partial sums go to a [row][env] table.
//   * phase B: ROW warps (warp == env) run the terrain scan -- six height points and their noise scales
//     stay in registers, the six gathers go out back to back, Philox runs while they are in flight --
//     then the observation head; meanwhile ONE warp (lane == env) assembles the reward terms.
//   * code footprint kept inside the 32 KB instruction cache level (loops stay rolled where only one warp
//     runs them).
//
// Arithmetic is the v4 kernel's, expression for expression (DESIGN.md "rounding model"); only per-env
// sums over DOFs / feet run in a fixed index order now (fp32 tolerance class, not bit-exact class).
#include <cuda_runtime.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include "elg_common.cuh"
#include "elg_async.cuh"

namespace elg {

constexpr int kFastMaxCap = 28;                     // row warps per CTA
constexpr int kTaskWarps = 4;                       // spare warps behind the row warps (the first assembles the rewards)
constexpr int kFastMaxIn = 16 + ELG_NUM_REWARD_TERMS;
constexpr int kFastMaxOut = 20 + ELG_NUM_REWARD_TERMS;
constexpr int kNJ = 6;                              // height points per lane (H <= 192)

struct FastCopy {
  const void* g;   // global base of the array (env 0)
  int32_t soff;    // shared-memory byte offset of slot 0
  int32_t bpe;     // bytes per environment
};

struct FastPlan {
  int cap, nchunks, quads_base, quads_rem;
  int n_in, n_out, n_out_early, nterms;
  // staged per-env arrays, byte offsets into dynamic shared memory
  int root, dof, act, lact, ldv, tq, cf, lrv, vec5, cmd, air, con, lc, ep, gidx, gprev, fpos, fvel, sums, rew, mh, obs;
  int buf_bytes;   // one staging buffer (everything above); two of them when a CTA handles more than one chunk
  int part, yaw;   // partial-sum table [rows][32], yaw frames [cap] x (float4 + float): one copy, behind the buffers
  int bytes;
  float r_hscale;  // RN(1 / horizontal_scale)
  float r_dt;      // RN(1 / dt)
  float pen_sq, term_sq;   // largest sums of squares whose IEEE sqrt is still <= 0.1 / <= 1.0 (contact thresholds)
  int flags;       // diagnostic switches (elg_set_step_tuning threads_per_cta): 16 = launch without PDL
  long long* dbg;  // diagnostic: clock64 stamps of CTA 0 (elg_set_step_debug), or NULL
  int8_t term_ids[ELG_NUM_REWARD_TERMS];
  FastCopy in[kFastMaxIn];
  FastCopy out[kFastMaxOut];
};

// phase-A tasks (warp == task): feet first, their strided global rows take longest; a foot is two tasks -- timers / contact
// logic / gather (kTaskFeet + f) and the force / velocity norms (kTaskFeetB + f) -- because it was the longest chain by far
enum { kTaskFeet = 0, kTaskRot = 4, kTaskCmd = 9, kTaskDof = 10, kTaskBody = 16, kTaskFeetB = 19, kNumTasks = 23 };
constexpr int kDofWarps = 6, kBodyWarps = 3;
constexpr int kBarHsum = 1;   // named barrier (0 is __syncthreads)

// rows of the partial-sum table: [row][32], lane == env.  Registry ids first (final per-term values).
enum { kDAr = 0, kDDa, kDDv, kDTq, kDSs, kDPl, kDVl, kDTl, kNumDofSums };
enum { kFTz = 0, kFTn, kFAir, kFCf, kFSlip, kFLift, kFJump, kFStum, kFDown, kFGs, kNumFeetSums };
enum {
  kPDof = ELG_NUM_REWARD_TERMS,                 // [sum][dof warp]
  kPFeet = kPDof + kNumDofSums * kDofWarps,     // [sum][foot]
  kPHits = kPFeet + kNumFeetSums * 4,           // [body warp]
  kPTermHit = kPHits + kBodyWarps,              // [body warp]
  kPHsum = kPTermHit + kBodyWarps,
  kPRows
};

__device__ __forceinline__ bool on(const ElgStepParams& pr, int t) { return (pr.reward_mask >> t) & 1u; }

// kNoise: ElgNoiseMode, kClip: clip_observations > 0 -- compile-time so that the per-point code carries no mode tests
// kLoop: persistent form (a CTA walks over several chunks, double-buffered staging); false: one chunk per CTA, no loop state
// kRollout: post_physics_step_rollout form (rollout_mode) -- compile-time so that the headline kernel carries none of its tests
template <int kNoise, bool kClip, bool kLoop, bool kRollout>
__global__ void __launch_bounds__(32 * (kFastMaxCap + kTaskWarps), 1)
elg_step_fast_kernel(const __grid_constant__ ElgDims dm, const __grid_constant__ ElgStepParams pr,
                     const __grid_constant__ ElgStepBuffers bf, const __grid_constant__ FastPlan L) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_bar[2];   // one mbarrier per staging buffer

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nwarps = blockDim.x >> 5;
  const int cap = L.cap;
  const int H = dm.num_height_points, O = dm.num_obs, B = dm.num_bodies, C = dm.num_commands;
  constexpr int D = 12, F = 4, head = 12 + 3 * D;
  const float clip_obs = pr.clip_observations;
  const bool need_hsum = on(pr, ELG_REW_BASE_HEIGHT) && H > 0;
  // rollout mode (post_physics_step_rollout, batch_rollout/robot_batch_rollout.py:763-817): no episode counter, heading command,
  // height scan (measured_heights is an input), termination or episode sums
  constexpr bool rollout = kRollout;
  const bool heights_live = H > 0 && !pr.terrain_is_plane && !rollout;
  const bool gait = bf.gait_idx != nullptr && bf.gait_prev_foot_z != nullptr;
#ifdef ELG_STEP_STAMPS   // diagnostic build only (scripts/step_stamps.py): the product kernel carries no stamp code
  const bool dbg_on = L.dbg != nullptr && blockIdx.x == 0;
#define STAMP(i, w) if (dbg_on && warp == (w) && lane == 0) L.dbg[i] = clock64();
#else
#define STAMP(i, w)
#endif
  STAMP(0, 0)
  auto chunk_range = [&](int chunk, int& env0, int& n) {   // whole quads of envs, balanced to within one quad
    const int q_lo = chunk * L.quads_base + min(chunk, L.quads_rem);
    const int q_n = L.quads_base + (chunk < L.quads_rem ? 1 : 0);
    env0 = q_lo * 4;
    n = min(dm.num_envs, (q_lo + q_n) * 4) - env0;   // multiple of 4 (host guarantees N % 4 == 0)
  };

  // x / dt as the tail of the IEEE division sequence (q = x r, two FMA residual corrections; r = RN(1 / dt) from the host):
  // same value as the division for normal-range operands, 5 instructions, no out-of-line special-operand path
  auto div_dt = [&](float x) {
    float q = __fmul_rn(x, L.r_dt);
    float e = __fmaf_rn(-pr.dt, q, x);
    q = __fmaf_rn(e, L.r_dt, q);
    e = __fmaf_rn(-pr.dt, q, x);
    return __fmaf_rn(e, L.r_dt, q);
  };

  if (tid == 0) {
    mbar_init(&s_bar[0], L.n_in);
    mbar_init(&s_bar[1], L.n_in);
  }
  pdl_launch_dependents();
  // the copy-table entries this warp will issue: fetched before the wait so that the constant-bank miss is off the load path
  FastCopy my_in = L.in[warp < L.n_in ? warp : 0];
  __syncthreads();
  pdl_wait();   // nothing above reads or writes global memory
  STAMP(1, 0)

  // ---- TMA loads of one chunk into staging buffer b: lane 0 of warp w issues copy-table entries w, w + W, ...
  auto issue_loads = [&](int chunk, int b) {
    if (lane == 0) {
      int env0, n;
      chunk_range(chunk, env0, n);
      for (int i = warp; i < L.n_in; i += nwarps) {
        const FastCopy d = i == warp ? my_in : L.in[i];
        const uint32_t bytes = (uint32_t)(n * d.bpe);
        mbar_expect_tx(&s_bar[b], bytes);
        bulk_g2s(smem_raw + b * L.buf_bytes + d.soff, static_cast<const uint8_t*>(d.g) + (size_t)env0 * d.bpe, bytes, &s_bar[b]);
      }
    }
  };
  issue_loads(blockIdx.x, 0);

  float* const s_part = reinterpret_cast<float*>(smem_raw + L.part);
#define PART(r, e) s_part[(r) * 32 + (e)]
  float4* const s_yaw = reinterpret_cast<float4*>(smem_raw + L.yaw);   // (zz, ww, X, Y)
  float* const s_yz = reinterpret_cast<float*>(smem_raw + L.yaw + cap * 16);   // Z

  // this lane's height points p = lane + 32 j: registers for the whole kernel
  float gx[kNJ], gy[kNJ];
  const int p_last = min(lane + 32 * (kNJ - 1), H - 1);   // H > 32 (kNJ - 1): only the last round is ragged, its surplus lanes shadow point H - 1
  if (warp < cap && heights_live) {
    const float* hp = bf.height_points + 3 * lane;
#pragma unroll
    for (int j = 0; j < kNJ - 1; ++j) {
      gx[j] = __ldg(hp + 96 * j);
      gy[j] = __ldg(hp + 96 * j + 1);
    }
    gx[kNJ - 1] = __ldg(bf.height_points + 3 * p_last);
    gy[kNJ - 1] = __ldg(bf.height_points + 3 * p_last + 1);
  }

  // =========================================================================================================
  // chunks of this CTA: staging buffer it & 1; the loads of the next chunk go out behind barrier B1 of the current one
  // and land under its terrain scan, the stores of the previous chunk drain under the current phase A
  // =========================================================================================================
  bool stores_pending = false;   // (lane 0 of the warps that issue stores)
#pragma unroll 1
  for (int it = 0, chunk = blockIdx.x; kLoop ? chunk < L.nchunks : it < 1; ++it, chunk += gridDim.x) {
  const int bsel = kLoop ? (it & 1) : 0;
  const uint32_t parity = kLoop ? ((uint32_t)(it >> 1) & 1u) : 0u;
  uint8_t* const sbuf = smem_raw + bsel * L.buf_bytes;
  int env0, n;
  chunk_range(chunk, env0, n);
#define SM_F(off) reinterpret_cast<float*>(sbuf + (off))
  float* const s_obs = SM_F(L.obs);
  const float* const s_root = SM_F(L.root);
  const float* const s_cf = SM_F(L.cf);

  // ---- row warps, while the bulk copies are in flight: this lane's height points p = lane + 32 j and everything about
  // the observation noise that does not depend on the state -- nz = (2 u - 1) * noise_scale per height point / head entry.
  // Philox: 16-bit samples, 8 per 128-bit block; point j uses sample j, head entry k = lane + 32 m sample nj + m of
  // block 1 of (env, lane) (elg_common.cuh).
  const bool row_warp = warp < n;
  float nz[kNJ], nzh[2];
  if (row_warp) {
    const int env = env0 + warp;
    if (kNoise != ELG_NOISE_OFF) {
      const int nj = (H + 31) >> 5;
      float ns[kNJ], nsd[2], u[kNJ], ud[2];
      const bool ns_on = bf.noise_scale_vec != nullptr;
#pragma unroll
      for (int j = 0; j < kNJ; ++j) ns[j] = 0.0f;
      if (ns_on && H > 0) {
        const float* nsb = bf.noise_scale_vec + head + lane;
#pragma unroll
        for (int j = 0; j < kNJ - 1; ++j) ns[j] = __ldg(nsb + 32 * j);
        ns[kNJ - 1] = __ldg(bf.noise_scale_vec + head + p_last);
      }
      nsd[0] = ns_on ? __ldg(bf.noise_scale_vec + lane) : 0.0f;
      nsd[1] = ns_on ? __ldg(bf.noise_scale_vec + min(lane + 32, head - 1)) : 0.0f;
      if (kNoise == ELG_NOISE_PHILOX) {
        const uint4 blk = noise_block(pr.noise_seed, pr.noise_offset + (bf.step_counter ? *bf.step_counter : 0ull), env, lane, 1);
#pragma unroll
        for (int j = 0; j < kNJ; ++j) u[j] = sym16(blk, j);   // 2u - 1, exact
#pragma unroll
        for (int m = 0; m < 2; ++m) ud[m] = sym16(blk, (nj + m) & 7);
      } else {
        const float* nu = bf.noise_u + (size_t)env * O;
#pragma unroll
        for (int j = 0; j < kNJ; ++j) u[j] = 0.0f;
        if (H > 0) {
#pragma unroll
          for (int j = 0; j < kNJ - 1; ++j) u[j] = 2.0f * __ldg(nu + head + lane + 32 * j) - 1.0f;
          u[kNJ - 1] = 2.0f * __ldg(nu + head + p_last) - 1.0f;
        }
        ud[0] = 2.0f * __ldg(nu + lane) - 1.0f;
        ud[1] = 2.0f * __ldg(nu + min(lane + 32, head - 1)) - 1.0f;
      }
#pragma unroll
      for (int j = 0; j < kNJ; ++j) nz[j] = u[j] * ns[j];
#pragma unroll
      for (int m = 0; m < 2; ++m) nzh[m] = ud[m] * nsd[m];
    }
  }

  // =========================================================================================================
  // phase A: warp == item, lane == env slot
  // =========================================================================================================
  const bool live = lane < n;
  const int e = live ? lane : n - 1;   // surplus lanes shadow the last env (loads stay in range, stores are guarded)
  const int genv = env0 + e;
  // Task inputs that do not come through the bulk copies are fetched now, while those are in flight: feet tasks their strided
  // rigid_body_state row (52-byte rows, 6 useful floats), DOF tasks their default angles and soft limits.
  const int task0 = warp;   // first phase-A task of this warp
  const bool lim_terms = on(pr, ELG_REW_DOF_POS_LIMITS) | on(pr, ELG_REW_DOF_VEL_LIMITS) | on(pr, ELG_REW_TORQUE_LIMITS);
  float pre[10] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
  auto prefetch = [&](int task) {
    if (task < kTaskRot) {
      const float* row = bf.rigid_body_state + ((size_t)genv * B + dm.feet_idx[task - kTaskFeet]) * 13;
      pre[0] = __ldg(row + 0); pre[1] = __ldg(row + 1); pre[2] = __ldg(row + 2);
      pre[3] = __ldg(row + 7); pre[4] = __ldg(row + 8); pre[5] = __ldg(row + 9);
    } else if (task >= kTaskFeetB) {
      const float* row = bf.rigid_body_state + ((size_t)genv * B + dm.feet_idx[task - kTaskFeetB]) * 13;
      pre[3] = __ldg(row + 7); pre[4] = __ldg(row + 8); pre[5] = __ldg(row + 9);
    } else if (task >= kTaskDof && task < kTaskBody) {
      const int j = 2 * (task - kTaskDof);
      pre[0] = __ldg(bf.default_dof_pos + j);
      pre[1] = __ldg(bf.default_dof_pos + j + 1);
      if (lim_terms) {
        pre[2] = __ldg(bf.dof_pos_limits + 2 * j);     pre[3] = __ldg(bf.dof_pos_limits + 2 * j + 1);
        pre[4] = __ldg(bf.dof_pos_limits + 2 * j + 2); pre[5] = __ldg(bf.dof_pos_limits + 2 * j + 3);
        pre[6] = __ldg(bf.dof_vel_limits + j);         pre[7] = __ldg(bf.dof_vel_limits + j + 1);
        pre[8] = __ldg(bf.torque_limits + j);          pre[9] = __ldg(bf.torque_limits + j + 1);
      }
    }
  };
  prefetch(task0);
  STAMP(11, 0)
  mbar_wait(&s_bar[bsel], parity);
  STAMP(2, 0)
  STAMP(12, nwarps - 1)

#pragma unroll 1
  for (int task = task0; task < kNumTasks; task += nwarps) {
    if (task < kTaskRot) {
      // ---------------- foot f (legged_robot_rew_mixin.py:58-81, :121-212; gait_scheduler.py:74-81)
      // Terms that sort before feet_air_time read the OLD timers, terms after it the updated ones and the rebound
      // last_contacts.
      const int f = task - kTaskFeet, fi = e * F + f;
      if (task != task0) prefetch(task);   // second round of a narrow CTA: fetch now
      const float* fr = pre;
      float* const s_air = SM_F(L.air);
      float* const s_con = SM_F(L.con);
      uint8_t* const s_lc = sbuf + L.lc;
      const float* cf = s_cf + (e * B + dm.feet_idx[f]) * 3;
      const float fz = cf[2];
      const float pz = fr[2];
      float air = s_air[fi], con = s_con[fi];
      const bool last_c = s_lc[fi] != 0;
      const bool contact = fz > 1.0f;
      const bool touching = con > 1e-3f;   // base_foot_height: nanmean over touching feet (old timers)
      PART(kPFeet + 4 * kFTz + f, lane) = touching ? pz : 0.0f;
      PART(kPFeet + 4 * kFTn + f, lane) = touching ? 1.0f : 0.0f;
      bool lc_after = last_c;
      if (on(pr, ELG_REW_FEET_AIR_TIME)) {
        const bool filt = contact | last_c;
        const bool first = (air > 0.0f) && filt;
        air += pr.dt;
        con += pr.dt;
        PART(kPFeet + 4 * kFAir + f, lane) = (air - 0.5f) * (first ? 1.0f : 0.0f);
        air *= filt ? 0.0f : 1.0f;
        con *= filt ? 1.0f : 0.0f;
        if (live) {
          s_air[fi] = air;
          s_con[fi] = con;
          s_lc[fi] = contact ? 1 : 0;
        }
        lc_after = contact;
      }
      const bool filt2 = contact | lc_after;
      PART(kPFeet + 4 * kFJump + f, lane) = (filt2 ? 0.0f : 1.0f) * (air - 0.5f);
      PART(kPFeet + 4 * kFDown + f, lane) = fz < 1.0f ? 0.0f : 1.0f;   // number of feet that are NOT up
      if (gait) {
        float* const s_gprev = SM_F(L.gprev);
        if (on(pr, ELG_REW_GAIT_SCHEDULER)) {
          float ph = SM_F(L.gidx)[e] + pr.gait_foot_phases[f];
          ph = ph - floorf(ph);                                 // torch.remainder(x, 1.0)
          const float target = ph < 0.5f ? pr.gait_swing_height * sinf(6.283185307179586f * ph) : 0.0f;
          const float dz = target - s_gprev[fi];
          PART(kPFeet + 4 * kFGs + f, lane) = dz * dz;
        }
        if (live) s_gprev[fi] = pz;   // GaitScheduler.step keeps this step's feet
      }
      if (live) {
        float* fp = SM_F(L.fpos) + fi * 3;
        float* fv = SM_F(L.fvel) + fi * 3;
        fp[0] = fr[0]; fp[1] = fr[1]; fp[2] = fr[2];
        fv[0] = fr[3]; fv[1] = fr[4]; fv[2] = fr[5];
      }
    } else if (task < kTaskCmd) {
      // ---------------- rotation r: base-frame velocities, gravity, acceleration EMAs (legged_robot.py:128-134),
      // observation entries [3r, 3r + 3) for r < 3, root-velocity history (:150)
      const int r = task - kTaskRot;
      const float* rs = s_root + e * 13;
      const Quat q = {rs[3], rs[4], rs[5], rs[6]};
      float* lrv = SM_F(L.lrv) + e * 6 + (r == 4 ? 3 : 0);
      Vec3 v;
      if (r == 2) {
        v = Vec3{pr.gravity_vec[0], pr.gravity_vec[1], pr.gravity_vec[2]};
      } else {
        const int k = (r == 1 || r == 4) ? 10 : 7;
        v = Vec3{rs[k], rs[k + 1], rs[k + 2]};
        if (r >= 3) {
          const float ox = lrv[0], oy = lrv[1], oz = lrv[2];
          if (live) { lrv[0] = v.x; lrv[1] = v.y; lrv[2] = v.z; }
          v.x -= ox; v.y -= oy; v.z -= oz;
        }
      }
      Vec3 o = quat_rotate_inverse(q, v);
      float* dst = SM_F(L.vec5) + (r * cap + e) * 3;
      if (r >= 3) {
        const float ema = pr.acc_ema, w1 = pr.acc_ema_c;
        o.x = dst[0] * ema + div_dt(w1 * o.x);
        o.y = dst[1] * ema + div_dt(w1 * o.y);
        o.z = dst[2] * ema + div_dt(w1 * o.z);
      }
      if (live) {
        dst[0] = o.x; dst[1] = o.y; dst[2] = o.z;
        if (r < 3) {
          const float sc = r == 0 ? pr.obs_scale_lin_vel : r == 1 ? pr.obs_scale_ang_vel : 1.0f;
          float* hrow = s_obs + e * O + 3 * r;
          if (r == 2) { hrow[0] = o.x; hrow[1] = o.y; hrow[2] = o.z; }
          else { hrow[0] = o.x * sc; hrow[1] = o.y * sc; hrow[2] = o.z * sc; }
        }
      }
    } else if (task == kTaskCmd) {
      // ---------------- yaw frame for the terrain scan, episode counter, heading command, command observations
      const float* rs = s_root + e * 13;
      const Quat q = {rs[3], rs[4], rs[5], rs[6]};
      if (H > 0 && !rollout) {
        // normalize((0,0,qz,qw)): torch's 4-wide norm is the plain sequential sum (no FMA), clamp(min=1e-9)
        float nrm = __fsqrt_rn(add_r(mul_r(q.z, q.z), mul_r(q.w, q.w)));
        nrm = fmaxf(nrm, 1e-9f);
        if (live) {
          s_yaw[e] = make_float4(div_r(q.z, nrm), div_r(q.w, nrm), rs[0], rs[1]);
          s_yz[e] = rs[2];
        }
      }
      int64_t* const s_ep = reinterpret_cast<int64_t*>(sbuf + L.ep);
      if (live && !rollout) s_ep[e] += 1;   // episode counter (legged_robot.py:122)
      float* cmd = SM_F(L.cmd) + e * C;
      float cmd2 = cmd[2];
      if (pr.heading_command && !rollout) {   // (legged_robot.py:394-398); forward = quat_apply(q, (1,0,0))
        const float fx = 1.0f + (q.y * (-2.0f * q.y) - q.z * (2.0f * q.z));
        const float fy = q.w * (2.0f * q.z) + (q.z * 0.0f - q.x * (-2.0f * q.y));
        const float heading = atan2f(fy, fx);
        cmd2 = fminf(fmaxf(0.5f * wrap_to_pi(cmd[3] - heading), -1.0f), 1.0f);
        if (live) cmd[2] = cmd2;
      }
      if (live) {
        float* hrow = s_obs + e * O;
        hrow[9] = cmd[0] * pr.commands_scale[0]; hrow[10] = cmd[1] * pr.commands_scale[1]; hrow[11] = cmd2 * pr.commands_scale[2];
      }
    } else if (task < kTaskBody) {
      // ---------------- DOFs 2d, 2d + 1: reward partials, observation entries, history
      // (legged_robot_rew_mixin.py:84-114, legged_robot.py:237-244, :148-149)
      const int d = task - kTaskDof;
      if (task != task0) prefetch(task);
      const float2* s_dof = reinterpret_cast<const float2*>(sbuf + L.dof);
      float* const s_act = SM_F(L.act);
      float* const s_lact = SM_F(L.lact);
      float* const s_ldv = SM_F(L.ldv);
      const float* const s_tq = SM_F(L.tq);
      float* hrow = s_obs + e * O;
      float q_ar = 0.0f, q_da = 0.0f, q_dv = 0.0f, q_tq = 0.0f, q_ss = 0.0f, q_pl = 0.0f, q_vl = 0.0f, q_tl = 0.0f;
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = 2 * d + jj, fi = e * D + j;
        const float q0 = pre[jj];
        const float2 pv = s_dof[fi];
        const float pos = pv.x, vel = pv.y;
        const float a = s_act[fi];
        const float la = s_lact[fi], lv = s_ldv[fi], tq = s_tq[fi];
        const float da = la - a;
        q_ar += da * da;
        const float dv = div_dt(lv - vel);
        q_da += dv * dv;
        q_dv += vel * vel;
        q_tq += tq * tq;
        q_ss += fabsf(pos - q0);
        if (lim_terms) {
          q_pl += -fminf(pos - pre[2 + 2 * jj], 0.0f) + fmaxf(pos - pre[3 + 2 * jj], 0.0f);
          q_vl += fminf(fmaxf(fabsf(vel) - pre[6 + jj] * pr.soft_dof_vel_limit, 0.0f), 1.0f);
          q_tl += fmaxf(fabsf(tq) - pre[8 + jj] * pr.soft_torque_limit, 0.0f);
        }
        if (live) {
          hrow[12 + j] = (pos - q0) * pr.obs_scale_dof_pos;
          hrow[12 + D + j] = vel * pr.obs_scale_dof_vel;
          hrow[12 + 2 * D + j] = a;
          s_lact[fi] = a;
          s_ldv[fi] = vel;
        }
      }
      PART(kPDof + kDofWarps * kDAr + d, lane) = q_ar;
      PART(kPDof + kDofWarps * kDDa + d, lane) = q_da;
      PART(kPDof + kDofWarps * kDDv + d, lane) = q_dv;
      PART(kPDof + kDofWarps * kDTq + d, lane) = q_tq;
      PART(kPDof + kDofWarps * kDSs + d, lane) = q_ss;
      if (lim_terms) {
        PART(kPDof + kDofWarps * kDPl + d, lane) = q_pl;
        PART(kPDof + kDofWarps * kDVl + d, lane) = q_vl;
        PART(kPDof + kDofWarps * kDTl + d, lane) = q_tl;
      }
    } else if (task >= kTaskFeetB) {
      // ---------------- foot f, second half: contact-force, slip and stumble terms (legged_robot_rew_mixin.py:121-148, :208-212).
      // filt2 = contact | last_contacts AFTER feet_air_time rebinds it: with that term on it is just `contact`; with it off
      // last_contacts is not written by the first half, so reading it here does not race.
      const int f = task - kTaskFeetB, fi = e * F + f;
      if (task != task0) prefetch(task);
      const float* cf = s_cf + (e * B + dm.feet_idx[f]) * 3;
      const float fxx = cf[0], fyy = cf[1], fz = cf[2];
      const float vx = pre[3], vy = pre[4], vz = pre[5];
      const bool contact = fz > 1.0f;
      const bool filt2 = contact | (on(pr, ELG_REW_FEET_AIR_TIME) ? false : (sbuf + L.lc)[fi] != 0);
      PART(kPFeet + 4 * kFCf + f, lane) = fmaxf(norm3_tz(fxx, fyy, fz) - pr.max_contact_force, 0.0f);
      const float vn = norm2_tz(vx, vy);
      PART(kPFeet + 4 * kFSlip + f, lane) = (filt2 ? 1.0f : 0.0f) * (vn * vn);
      const bool stumble = norm2_tz(fxx, fyy) > mul_r(5.0f, fabsf(fz));
      PART(kPFeet + 4 * kFLift + f, lane) = (stumble ? 1.0f : 0.0f) * vz;
      PART(kPFeet + 4 * kFStum + f, lane) = stumble ? 1.0f : 0.0f;
    } else {
      // ---------------- contact bodies g, g + 3, ...: collision count and termination contacts
      // (legged_robot_rew_mixin.py:117-119, legged_robot.py:155-160); penalised bodies first, then termination bodies
      const int g = task - kTaskBody;
      const int P = dm.num_penalised, PT = P + dm.num_termination;
      int hits = 0;
      bool thit = false;
#pragma unroll 1
      for (int b = g; b < PT; b += kBodyWarps) {
        const int body = b < P ? dm.penalised_idx[b] : dm.termination_idx[b - P];
        const float* f = s_cf + (e * B + body) * 3;
        const float ss = sumsq3_t(f[0], f[1], f[2]);   // sqrt(ss) > t  <=>  ss > max{s : sqrt_rn(s) <= t}: no sqrt here
        if (b < P) hits += ss > L.pen_sq ? 1 : 0;
        else thit |= ss > L.term_sq;
      }
      PART(kPHits + g, lane) = (float)hits;
      PART(kPTermHit + g, lane) = thit ? 1.0f : 0.0f;
    }
  }
#ifdef ELG_STEP_STAMPS
  if (dbg_on && lane == 0) L.dbg[32 + warp] = clock64();   // per-warp end of phase A
#endif
  STAMP(3, 0)
  if (kLoop && stores_pending) {   // the previous chunk's stores have read the other staging buffer out (they had phase A to do so)
    bulk_wait_read_all();
    stores_pending = false;
  }
  fence_async_smem();   // phase-A results -> visible to the async (TMA) proxy: the early stores below read them
  __syncthreads();   // (B1) derived state, partial sums, raw observation heads, yaw frames are in shared memory
  STAMP(4, 0)
  if (warp > cap && lane == 0) {   // the spare warps behind the assembly warp: outputs that phase B does not write any more
    for (int i = warp - cap - 1; i < L.n_out_early; i += kTaskWarps - 1) {
      const FastCopy d = L.out[i];
      bulk_s2g(static_cast<uint8_t*>(const_cast<void*>(d.g)) + (size_t)env0 * d.bpe, sbuf + d.soff, (uint32_t)(n * d.bpe));
      stores_pending = true;
    }
    if (stores_pending) bulk_commit();
  }
  if (kLoop && chunk + (int)gridDim.x < L.nchunks) issue_loads(chunk + gridDim.x, bsel ^ 1);

  if (row_warp) {
    // =====================================================================================================
    // phase B, ROW warp: terrain scan (legged_robot.py:900-938), height observations, observation-head noise
    // =====================================================================================================
    const int slot = warp;
    float* const orow = s_obs + slot * O;                                  // raw observation head staged in phase A
    float* const gobs = bf.obs_buf + (size_t)(env0 + slot) * O;             // final values go straight to global memory:
    float* const gmh = rollout ? nullptr : bf.measured_heights + (size_t)(env0 + slot) * H;   // 128 contiguous bytes per warp store,
    // issued as each round finishes (no staging, nothing left to drain through the TMA at the end of the launch)
    if (H > 0) {
      float* const mh = SM_F(L.mh) + slot * H;
      const float4 yf = rollout ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : s_yaw[slot];
      const float rootz = rollout ? s_root[slot * 13 + 2] : s_yz[slot];
      const float zc = sub_r(rootz, 0.5f);
      float hv[kNJ];
      if (rollout) {
#pragma unroll
        for (int j = 0; j < kNJ; ++j) hv[j] = mh[min(lane + 32 * j, H - 1)];   // the heights the main step measured (bulk-copied in)
      } else if (heights_live) {
        const float zz = yf.x, ww = yf.y;
        const f32x2 rr = pack2(L.r_hscale, L.r_hscale);
        const f32x2 nc = pack2(-pr.horizontal_scale, -pr.horizontal_scale);
        const f32x2 bord = pack2(pr.border_size, pr.border_size);
        // quat_apply with the yaw-only quaternion (0, 0, zz, ww): t = 2 (q x b) = (-2 zz by, 2 zz bx) (2 RN(x) == RN(2 x), so the
        // doubling is folded into the multiplier), p = b + ww t + q x t with q x t = (-zz ty, zz tx).  The packed halves are
        // TWO POINTS of this lane (not x / y of one point): every multiplier is then a plain broadcast, tx / ty are computed
        // once and serve both coordinates -- 20 packed instructions per two points instead of 26, and no register shuffling.
        const float z2 = mul_r(zz, 2.0f);
        const f32x2 m_tx = pack2(-z2, -z2), m_ty = pack2(z2, z2), m_w = pack2(ww, ww), m_nz = pack2(-zz, -zz), m_pz = pack2(zz, zz);
        const f32x2 X2 = pack2(yf.z, yf.z), Y2 = pack2(yf.w, yf.w);
        const unsigned cols = (unsigned)pr.hf_cols, rmax = (unsigned)(pr.hf_rows - 2), cmax = (unsigned)(pr.hf_cols - 2);
        const float* __restrict__ hmin = bf.height_field_min;
        auto div_h = [&](f32x2 pt) {   // correctly rounded pt / horizontal_scale: q0 = x r, two FMA residual corrections (Markstein)
          f32x2 q = mul2(pt, rr);
          f32x2 er = fma2(nc, q, pt);
          q = fma2(er, rr, q);
          er = fma2(nc, q, pt);
          return fma2(er, rr, q);
        };
#pragma unroll
        for (int k = 0; k < kNJ / 2; ++k) {
          const f32x2 BX = pack2(gx[2 * k], gx[2 * k + 1]), BY = pack2(gy[2 * k], gy[2 * k + 1]);
          const f32x2 TX = mul2(m_tx, BY), TY = mul2(m_ty, BX);
          f32x2 PX = madd2_unfused(BX, m_w, TX);     // bx + ww tx
          PX = madd2_unfused(PX, m_nz, TY);          //    - zz ty
          PX = add2(add2(PX, X2), bord);             //    + base x, + border_size
          f32x2 PY = madd2_unfused(BY, m_w, TY);     // by + ww ty
          PY = madd2_unfused(PY, m_pz, TX);          //    + zz tx
          PY = add2(add2(PY, Y2), bord);
          float qx0, qx1, qy0, qy1;
          unpack2(div_h(PX), qx0, qx1);
          unpack2(div_h(PY), qy0, qy1);
          // .long() truncates toward zero, then clip(0, max): the saturating unsigned conversion already maps everything
          // below 1 (negatives, NaN) to cell 0 and everything too large to UINT_MAX, so one min finishes the clip
          const unsigned ix0 = min(__float2uint_rz(qx0), rmax), iy0 = min(__float2uint_rz(qy0), cmax);
          const unsigned ix1 = min(__float2uint_rz(qx1), rmax), iy1 = min(__float2uint_rz(qy1), cmax);
          hv[2 * k] = __ldg(hmin + (ix0 * cols + iy0));
          hv[2 * k + 1] = __ldg(hmin + (ix1 * cols + iy1));
        }
      } else {
#pragma unroll
        for (int j = 0; j < kNJ; ++j) hv[j] = 0.0f;   // plane terrain (legged_robot.py:913-914)
      }
      STAMP(5, 0)
      float hsum = 0.0f;
#pragma unroll
      for (int j = 0; j < kNJ; ++j) {
        const int p = lane + 32 * j;
        if (j < kNJ - 1 || p < H) {   // H > 32 (kNJ - 1): only the last round is ragged
          const float h = hv[j];
          if (!rollout) gmh[p] = h;
          if (need_hsum) hsum += sub_r(rootz, h);
          float v = mul_r(fminf(fmaxf(sub_r(zc, h), -1.0f), 1.0f), pr.obs_scale_height);
          if (kNoise != ELG_NOISE_OFF) v = v + nz[j];
          if (kClip && (L.flags & 32) == 0) v = fminf(fmaxf(v, -clip_obs), clip_obs);   // (flag 32: the host proved the clip a no-op)
          gobs[head + p] = v;
        }
      }
      if (need_hsum) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) hsum += __shfl_xor_sync(0xffffffffu, hsum, o);
        if (lane == 0) PART(kPHsum, slot) = hsum;
        __threadfence_block();
        named_bar_arrive(kBarHsum, 32 * (n + 1));
      }
    }
    STAMP(6, 0)
    // ---- observation head: noise + clip on the raw entries staged in phase A (legged_robot.py:234-252, :107-108)
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const int k = lane + 32 * m;
      if (k < head) {
        float v = orow[k];
        if (kNoise != ELG_NOISE_OFF) v = v + nzh[m];
        if (kClip) v = fminf(fmaxf(v, -clip_obs), clip_obs);
        gobs[k] = v;
      }
    }
    STAMP(7, 0)
  } else if (warp == cap) {
    // =====================================================================================================
    // phase B, assembly warp (lane == env): termination, the reward registry, ordered sum, episode sums
    // (legged_robot.py:155-160, :215-232; legged_robot_rew_mixin.py:41-234)
    // =====================================================================================================
    const float rootz = s_root[e * 13 + 2];
    const float* s_vec5 = SM_F(L.vec5);
    const float* blv = s_vec5 + (0 * cap + e) * 3;
    const float* bav = s_vec5 + (1 * cap + e) * 3;
    const float* pg = s_vec5 + (2 * cap + e) * 3;
    const float* cmd = SM_F(L.cmd) + e * C;
    const float cmd0 = cmd[0], cmd1 = cmd[1], cmd2 = cmd[2], cmd3 = C > 3 ? cmd[3] : 0.0f;
    const int PT = dm.num_penalised + dm.num_termination;
    bool reset, time_out;
    if (rollout) {   // no check_termination in the rollout step: the termination term reads the flags as they are
      reset = bf.reset_buf[genv] != 0;
      time_out = bf.time_out_buf[genv] != 0;
    } else {
      const int64_t ep = reinterpret_cast<const int64_t*>(sbuf + L.ep)[e];
      const bool contact_term = (PART(kPTermHit, e) + PART(kPTermHit + 1, e) + PART(kPTermHit + 2, e)) != 0.0f;
      time_out = ep > pr.max_episode_length;
      // main / rollout layout (batch_rollout/robot_batch_rollout.py:857-866): time-outs reset the main rows only
      const bool to_resets = pr.rows_per_main <= 0 || genv % pr.rows_per_main == 0;
      reset = contact_term | (time_out & to_resets);
      if (live) {
        bf.reset_buf[genv] = reset ? 1 : 0;
        bf.time_out_buf[genv] = time_out ? 1 : 0;
      }
    }
    const float cmd_xy = norm2_tz(cmd0, cmd1);
    auto dsum = [&](int k) {
      const float* p = &PART(kPDof + kDofWarps * k, e);
      return ((p[0] + p[32]) + (p[64] + p[96])) + (p[128] + p[160]);
    };
    auto fsum = [&](int k) {
      const float* p = &PART(kPFeet + 4 * k, e);
      return (p[0] + p[32]) + (p[64] + p[96]);
    };
    if (gait) {   // GaitScheduler.step (gait_scheduler.py:63-72) runs after the env step
      float* s_gidx = SM_F(L.gidx);
      const float g = s_gidx[e] + pr.gait_increment;
      if (live) s_gidx[e] = g - floorf(g);
    }
    // The registry in alphabetical (= enum) order, straight-line: value -> scaled term -> episode sum -> running fp32 sum
    // (legged_robot.py:220-232).  Row ti of the staged episode sums belongs to the ti-th enabled term.
    float* const s_sums = SM_F(L.sums);
    float total = 0.0f, r_term = 0.0f;
    int ti = 0;
#define SUM_UPDATE(T, r_) if (live && !rollout) s_sums[ti * cap + e] += (r_);
#define TERM(T, VALUE)                                   \
    if (on(pr, T)) {                                       \
      const float r_ = (VALUE) * pr.reward_scales[T];      \
      SUM_UPDATE(T, r_)                                    \
      ++ti;                                                \
      if (T == ELG_REW_TERMINATION) r_term = r_;           \
      else total += r_;                                    \
    }
    TERM(ELG_REW_ACTION_RATE, dsum(kDAr))
    TERM(ELG_REW_ANG_VEL_XY, bav[0] * bav[0] + bav[1] * bav[1])
    TERM(ELG_REW_BASE_FOOT_HEIGHT, ([&] {
           const float cnt = fsum(kFTn);
           const float ground = cnt > 0.0f ? fsum(kFTz) / cnt : rootz - pr.base_height_target;
           const float rel = rootz - ground - pr.base_height_target;
           return rel * rel;
         })())
    TERM(ELG_REW_BASE_HEIGHT, ([&] {
           if (!need_hsum) return 0.0f;
           named_bar_sync(kBarHsum, 32 * (n + 1));   // the row warps' height sums
           const float d = PART(kPHsum, e) / (float)H - pr.base_height_target;
           return d * d;
         })())
    TERM(ELG_REW_COLLISION, PT > 0 ? (PART(kPHits, e) + PART(kPHits + 1, e)) + PART(kPHits + 2, e) : 0.0f)
    TERM(ELG_REW_DOF_ACC, dsum(kDDa))
    TERM(ELG_REW_DOF_POS_LIMITS, dsum(kDPl))
    TERM(ELG_REW_DOF_VEL, dsum(kDDv))
    TERM(ELG_REW_DOF_VEL_LIMITS, dsum(kDVl))
    TERM(ELG_REW_FEET_AIR_TIME, fsum(kFAir) * (cmd_xy > 0.1f ? 1.0f : 0.0f))
    TERM(ELG_REW_FEET_CONTACT_FORCES, fsum(kFCf))
    TERM(ELG_REW_FEET_SLIP, fsum(kFSlip))
    TERM(ELG_REW_FEET_STUMBLE, fsum(kFStum) > 0.0f ? 1.0f : 0.0f)
    TERM(ELG_REW_FEET_STUMBLE_LIFTUP, fsum(kFLift))
    TERM(ELG_REW_FOUR_FOOTUP, fsum(kFDown) == 0.0f ? 0.1f : 0.0f)
    TERM(ELG_REW_GAIT_2_STEP, ([&] {
           // gait_2_step (legged_robot_rew_mixin.py:170-206): FL/RR and FR/RL in phase, the rest anti-phase (updated timers)
           const float4 ar = reinterpret_cast<const float4*>(sbuf + L.air)[e];
           const float4 cn = reinterpret_cast<const float4*>(sbuf + L.con)[e];
           auto sq4 = [](float a, float b) { const float d = a - b; return fminf(d * d, 4.0f); };
           const float s = ((sq4(ar.x, ar.w) + sq4(cn.x, cn.w)) + (sq4(ar.y, ar.z) + sq4(cn.y, cn.z))) / 2.0f;
           const float a = ((sq4(ar.x, cn.y) + sq4(cn.x, ar.y)) + (sq4(ar.x, cn.z) + sq4(cn.x, ar.z)) + (sq4(ar.w, cn.z) + sq4(cn.w, ar.z)) +
                            (sq4(ar.w, cn.y) + sq4(cn.w, ar.y))) / 4.0f;
           const float yawish = pr.heading_command ? cmd3 : cmd2;
           const bool moving = (cmd_xy > pr.speed_min) | (fabsf(yawish) >= pr.speed_min / 2.0f);
           return (s + a) * (moving ? 1.0f : 0.0f);
         })())
    TERM(ELG_REW_GAIT_SCHEDULER, gait ? fsum(kFGs) : 0.0f)
    TERM(ELG_REW_JUMP_AIR, fmaxf(fsum(kFJump) - (float)F / 2.0f, 0.0f))
    TERM(ELG_REW_LIN_VEL_Z, blv[2] * blv[2])
    TERM(ELG_REW_ORIENTATION, pg[0] * pg[0] + pg[1] * pg[1])
    TERM(ELG_REW_STAND_STILL, dsum(kDSs) * (cmd_xy < pr.stand_still_threshold ? 1.0f : 0.0f))
    TERM(ELG_REW_TERMINATION, (reset && !time_out) ? 1.0f : 0.0f)
    TERM(ELG_REW_TORQUE_LIMITS, dsum(kDTl))
    TERM(ELG_REW_TORQUES, dsum(kDTq))
    TERM(ELG_REW_TRACKING_ANG_VEL, ([&] {
           const float ez = cmd2 - bav[2];
           return expf(-(ez * ez) / pr.tracking_sigma);
         })())
    TERM(ELG_REW_TRACKING_LIN_VEL, ([&] {
           const float ex = cmd0 - blv[0], ey = cmd1 - blv[1];
           return expf(-(ex * ex + ey * ey) / pr.tracking_sigma);
         })())
#undef TERM
    if (bf.extra_reward) total += bf.extra_reward[genv];
    if (pr.only_positive_rewards) total = fmaxf(total, 0.0f);
    if (on(pr, ELG_REW_TERMINATION)) total += r_term;
    if (live) SM_F(L.rew)[e] = total;
    if (rollout && live && bf.rollout_rew_out && pr.rows_per_main > 1) {   // column of rollout_batch's reward table
      const int k = genv / pr.rows_per_main, r = genv - k * pr.rows_per_main;
      if (r > 0) bf.rollout_rew_out[((size_t)k * (pr.rows_per_main - 1) + (r - 1)) * pr.rollout_rew_stride] = total;
    }
    STAMP(8, cap)
  }
  // ------------------------------- write back: one cp.async.bulk per output array -------------------------------
  fence_async_smem();   // this thread's generic-proxy writes -> visible to the async (TMA) proxy
  __syncthreads();      // (B2)
  STAMP(9, 0)
  if (lane == 0) {
    bool any = false;
    for (int i = L.n_out_early + warp; i < L.n_out; i += nwarps) {
      const FastCopy d = L.out[i];
      bulk_s2g(static_cast<uint8_t*>(const_cast<void*>(d.g)) + (size_t)env0 * d.bpe, sbuf + d.soff, (uint32_t)(n * d.bpe));
      any = true;
    }
    if (any) {
      bulk_commit();
      stores_pending = true;
    }
  }
#undef SM_F
  }   // chunk loop
#undef PART
  if (stores_pending) bulk_wait_read_all();   // shared memory must outlive the reads; global visibility comes with kernel completion
  STAMP(10, 0)
#undef STAMP
}

}  // namespace elg

// =================================================================================================
// host side: eligibility, shared-memory plan, copy tables, launch (called by elg_post_physics_step)
// =================================================================================================
namespace elg {

// returns 1 when the launch was taken (rc holds the status), 0 when the configuration is not covered by the fast kernel
int launch_step_fast(const ElgDims* dims, const ElgStepParams* prm, const ElgStepBuffers* buf, uint32_t phase, int cap_override,
                     int flags, long long* dbg, void* stream, int* rc) {
  const int N = dims->num_envs, D = dims->num_dof, F = dims->num_feet, B = dims->num_bodies;
  const int H = dims->num_height_points, O = dims->num_obs, C = dims->num_commands;
  const int head = 12 + 3 * D;
  const bool rollout = prm->rollout_mode != 0;
  const uint32_t rollout_phase = ELG_PHASE_DERIVE | ELG_PHASE_REWARD | ELG_PHASE_OBS | ELG_PHASE_HISTORY;   // post_physics_step_rollout
  if (D != 12 || F != 4 || phase != (rollout ? rollout_phase : ELG_PHASE_FUSED)) return 0;
  if (N % 4 != 0 || O != head + H || prm->height_points_env_stride != 0) return 0;
  if (H != 0 && (H <= 32 * (kNJ - 1) || H > 32 * kNJ)) return 0;   // the scan is unrolled for ceil(H / 32) == kNJ
  if (H > 0 && !rollout && !prm->terrain_is_plane && buf->height_field_min == nullptr) return 0;
  if (C > 8 || B > 64) return 0;
  // upside-down termination and the hexapod gait pattern live in the generic kernel (the rollout-mode step has no termination:
  // the robot-specific main / rollout classes keep the lean kernel for their horizon loop)
  if ((prm->terminate_upside_down && !rollout) || prm->gait_2_step_hexapod) return 0;
  const bool gait = buf->gait_idx && buf->gait_prev_foot_z;
  const bool air_on = (prm->reward_mask >> ELG_REW_FEET_AIR_TIME) & 1u;
  const int sms = sm_count();
  if (sms <= 0) return 0;

  // ---- chunking: whole quads of envs, one chunk per CTA; small N: one balanced chunk per SM
  // (measured, scripts/step_large.py: at many chunks per SM the kernel is issue-bound -- smaller CTAs with 3 or 4 resident
  //  per SM are no faster than one 28-env CTA per SM, so every launch uses the widest chunks that balance)
  const long long Q = (long long)N / 4;
  int cap, nchunks;
  if (cap_override >= 4 && cap_override <= kFastMaxCap && cap_override % 4 == 0) {
    cap = cap_override;
    nchunks = (int)((Q + cap / 4 - 1) / (cap / 4));
  } else {
    const long long per = kFastMaxCap / 4;
    long long nch = (Q + per - 1) / per;                 // fewest chunks of <= 28 envs ...
    if (nch < sms) nch = Q < sms ? Q : sms;              // ... but never fewer than one per SM while there are quads to hand out
    nchunks = (int)nch;
    cap = 4 * (int)((Q + nchunks - 1) / nchunks);
  }
  FastPlan L{};
  L.cap = cap;
  L.nchunks = nchunks;
  L.quads_base = (int)(Q / nchunks);
  L.quads_rem = (int)(Q % nchunks);
  L.r_hscale = 1.0f / prm->horizontal_scale;
  L.r_dt = 1.0f / prm->dt;
  {   // contact thresholds on the sum of squares: the largest float whose correctly rounded sqrt does not exceed t
    auto sq_threshold = [](float t) {
      float s = t * t;
      while (sqrtf(nextafterf(s, INFINITY)) <= t) s = nextafterf(s, INFINITY);
      while (sqrtf(s) > t) s = nextafterf(s, -INFINITY);
      return s;
    };
    static const float pen_sq = sq_threshold(0.1f), term_sq = sq_threshold(1.0f);
    L.pen_sq = pen_sq;
    L.term_sq = term_sq;
  }
  L.dbg = dbg;
  L.flags = flags & ~32;
  // height observations are clip(.., -1, 1) * scale + noise with |noise| <= noise_scale: when that bound sits below clip_observations
  // the final clip cannot change a value.  The host vouches for the bound through ElgStepParams.height_obs_bound (0: unknown).
  if (prm->height_obs_bound > 0.0f && prm->height_obs_bound <= prm->clip_observations) L.flags |= 32;
  int nt = 0;
  for (int t = 0; t < ELG_NUM_REWARD_TERMS; ++t)
    if ((prm->reward_mask >> t) & 1u) L.term_ids[nt++] = (int8_t)t;
  L.nterms = nt;
  int off = 0;
  auto take = [&](int bytes) { const int at = off; off += (bytes + 127) & ~127; return at; };
  L.root = take(cap * 52);
  L.dof = take(cap * 8 * D);
  L.act = take(cap * 4 * D);
  L.lact = take(cap * 4 * D);
  L.ldv = take(cap * 4 * D);
  L.tq = take(cap * 4 * D);
  L.cf = take(cap * B * 12);
  L.lrv = take(cap * 24);
  L.vec5 = take(5 * cap * 12);
  L.cmd = take(cap * C * 4);
  L.air = take(cap * 16);
  L.con = take(cap * 16);
  L.lc = take(cap * 4);
  L.ep = take(cap * 8);
  L.gidx = take(cap * 4);
  L.gprev = take(cap * 16);
  L.fpos = take(cap * 48);
  L.fvel = take(cap * 48);
  L.sums = take((nt > 0 ? nt : 1) * cap * 4);
  L.rew = take(cap * 4);
  L.mh = take(cap * (H > 0 ? H : 1) * 4);
  L.obs = take(cap * O * 4);
  L.buf_bytes = off;
  const int grid = nchunks < sms ? nchunks : sms;   // persistent: one CTA per SM, chunks round-robin
  if (nchunks > grid) off *= 2;                     // double-buffered staging
  L.part = take(kPRows * 32 * 4);
  L.yaw = take(cap * 20);
  L.bytes = off;
  if ((size_t)L.bytes + 1024 > (size_t)227 * 1024) return 0;

  int n_in = 0, n_out = 0;
  bool aligned = true;
  auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  auto in = [&](const void* g, int soff, int bpe) {
    L.in[n_in++] = FastCopy{g, soff, bpe};
    aligned = aligned && a16(g);
  };
  auto out = [&](void* g, int soff, int bpe) {
    L.out[n_out++] = FastCopy{g, soff, bpe};
    aligned = aligned && a16(g);
  };
  const int v3 = cap * 12;
  in(buf->root_states, L.root, 52);
  in(buf->dof_state, L.dof, 8 * D);
  in(buf->actions, L.act, 4 * D);
  in(buf->last_actions, L.lact, 4 * D);
  in(buf->last_dof_vel, L.ldv, 4 * D);
  in(buf->torques, L.tq, 4 * D);
  in(buf->contact_forces, L.cf, 12 * B);
  in(buf->last_root_vel, L.lrv, 24);
  in(buf->base_lin_acc, L.vec5 + 3 * v3, 12);
  in(buf->base_ang_acc, L.vec5 + 4 * v3, 12);
  in(buf->commands, L.cmd, 4 * C);
  in(buf->feet_air_time, L.air, 16);
  in(buf->feet_contact_time, L.con, 16);
  in(buf->last_contacts, L.lc, 4);
  if (!rollout) in(buf->episode_length_buf, L.ep, 8);
  if (rollout && H > 0) in(buf->measured_heights, L.mh, 4 * H);
  if (gait) {
    in(buf->gait_idx, L.gidx, 4);
    in(buf->gait_prev_foot_z, L.gprev, 16);
  }
  if (!rollout)
    for (int ti = 0; ti < nt; ++ti) in(buf->episode_sums + (size_t)L.term_ids[ti] * N, L.sums + ti * cap * 4, 4);

  // outputs that are final when phase A ends come first: they are stored right behind barrier B1, under the scan
  out(buf->base_lin_vel, L.vec5 + 0 * v3, 12);
  out(buf->base_ang_vel, L.vec5 + 1 * v3, 12);
  out(buf->projected_gravity, L.vec5 + 2 * v3, 12);
  out(buf->base_lin_acc, L.vec5 + 3 * v3, 12);
  out(buf->base_ang_acc, L.vec5 + 4 * v3, 12);
  out(buf->foot_positions, L.fpos, 48);
  out(buf->foot_velocities, L.fvel, 48);
  if (prm->heading_command && !rollout) out(buf->commands, L.cmd, 4 * C);
  if (!rollout) out(buf->episode_length_buf, L.ep, 8);
  if (air_on) {
    out(buf->feet_air_time, L.air, 16);
    out(buf->feet_contact_time, L.con, 16);
    out(buf->last_contacts, L.lc, 4);
  }
  if (gait) out(buf->gait_prev_foot_z, L.gprev, 16);
  out(buf->last_actions, L.lact, 4 * D);
  out(buf->last_dof_vel, L.ldv, 4 * D);
  out(buf->last_root_vel, L.lrv, 24);
  L.n_out_early = n_out;
  // written by the reward assembly (phase B): stored behind barrier B2
  if (!rollout)
    for (int ti = 0; ti < nt; ++ti) out(buf->episode_sums + (size_t)L.term_ids[ti] * N, L.sums + ti * cap * 4, 4);
  out(buf->rew_buf, L.rew, 4);
  if (gait) out(buf->gait_idx, L.gidx, 4);
  L.n_in = n_in;
  L.n_out = n_out;
  if (!aligned) return 0;   // TMA bulk staging needs 16-byte aligned array bases

  using Kern = void (*)(ElgDims, ElgStepParams, ElgStepBuffers, FastPlan);
  const bool clip = prm->clip_observations > 0.0f;
  Kern kern = nullptr;
  const bool loop = nchunks > grid;
#define ELG_PICK2(NOISE, CLIP)                                                                                          \
  kern = loop ? (rollout ? elg_step_fast_kernel<NOISE, CLIP, true, true> : elg_step_fast_kernel<NOISE, CLIP, true, false>)  \
              : (rollout ? elg_step_fast_kernel<NOISE, CLIP, false, true> : elg_step_fast_kernel<NOISE, CLIP, false, false>)
#define ELG_PICK(NOISE)              \
  if (clip) { ELG_PICK2(NOISE, true); } \
  else { ELG_PICK2(NOISE, false); }
  switch (prm->noise_mode) {
    case ELG_NOISE_OFF: ELG_PICK(ELG_NOISE_OFF); break;
    case ELG_NOISE_TENSOR: ELG_PICK(ELG_NOISE_TENSOR); break;
    default: ELG_PICK(ELG_NOISE_PHILOX); break;
  }
#undef ELG_PICK2
#undef ELG_PICK
  const int which = ((prm->noise_mode * 2 + (clip ? 1 : 0)) * 2 + (loop ? 1 : 0)) * 2 + (rollout ? 1 : 0);
  static SmemCache smem_cache[24] = {};
  size_t& smem_have = smem_slot(smem_cache[which]);
  if ((size_t)L.bytes > smem_have) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.bytes) != cudaSuccess) {
      *rc = set_error(ELG_ERR_CUDA, "cannot reserve dynamic shared memory for elg_step_fast_kernel");
      return 1;
    }
    smem_have = (size_t)L.bytes;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)(32 * (cap + kTaskWarps)));
  cfg.dynamicSmemBytes = (size_t)L.bytes;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (flags & 16) ? 0 : 1;   // experiment switch: plain stream-ordered launch
  if (cudaLaunchKernelEx(&cfg, kern, *dims, *prm, *buf, L) != cudaSuccess) {
    *rc = check_launch("elg_post_physics_step (fast)");
    if (*rc == ELG_OK) *rc = set_error(ELG_ERR_CUDA, "cudaLaunchKernelEx failed for elg_step_fast_kernel");
    return 1;
  }
  *rc = check_launch("elg_post_physics_step (fast)");
  return 1;
}

}  // namespace elg
