// elg_step.cu -- fused post-physics step, PD torques and terrain height scan for sm_100a.
//
// One launch of elg_step_kernel replaces the ~100 ATen launches of
// LeggedRobot.post_physics_step (envs/base/legged_robot.py:113-150 in the reference).
//
// Design (v4, see DESIGN.md "step kernel"):
//   * The environments are cut into CHUNKS of whole quads (4 envs), balanced so that every SM gets
//     the same number of quads to within one: 4096 envs -> 148 chunks of 24..28 envs, one CTA per
//     SM.  Large N: 12-env chunks, 13-warp CTAs, 2 per SM, each looping over its chunks.
//   * TMA in, TMA out.  The envs of a chunk are consecutive, so every per-env array is ONE
//     contiguous global range per chunk.  The host builds two copy tables (global base, shared
//     offset, bytes per env); warp 0 issues one cp.async.bulk (global -> shared, mbarrier
//     completion) per input entry, a lane each, and at the end one cp.async.bulk (shared ->
//     global) per output entry -- including the whole [n, 235] observation block and the [n, 187]
//     height block.  The compute code only touches shared memory.
//   * The step is latency-bound at 28 envs per SM (measured: profiles/README.md), so the work is
//     laid out for SHORT dependent chains: every (env, foot / rotation / dof / contact body) item
//     is a lane of a power-of-two segment and all warps take item passes at once (segmented
//     shuffle butterflies for the per-env sums); then one spare "scalar" warp (lane == env) does
//     commands, termination, the scalar reward terms and the ordered reward sum WHILE the row
//     warps (warp == env) run the 187-point terrain scan.  The scan gathers ONE float per point
//     from a min-of-3 table built once per terrain (elg_prepare_height_field); the x/y halves of
//     the terrain-cell chain use the packed FMUL2/FADD2/FFMA2 of sm_100a, every op individually
//     IEEE-rounded so the cell index stays bit-exact with torch.  Observation noise comes from
//     in-register Philox4x32-10, one 128-bit block per lane per env cut into eight 16-bit samples.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "elg_common.cuh"
#include "elg_async.cuh"

namespace elg {

constexpr int kMaxCap = 32;       // environments per chunk (multiple of 4) == warps per CTA
constexpr int kMaxStepThreads = 1024;
constexpr int kMaxIn = 30 + ELG_NUM_REWARD_TERMS;
constexpr int kMaxOut = 24 + ELG_NUM_REWARD_TERMS;

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
// shared-memory plan of one chunk, built by the host (elg_post_physics_step below).
// Offsets are BYTES from the start of dynamic shared memory; every region starts 128-byte aligned.
// ---------------------------------------------------------------------------------------------
struct CopyDesc {
  const void* g;   // global base of the array (env 0)
  int32_t soff;    // shared-memory byte offset of slot 0
  int32_t bpe;     // bytes per environment
};

struct StepPlan {
  // staged per-env arrays, [slot][per-env]
  int root, dof, act, lact, ldv, tq, cf, lrv, vec5, cmd, air, con, lc, ep, gidx, gprev, fpos, fvel, sums, rew, mh, obs;
  // CTA-constant tables and per-warp scratch
  int q0, plim, vlim, tlim, ns, grid, acc, idx;
  int bytes;
  // launch geometry
  int cap, nchunks, quads_base, quads_rem, use_bulk, obs_smem, nterms, hm;
  int n_in, n_out, in_bpe;
  long long* dbg;   // diagnostic: clock64 stamps of CTA 0 (elg_set_step_debug), or NULL
  int8_t term_ids[ELG_NUM_REWARD_TERMS];
  CopyDesc in[kMaxIn];
  CopyDesc out[kMaxOut];
};

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

// butterfly sums: lane 0 (indeed every lane below the width) ends up with the total of lanes [0, W)
template <int W>
__device__ __forceinline__ float bfly_sum(float v) {
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float lanes_sum(float v, int n) {   // n = number of contributing lanes (others hold 0)
  if (n > 16) return bfly_sum<32>(v);
  if (n > 8) return bfly_sum<16>(v);
  if (n > 4) return bfly_sum<8>(v);
  return bfly_sum<4>(v);
}

// ---------------------------------------------------------------------------------------------
// the fused kernel.  kD / kF > 0 fix the DOF / foot count at compile time (12 / 4 for every
// quadruped of the BASELINE configs); 0 = take them from ElgDims.
//
// Stages of one chunk (n <= 32 envs, W = min(32, n + 1) warps):
//   0  warp 0: one cp.async.bulk per copy-table entry (a lane each) into shared memory; foot lanes prefetch their
//      strided rigid_body_state rows; everybody waits on the chunk's mbarrier.
//   1  ITEM PASSES over all warps: every (env, foot), (env, rotation), (env, dof) and (env, contact body) is one lane
//      in a power-of-two segment (4 / 8 / 16 / 16 lanes per env for a quadruped); per-env reductions are segmented
//      shuffle butterflies, results land in the per-env accumulator table acc[term][env].
//   2  the SCALAR warp (warp n, lane == env) evaluates commands, termination, the scalar reward terms, the ordered
//      fp32 reward sum and the episode sums, while the ROW warps (warp == env) run the terrain scan and the height
//      part of the observation row.
//   3  row warps finish the head of the observation row (noise, clip); warp 0 issues one cp.async.bulk per output entry.
// ---------------------------------------------------------------------------------------------
template <int kD, int kF, bool kFast>
__global__ void __launch_bounds__(kMaxStepThreads, 1)
elg_step_kernel(const __grid_constant__ ElgDims dm, const __grid_constant__ ElgStepParams pr,
                const __grid_constant__ ElgStepBuffers bf, const __grid_constant__ StepPlan L, const uint32_t phase) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_bar;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nthreads = blockDim.x, nwarps = nthreads >> 5;
  const bool dbg_on = L.dbg != nullptr && blockIdx.x == 0;
#define STAMP(i, w) if (dbg_on && warp == (w) && lane == 0) L.dbg[i] = clock64();
  STAMP(0, 0)
  const int N = dm.num_envs, D = kD ? kD : dm.num_dof, B = dm.num_bodies, F = kF ? kF : dm.num_feet;
  const int H = dm.num_height_points, O = dm.num_obs, C = dm.num_commands, P = dm.num_penalised, T = dm.num_termination;
  const int PT = P + T;
  const int cap = L.cap;
  const int head = 12 + 3 * D;
  // kFast: the whole step in one launch (ELG_PHASE_FUSED) with the common data layout -- one [H,3] grid shared by all
  // envs, min-of-3 height table present, observation rows exactly head + H wide, every array TMA-aligned.  The flags
  // below fold at compile time, which keeps the instruction footprint of the hot instantiation inside the
  // instruction caches; everything else runs through the kFast == false instantiation.
  const bool do_derive = kFast || (phase & ELG_PHASE_DERIVE), do_term = kFast || (phase & ELG_PHASE_TERMINATION);
  const bool do_reward = kFast || (phase & ELG_PHASE_REWARD), do_obs = kFast || (phase & ELG_PHASE_OBS);
  const bool do_hist = kFast || (phase & ELG_PHASE_HISTORY);
  const bool need_hsum = do_reward && term_on(pr, ELG_REW_BASE_HEIGHT) && H > 0;
  const bool gait = bf.gait_idx != nullptr && bf.gait_prev_foot_z != nullptr;
  const bool lim_terms = term_on(pr, ELG_REW_DOF_POS_LIMITS) | term_on(pr, ELG_REW_DOF_VEL_LIMITS) | term_on(pr, ELG_REW_TORQUE_LIMITS);
  const bool rollout = !kFast && pr.rollout_mode != 0;
  const bool heights_live = H > 0 && do_derive && !pr.terrain_is_plane && !rollout;
  const bool mh_given = !do_derive || rollout;   // measured_heights is an input of this launch
  const bool shared_grid = kFast || pr.height_points_env_stride == 0;
  const int noise_mode = pr.noise_mode;
  const float clip_obs = pr.clip_observations;
  // lanes per env of the four item types (powers of two, so that a segment never straddles a warp)
  const int LD = kD ? (kD <= 16 ? 16 : 32) : (D <= 16 ? 16 : 32);
  const int LF = kF ? (kF <= 4 ? 4 : 8) : (F <= 4 ? 4 : 8);
  const int LP = PT <= 16 ? 16 : 32;
  constexpr int LR = 8;
  const int lgD = LD == 16 ? 4 : 5, lgF = LF == 4 ? 2 : 3, lgP = LP == 16 ? 4 : 5;

#define SM_F(off) reinterpret_cast<float*>(smem_raw + (off))
  float* const s_root = SM_F(L.root);   float* const s_dof = SM_F(L.dof);    float* const s_act = SM_F(L.act);
  float* const s_lact = SM_F(L.lact);   float* const s_ldv = SM_F(L.ldv);    float* const s_tq = SM_F(L.tq);
  float* const s_cf = SM_F(L.cf);       float* const s_lrv = SM_F(L.lrv);    float* const s_vec5 = SM_F(L.vec5);
  float* const s_cmd = SM_F(L.cmd);     float* const s_air = SM_F(L.air);    float* const s_con = SM_F(L.con);
  uint8_t* const s_lc = smem_raw + L.lc;
  int64_t* const s_ep = reinterpret_cast<int64_t*>(smem_raw + L.ep);
  float* const s_gidx = SM_F(L.gidx);   float* const s_gprev = SM_F(L.gprev);
  float* const s_fpos = SM_F(L.fpos);   float* const s_fvel = SM_F(L.fvel);
  float* const s_sums = SM_F(L.sums);   float* const s_rew = SM_F(L.rew);
  float* const s_mh = SM_F(L.mh);       float* const s_obs = SM_F(L.obs);
  float* const s_q0 = SM_F(L.q0);       float* const s_plim = SM_F(L.plim);  float* const s_vlim = SM_F(L.vlim);
  float* const s_tlim = SM_F(L.tlim);   float* const s_ns = SM_F(L.ns);
  float4* const s_grid = reinterpret_cast<float4*>(smem_raw + L.grid);
  int* const s_feet = reinterpret_cast<int*>(smem_raw + L.idx);   // feet_idx[8] | contact bodies: penalised then termination [24]
  int* const s_body = s_feet + ELG_MAX_FEET;
  float* const s_acc = SM_F(L.acc);     // [kAccRows][32]: raw reward terms by registry id, then helper rows, lane == env
#undef SM_F
#define ACC(t, e) s_acc[(t) * 32 + (e)]
  // helper rows behind the registry ids
  constexpr int kFootZ = ELG_NUM_REWARD_TERMS, kFootN = kFootZ + 1, kHits = kFootZ + 2, kTermHit = kFootZ + 3, kHsum = kFootZ + 4;

  auto feet_of = [&](int f) {   // warp-uniform constant-bank reads, lane-selected (used before the shared table exists)
    int v = 0;
    for (int i = 0; i < F; ++i) v = (f == i) ? dm.feet_idx[i] : v;
    return v;
  };
  auto chunk_range = [&](int chunk, int& env0, int& nenv) {
    const int q_lo = chunk * L.quads_base + min(chunk, L.quads_rem);
    const int q_n = L.quads_base + (chunk < L.quads_rem ? 1 : 0);
    env0 = q_lo * 4;
    nenv = min(N, (q_lo + q_n) * 4) - env0;
  };
  auto issue_loads = [&](int env0, int nenv) {   // every warp: lane 0 issues table entries warp, warp + nwarps, ...
    if (tid == 0) mbar_expect_tx(&s_bar, (uint32_t)(nenv * L.in_bpe));
    if (lane == 0)
      for (int i = warp; i < L.n_in; i += nwarps) {
        const CopyDesc d = L.in[i];
        bulk_g2s(smem_raw + d.soff, static_cast<const uint8_t*>(d.g) + (size_t)env0 * d.bpe, (uint32_t)(nenv * d.bpe), &s_bar);
      }
  };

  int chunk = blockIdx.x;
  int env0 = 0, nenv = 0;
  if (chunk < L.nchunks) chunk_range(chunk, env0, nenv);
  bool bulk = L.use_bulk && (nenv & 3) == 0;
  if (tid == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  if (chunk < L.nchunks && bulk) issue_loads(env0, nenv);

  // ------------------------------- CTA-constant tables (overlap the loads in flight) -------------------------------
  // (kernel parameters live in the constant bank: lane-divergent indexing there is serialised, so the index tables
  //  are copied to shared memory with warp-uniform reads)
  if (warp == nwarps - 1) {
    // one warp, once per CTA, off the critical path (everybody else is waiting for the bulk loads)
    if (lane < F) s_feet[lane] = dm.feet_idx[lane];
    if (lane < P) s_body[lane] = dm.penalised_idx[lane];
    if (lane < T) s_body[P + lane] = dm.termination_idx[lane];
  }
#pragma unroll 1
  for (int j = tid; j < D; j += nthreads) {
    s_q0[j] = __ldg(bf.default_dof_pos + j);
    if (lim_terms) {
      s_plim[2 * j] = __ldg(bf.dof_pos_limits + 2 * j);
      s_plim[2 * j + 1] = __ldg(bf.dof_pos_limits + 2 * j + 1);
      s_vlim[j] = __ldg(bf.dof_vel_limits + j) * pr.soft_dof_vel_limit;
      s_tlim[j] = __ldg(bf.torque_limits + j) * pr.soft_torque_limit;
    }
  }
  {
    const bool ns_on = bf.noise_scale_vec != nullptr && noise_mode != ELG_NOISE_OFF;
#pragma unroll 1
    for (int k = tid; k < O; k += nthreads) s_ns[k] = ns_on ? __ldg(bf.noise_scale_vec + k) : 0.0f;
  }
  if (H > 0 && shared_grid && bf.height_points)
#pragma unroll 1
    for (int p = tid; p < H; p += nthreads) {
      const float bx = __ldg(bf.height_points + 3 * p), by = __ldg(bf.height_points + 3 * p + 1);
      s_grid[p] = make_float4(bx, by, by, bx);
    }
  STAMP(1, 0)
  __syncthreads();   // tables + mbarrier initialisation visible to every warp
  STAMP(2, 0)

  uint32_t parity = 0;
  bool stores_pending = false;   // lane 0 of every warp: its bulk store group may still be reading shared memory

  for (; chunk < L.nchunks; chunk += gridDim.x) {
    if (chunk != (int)blockIdx.x) {
      chunk_range(chunk, env0, nenv);
      bulk = L.use_bulk && (nenv & 3) == 0;
      if (stores_pending) bulk_wait_read_all();   // this thread's previous stores have left shared memory ...
      stores_pending = false;
      __syncthreads();                              // ... and so have everybody else's
      if (bulk) issue_loads(env0, nenv);
    }
    const int n = nenv;
    // item-pass geometry: [foot | rotation | dof | body] passes, one warp each
    const int wF = (F > 0) ? ((n << lgF) + 31) >> 5 : 0;
    const int wR = (n * LR + 31) >> 5;
    const int wD = ((n << lgD) + 31) >> 5;
    const int wP = (PT > 0) ? ((n << lgP) + 31) >> 5 : 0;
    const int n_passes = wF + wR + wD + wP;

    // ---- stage 0: foot lanes prefetch their rigid_body_state rows (52-byte rows, 6 useful floats) while the bulk
    // copies are in flight; kept in registers until the wait below (shared memory may still feed the previous stores)
    float fr[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    const int f_e = (warp * 32 + lane) >> lgF, f_f = lane & (LF - 1);
    const bool foot_lane = warp < wF && f_e < n && f_f < F;
    if (foot_lane && do_derive) {
      const float* row = bf.rigid_body_state + ((size_t)(env0 + f_e) * B + feet_of(f_f)) * 13;
      fr[0] = __ldg(row + 0); fr[1] = __ldg(row + 1); fr[2] = __ldg(row + 2);
      fr[3] = __ldg(row + 7); fr[4] = __ldg(row + 8); fr[5] = __ldg(row + 9);
    }
    if (kFast || bulk) {
      mbar_wait(&s_bar, parity);
      parity ^= 1u;
    } else {
      for (int i = 0; i < L.n_in; ++i) {
        const CopyDesc d = L.in[i];
        coop_copy(smem_raw + d.soff, static_cast<const uint8_t*>(d.g) + (size_t)env0 * d.bpe, (uint32_t)(nenv * d.bpe), tid, nthreads);
      }
      __syncthreads();
    }

    STAMP(3, 0)
    // =========================== stage 1: item passes ===========================
    for (int wp = warp; wp < n_passes; wp += nwarps) {
      if (wp < wF) {
        // ---- (env, foot) (legged_robot_rew_mixin.py:58-81, :121-212; gait_scheduler.py:74-81)
        // Terms that sort before feet_air_time read the OLD timers, terms after it the updated ones
        // and the rebound last_contacts.
        const int e = (wp * 32 + lane) >> lgF, f = lane & (LF - 1);
        const bool ok = e < n && f < F;
        const int fi = e * F + f;
        float f_tz = 0.0f, f_tn = 0.0f, f_air = 0.0f, f_cf = 0.0f, f_slip = 0.0f, f_lift = 0.0f, f_jump = 0.0f, f_stum = 0.0f,
              f_down = 0.0f, f_gs = 0.0f;
        const bool air_on = term_on(pr, ELG_REW_FEET_AIR_TIME);
        const bool gs_on = term_on(pr, ELG_REW_GAIT_SCHEDULER) && gait;
        if (ok) {
          if (do_derive) {
            if (wp != warp) {   // not the prefetched pass (only for > 32 * 32 / LF feet lanes): load now
              const float* row = bf.rigid_body_state + ((size_t)(env0 + e) * B + s_feet[f]) * 13;
              fr[0] = __ldg(row + 0); fr[1] = __ldg(row + 1); fr[2] = __ldg(row + 2);
              fr[3] = __ldg(row + 7); fr[4] = __ldg(row + 8); fr[5] = __ldg(row + 9);
            }
            s_fpos[fi * 3] = fr[0]; s_fpos[fi * 3 + 1] = fr[1]; s_fpos[fi * 3 + 2] = fr[2];
            s_fvel[fi * 3] = fr[3]; s_fvel[fi * 3 + 1] = fr[4]; s_fvel[fi * 3 + 2] = fr[5];
          } else if (do_reward) {
            fr[2] = s_fpos[fi * 3 + 2];
            fr[3] = s_fvel[fi * 3]; fr[4] = s_fvel[fi * 3 + 1]; fr[5] = s_fvel[fi * 3 + 2];
          }
          if (do_reward) {
            const float* cf = s_cf + (e * B + s_feet[f]) * 3;
            const float fxx = cf[0], fyy = cf[1], fz = cf[2];
            const float pz = fr[2], vx = fr[3], vy = fr[4], vz = fr[5];
            float air = s_air[fi], con = s_con[fi];
            const bool last_c = s_lc[fi] != 0;
            const bool contact = fz > 1.0f;
            const bool touching = con > 1e-3f;   // base_foot_height: nanmean over touching feet (old timers)
            f_tz = touching ? pz : 0.0f;
            f_tn = touching ? 1.0f : 0.0f;
            bool lc_after = last_c;
            if (air_on) {
              const bool filt = contact | last_c;
              const bool first = (air > 0.0f) && filt;
              air += pr.dt;
              con += pr.dt;
              f_air = (air - 0.5f) * (first ? 1.0f : 0.0f);
              air *= filt ? 0.0f : 1.0f;
              con *= filt ? 1.0f : 0.0f;
              s_air[fi] = air;
              s_con[fi] = con;
              s_lc[fi] = contact ? 1 : 0;
              lc_after = contact;
            }
            const bool filt2 = contact | lc_after;
            f_cf = fmaxf(norm3_t(fxx, fyy, fz) - pr.max_contact_force, 0.0f);
            const float vn = norm2_t(vx, vy);
            f_slip = (filt2 ? 1.0f : 0.0f) * (vn * vn);
            const bool stumble = norm2_t(fxx, fyy) > mul_r(5.0f, fabsf(fz));
            f_lift = (stumble ? 1.0f : 0.0f) * vz;
            f_jump = (filt2 ? 0.0f : 1.0f) * (air - 0.5f);
            f_stum = stumble ? 1.0f : 0.0f;
            f_down = fz < 1.0f ? 0.0f : 1.0f;   // number of feet that are NOT up
            if (gs_on) {
              float ph = s_gidx[e] + pr.gait_foot_phases[f];
              ph = ph - floorf(ph);                                 // torch.remainder(x, 1.0)
              const float target = ph < 0.5f ? pr.gait_swing_height * sinf(6.283185307179586f * ph) : 0.0f;
              const float dz = target - s_gprev[fi];
              f_gs = dz * dz;
            }
            if (gait) s_gprev[fi] = pz;   // GaitScheduler.step keeps this step's feet
          }
        }
        if (do_reward) {
          const bool lead = ok && f == 0;
#define FOOT_RED(on, var, row)                                   \
  if (on) {                                                      \
    var = (LF == 4) ? bfly_sum<4>(var) : bfly_sum<8>(var);       \
    if (lead) ACC(row, e) = var;                                 \
  }
          FOOT_RED(term_on(pr, ELG_REW_BASE_FOOT_HEIGHT), f_tz, kFootZ)
          FOOT_RED(term_on(pr, ELG_REW_BASE_FOOT_HEIGHT), f_tn, kFootN)
          FOOT_RED(air_on, f_air, ELG_REW_FEET_AIR_TIME)
          FOOT_RED(term_on(pr, ELG_REW_FEET_CONTACT_FORCES), f_cf, ELG_REW_FEET_CONTACT_FORCES)
          FOOT_RED(term_on(pr, ELG_REW_FEET_SLIP), f_slip, ELG_REW_FEET_SLIP)
          FOOT_RED(term_on(pr, ELG_REW_FEET_STUMBLE_LIFTUP), f_lift, ELG_REW_FEET_STUMBLE_LIFTUP)
          FOOT_RED(term_on(pr, ELG_REW_JUMP_AIR), f_jump, ELG_REW_JUMP_AIR)
          FOOT_RED(term_on(pr, ELG_REW_FEET_STUMBLE), f_stum, ELG_REW_FEET_STUMBLE)
          FOOT_RED(term_on(pr, ELG_REW_FOUR_FOOTUP), f_down, ELG_REW_FOUR_FOOTUP)
          FOOT_RED(gs_on, f_gs, ELG_REW_GAIT_SCHEDULER)
#undef FOOT_RED
        }
      } else if (wp < wF + wR) {
        // ---- (env, rotation): base-frame velocities, gravity, acceleration EMAs (:128-134); root-velocity history (:150)
        const int e = ((wp - wF) * 32 + lane) >> 3, r = lane & (LR - 1);
        const bool ok = e < n;
        const float* rs = s_root + e * 13;
        if (ok && do_derive && r < 5) {
          const Quat q = {rs[3], rs[4], rs[5], rs[6]};
          Vec3 v;
          if (r == 2) {
            v = Vec3{pr.gravity_vec[0], pr.gravity_vec[1], pr.gravity_vec[2]};
          } else {
            const int k = (r == 1 || r == 4) ? 10 : 7;
            v = Vec3{rs[k], rs[k + 1], rs[k + 2]};
            if (r >= 3) {
              const float* lrv = s_lrv + e * 6 + (r - 3) * 3;
              v.x -= lrv[0]; v.y -= lrv[1]; v.z -= lrv[2];
            }
          }
          Vec3 o = quat_rotate_inverse(q, v);
          float* dst = s_vec5 + (r * cap + e) * 3;
          if (r >= 3) {
            const float ema = pr.acc_ema, w1 = pr.acc_ema_c;
            o.x = dst[0] * ema + (w1 * o.x) / pr.dt;
            o.y = dst[1] * ema + (w1 * o.y) / pr.dt;
            o.z = dst[2] * ema + (w1 * o.z) / pr.dt;
          }
          dst[0] = o.x; dst[1] = o.y; dst[2] = o.z;
        }
        __syncwarp();
        if (ok && do_hist && r < 6) s_lrv[e * 6 + r] = rs[7 + r];
        if (ok && do_derive && !rollout && r == 7) s_ep[e] += 1;   // episode counter (legged_robot.py:122)
      } else if (wp < wF + wR + wD) {
        // ---- (env, dof): per-dof reward partials, observation entries, history (:84-114, :237-244, :148-149)
        const int e = ((wp - wF - wR) * 32 + lane) >> lgD, j = lane & (LD - 1);
        const bool ok = e < n && j < D;
        float q_ar = 0.0f, q_da = 0.0f, q_dv = 0.0f, q_tq = 0.0f, q_ss = 0.0f, q_pl = 0.0f, q_vl = 0.0f, q_tl = 0.0f;
        if (ok) {
          const int fi = e * D + j;
          const float2 pv = *reinterpret_cast<const float2*>(s_dof + 2 * fi);
          const float pos = pv.x, vel = pv.y;
          const float a = s_act[fi], q0 = s_q0[j];
          if (do_reward) {
            const float la = s_lact[fi], lv = s_ldv[fi], tq = s_tq[fi];
            const float da = la - a;
            q_ar = da * da;
            const float dv = (lv - vel) / pr.dt;
            q_da = dv * dv;
            q_dv = vel * vel;
            q_tq = tq * tq;
            q_ss = fabsf(pos - q0);
            if (lim_terms) {
              q_pl = -fminf(pos - s_plim[2 * j], 0.0f) + fmaxf(pos - s_plim[2 * j + 1], 0.0f);
              q_vl = fminf(fmaxf(fabsf(vel) - s_vlim[j], 0.0f), 1.0f);
              q_tl = fmaxf(fabsf(tq) - s_tlim[j], 0.0f);
            }
          }
          if (do_obs) {
            float* hrow = s_obs + e * (L.obs_smem ? O : head);
            hrow[12 + j] = (pos - q0) * pr.obs_scale_dof_pos;
            hrow[12 + D + j] = vel * pr.obs_scale_dof_vel;
            hrow[12 + 2 * D + j] = a;
          }
          if (do_hist) {
            s_lact[fi] = a;
            s_ldv[fi] = vel;
          }
        }
        if (do_reward) {
          const bool lead = e < n && j == 0;
#define DOF_RED(t, var)                                            \
  if (term_on(pr, t)) {                                            \
    var = (LD == 16) ? bfly_sum<16>(var) : bfly_sum<32>(var);      \
    if (lead) ACC(t, e) = var;                                     \
  }
          DOF_RED(ELG_REW_ACTION_RATE, q_ar)
          DOF_RED(ELG_REW_DOF_ACC, q_da)
          DOF_RED(ELG_REW_DOF_VEL, q_dv)
          DOF_RED(ELG_REW_TORQUES, q_tq)
          DOF_RED(ELG_REW_STAND_STILL, q_ss)
          DOF_RED(ELG_REW_DOF_POS_LIMITS, q_pl)
          DOF_RED(ELG_REW_DOF_VEL_LIMITS, q_vl)
          DOF_RED(ELG_REW_TORQUE_LIMITS, q_tl)
#undef DOF_RED
        }
      } else {
        // ---- (env, body): collision count and termination contacts (:117-119, legged_robot.py:155-160)
        const int e = ((wp - wF - wR - wD) * 32 + lane) >> lgP, b = lane & (LP - 1);
        bool hit = false;
        if (e < n && b < PT && (do_term || do_reward)) {
          const int body = s_body[b];
          const float* f = s_cf + (e * B + body) * 3;
          hit = norm3_t(f[0], f[1], f[2]) > (b < P ? 0.1f : 1.0f);
        }
        unsigned m = __ballot_sync(0xffffffffu, hit);
        if (LP == 16) m = (m >> (lane & 16)) & 0xffffu;
        if (e < n && b == 0) {
          ACC(kHits, e) = (float)__popc(m & ((1u << P) - 1u));
          ACC(kTermHit, e) = (m >> P) != 0u ? 1.0f : 0.0f;
        }
      }
    }
    STAMP(4, 0)
    __syncthreads();   // (B1) derived state, accumulators, raw observation heads are in shared memory
    STAMP(5, 0)

    const int ewarp = (n < nwarps) ? n : 0;   // the scalar warp: a spare warp when there is one
    const int slot = warp;
    const bool row_warp = slot < n;
    const int env = env0 + slot;
    float hsum = 0.0f;
    uint4 blk = make_uint4(0, 0, 0, 0);
    int blk_id = -1;
    const int nj = (H + 31) >> 5;
    const bool philox = do_obs && noise_mode == ELG_NOISE_PHILOX;
    const uint64_t noise_off = pr.noise_offset + (bf.step_counter ? *bf.step_counter : 0ull);
    const bool obs_to_smem = kFast || L.obs_smem;
    float* const orow = s_obs + slot * (obs_to_smem ? O : head);   // staged observation row (head only when rows are user-extended)
    float* const grow = bf.obs_buf + (size_t)env * O;

    // =========================== stage 2a: scalar warp, lane == env ===========================
    bool scalar_pending = warp == ewarp && do_reward | do_obs | do_term | do_derive;
    auto scalar_stage = [&]() {
      const int e = lane;
      if (e >= n) return;
      const int genv = env0 + e;
      const float* rs = s_root + e * 13;
      const float rootz = rs[2];
      const float* blv = s_vec5 + (0 * cap + e) * 3;
      const float* bav = s_vec5 + (1 * cap + e) * 3;
      const float* pg = s_vec5 + (2 * cap + e) * 3;
      float* cmd = s_cmd + e * C;
      float cmd2 = cmd[2];
      if (do_derive && !rollout && pr.heading_command) {   // (legged_robot.py:394-398); forward = quat_apply(q, (1,0,0))
        const Quat q = {rs[3], rs[4], rs[5], rs[6]};
        const float fx = 1.0f + (q.y * (-2.0f * q.y) - q.z * (2.0f * q.z));
        const float fy = q.w * (2.0f * q.z) + (q.z * 0.0f - q.x * (-2.0f * q.y));
        const float heading = atan2f(fy, fx);
        cmd2 = fminf(fmaxf(0.5f * wrap_to_pi(cmd[3] - heading), -1.0f), 1.0f);
        cmd[2] = cmd2;
      }
      const float cmd0 = cmd[0], cmd1 = cmd[1], cmd3 = C > 3 ? cmd[3] : 0.0f;
      bool reset = false, time_out = false;
      if (do_term) {
        const bool contact_term = PT > 0 && ACC(kTermHit, e) != 0.0f;
        time_out = s_ep[e] > pr.max_episode_length;
        // main / rollout layout (batch_rollout/robot_batch_rollout.py:857-866): time-outs reset the main rows only
        const bool to_resets = pr.rows_per_main <= 0 || genv % pr.rows_per_main == 0;
        // upside-down robots: every row (1: elspider.py:340-345, elspider_air_batch_rollout.py:176) or the main rows of the
        // main / rollout layout only (2: anymal_c_batch_rollout.py:192-199, go2_batch_rollout.py:200)
        const bool upside_down = pr.terminate_upside_down != 0 && pg[2] > 0.0f && (pr.terminate_upside_down != 2 || to_resets);
        reset = contact_term | (time_out & to_resets) | upside_down;
        bf.reset_buf[genv] = reset ? 1 : 0;
        bf.time_out_buf[genv] = time_out ? 1 : 0;
      } else if (do_reward) {
        reset = bf.reset_buf[genv] != 0;
        time_out = bf.time_out_buf[genv] != 0;
      }
      if (do_obs) {
        float* hrow = s_obs + e * (obs_to_smem ? O : head);
        hrow[0] = blv[0] * pr.obs_scale_lin_vel; hrow[1] = blv[1] * pr.obs_scale_lin_vel; hrow[2] = blv[2] * pr.obs_scale_lin_vel;
        hrow[3] = bav[0] * pr.obs_scale_ang_vel; hrow[4] = bav[1] * pr.obs_scale_ang_vel; hrow[5] = bav[2] * pr.obs_scale_ang_vel;
        hrow[6] = pg[0]; hrow[7] = pg[1]; hrow[8] = pg[2];
        hrow[9] = cmd0 * pr.commands_scale[0]; hrow[10] = cmd1 * pr.commands_scale[1]; hrow[11] = cmd2 * pr.commands_scale[2];
      }
      if (!do_reward) return;
      const float cmd_xy = norm2_t(cmd0, cmd1);
      if (term_on(pr, ELG_REW_STAND_STILL)) ACC(ELG_REW_STAND_STILL, e) *= (cmd_xy < pr.stand_still_threshold ? 1.0f : 0.0f);
      if (term_on(pr, ELG_REW_LIN_VEL_Z)) ACC(ELG_REW_LIN_VEL_Z, e) = blv[2] * blv[2];
      if (term_on(pr, ELG_REW_ANG_VEL_XY)) ACC(ELG_REW_ANG_VEL_XY, e) = bav[0] * bav[0] + bav[1] * bav[1];
      if (term_on(pr, ELG_REW_ORIENTATION)) ACC(ELG_REW_ORIENTATION, e) = pg[0] * pg[0] + pg[1] * pg[1];
      if (term_on(pr, ELG_REW_TRACKING_LIN_VEL)) {
        const float ex = cmd0 - blv[0], ey = cmd1 - blv[1];
        ACC(ELG_REW_TRACKING_LIN_VEL, e) = expf(-(ex * ex + ey * ey) / pr.tracking_sigma);
      }
      if (term_on(pr, ELG_REW_TRACKING_ANG_VEL)) {
        const float ez = cmd2 - bav[2];
        ACC(ELG_REW_TRACKING_ANG_VEL, e) = expf(-(ez * ez) / pr.tracking_sigma);
      }
      if (term_on(pr, ELG_REW_TERMINATION)) ACC(ELG_REW_TERMINATION, e) = (reset && !time_out) ? 1.0f : 0.0f;
      if (term_on(pr, ELG_REW_COLLISION)) ACC(ELG_REW_COLLISION, e) = PT > 0 ? ACC(kHits, e) : 0.0f;
      if (term_on(pr, ELG_REW_BASE_FOOT_HEIGHT)) {
        const float cnt = ACC(kFootN, e);
        const float ground = cnt > 0.0f ? ACC(kFootZ, e) / cnt : rootz - pr.base_height_target;
        const float rel = rootz - ground - pr.base_height_target;
        ACC(ELG_REW_BASE_FOOT_HEIGHT, e) = rel * rel;
      }
      if (term_on(pr, ELG_REW_FEET_AIR_TIME)) ACC(ELG_REW_FEET_AIR_TIME, e) *= (cmd_xy > 0.1f ? 1.0f : 0.0f);
      if (term_on(pr, ELG_REW_JUMP_AIR)) ACC(ELG_REW_JUMP_AIR, e) = fmaxf(ACC(ELG_REW_JUMP_AIR, e) - (float)F / 2.0f, 0.0f);
      if (term_on(pr, ELG_REW_FEET_STUMBLE)) ACC(ELG_REW_FEET_STUMBLE, e) = ACC(ELG_REW_FEET_STUMBLE, e) > 0.0f ? 1.0f : 0.0f;
      if (term_on(pr, ELG_REW_FOUR_FOOTUP)) ACC(ELG_REW_FOUR_FOOTUP, e) = ACC(ELG_REW_FOUR_FOOTUP, e) == 0.0f ? 0.1f : 0.0f;
      if (term_on(pr, ELG_REW_GAIT_SCHEDULER) && !gait) ACC(ELG_REW_GAIT_SCHEDULER, e) = 0.0f;
      if (term_on(pr, ELG_REW_GAIT_2_STEP)) {
        const float* ar = s_air + e * F;
        const float* cn = s_con + e * F;
        auto sq4 = [](float a, float b) { const float d = a - b; return fminf(d * d, 4.0f); };
        auto sync = [&](int i, int j) { return sq4(ar[i], ar[j]) + sq4(cn[i], cn[j]); };
        auto anti = [&](int i, int j) { return sq4(ar[i], cn[j]) + sq4(cn[i], ar[j]); };
        float s, a;
        if (pr.gait_2_step_hexapod) {
          // ElSpider (elspider.py:365-408): feet LB LF LM RB RF RM; tripods (0,1,5) and (2,3,4) in phase, anti-phase across
          const float g1 = ((sync(0, 1) + sync(0, 5)) + sync(1, 5)) / 3.0f;
          const float g2 = ((sync(2, 3) + sync(2, 4)) + sync(3, 4)) / 3.0f;
          a = ((((((((anti(0, 2) + anti(0, 3)) + anti(0, 4)) + anti(1, 2)) + anti(1, 3)) + anti(1, 4)) + anti(5, 2)) + anti(5, 3)) + anti(5, 4)) / 9.0f;
          s = (g1 + g2) / 2.0f;
        } else {
          // gait_2_step (legged_robot_rew_mixin.py:170-206): FL/RR and FR/RL in phase, the rest anti-phase
          s = (sync(0, 3) + sync(1, 2)) / 2.0f;
          a = (((anti(0, 1) + anti(0, 2)) + anti(3, 2)) + anti(3, 1)) / 4.0f;
        }
        const float yawish = pr.heading_command ? cmd3 : cmd2;
        const bool moving = (cmd_xy > pr.speed_min) | (fabsf(yawish) >= pr.speed_min / 2.0f);
        ACC(ELG_REW_GAIT_2_STEP, e) = (s + a) * (moving ? 1.0f : 0.0f);
      }
      if (term_on(pr, ELG_REW_BASE_HEIGHT)) {
        const float d = (need_hsum ? ACC(kHsum, e) / (float)H : 0.0f) - pr.base_height_target;
        ACC(ELG_REW_BASE_HEIGHT, e) = need_hsum ? d * d : 0.0f;
      }
      if (gait) {   // GaitScheduler.step (gait_scheduler.py:63-72) runs after the env step
        const float g = s_gidx[e] + pr.gait_increment;
        s_gidx[e] = g - floorf(g);
      }
      // scaled terms + episode sums, then the fp32 sum in registry (alphabetical) order (legged_robot.py:220-232)
      float total = 0.0f, r_term = 0.0f;
      for (int ti = 0; ti < L.nterms; ++ti) {
        const int t = L.term_ids[ti];
        const float r = ACC(t, e) * pr.reward_scales[t];
        if (!rollout) s_sums[ti * cap + e] += r;
        if (t == ELG_REW_TERMINATION) r_term = r;   // added after the clip
        else total += r;
      }
      if (bf.extra_reward) total += bf.extra_reward[genv];
      if (pr.only_positive_rewards) total = fmaxf(total, 0.0f);
      if (term_on(pr, ELG_REW_TERMINATION)) total += r_term;
      s_rew[e] = total;
      if (rollout && bf.rollout_rew_out && pr.rows_per_main > 1) {   // column of rollout_batch's reward table (robot_traj_grad_sampling.py:262-266)
        const int k = genv / pr.rows_per_main, r = genv - k * pr.rows_per_main;
        if (r > 0) bf.rollout_rew_out[((size_t)k * (pr.rows_per_main - 1) + (r - 1)) * pr.rollout_rew_stride] = total;
      }
    };
#pragma unroll 1
    for (int round = 0; round < 2; ++round) {
      if (scalar_pending && (round == 1 || !need_hsum)) { scalar_stage(); scalar_pending = false; }
      STAMP(6 + 6 * round, ewarp)
      if (round == 1) break;

    // =========================== stage 2b: row warps -- terrain scan (legged_robot.py:900-938) + height observations ===========================
    // Height point p = lane + 32 j.  Noise: 16-bit samples, 8 per Philox block; point j uses sample j % 8 of block
    // 1 + j / 8 of this lane; the head entries (below) use the free samples of the last height block when they fit.
      if (row_warp && H > 0 && (do_derive || do_obs || need_hsum)) {
      const float* rs = s_root + slot * 13;
      const float rootz = rs[2];
      const float zc = sub_r(rootz, 0.5f);
      float* mh = s_mh + slot * H;
      const int16_t* __restrict__ hs = bf.height_samples;
      const float* __restrict__ hmin = bf.height_field_min;
      const int cols = pr.hf_cols, rmax = pr.hf_rows - 2, cmax = pr.hf_cols - 2;
      f32x2 rr = 0, nc = 0, bord = 0, c_t = 0, c_ts = 0, c_w = 0, c_u = 0, xy = 0;
      if (heights_live) {
        const float r_h = __frcp_rn(pr.horizontal_scale);
        rr = pack2(r_h, r_h);
        nc = pack2(-pr.horizontal_scale, -pr.horizontal_scale);
        bord = pack2(pr.border_size, pr.border_size);
        float nrm = __fsqrt_rn(add_r(mul_r(rs[5], rs[5]), mul_r(rs[6], rs[6])));
        nrm = fmaxf(nrm, 1e-9f);
        const float zz = div_r(rs[5], nrm), ww = div_r(rs[6], nrm);
        // t = 2 (q x b) = (-2 zz by, 2 zz bx); 2*RN(x) == RN(2x), so the doubling is folded into the multiplier
        c_t = pack2(-mul_r(zz, 2.0f), mul_r(zz, 2.0f));    // times (by, bx) -> (tx, ty)
        c_ts = pack2(mul_r(zz, 2.0f), -mul_r(zz, 2.0f));   // times (bx, by) -> (ty, tx)
        c_w = pack2(ww, ww);
        c_u = pack2(-zz, zz);
        xy = pack2(rs[0], rs[1]);
      }
      const float* hp_env = shared_grid ? nullptr : bf.height_points + (size_t)env * pr.height_points_env_stride;
      auto cell_of = [&](int p) -> int {
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
        f32x2 pt = madd2_unfused(b, c_w, t);    // b + w t
        pt = madd2_unfused(pt, c_u, ts);        // + q_xyz x t = (-zz ty, zz tx)
        pt = add2(add2(pt, xy), bord);          // + base xy, + border_size
        // correctly rounded pt / horizontal_scale: q0 = x r, two FMA residual corrections (Markstein)
        f32x2 q = mul2(pt, rr);
        f32x2 er = fma2(nc, q, pt);
        q = fma2(er, rr, q);
        er = fma2(nc, q, pt);
        q = fma2(er, rr, q);
        float qx, qy;
        unpack2(q, qx, qy);
        int ix = __float2int_rz(qx), iy = __float2int_rz(qy);
        ix = min(max(ix, 0), rmax);
        iy = min(max(iy, 0), cmax);
        return ix * cols + iy;
      };
      for (int jb = 0; jb < nj; jb += 8) {
        if (philox) {
          blk_id = 1 + (jb >> 3);
          blk = noise_block(pr.noise_seed, noise_off, env, lane, blk_id);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int j0 = jb + 4 * half;
          if (j0 >= nj) break;
          float hv[4];
          if (heights_live) {
            if (kFast || hmin) {            // precomputed fp32 min-of-3 field (elg_prepare_height_field): one gather per point
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int p = lane + 32 * (j0 + u);
                hv[u] = (p < H) ? __ldg(hmin + cell_of(p)) : 0.0f;
              }
            } else {
              int a0[4], a1[4], a2[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int p = lane + 32 * (j0 + u);
                a0[u] = a1[u] = a2[u] = 0;
                if (p < H) {
                  const int16_t* cell = hs + cell_of(p);
                  a0[u] = __ldg(cell);
                  a1[u] = __ldg(cell + cols);
                  a2[u] = __ldg(cell + 1);
                }
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) hv[u] = mul_r((float)min(min(a0[u], a1[u]), a2[u]), pr.vertical_scale);
            }
          } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int p = lane + 32 * (j0 + u);
              hv[u] = (mh_given && p < H) ? mh[p] : 0.0f;   // plane terrain: zeros (legged_robot.py:913-914)
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int p = lane + 32 * (j0 + u);
            if (p < H) {
              const float h = hv[u];
              if (!mh_given) mh[p] = h;
              hsum += sub_r(rootz, h);
              if (do_obs) {
                const int k = head + p;
                float v = mul_r(fminf(fmaxf(sub_r(zc, h), -1.0f), 1.0f), pr.obs_scale_height);
                if (noise_mode != ELG_NOISE_OFF) {
                  float t;
                  if (philox) t = fmaf(sample16(blk, 4 * half + u), 1.0f / 32768.0f, -1.0f);       // 2u - 1, exact
                  else t = 2.0f * __ldg(bf.noise_u + (size_t)env * O + k) - 1.0f;
                  v = v + t * s_ns[k];
                }
                if (clip_obs > 0.0f) v = fminf(fmaxf(v, -clip_obs), clip_obs);
                if (kFast || obs_to_smem) orow[k] = v;
                else grow[k] = v;
              }
            }
          }
        }
      }
      if (need_hsum) {
        hsum = bfly_sum<32>(hsum);
        if (lane == 0) ACC(kHsum, slot) = hsum;
      }
    }
      STAMP(7, 0)
      __syncthreads();   // (B2) terrain scan done (and the scalar stage, unless it needs the height sums)
      STAMP(8, 0)
    }
    if (need_hsum) __syncthreads();

    // =========================== stage 3: observation head: noise + clip (legged_robot.py:234-252, :107-108) ===========================
    if (row_warp && do_obs) {
      const int hm = L.hm;
      // head entry k = lane + 32 m uses sample (nj % 8) + m of the last height block when hm fits behind the height samples
      const bool share = (nj & 7) + hm <= 8;
      const int bid = share ? 1 + (nj >> 3) : 0;
      const int s0 = share ? (nj & 7) : 0;
      for (int m = 0; m < hm; ++m) {
        const int k = lane + 32 * m;
        if (philox && ((m & 7) == 0) && !(m == 0 && blk_id == bid)) {
          blk_id = bid + (m >> 3);
          blk = noise_block(pr.noise_seed, noise_off, env, lane, blk_id);
        }
        if (k < head) {
          float v = orow[k];
          if (noise_mode != ELG_NOISE_OFF) {
            float t;
            if (philox) t = fmaf(sample16_dyn(blk, (s0 + m) & 7), 1.0f / 32768.0f, -1.0f);
            else t = 2.0f * __ldg(bf.noise_u + (size_t)env * O + k) - 1.0f;
            v = v + t * s_ns[k];
          }
          if (clip_obs > 0.0f) v = fminf(fmaxf(v, -clip_obs), clip_obs);
          if (kFast || obs_to_smem) orow[k] = v;
          else grow[k] = v;
        }
      }
    }

    STAMP(9, 0)
    // ------------------------------- write back -------------------------------
    if (kFast || bulk) {
      fence_async_smem();   // this thread's generic-proxy writes -> visible to the async (TMA) proxy
      __syncthreads();
      if (lane == 0) {
        for (int i = warp; i < L.n_out; i += nwarps) {
          const CopyDesc d = L.out[i];
          bulk_s2g(static_cast<uint8_t*>(const_cast<void*>(d.g)) + (size_t)env0 * d.bpe, smem_raw + d.soff, (uint32_t)(nenv * d.bpe));
        }
        bulk_commit();
        stores_pending = true;
        STAMP(10, 0)
      }
    } else {
      __syncthreads();
      for (int i = 0; i < L.n_out; ++i) {
        const CopyDesc d = L.out[i];
        coop_copy(static_cast<uint8_t*>(const_cast<void*>(d.g)) + (size_t)env0 * d.bpe, smem_raw + d.soff, (uint32_t)(nenv * d.bpe), tid, nthreads);
      }
      __syncthreads();
    }
  }
  if (stores_pending) bulk_wait_all();
  STAMP(11, 0)
#undef STAMP
#undef ACC
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
  pdl_launch_dependents();   // programmatic dependent launch: scheduled while the previous kernel drains, but nothing is
  pdl_wait();                // read before that kernel's memory is visible
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

// the same arithmetic, four consecutive (env, dof) values per thread with 16-byte accesses (all envs, D % 4 == 0, 16-byte aligned
// arrays, < 2^31 values): one 32-bit remainder per thread instead of a 64-bit division per value
__global__ void __launch_bounds__(256)
elg_torques4_kernel(const uint32_t n4, const uint32_t D, const int control_type, const float action_scale, const float sim_dt,
                    const float4* __restrict__ actions, const float4* __restrict__ dof_state, const float4* __restrict__ last_dof_vel,
                    const float* __restrict__ p_gains, const float* __restrict__ d_gains, const float* __restrict__ torque_limits,
                    const float* __restrict__ default_dof_pos, float4* __restrict__ torques) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  if (i >= n4) return;
  const uint32_t j = (4u * i) % D;
  const float4 a4 = actions[i];
  const float4 lim4 = __ldg(reinterpret_cast<const float4*>(torque_limits + j));
  const float a[4] = {mul_r(a4.x, action_scale), mul_r(a4.y, action_scale), mul_r(a4.z, action_scale), mul_r(a4.w, action_scale)};
  const float lim[4] = {lim4.x, lim4.y, lim4.z, lim4.w};
  float tq[4];
  if (control_type == ELG_CONTROL_T) {
#pragma unroll
    for (int k = 0; k < 4; ++k) tq[k] = a[k];
  } else {
    const float4 s0 = dof_state[2 * i], s1 = dof_state[2 * i + 1];
    const float pos[4] = {s0.x, s0.z, s1.x, s1.z}, vel[4] = {s0.y, s0.w, s1.y, s1.w};
    const float4 pg4 = __ldg(reinterpret_cast<const float4*>(p_gains + j)), dg4 = __ldg(reinterpret_cast<const float4*>(d_gains + j));
    const float pg[4] = {pg4.x, pg4.y, pg4.z, pg4.w}, dg[4] = {dg4.x, dg4.y, dg4.z, dg4.w};
    if (control_type == ELG_CONTROL_P) {
      const float4 q4 = __ldg(reinterpret_cast<const float4*>(default_dof_pos + j));
      const float q0[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) tq[k] = sub_r(mul_r(pg[k], sub_r(add_r(a[k], q0[k]), pos[k])), mul_r(dg[k], vel[k]));
    } else {
      const float4 l4 = last_dof_vel[i];
      const float lv[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) tq[k] = sub_r(mul_r(pg[k], sub_r(a[k], vel[k])), div_r(mul_r(dg[k], sub_r(vel[k], lv[k])), sim_dt));
    }
  }
  torques[i] = make_float4(fminf(fmaxf(tq[0], -lim[0]), lim[0]), fminf(fmaxf(tq[1], -lim[1]), lim[1]), fminf(fmaxf(tq[2], -lim[2]), lim[2]),
                           fminf(fmaxf(tq[3], -lim[3]), lim[3]));
}

// ---------------------------------------------------------------------------------------------
// step_rollout's action hand-over (robot_batch_rollout.py:643-656; robot_traj_grad_sampling.py:326-345): one thread per value
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
elg_rollout_actions_kernel(const float* __restrict__ src, const int M, const int R, const int A, const int64_t src_stride, const float clip,
                           const float* __restrict__ lower, const float* __restrict__ range, float* __restrict__ actions) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  if (i >= (int64_t)M * R * A) return;
  const int64_t row = i / A;
  const int j = (int)(i - row * A);
  const int64_t k = row / R;
  float a = src[row * src_stride + j];
  if (lower) a = add_r(__ldg(lower + j), div_r(mul_r(add_r(fminf(fmaxf(a, -1.0f), 1.0f), 1.0f), __ldg(range + j)), 2.0f));
  actions[(row + k + 1) * A + j] = fminf(fmaxf(a, -clip), clip);
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

// min-of-3 cells as fp32 metres, tabulated once per terrain (legged_robot.py:932-938)
__global__ void __launch_bounds__(256)
elg_height_min_kernel(const int16_t* __restrict__ hs, const int rows, const int cols, const float vs, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int r = i / cols, c = i - r * cols;
  float v = 0.0f;
  if (r <= rows - 2 && c <= cols - 2) v = mul_r((float)min(min((int)hs[i], (int)hs[i + cols]), (int)hs[i + 1]), vs);
  out[i] = v;
}

}  // namespace elg

// =================================================================================================
// C ABI
// =================================================================================================
namespace elg {
thread_local char g_err[256] = "";
int set_error(int code, const char* msg) {
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
int device_index() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
  return dev;
}
int sm_count() {
  static int sms[kMaxDevices] = {};
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  const int slot = (dev >= 0 && dev < kMaxDevices) ? dev : 0;
  if (sms[slot] == 0 || dev != slot) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 0;
    if (dev != slot) return v;
    sms[slot] = v;
  }
  return sms[slot];
}
}  // namespace elg

namespace {
struct StepTune { int cap, threads, ctas_per_sm, no_bulk, no_fast; };
StepTune g_tune = {0, 0, 0, 0, 0};
long long* g_step_dbg = nullptr;
int fail(int code, const char* msg) { return elg::set_error(code, msg); }
int check_launch(const char* what) { return elg::check_launch(what); }
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
const char* elg_last_error(void) { return elg::g_err; }
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
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(threads);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  const bool pv = prm->control_type != ELG_CONTROL_T;
  if (!env_ids && dims->num_dof % 4 == 0 && total < ((int64_t)1 << 31) && a16(actions) && a16(torques) && a16(torque_limits) &&
      (!pv || (a16(dof_state) && a16(p_gains) && a16(d_gains) && a16(default_dof_pos))) &&
      (prm->control_type != ELG_CONTROL_V || a16(last_dof_vel))) {
    const uint32_t n4 = (uint32_t)(total / 4);
    cfg.gridDim = dim3((n4 + threads - 1) / threads);
    cudaLaunchKernelEx(&cfg, elg::elg_torques4_kernel, n4, (uint32_t)dims->num_dof, (int)prm->control_type, prm->action_scale, prm->sim_dt,
                       reinterpret_cast<const float4*>(actions), reinterpret_cast<const float4*>(dof_state), reinterpret_cast<const float4*>(last_dof_vel),
                       p_gains, d_gains, torque_limits, default_dof_pos, reinterpret_cast<float4*>(torques));
    return check_launch("elg_compute_torques");
  }
  cudaLaunchKernelEx(&cfg, elg::elg_torques_kernel, rows, (int)dims->num_dof, (int)prm->control_type, prm->action_scale, prm->sim_dt,
                     actions, dof_state, last_dof_vel, p_gains, d_gains, torque_limits, default_dof_pos, torques, env_ids);
  return check_launch("elg_compute_torques");
}

int elg_rollout_actions(const float* rollout_actions, int32_t num_main, int32_t rollouts_per_main, int32_t num_actions, int64_t src_row_stride,
                        float clip_actions, const float* joint_lower, const float* joint_range, float* actions, void* stream) {
  if (num_main < 0 || rollouts_per_main < 0 || num_actions < 1) return fail(ELG_ERR_INVALID_ARGUMENT, "elg_rollout_actions: bad sizes");
  if (src_row_stride == 0) src_row_stride = num_actions;
  if (src_row_stride < num_actions) return fail(ELG_ERR_INVALID_ARGUMENT, "elg_rollout_actions: src_row_stride < num_actions");
  if ((joint_lower == nullptr) != (joint_range == nullptr)) return fail(ELG_ERR_INVALID_ARGUMENT, "joint_lower and joint_range go together");
  const int64_t total = (int64_t)num_main * rollouts_per_main * num_actions;
  if (total == 0) return ELG_OK;
  if (!rollout_actions || !actions) return fail(ELG_ERR_NULL_POINTER, "elg_rollout_actions: a buffer is NULL");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((total + 255) / 256));
  cfg.blockDim = dim3(256);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, elg::elg_rollout_actions_kernel, rollout_actions, (int)num_main, (int)rollouts_per_main, (int)num_actions, src_row_stride,
                     clip_actions, joint_lower, joint_range, actions);
  return check_launch("elg_rollout_actions");
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
  if (((prm->reward_mask >> ELG_REW_GAIT_2_STEP) & 1u) && dims->num_feet < (prm->gait_2_step_hexapod ? 6 : 4))
    return fail(ELG_ERR_INVALID_ARGUMENT, "_reward_gait_2_step indexes feet 0..3 (0..5 for the hexapod form)");
  const uint32_t lim = (1u << ELG_REW_DOF_POS_LIMITS) | (1u << ELG_REW_DOF_VEL_LIMITS) | (1u << ELG_REW_TORQUE_LIMITS);
  if ((prm->reward_mask & lim) && (!buf->dof_pos_limits || !buf->dof_vel_limits || !buf->torque_limits))
    return fail(ELG_ERR_NULL_POINTER, "limit reward terms need dof_pos_limits, dof_vel_limits and torque_limits");
  if (prm->noise_mode == ELG_NOISE_TENSOR && (!buf->noise_u || !buf->noise_scale_vec)) return fail(ELG_ERR_NULL_POINTER, "ELG_NOISE_TENSOR needs noise_u and noise_scale_vec");
  if (prm->noise_mode == ELG_NOISE_PHILOX && !buf->noise_scale_vec) return fail(ELG_ERR_NULL_POINTER, "ELG_NOISE_PHILOX needs noise_scale_vec");
  if (prm->noise_mode < ELG_NOISE_OFF || prm->noise_mode > ELG_NOISE_PHILOX) return fail(ELG_ERR_INVALID_ARGUMENT, "bad noise_mode");
  const int N = dims->num_envs;
  if (N == 0) return ELG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (g_tune.no_fast == 0) {   // common quadruped layout, whole step in one launch: the lean kernel of elg_step_fast.cu
    int rc = ELG_OK;
    if (elg::launch_step_fast(dims, prm, buf, phase, g_tune.cap, g_tune.threads, g_step_dbg, stream, &rc)) return rc;
  }
  const int D = dims->num_dof, F = dims->num_feet, B = dims->num_bodies, H = dims->num_height_points, O = dims->num_obs, C = dims->num_commands;
  const bool do_derive = phase & ELG_PHASE_DERIVE, do_term = phase & ELG_PHASE_TERMINATION, do_reward = phase & ELG_PHASE_REWARD;
  const bool do_obs = phase & ELG_PHASE_OBS, do_hist = phase & ELG_PHASE_HISTORY;
  const bool gait = buf->gait_idx && buf->gait_prev_foot_z;
  const bool need_hsum = do_reward && ((prm->reward_mask >> ELG_REW_BASE_HEIGHT) & 1u) && H > 0;
  const bool air_on = (prm->reward_mask >> ELG_REW_FEET_AIR_TIME) & 1u;
  const bool rollout = prm->rollout_mode != 0;

  static int sms = 0;
  if (sms == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
      return fail(ELG_ERR_CUDA, "cannot query the SM count");
    sms = v;
  }
  // ---- chunking: whole quads of envs, one warp per env, balanced over the SMs (see the header comment of this file)
  const long long Q = ((long long)N + 3) / 4;
  int cap, grid, nchunks;
  if (g_tune.cap > 0) {   // explicit tuning (elg_set_step_tuning): envs per chunk, CTAs per SM
    cap = g_tune.cap;
    long long nch = (Q + cap / 4 - 1) / (cap / 4);
    const long long g = (long long)sms * g_tune.ctas_per_sm;
    if (nch > g) nch = (nch + g - 1) / g * g;
    if (nch > Q) nch = Q;
    nchunks = (int)nch;
    grid = (int)(nch < g ? nch : g);
  } else if ((Q + 7) / 8 <= sms) {   // at most one chunk per SM, every chunk <= 32 envs
    nchunks = (int)(Q < sms ? Q : sms);
    cap = 4 * (int)((Q + nchunks - 1) / nchunks);
    grid = nchunks;
  } else {                           // many chunks: 12-env CTAs (13 warps), 2 per SM, each looping over its chunks
    cap = 12;
    const long long g = (long long)sms * 2;
    long long nch = (Q + 2) / 3;
    if (nch > g) nch = (nch + g - 1) / g * g;
    if (nch > Q) nch = Q;
    nchunks = (int)nch;
    grid = (int)(nch < g ? nch : g);
  }
  const int threads = 32 * (cap < 32 ? cap + 1 : 32);   // one row warp per env + the scalar warp

  // ---- shared-memory plan + copy tables
  elg::StepPlan L{};
  L.cap = cap;
  L.nchunks = nchunks;
  L.quads_base = (int)(Q / nchunks);
  L.quads_rem = (int)(Q % nchunks);
  L.obs_smem = (O == head + H) ? 1 : 0;
  L.hm = (head + 31) / 32;
  int nt = 0;
  for (int t = 0; t < ELG_NUM_REWARD_TERMS; ++t)
    if ((prm->reward_mask >> t) & 1u) L.term_ids[nt++] = (int8_t)t;
  L.nterms = nt;
  int off = 0;
  auto take = [&](int bytes) { const int at = off; off += (bytes + 127) & ~127; return at; };
  L.root = take(cap * 13 * 4);
  L.dof = take(cap * 2 * D * 4);
  L.act = take(cap * D * 4);
  L.lact = take(cap * D * 4);
  L.ldv = take(cap * D * 4);
  L.tq = take(cap * D * 4);
  L.cf = take(cap * B * 12);
  L.lrv = take(cap * 24);
  L.vec5 = take(5 * cap * 12);     // base_lin_vel, base_ang_vel, projected_gravity, base_lin_acc, base_ang_acc
  L.cmd = take(cap * C * 4);
  L.air = take(cap * F * 4);
  L.con = take(cap * F * 4);
  L.lc = take(cap * F);
  L.ep = take(cap * 8);
  L.gidx = take(cap * 4);
  L.gprev = take(cap * F * 4);
  L.fpos = take(cap * F * 12);
  L.fvel = take(cap * F * 12);
  L.sums = take((nt > 0 ? nt : 1) * cap * 4);
  L.rew = take(cap * 4);
  L.mh = take(cap * H * 4);
  L.obs = take(cap * (L.obs_smem ? O : head) * 4);
  L.q0 = take(D * 4);
  L.plim = take(2 * D * 4);
  L.vlim = take(D * 4);
  L.tlim = take(D * 4);
  L.ns = take(O * 4);
  L.grid = take(H * 16);
  L.acc = take((ELG_NUM_REWARD_TERMS + 5) * 32 * 4);
  L.idx = take((ELG_MAX_FEET + ELG_MAX_PENALISED + ELG_MAX_TERMINATION) * 4);
  L.bytes = off;
  if ((size_t)L.bytes + 1024 > (size_t)227 * 1024)
    return fail(ELG_ERR_UNSUPPORTED, "robot dimensions do not fit the shared-memory plan of elg_step_kernel");

  int n_in = 0, n_out = 0, in_bpe = 0;
  bool aligned = true;
  auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  auto in = [&](const void* g, int soff, int bpe) {
    L.in[n_in++] = elg::CopyDesc{g, soff, bpe};
    in_bpe += bpe;
    aligned = aligned && a16(g);
  };
  auto out = [&](void* g, int soff, int bpe) {
    L.out[n_out++] = elg::CopyDesc{g, soff, bpe};
    aligned = aligned && a16(g);
  };
  const int v3 = cap * 12;
  in(buf->root_states, L.root, 52);
  in(buf->dof_state, L.dof, 8 * D);
  in(buf->actions, L.act, 4 * D);
  if (do_reward) {
    in(buf->last_actions, L.lact, 4 * D);
    in(buf->last_dof_vel, L.ldv, 4 * D);
    in(buf->torques, L.tq, 4 * D);
  }
  if (do_term || do_reward) in(buf->contact_forces, L.cf, 12 * B);
  if (do_derive) {
    in(buf->last_root_vel, L.lrv, 24);
    in(buf->base_lin_acc, L.vec5 + 3 * v3, 12);
    in(buf->base_ang_acc, L.vec5 + 4 * v3, 12);
    if (rollout && H > 0 && (do_obs || need_hsum)) in(buf->measured_heights, L.mh, 4 * H);
  } else {
    in(buf->base_lin_vel, L.vec5 + 0 * v3, 12);
    in(buf->base_ang_vel, L.vec5 + 1 * v3, 12);
    in(buf->projected_gravity, L.vec5 + 2 * v3, 12);
    if (H > 0 && (do_obs || need_hsum)) in(buf->measured_heights, L.mh, 4 * H);
    if (F > 0 && do_reward) {
      in(buf->foot_positions, L.fpos, 12 * F);
      in(buf->foot_velocities, L.fvel, 12 * F);
    }
  }
  in(buf->commands, L.cmd, 4 * C);
  if (F > 0 && do_reward) {
    in(buf->feet_air_time, L.air, 4 * F);
    in(buf->feet_contact_time, L.con, 4 * F);
    in(buf->last_contacts, L.lc, F);
  }
  if ((do_derive && !rollout) || do_term) in(buf->episode_length_buf, L.ep, 8);
  if (gait && do_reward) {
    in(buf->gait_idx, L.gidx, 4);
    in(buf->gait_prev_foot_z, L.gprev, 4 * F);
  }
  if (do_reward && !rollout)
    for (int ti = 0; ti < nt; ++ti) in(buf->episode_sums + (size_t)L.term_ids[ti] * N, L.sums + ti * cap * 4, 4);

  if (do_derive) {
    out(buf->base_lin_vel, L.vec5 + 0 * v3, 12);
    out(buf->base_ang_vel, L.vec5 + 1 * v3, 12);
    out(buf->projected_gravity, L.vec5 + 2 * v3, 12);
    out(buf->base_lin_acc, L.vec5 + 3 * v3, 12);
    out(buf->base_ang_acc, L.vec5 + 4 * v3, 12);
    if (F > 0) {
      out(buf->foot_positions, L.fpos, 12 * F);
      out(buf->foot_velocities, L.fvel, 12 * F);
    }
    if (prm->heading_command && !rollout) out(buf->commands, L.cmd, 4 * C);
    if (!rollout) out(buf->episode_length_buf, L.ep, 8);
    if (H > 0 && !rollout) out(buf->measured_heights, L.mh, 4 * H);
  }
  if (do_reward) {
    if (air_on && F > 0) {
      out(buf->feet_air_time, L.air, 4 * F);
      out(buf->feet_contact_time, L.con, 4 * F);
      out(buf->last_contacts, L.lc, F);
    }
    if (!rollout)
      for (int ti = 0; ti < nt; ++ti) out(buf->episode_sums + (size_t)L.term_ids[ti] * N, L.sums + ti * cap * 4, 4);
    out(buf->rew_buf, L.rew, 4);
    if (gait) {
      out(buf->gait_idx, L.gidx, 4);
      out(buf->gait_prev_foot_z, L.gprev, 4 * F);
    }
  }
  if (do_obs && L.obs_smem) out(buf->obs_buf, L.obs, 4 * O);
  if (do_hist) {
    out(buf->last_actions, L.lact, 4 * D);
    out(buf->last_dof_vel, L.ldv, 4 * D);
    out(buf->last_root_vel, L.lrv, 24);
  }
  L.dbg = g_step_dbg;
  L.n_in = n_in;
  L.n_out = n_out;
  L.in_bpe = in_bpe;
  // TMA bulk staging needs 16-byte aligned array bases; sub-chunk offsets and sizes are multiples of 16 by construction
  // (whole quads of envs) when N and the foot count are multiples of 4.  Otherwise: cooperative element-wise staging.
  L.use_bulk = (N % 4 == 0) && (F % 4 == 0) && aligned && g_tune.no_bulk == 0;

  const size_t smem = (size_t)L.bytes;
  const bool quad = (D == 12 && F == 4);
  auto kern = quad ? elg::elg_step_kernel<12, 4, false> : elg::elg_step_kernel<0, 0, false>;
  const int which = quad ? 1 : 0;
  static elg::SmemCache smem_cache[3] = {};
  size_t& smem_have = elg::smem_slot(smem_cache[which]);
  if (smem > smem_have) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return fail(ELG_ERR_CUDA, "cannot reserve dynamic shared memory for elg_step_kernel");
    smem_have = smem;
  }
  kern<<<grid, threads, smem, st>>>(*dims, *prm, *buf, L, phase);
  return check_launch("elg_post_physics_step");
}

int elg_set_step_debug(long long* device_stamps) {
  g_step_dbg = device_stamps;
  return ELG_OK;
}

int elg_set_step_tuning(int envs_per_chunk, int threads_per_cta, int ctas_per_sm, int disable_bulk) {
  if (envs_per_chunk == 0) {
    g_tune = StepTune{0, threads_per_cta, ctas_per_sm, disable_bulk & 1, (disable_bulk >> 1) & 1};
    return ELG_OK;
  }
  if (envs_per_chunk < 4 || envs_per_chunk > elg::kMaxCap || envs_per_chunk % 4 != 0)
    return fail(ELG_ERR_INVALID_ARGUMENT, "envs_per_chunk must be a multiple of 4 in [4, 32]");
  (void)threads_per_cta;   // v3: one warp per env, the CTA width follows envs_per_chunk
  if (ctas_per_sm < 1 || ctas_per_sm > 16) return fail(ELG_ERR_INVALID_ARGUMENT, "ctas_per_sm must be in [1, 16]");
  g_tune = StepTune{envs_per_chunk, threads_per_cta, ctas_per_sm, disable_bulk & 1, (disable_bulk >> 1) & 1};
  return ELG_OK;
}

int elg_prepare_height_field(const int16_t* height_samples, int32_t rows, int32_t cols, float vertical_scale, float* out, void* stream) {
  if (!height_samples || !out) return fail(ELG_ERR_NULL_POINTER, "height_samples/out is NULL");
  if (rows < 2 || cols < 2 || (long long)rows * cols > 0x7fffffffLL) return fail(ELG_ERR_INVALID_ARGUMENT, "rows, cols must be >= 2 and rows*cols < 2^31");
  const int n = rows * cols;
  elg::elg_height_min_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(height_samples, rows, cols, vertical_scale, out);
  return check_launch("elg_prepare_height_field");
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
