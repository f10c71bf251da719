// elg_step.cu -- fused post-physics step, PD torques and terrain height scan for sm_100a.
//
// One launch of elg_step_kernel replaces the ~100 ATen launches of
// LeggedRobot.post_physics_step (envs/base/legged_robot.py:113-150 in the reference).
//
// Work decomposition (one CTA = 9 warps = EPB environments):
//   warp 0      "state warp": one THREAD per environment.  Everything that is O(D + F + B) per env:
//               base-frame velocities / accelerations / projected gravity, heading command,
//               termination, the whole _reward_* registry with its alphabetical fp32 sum and the
//               episode sums, feet timers, history copies.  It leaves the first 12+3D observation
//               entries in shared memory.
//   warps 1..8  "row warps": one WARP per environment (EPB/8 environments each, in turn), lanes over
//               the H height points: yaw-rotate the sampling grid, terrain cell lookup (min of three
//               int16 samples), measured_heights and the height part of the observation row, then --
//               after the CTA barrier -- the head of the row from shared memory.  Rows are written
//               with consecutive lanes on consecutive floats (coalesced 128 B stores); observation
//               noise is generated in registers (Philox4x32-10) or read from a caller tensor.
// HBM traffic is the algorithmic minimum: every input element is read once (L1 serves the strided
// re-use inside the state warp) and every output element written once; the 1.6 MB height field is
// L2 resident.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "elg_common.cuh"

namespace elg {

constexpr int kRowWarps = 8;
constexpr int kStepThreads = (1 + kRowWarps) * kWarp;  // 288

__device__ __forceinline__ bool term_on(const ElgStepParams& pr, int t) { return (pr.reward_mask >> t) & 1u; }

// ---------------------------------------------------------------------------------------------
// height scan of one environment by one warp  (LeggedRobot._get_heights, legged_robot.py:900-938)
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
// observation post-processing shared by head and height entries (legged_robot.py:250-252, :107-108)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float finish_obs(float v, float u, float ns, const ElgStepParams& pr) {
  if (pr.noise_mode != ELG_NOISE_OFF) v = v + (2.0f * u - 1.0f) * ns;
  if (pr.clip_observations > 0.0f) v = fminf(fmaxf(v, -pr.clip_observations), pr.clip_observations);
  return v;
}

// ---------------------------------------------------------------------------------------------
// shared-memory staging of one CTA's environments
// ---------------------------------------------------------------------------------------------
// The EPB environments of a CTA are consecutive, so every per-env input array is ONE contiguous
// global range per CTA.  Thread 0 issues one TMA bulk copy (cp.async.bulk global -> shared,
// completion on an mbarrier) per array; the whole CTA then works out of shared memory, i.e. the
// ~60 dependent DRAM round trips of a per-thread gather collapse into a single one.
struct StageLayout {   // offsets in 4-byte words; every region starts 16-byte aligned
  int root, dof, act, lact, ldv, tq, cf, feet, lrv, lacc, aacc, cmd, air, con, lc, ep, gidx, gprev, head, hsum, sums, accs, words;
};
__host__ __device__ inline int up4(int w) { return (w + 3) & ~3; }
__host__ __device__ inline StageLayout make_layout(const ElgDims& d, int epb) {
  StageLayout L;
  int o = 0;
  const int D = d.num_dof, F = d.num_feet;
  L.root = o;  o += up4(epb * 13);
  L.dof = o;   o += up4(epb * 2 * D);
  L.act = o;   o += up4(epb * D);
  L.lact = o;  o += up4(epb * D);
  L.ldv = o;   o += up4(epb * D);
  L.tq = o;    o += up4(epb * D);
  L.cf = o;    o += up4(epb * d.num_bodies * 3);
  L.feet = o;  o += up4(epb * F * 6);          // per foot: pos xyz, lin vel xyz (gathered rows)
  L.lrv = o;   o += up4(epb * 6);
  L.lacc = o;  o += up4(epb * 3);
  L.aacc = o;  o += up4(epb * 3);
  L.cmd = o;   o += up4(epb * d.num_commands);
  L.air = o;   o += up4(epb * F);
  L.con = o;   o += up4(epb * F);
  L.lc = o;    o += up4((epb * F + 3) / 4);    // bytes
  L.ep = o;    o += up4(epb * 2);              // int64
  L.gidx = o;  o += up4(epb);
  L.gprev = o; o += up4(epb * F);
  L.head = o;  o += up4(epb * (12 + 3 * D + 1));
  L.hsum = o;  o += up4(epb);
  L.sums = o;  o += up4(epb * ELG_NUM_REWARD_TERMS);   // episode sums being updated (cp.async prefetch)
  L.accs = o;  o += up4(epb * ELG_NUM_REWARD_TERMS);   // raw reward terms, [term][slot]
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

// ---------------------------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------------------------
constexpr int kMaxPts = 8;   // height points per lane and pass (covers H <= 256 in one pass)

template <int EPB>
__global__ void __launch_bounds__(kStepThreads, 4)
elg_step_kernel(const __grid_constant__ ElgDims dm, const __grid_constant__ ElgStepParams pr,
                const __grid_constant__ ElgStepBuffers bf, const uint32_t phase, const int use_bulk) {
  extern __shared__ __align__(16) float smem[];
  __shared__ __align__(8) uint64_t s_bar;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int env0 = blockIdx.x * EPB;
  const int N = dm.num_envs, D = dm.num_dof, B = dm.num_bodies, F = dm.num_feet, H = dm.num_height_points;
  const int O = dm.num_obs, C = dm.num_commands;
  const int head = 12 + 3 * D;
  const int nenv = min(EPB, N - env0);
  const bool do_derive = phase & ELG_PHASE_DERIVE, do_term = phase & ELG_PHASE_TERMINATION;
  const bool do_reward = phase & ELG_PHASE_REWARD, do_obs = phase & ELG_PHASE_OBS, do_hist = phase & ELG_PHASE_HISTORY;
  const bool need_hsum = do_reward && term_on(pr, ELG_REW_BASE_HEIGHT) && H > 0;   // CTA-uniform
  const bool gait = bf.gait_idx != nullptr && bf.gait_prev_foot_z != nullptr;
  const StageLayout L = make_layout(dm, EPB);

  float* s_root = smem + L.root;   float* s_dof = smem + L.dof;    float* s_act = smem + L.act;
  float* s_lact = smem + L.lact;   float* s_ldv = smem + L.ldv;    float* s_tq = smem + L.tq;
  float* s_cf = smem + L.cf;       float* s_feet = smem + L.feet;  float* s_lrv = smem + L.lrv;
  float* s_lacc = smem + L.lacc;   float* s_aacc = smem + L.aacc;  float* s_cmd = smem + L.cmd;
  float* s_air = smem + L.air;     float* s_con = smem + L.con;
  uint8_t* s_lc = reinterpret_cast<uint8_t*>(smem + L.lc);
  int64_t* s_ep = reinterpret_cast<int64_t*>(smem + L.ep);
  float* s_gidx = smem + L.gidx;   float* s_gprev = smem + L.gprev;
  float* s_head = smem + L.head;   float* s_hsum = smem + L.hsum;
  float* s_sums = smem + L.sums;   float* s_accs = smem + L.accs;
  const int head_pitch = head + 1;   // odd pitch when D is even: the state warp writes column-wise

  // ------------------------------- stage inputs -------------------------------
  const bool bulk = use_bulk && nenv == EPB;
  if (bulk) {
    if (tid == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (tid == 0) {
      const size_t e = (size_t)env0;
      uint32_t bytes = 4u * EPB * (13 + 2 * D + 4 * D + B * 3 + 6 + 3 + 3 + C + 2 * F) + EPB * F + 8u * EPB;
      if (gait) bytes += 4u * EPB * (1 + F);
      mbar_expect_tx(&s_bar, bytes);
      bulk_g2s(s_root, bf.root_states + e * 13, 4u * EPB * 13, &s_bar);
      bulk_g2s(s_dof, bf.dof_state + e * 2 * D, 4u * EPB * 2 * D, &s_bar);
      bulk_g2s(s_act, bf.actions + e * D, 4u * EPB * D, &s_bar);
      bulk_g2s(s_lact, bf.last_actions + e * D, 4u * EPB * D, &s_bar);
      bulk_g2s(s_ldv, bf.last_dof_vel + e * D, 4u * EPB * D, &s_bar);
      bulk_g2s(s_tq, bf.torques + e * D, 4u * EPB * D, &s_bar);
      bulk_g2s(s_cf, bf.contact_forces + e * B * 3, 4u * EPB * B * 3, &s_bar);
      bulk_g2s(s_lrv, bf.last_root_vel + e * 6, 4u * EPB * 6, &s_bar);
      bulk_g2s(s_lacc, bf.base_lin_acc + e * 3, 4u * EPB * 3, &s_bar);
      bulk_g2s(s_aacc, bf.base_ang_acc + e * 3, 4u * EPB * 3, &s_bar);
      bulk_g2s(s_cmd, bf.commands + e * C, 4u * EPB * C, &s_bar);
      bulk_g2s(s_air, bf.feet_air_time + e * F, 4u * EPB * F, &s_bar);
      bulk_g2s(s_con, bf.feet_contact_time + e * F, 4u * EPB * F, &s_bar);
      bulk_g2s(s_lc, bf.last_contacts + e * F, (uint32_t)(EPB * F), &s_bar);
      bulk_g2s(s_ep, bf.episode_length_buf + e, 8u * EPB, &s_bar);
      if (gait) {
        bulk_g2s(s_gidx, bf.gait_idx + e, 4u * EPB, &s_bar);
        bulk_g2s(s_gprev, bf.gait_prev_foot_z + e * F, 4u * EPB * F, &s_bar);
      }
    }
  } else {
    // ragged tail CTA or unaligned caller tensors: cooperative element-wise staging
    auto stage = [&](float* dst, const float* src, int per_env) {
      const float* g = src + (size_t)env0 * per_env;
      for (int i = tid; i < nenv * per_env; i += kStepThreads) dst[i] = g[i];
    };
    stage(s_root, bf.root_states, 13);       stage(s_dof, bf.dof_state, 2 * D);
    stage(s_act, bf.actions, D);             stage(s_lact, bf.last_actions, D);
    stage(s_ldv, bf.last_dof_vel, D);        stage(s_tq, bf.torques, D);
    stage(s_cf, bf.contact_forces, B * 3);   stage(s_lrv, bf.last_root_vel, 6);
    stage(s_lacc, bf.base_lin_acc, 3);       stage(s_aacc, bf.base_ang_acc, 3);
    stage(s_cmd, bf.commands, C);            stage(s_air, bf.feet_air_time, F);
    stage(s_con, bf.feet_contact_time, F);
    for (int i = tid; i < nenv * F; i += kStepThreads) s_lc[i] = bf.last_contacts[(size_t)env0 * F + i];
    for (int i = tid; i < nenv; i += kStepThreads) s_ep[i] = bf.episode_length_buf[env0 + i];
    if (gait) { stage(s_gidx, bf.gait_idx, 1); stage(s_gprev, bf.gait_prev_foot_z, F); }
  }
  // feet rows of rigid_body_state are a strided gather (52-byte rows): plain loads, 6 useful floats per row
  for (int r = tid; r < nenv * F * 6; r += kStepThreads) {
    const int c = r % 6, ef = r / 6;
    const int e = ef / F, f = ef - e * F;
    s_feet[r] = __ldg(bf.rigid_body_state + ((size_t)(env0 + e) * B + dm.feet_idx[f]) * 13 + (c < 3 ? c : c + 4));
  }
  if (bulk) mbar_wait(&s_bar, 0);
  __syncthreads();

  if (warp > 0) {
    // =========================== row warps: heights + observation rows ===========================
    if (H > 0 && (do_derive || do_obs || need_hsum)) {
      const int16_t* __restrict__ hs = bf.height_samples;
      float* __restrict__ mh = bf.measured_heights;
      float* __restrict__ obs = bf.obs_buf;
      for (int slot = warp - 1; slot < nenv; slot += kRowWarps) {
        const int env = env0 + slot;
        const float* rs = s_root + slot * 13;
        float hsum = 0.0f;
        YawFrame fr;
        fr.X = rs[0]; fr.Y = rs[1]; fr.Z = rs[2];
        if (do_derive) {
          float n = __fsqrt_rn(add_r(mul_r(rs[5], rs[5]), mul_r(rs[6], rs[6])));
          n = fmaxf(n, 1e-9f);
          fr.zz = div_r(rs[5], n);
          fr.ww = div_r(rs[6], n);
        }
        const float rootz = fr.Z;
        const float zc = sub_r(rootz, 0.5f);
        const float* hp = bf.height_points + (size_t)env * pr.height_points_env_stride;
        for (int p0 = 0; p0 < H; p0 += kMaxPts * kWarp) {
          float h[kMaxPts];
          if (do_derive && !pr.terrain_is_plane) {
            int a0[kMaxPts], a1[kMaxPts], a2[kMaxPts];
#pragma unroll
            for (int i = 0; i < kMaxPts; ++i) {
              const int p = p0 + i * kWarp + lane;
              a0[i] = a1[i] = a2[i] = 0;
              if (p < H) {
                int ix, iy;
                terrain_cell(fr, __ldg(hp + 3 * p), __ldg(hp + 3 * p + 1), pr, ix, iy);
                const int16_t* q = hs + (size_t)ix * pr.hf_cols + iy;
                a0[i] = __ldg(q); a1[i] = __ldg(q + pr.hf_cols); a2[i] = __ldg(q + 1);
              }
            }
#pragma unroll
            for (int i = 0; i < kMaxPts; ++i) h[i] = mul_r((float)min(min(a0[i], a1[i]), a2[i]), pr.vertical_scale);
          } else {
#pragma unroll
            for (int i = 0; i < kMaxPts; ++i) {
              const int p = p0 + i * kWarp + lane;
              h[i] = (!do_derive && p < H) ? mh[(size_t)env * H + p] : 0.0f;
            }
          }
          uint4 rnd = make_uint4(0, 0, 0, 0);
#pragma unroll
          for (int i = 0; i < kMaxPts; ++i) {
            const int p = p0 + i * kWarp + lane;
            if (p < H) {
              if (do_derive) mh[(size_t)env * H + p] = h[i];
              hsum += sub_r(rootz, h[i]);
              if (do_obs) {
                const int k = head + p;
                float v = mul_r(fminf(fmaxf(sub_r(zc, h[i]), -1.0f), 1.0f), pr.obs_scale_height);
                float u = 0.0f, ns = 0.0f;
                if (pr.noise_mode == ELG_NOISE_TENSOR) {
                  u = __ldg(bf.noise_u + (size_t)env * O + k);
                  ns = __ldg(bf.noise_scale_vec + k);
                } else if (pr.noise_mode == ELG_NOISE_PHILOX) {
                  const int kk = k >> 5;
                  if ((kk & 3) == 0 || i == 0) rnd = noise_block(pr.noise_seed, pr.noise_offset, env, k & 31, kk >> 2);
                  u = u01(pick(rnd, kk & 3));
                  ns = __ldg(bf.noise_scale_vec + k);
                }
                obs[(size_t)env * O + k] = finish_obs(v, u, ns, pr);
              }
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
    if (need_hsum) __syncthreads();   // (A) heights -> state warp
  } else {
    // =========================== state warp: one thread per environment ===========================
    const int slot = lane;
    const int env = env0 + slot;
    const bool active = slot < nenv;
    Vec3 blv = {0, 0, 0}, bav = {0, 0, 0}, pg = {0, 0, 0};
    float cmd0 = 0, cmd1 = 0, cmd2 = 0, cmd3 = 0;
    // raw (unscaled) value of every built-in term lives in shared memory, [term][slot]
#define acc(t) s_accs[(t) * EPB + (slot & (EPB - 1))]
    bool reset = false, time_out = false;
    const float* rs = s_root + (active ? slot : 0) * 13;

    if (active) {
      if (do_reward) {   // prefetch the episode sums straight into shared memory (LDGSTS); consumed at the very end
#pragma unroll
        for (int t = 0; t < ELG_NUM_REWARD_TERMS; ++t)
          if (term_on(pr, t))
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(s_sums + t * EPB + slot)),
                         "l"(bf.episode_sums + (size_t)t * N + env) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      if (do_derive) {
        const Quat q = {rs[3], rs[4], rs[5], rs[6]};
        const Vec3 lin = {rs[7], rs[8], rs[9]}, ang = {rs[10], rs[11], rs[12]};
        // ---- episode counter + derived base state (legged_robot.py:122-134)
        s_ep[slot] += 1;
        bf.episode_length_buf[env] = s_ep[slot];
        const float* lrv = s_lrv + slot * 6;
        blv = quat_rotate_inverse(q, lin);
        bav = quat_rotate_inverse(q, ang);
        pg = quat_rotate_inverse(q, Vec3{pr.gravity_vec[0], pr.gravity_vec[1], pr.gravity_vec[2]});
        {
          const Vec3 dl = quat_rotate_inverse(q, Vec3{lin.x - lrv[0], lin.y - lrv[1], lin.z - lrv[2]});
          const Vec3 da = quat_rotate_inverse(q, Vec3{ang.x - lrv[3], ang.y - lrv[4], ang.z - lrv[5]});
          const float ema = pr.acc_ema, w1 = pr.acc_ema_c;
          const float* la0 = s_lacc + slot * 3;
          const float* aa0 = s_aacc + slot * 3;
          float* la = bf.base_lin_acc + (size_t)env * 3;
          float* aa = bf.base_ang_acc + (size_t)env * 3;
          la[0] = la0[0] * ema + (w1 * dl.x) / pr.dt;
          la[1] = la0[1] * ema + (w1 * dl.y) / pr.dt;
          la[2] = la0[2] * ema + (w1 * dl.z) / pr.dt;
          aa[0] = aa0[0] * ema + (w1 * da.x) / pr.dt;
          aa[1] = aa0[1] * ema + (w1 * da.y) / pr.dt;
          aa[2] = aa0[2] * ema + (w1 * da.z) / pr.dt;
        }
        float* o3;
        o3 = bf.base_lin_vel + (size_t)env * 3;      o3[0] = blv.x; o3[1] = blv.y; o3[2] = blv.z;
        o3 = bf.base_ang_vel + (size_t)env * 3;      o3[0] = bav.x; o3[1] = bav.y; o3[2] = bav.z;
        o3 = bf.projected_gravity + (size_t)env * 3; o3[0] = pg.x;  o3[1] = pg.y;  o3[2] = pg.z;
        // ---- feet gather (legged_robot.py:136-137)
        for (int f = 0; f < F; ++f) {
          const float* ft = s_feet + (slot * F + f) * 6;
          float* fp = bf.foot_positions + ((size_t)env * F + f) * 3;
          float* fv = bf.foot_velocities + ((size_t)env * F + f) * 3;
          fp[0] = ft[0]; fp[1] = ft[1]; fp[2] = ft[2];
          fv[0] = ft[3]; fv[1] = ft[4]; fv[2] = ft[5];
        }
        // ---- heading command (legged_robot.py:394-398); forward = quat_apply(q, (1,0,0))
        if (pr.heading_command) {
          const float fx = 1.0f + (q.y * (-2.0f * q.y) - q.z * (2.0f * q.z));
          const float fy = q.w * (2.0f * q.z) + (q.z * 0.0f - q.x * (-2.0f * q.y));
          const float heading = atan2f(fy, fx);
          const float c2 = fminf(fmaxf(0.5f * wrap_to_pi(s_cmd[slot * C + 3] - heading), -1.0f), 1.0f);
          s_cmd[slot * C + 2] = c2;
          bf.commands[(size_t)env * C + 2] = c2;
        }
      } else if (do_reward || do_obs) {
        const float* p3;
        p3 = bf.base_lin_vel + (size_t)env * 3;      blv = Vec3{p3[0], p3[1], p3[2]};
        p3 = bf.base_ang_vel + (size_t)env * 3;      bav = Vec3{p3[0], p3[1], p3[2]};
        p3 = bf.projected_gravity + (size_t)env * 3; pg = Vec3{p3[0], p3[1], p3[2]};
      }
      cmd0 = s_cmd[slot * C]; cmd1 = s_cmd[slot * C + 1]; cmd2 = s_cmd[slot * C + 2];
      if (C > 3) cmd3 = s_cmd[slot * C + 3];

      // ---- termination (legged_robot.py:155-160)
      if (do_term) {
        bool contact_term = false;
        for (int t = 0; t < dm.num_termination; ++t) {
          const float* f = s_cf + (slot * B + dm.termination_idx[t]) * 3;
          contact_term |= norm3_t(f[0], f[1], f[2]) > 1.0f;
        }
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
        // ---- per-DOF partial sums (legged_robot_rew_mixin.py:84-114, :96-98, :219-221)
        float s_action_rate = 0, s_dof_acc = 0, s_pos_lim = 0, s_dof_vel = 0, s_vel_lim = 0, s_still = 0, s_tq_lim = 0, s_tq2 = 0;
        const bool lim_terms = term_on(pr, ELG_REW_DOF_POS_LIMITS) | term_on(pr, ELG_REW_DOF_VEL_LIMITS) | term_on(pr, ELG_REW_TORQUE_LIMITS);
        for (int j = 0; j < D; ++j) {
          const int e = slot * D + j;
          const float pos = s_dof[2 * e], vel = s_dof[2 * e + 1];
          const float a = s_act[e], la = s_lact[e], lv = s_ldv[e], tq = s_tq[e];
          const float da = la - a;
          s_action_rate += da * da;
          const float dv = (lv - vel) / pr.dt;
          s_dof_acc += dv * dv;
          s_dof_vel += vel * vel;
          s_tq2 += tq * tq;
          s_still += fabsf(pos - __ldg(bf.default_dof_pos + j));
          if (lim_terms) {
            const float lo = __ldg(bf.dof_pos_limits + 2 * j), hi = __ldg(bf.dof_pos_limits + 2 * j + 1);
            s_pos_lim += -fminf(pos - lo, 0.0f) + fmaxf(pos - hi, 0.0f);
            s_vel_lim += fminf(fmaxf(fabsf(vel) - __ldg(bf.dof_vel_limits + j) * pr.soft_dof_vel_limit, 0.0f), 1.0f);
            s_tq_lim += fmaxf(fabsf(tq) - __ldg(bf.torque_limits + j) * pr.soft_torque_limit, 0.0f);
          }
        }
        acc(ELG_REW_ACTION_RATE) = s_action_rate;
        acc(ELG_REW_DOF_ACC) = s_dof_acc;
        acc(ELG_REW_DOF_POS_LIMITS) = s_pos_lim;
        acc(ELG_REW_DOF_VEL) = s_dof_vel;
        acc(ELG_REW_DOF_VEL_LIMITS) = s_vel_lim;
        acc(ELG_REW_STAND_STILL) = s_still * (cmd_xy < pr.stand_still_threshold ? 1.0f : 0.0f);
        acc(ELG_REW_TORQUE_LIMITS) = s_tq_lim;
        acc(ELG_REW_TORQUES) = s_tq2;
        // ---- base terms
        acc(ELG_REW_LIN_VEL_Z) = blv.z * blv.z;
        acc(ELG_REW_ANG_VEL_XY) = bav.x * bav.x + bav.y * bav.y;
        acc(ELG_REW_ORIENTATION) = pg.x * pg.x + pg.y * pg.y;
        {
          const float ex = cmd0 - blv.x, ey = cmd1 - blv.y, ez = cmd2 - bav.z;
          acc(ELG_REW_TRACKING_LIN_VEL) = expf(-(ex * ex + ey * ey) / pr.tracking_sigma);
          acc(ELG_REW_TRACKING_ANG_VEL) = expf(-(ez * ez) / pr.tracking_sigma);
        }
        acc(ELG_REW_TERMINATION) = (reset && !time_out) ? 1.0f : 0.0f;
        // ---- collision (legged_robot_rew_mixin.py:117-119)
        if (term_on(pr, ELG_REW_COLLISION)) {
          float n = 0.0f;
          for (int p = 0; p < dm.num_penalised; ++p) {
            const float* f = s_cf + (slot * B + dm.penalised_idx[p]) * 3;
            n += norm3_t(f[0], f[1], f[2]) > 0.1f ? 1.0f : 0.0f;
          }
          acc(ELG_REW_COLLISION) = n;
        }
        // ---- feet (legged_robot_rew_mixin.py:58-81, :121-212; gait_scheduler.py:74-81)
        // Terms that sort before feet_air_time read the OLD timers, terms after it the updated ones
        // and the rebound last_contacts.
        {
          const bool air_on = term_on(pr, ELG_REW_FEET_AIR_TIME);
          const bool gs_on = term_on(pr, ELG_REW_GAIT_SCHEDULER) && gait;
          float bfh_sum = 0.0f, bfh_cnt = 0.0f;
          float r_air = 0.0f, r_cf = 0.0f, r_slip = 0.0f, r_lift = 0.0f, r_jump = 0.0f, r_gs = 0.0f;
          bool any_stumble = false, all_up = true;
          float a0 = 0, a1 = 0, a2 = 0, a3 = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0;   // updated timers of feet 0..3
          const float gait_phase0 = gs_on ? s_gidx[slot] : 0.0f;
          for (int f = 0; f < F; ++f) {
            const float* cf = s_cf + (slot * B + dm.feet_idx[f]) * 3;
            const float* ft = s_feet + (slot * F + f) * 6;
            const float fxx = cf[0], fyy = cf[1], fz = cf[2];
            const float pz = ft[2], vx = ft[3], vy = ft[4], vz = ft[5];
            const int sf = slot * F + f;
            const size_t ef = (size_t)env * F + f;
            float air = s_air[sf], con = s_con[sf];
            const bool last_c = s_lc[sf] != 0;
            const bool contact = fz > 1.0f;
            if (con > 1e-3f) { bfh_sum += pz; bfh_cnt += 1.0f; }   // base_foot_height: nanmean over touching feet
            bool lc_after = last_c;
            if (air_on) {
              const bool filt = contact | last_c;
              const bool first = (air > 0.0f) && filt;
              air += pr.dt;
              con += pr.dt;
              r_air += (air - 0.5f) * (first ? 1.0f : 0.0f);
              air *= filt ? 0.0f : 1.0f;
              con *= filt ? 1.0f : 0.0f;
              bf.feet_air_time[ef] = air;
              bf.feet_contact_time[ef] = con;
              bf.last_contacts[ef] = contact ? 1 : 0;
              lc_after = contact;
            }
            const bool filt2 = contact | lc_after;
            r_cf += fmaxf(norm3_t(fxx, fyy, fz) - pr.max_contact_force, 0.0f);
            {
              const float vn = norm2_t(vx, vy);
              r_slip += (filt2 ? 1.0f : 0.0f) * (vn * vn);
            }
            const bool stumble = norm2_t(fxx, fyy) > mul_r(5.0f, fabsf(fz));
            any_stumble |= stumble;
            r_lift += (stumble ? 1.0f : 0.0f) * vz;
            all_up &= fz < 1.0f;
            r_jump += (filt2 ? 0.0f : 1.0f) * (air - 0.5f);
            if (f == 0) { a0 = air; c0 = con; } else if (f == 1) { a1 = air; c1 = con; }
            else if (f == 2) { a2 = air; c2 = con; } else if (f == 3) { a3 = air; c3 = con; }
            if (gs_on) {
              float ph = gait_phase0 + pr.gait_foot_phases[f];
              ph = ph - floorf(ph);                                 // torch.remainder(x, 1.0)
              const float target = ph < 0.5f ? pr.gait_swing_height * sinf(6.283185307179586f * ph) : 0.0f;
              const float dz = target - s_gprev[sf];
              r_gs += dz * dz;
            }
            if (gait) bf.gait_prev_foot_z[ef] = pz;   // GaitScheduler.step keeps this step's feet
          }
          {
            const float ground = bfh_cnt > 0.0f ? bfh_sum / bfh_cnt : rootz - pr.base_height_target;
            const float rel = rootz - ground - pr.base_height_target;
            acc(ELG_REW_BASE_FOOT_HEIGHT) = rel * rel;
          }
          acc(ELG_REW_FEET_AIR_TIME) = r_air * (cmd_xy > 0.1f ? 1.0f : 0.0f);
          acc(ELG_REW_FEET_CONTACT_FORCES) = r_cf;
          acc(ELG_REW_FEET_SLIP) = r_slip;
          acc(ELG_REW_FEET_STUMBLE) = any_stumble ? 1.0f : 0.0f;
          acc(ELG_REW_FEET_STUMBLE_LIFTUP) = r_lift;
          acc(ELG_REW_FOUR_FOOTUP) = all_up ? 0.1f : 0.0f;
          acc(ELG_REW_JUMP_AIR) = fmaxf(r_jump - (float)F / 2.0f, 0.0f);
          acc(ELG_REW_GAIT_SCHEDULER) = r_gs;
          {
            // gait_2_step (legged_robot_rew_mixin.py:170-206): FL/RR and FR/RL in phase, the rest anti-phase
            auto sq4 = [](float a, float b) { const float d = a - b; return fminf(d * d, 4.0f); };
            const float s = ((sq4(a0, a3) + sq4(c0, c3)) + (sq4(a1, a2) + sq4(c1, c2))) / 2.0f;
            const float a = ((sq4(a0, c1) + sq4(c0, a1)) + (sq4(a0, c2) + sq4(c0, a2)) + (sq4(a3, c2) + sq4(c3, a2)) +
                             (sq4(a3, c1) + sq4(c3, a1))) / 4.0f;
            const float yawish = pr.heading_command ? cmd3 : cmd2;
            const bool moving = (cmd_xy > pr.speed_min) | (fabsf(yawish) >= pr.speed_min / 2.0f);
            acc(ELG_REW_GAIT_2_STEP) = (s + a) * (moving ? 1.0f : 0.0f);
          }
        }
        if (gait) {   // GaitScheduler.step (gait_scheduler.py:63-72) runs after the env step
          const float g = s_gidx[slot] + pr.gait_increment;
          bf.gait_idx[env] = g - floorf(g);
        }
      }
    }

    if (need_hsum) __syncthreads();   // (A)
    if (active && do_reward) {
      if (need_hsum) {
        const float d = s_hsum[slot] / (float)H - pr.base_height_target;
        acc(ELG_REW_BASE_HEIGHT) = d * d;
      }
      // ---- weighted sum in registry (alphabetical) order (legged_robot.py:220-232)
      asm volatile("cp.async.wait_all;" ::: "memory");
      float total = 0.0f;
#pragma unroll
      for (int t = 0; t < ELG_NUM_REWARD_TERMS; ++t) {
        if (t == ELG_REW_TERMINATION) continue;
        if (term_on(pr, t)) {
          const float r = acc(t) * pr.reward_scales[t];
          total += r;
          bf.episode_sums[(size_t)t * N + env] = s_sums[t * EPB + slot] + r;
        }
      }
      if (bf.extra_reward) total += bf.extra_reward[env];
      if (pr.only_positive_rewards) total = fmaxf(total, 0.0f);
      if (term_on(pr, ELG_REW_TERMINATION)) {
        const float r = acc(ELG_REW_TERMINATION) * pr.reward_scales[ELG_REW_TERMINATION];
        total += r;
        bf.episode_sums[(size_t)ELG_REW_TERMINATION * N + env] = s_sums[ELG_REW_TERMINATION * EPB + slot] + r;
      }
      bf.rew_buf[env] = total;
    }

    if (active && (do_obs || do_hist)) {
      // ---- observation head into shared memory (legged_robot.py:237-244), history (:148-150)
      float* hrow = s_head + slot * head_pitch;
      if (do_obs) {
        hrow[0] = blv.x * pr.obs_scale_lin_vel; hrow[1] = blv.y * pr.obs_scale_lin_vel; hrow[2] = blv.z * pr.obs_scale_lin_vel;
        hrow[3] = bav.x * pr.obs_scale_ang_vel; hrow[4] = bav.y * pr.obs_scale_ang_vel; hrow[5] = bav.z * pr.obs_scale_ang_vel;
        hrow[6] = pg.x; hrow[7] = pg.y; hrow[8] = pg.z;
        hrow[9] = cmd0 * pr.commands_scale[0]; hrow[10] = cmd1 * pr.commands_scale[1]; hrow[11] = cmd2 * pr.commands_scale[2];
      }
      for (int j = 0; j < D; ++j) {
        const int e = slot * D + j;
        const float pos = s_dof[2 * e], vel = s_dof[2 * e + 1], a = s_act[e];
        if (do_obs) {
          hrow[12 + j] = (pos - __ldg(bf.default_dof_pos + j)) * pr.obs_scale_dof_pos;
          hrow[12 + D + j] = vel * pr.obs_scale_dof_vel;
          hrow[12 + 2 * D + j] = a;
        }
        if (do_hist) {
          bf.last_actions[(size_t)env * D + j] = a;
          bf.last_dof_vel[(size_t)env * D + j] = vel;
        }
      }
      if (do_hist) {
        float* lrv = bf.last_root_vel + (size_t)env * 6;
#pragma unroll
        for (int k = 0; k < 6; ++k) lrv[k] = rs[7 + k];
      }
    }
  }

  if (!do_obs) return;
  __syncthreads();   // (B) observation heads are in shared memory
  if (warp > 0) {
    float* __restrict__ obs = bf.obs_buf;
    for (int slot = warp - 1; slot < nenv; slot += kRowWarps) {
      const int env = env0 + slot;
      for (int k = lane; k < head; k += kWarp) {
        float u = 0.0f, ns = 0.0f;
        if (pr.noise_mode == ELG_NOISE_TENSOR) {
          u = __ldg(bf.noise_u + (size_t)env * O + k);
          ns = __ldg(bf.noise_scale_vec + k);
        } else if (pr.noise_mode == ELG_NOISE_PHILOX) {
          const int kk = k >> 5;
          const uint4 rnd = noise_block(pr.noise_seed, pr.noise_offset, env, k & 31, kk >> 2);
          u = u01(pick(rnd, kk & 3));
          ns = __ldg(bf.noise_scale_vec + k);
        }
        obs[(size_t)env * O + k] = finish_obs(s_head[slot * head_pitch + k], u, ns, pr);
      }
    }
  }
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
  // TMA bulk staging needs 16-byte aligned array bases (chunk offsets/sizes are multiples of 16 by construction
  // for full CTAs when F is even); otherwise the kernel falls back to element-wise staging
  auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  int use_bulk = a16(buf->root_states) && a16(buf->dof_state) && a16(buf->actions) && a16(buf->last_actions) &&
                 a16(buf->last_dof_vel) && a16(buf->torques) && a16(buf->contact_forces) && a16(buf->last_root_vel) &&
                 a16(buf->base_lin_acc) && a16(buf->base_ang_acc) && a16(buf->commands) && a16(buf->feet_air_time) &&
                 a16(buf->feet_contact_time) && a16(buf->last_contacts) && a16(buf->episode_length_buf) &&
                 (!buf->gait_idx || a16(buf->gait_idx)) && (!buf->gait_prev_foot_z || a16(buf->gait_prev_foot_z)) &&
                 (dims->num_feet % 2 == 0) && dims->num_feet > 0;
  // environments per CTA: few for small N (several CTAs on each of the 148 SMs), many for large N
  auto launch = [&](auto kernel, int epb) -> int {
    const size_t smem = (size_t)elg::make_layout(*dims, epb).words * 4;
    if (smem > 48 * 1024) {
      if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return fail(ELG_ERR_CUDA, "cannot reserve dynamic shared memory for elg_step_kernel");
    }
    kernel<<<(N + epb - 1) / epb, elg::kStepThreads, smem, st>>>(*dims, *prm, *buf, phase, use_bulk);
    return ELG_OK;
  };
  int rc;
  if (N <= 12288) rc = launch(elg::elg_step_kernel<8>, 8);
  else if (N <= 32768) rc = launch(elg::elg_step_kernel<16>, 16);
  else rc = launch(elg::elg_step_kernel<32>, 32);
  if (rc) return rc;
  return check_launch("elg_post_physics_step");
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
