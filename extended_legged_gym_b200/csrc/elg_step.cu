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
constexpr int kMaxHead = 12 + 3 * ELG_MAX_DOF;         // observation entries produced by the state warp

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
// the fused kernel
// ---------------------------------------------------------------------------------------------
template <int EPB>
__global__ void __launch_bounds__(kStepThreads, 3)
elg_step_kernel(const __grid_constant__ ElgDims dm, const __grid_constant__ ElgStepParams pr,
                const __grid_constant__ ElgStepBuffers bf, const uint32_t phase) {
  __shared__ float s_head[EPB][kMaxHead + 1];  // +1: odd row pitch, the state warp writes column-wise
  __shared__ float s_hsum[EPB];                // sum over points of (z - height), for _reward_base_height

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int env0 = blockIdx.x * EPB;
  const int N = dm.num_envs, D = dm.num_dof, B = dm.num_bodies, F = dm.num_feet, H = dm.num_height_points;
  const int O = dm.num_obs, C = dm.num_commands;
  const int head = 12 + 3 * D;
  const bool do_derive = phase & ELG_PHASE_DERIVE, do_term = phase & ELG_PHASE_TERMINATION;
  const bool do_reward = phase & ELG_PHASE_REWARD, do_obs = phase & ELG_PHASE_OBS, do_hist = phase & ELG_PHASE_HISTORY;
  const bool need_hsum = do_reward && term_on(pr, ELG_REW_BASE_HEIGHT) && H > 0;   // CTA-uniform

  if (warp > 0) {
    // =========================== row warps: heights + observation rows ===========================
    if (H > 0 && (do_derive || do_obs || need_hsum)) {
      for (int slot = warp - 1; slot < EPB; slot += kRowWarps) {
        const int env = env0 + slot;
        if (env >= N) break;
        const float* rs = bf.root_states + (size_t)env * 13;
        float hsum = 0.0f;
        YawFrame fr;
        if (do_derive) fr = make_yaw_frame(rs);
        const float rootz = __ldg(rs + 2);              // re-read: reset_idx may have moved the robot
        const float zc = sub_r(rootz, 0.5f);
        const float* hp = bf.height_points + (size_t)env * pr.height_points_env_stride;
        uint4 rnd = make_uint4(0, 0, 0, 0);
        for (int p = lane, it = 0; p < H; p += kWarp, ++it) {
          float h;
          if (do_derive) {
            if (pr.terrain_is_plane) {
              h = 0.0f;
            } else {
              int ix, iy;
              terrain_cell(fr, __ldg(hp + 3 * p), __ldg(hp + 3 * p + 1), pr, ix, iy);
              h = cell_height(bf.height_samples, ix, iy, pr);
            }
            bf.measured_heights[(size_t)env * H + p] = h;
          } else {
            h = bf.measured_heights[(size_t)env * H + p];
          }
          hsum += sub_r(rootz, h);
          if (do_obs) {
            const int k = head + p;
            float v = mul_r(fminf(fmaxf(sub_r(zc, h), -1.0f), 1.0f), pr.obs_scale_height);
            float u = 0.0f, ns = 0.0f;
            if (pr.noise_mode == ELG_NOISE_TENSOR) {
              u = __ldg(bf.noise_u + (size_t)env * O + k);
              ns = __ldg(bf.noise_scale_vec + k);
            } else if (pr.noise_mode == ELG_NOISE_PHILOX) {
              const int kk = k >> 5;
              if ((kk & 3) == 0 || it == 0) rnd = noise_block(pr.noise_seed, pr.noise_offset, env, k & 31, kk >> 2);
              u = u01(pick(rnd, kk & 3));
              ns = __ldg(bf.noise_scale_vec + k);
            }
            bf.obs_buf[(size_t)env * O + k] = finish_obs(v, u, ns, pr);
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
    const bool active = slot < EPB && env < N;
    Vec3 blv = {0, 0, 0}, bav = {0, 0, 0}, pg = {0, 0, 0};
    float cmd0 = 0, cmd1 = 0, cmd2 = 0, cmd3 = 0;
    float acc[ELG_NUM_REWARD_TERMS];   // raw (unscaled) value of every built-in term
#pragma unroll
    for (int t = 0; t < ELG_NUM_REWARD_TERMS; ++t) acc[t] = 0.0f;
    bool reset = false, time_out = false;
    const float* rs = bf.root_states + (size_t)(active ? env : 0) * 13;

    if (active) {
      if (do_derive) {
        const Quat q = {rs[3], rs[4], rs[5], rs[6]};
        const Vec3 lin = {rs[7], rs[8], rs[9]}, ang = {rs[10], rs[11], rs[12]};
        // ---- episode counter + derived base state (legged_robot.py:122-134)
        bf.episode_length_buf[env] += 1;
        const float* lrv = bf.last_root_vel + (size_t)env * 6;
        blv = quat_rotate_inverse(q, lin);
        bav = quat_rotate_inverse(q, ang);
        pg = quat_rotate_inverse(q, Vec3{pr.gravity_vec[0], pr.gravity_vec[1], pr.gravity_vec[2]});
        {
          const Vec3 dl = quat_rotate_inverse(q, Vec3{lin.x - lrv[0], lin.y - lrv[1], lin.z - lrv[2]});
          const Vec3 da = quat_rotate_inverse(q, Vec3{ang.x - lrv[3], ang.y - lrv[4], ang.z - lrv[5]});
          const float ema = pr.acc_ema, w1 = pr.acc_ema_c;
          float* la = bf.base_lin_acc + (size_t)env * 3;
          float* aa = bf.base_ang_acc + (size_t)env * 3;
          la[0] = la[0] * ema + (w1 * dl.x) / pr.dt;
          la[1] = la[1] * ema + (w1 * dl.y) / pr.dt;
          la[2] = la[2] * ema + (w1 * dl.z) / pr.dt;
          aa[0] = aa[0] * ema + (w1 * da.x) / pr.dt;
          aa[1] = aa[1] * ema + (w1 * da.y) / pr.dt;
          aa[2] = aa[2] * ema + (w1 * da.z) / pr.dt;
        }
        float* o3;
        o3 = bf.base_lin_vel + (size_t)env * 3;      o3[0] = blv.x; o3[1] = blv.y; o3[2] = blv.z;
        o3 = bf.base_ang_vel + (size_t)env * 3;      o3[0] = bav.x; o3[1] = bav.y; o3[2] = bav.z;
        o3 = bf.projected_gravity + (size_t)env * 3; o3[0] = pg.x;  o3[1] = pg.y;  o3[2] = pg.z;
        // ---- feet gather (legged_robot.py:136-137)
        for (int f = 0; f < F; ++f) {
          const float* rb = bf.rigid_body_state + ((size_t)env * B + dm.feet_idx[f]) * 13;
          float* fp = bf.foot_positions + ((size_t)env * F + f) * 3;
          float* fv = bf.foot_velocities + ((size_t)env * F + f) * 3;
          fp[0] = rb[0]; fp[1] = rb[1]; fp[2] = rb[2];
          fv[0] = rb[7]; fv[1] = rb[8]; fv[2] = rb[9];
        }
        // ---- heading command (legged_robot.py:394-398); forward = quat_apply(q, (1,0,0))
        if (pr.heading_command) {
          float* cmd = bf.commands + (size_t)env * C;
          const float fx = 1.0f + (q.y * (-2.0f * q.y) - q.z * (2.0f * q.z));
          const float fy = q.w * (2.0f * q.z) + (q.z * 0.0f - q.x * (-2.0f * q.y));
          const float heading = atan2f(fy, fx);
          cmd[2] = fminf(fmaxf(0.5f * wrap_to_pi(cmd[3] - heading), -1.0f), 1.0f);
        }
      } else if (do_reward || do_obs) {
        const float* p3;
        p3 = bf.base_lin_vel + (size_t)env * 3;      blv = Vec3{p3[0], p3[1], p3[2]};
        p3 = bf.base_ang_vel + (size_t)env * 3;      bav = Vec3{p3[0], p3[1], p3[2]};
        p3 = bf.projected_gravity + (size_t)env * 3; pg = Vec3{p3[0], p3[1], p3[2]};
      }
      if (do_reward || do_obs) {
        const float* cmd = bf.commands + (size_t)env * C;
        cmd0 = cmd[0]; cmd1 = cmd[1]; cmd2 = cmd[2];
        if (C > 3) cmd3 = cmd[3];
      }

      // ---- termination (legged_robot.py:155-160)
      if (do_term) {
        bool contact_term = false;
        for (int t = 0; t < dm.num_termination; ++t) {
          const float* f = bf.contact_forces + ((size_t)env * B + dm.termination_idx[t]) * 3;
          contact_term |= norm3_t(f[0], f[1], f[2]) > 1.0f;
        }
        time_out = bf.episode_length_buf[env] > pr.max_episode_length;
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
        float s_action_rate = 0, s_dof_acc = 0, s_pos_lim = 0, s_dof_vel = 0, s_vel_lim = 0, s_still = 0, s_tq_lim = 0, s_tq = 0;
        const bool lim_terms = term_on(pr, ELG_REW_DOF_POS_LIMITS) | term_on(pr, ELG_REW_DOF_VEL_LIMITS) | term_on(pr, ELG_REW_TORQUE_LIMITS);
        for (int j = 0; j < D; ++j) {
          const size_t e = (size_t)env * D + j;
          const float2 pv = *reinterpret_cast<const float2*>(bf.dof_state + 2 * e);
          const float a = bf.actions[e], la = bf.last_actions[e], lv = bf.last_dof_vel[e], tq = bf.torques[e];
          const float da = la - a;
          s_action_rate += da * da;
          const float dv = (lv - pv.y) / pr.dt;
          s_dof_acc += dv * dv;
          s_dof_vel += pv.y * pv.y;
          s_tq += tq * tq;
          s_still += fabsf(pv.x - __ldg(bf.default_dof_pos + j));
          if (lim_terms) {
            const float lo = __ldg(bf.dof_pos_limits + 2 * j), hi = __ldg(bf.dof_pos_limits + 2 * j + 1);
            s_pos_lim += -fminf(pv.x - lo, 0.0f) + fmaxf(pv.x - hi, 0.0f);
            s_vel_lim += fminf(fmaxf(fabsf(pv.y) - __ldg(bf.dof_vel_limits + j) * pr.soft_dof_vel_limit, 0.0f), 1.0f);
            s_tq_lim += fmaxf(fabsf(tq) - __ldg(bf.torque_limits + j) * pr.soft_torque_limit, 0.0f);
          }
        }
        acc[ELG_REW_ACTION_RATE] = s_action_rate;
        acc[ELG_REW_DOF_ACC] = s_dof_acc;
        acc[ELG_REW_DOF_POS_LIMITS] = s_pos_lim;
        acc[ELG_REW_DOF_VEL] = s_dof_vel;
        acc[ELG_REW_DOF_VEL_LIMITS] = s_vel_lim;
        acc[ELG_REW_STAND_STILL] = s_still * (cmd_xy < pr.stand_still_threshold ? 1.0f : 0.0f);
        acc[ELG_REW_TORQUE_LIMITS] = s_tq_lim;
        acc[ELG_REW_TORQUES] = s_tq;
        // ---- base terms
        acc[ELG_REW_LIN_VEL_Z] = blv.z * blv.z;
        acc[ELG_REW_ANG_VEL_XY] = bav.x * bav.x + bav.y * bav.y;
        acc[ELG_REW_ORIENTATION] = pg.x * pg.x + pg.y * pg.y;
        {
          const float ex = cmd0 - blv.x, ey = cmd1 - blv.y, ez = cmd2 - bav.z;
          acc[ELG_REW_TRACKING_LIN_VEL] = expf(-(ex * ex + ey * ey) / pr.tracking_sigma);
          acc[ELG_REW_TRACKING_ANG_VEL] = expf(-(ez * ez) / pr.tracking_sigma);
        }
        acc[ELG_REW_TERMINATION] = (reset && !time_out) ? 1.0f : 0.0f;
        // ---- collision (legged_robot_rew_mixin.py:117-119)
        if (term_on(pr, ELG_REW_COLLISION)) {
          float n = 0.0f;
          for (int p = 0; p < dm.num_penalised; ++p) {
            const float* f = bf.contact_forces + ((size_t)env * B + dm.penalised_idx[p]) * 3;
            n += norm3_t(f[0], f[1], f[2]) > 0.1f ? 1.0f : 0.0f;
          }
          acc[ELG_REW_COLLISION] = n;
        }
        // ---- feet (legged_robot_rew_mixin.py:58-81, :121-212; gait_scheduler.py:74-81)
        // Terms that sort before feet_air_time read the OLD timers, terms after it the updated ones
        // and the rebound last_contacts.
        {
          const bool air_on = term_on(pr, ELG_REW_FEET_AIR_TIME);
          const bool gs_on = term_on(pr, ELG_REW_GAIT_SCHEDULER) && bf.gait_prev_foot_z && bf.gait_idx;
          float bfh_sum = 0.0f, bfh_cnt = 0.0f;
          float r_air = 0.0f, r_cf = 0.0f, r_slip = 0.0f, r_lift = 0.0f, r_jump = 0.0f, r_gs = 0.0f;
          bool any_stumble = false, all_up = true;
          float a0 = 0, a1 = 0, a2 = 0, a3 = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0;   // updated timers of feet 0..3
          const float gait_phase0 = gs_on ? bf.gait_idx[env] : 0.0f;
          for (int f = 0; f < F; ++f) {
            const size_t body = (size_t)env * B + dm.feet_idx[f];
            const float* cf = bf.contact_forces + body * 3;
            const float* rb = bf.rigid_body_state + body * 13;
            const float fxx = cf[0], fyy = cf[1], fz = cf[2];
            const float pz = rb[2], vx = rb[7], vy = rb[8], vz = rb[9];
            const size_t ef = (size_t)env * F + f;
            float air = bf.feet_air_time[ef], con = bf.feet_contact_time[ef];
            const bool last_c = bf.last_contacts[ef] != 0;
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
              const float dz = target - bf.gait_prev_foot_z[ef];
              r_gs += dz * dz;
            }
            if (bf.gait_prev_foot_z) bf.gait_prev_foot_z[ef] = pz;   // GaitScheduler.step keeps this step's feet
          }
          {
            const float ground = bfh_cnt > 0.0f ? bfh_sum / bfh_cnt : rootz - pr.base_height_target;
            const float rel = rootz - ground - pr.base_height_target;
            acc[ELG_REW_BASE_FOOT_HEIGHT] = rel * rel;
          }
          acc[ELG_REW_FEET_AIR_TIME] = r_air * (cmd_xy > 0.1f ? 1.0f : 0.0f);
          acc[ELG_REW_FEET_CONTACT_FORCES] = r_cf;
          acc[ELG_REW_FEET_SLIP] = r_slip;
          acc[ELG_REW_FEET_STUMBLE] = any_stumble ? 1.0f : 0.0f;
          acc[ELG_REW_FEET_STUMBLE_LIFTUP] = r_lift;
          acc[ELG_REW_FOUR_FOOTUP] = all_up ? 0.1f : 0.0f;
          acc[ELG_REW_JUMP_AIR] = fmaxf(r_jump - (float)F / 2.0f, 0.0f);
          acc[ELG_REW_GAIT_SCHEDULER] = r_gs;
          {
            // gait_2_step (legged_robot_rew_mixin.py:170-206): FL/RR and FR/RL in phase, the rest anti-phase
            auto sq4 = [](float a, float b) { const float d = a - b; return fminf(d * d, 4.0f); };
            const float s = ((sq4(a0, a3) + sq4(c0, c3)) + (sq4(a1, a2) + sq4(c1, c2))) / 2.0f;
            const float a = ((sq4(a0, c1) + sq4(c0, a1)) + (sq4(a0, c2) + sq4(c0, a2)) + (sq4(a3, c2) + sq4(c3, a2)) +
                             (sq4(a3, c1) + sq4(c3, a1))) / 4.0f;
            const float yawish = pr.heading_command ? cmd3 : cmd2;
            const bool moving = (cmd_xy > pr.speed_min) | (fabsf(yawish) >= pr.speed_min / 2.0f);
            acc[ELG_REW_GAIT_2_STEP] = (s + a) * (moving ? 1.0f : 0.0f);
          }
        }
        if (bf.gait_idx) {   // GaitScheduler.step (gait_scheduler.py:63-72) runs after the env step
          const float g = bf.gait_idx[env] + pr.gait_increment;
          bf.gait_idx[env] = g - floorf(g);
        }
      }
    }

    if (need_hsum) __syncthreads();   // (A)
    if (active && do_reward) {
      if (need_hsum) {
        const float d = s_hsum[slot] / (float)H - pr.base_height_target;
        acc[ELG_REW_BASE_HEIGHT] = d * d;
      }
      // ---- weighted sum in registry (alphabetical) order (legged_robot.py:220-232)
      float total = 0.0f;
#pragma unroll
      for (int t = 0; t < ELG_NUM_REWARD_TERMS; ++t) {
        if (t == ELG_REW_TERMINATION) continue;
        if (term_on(pr, t)) {
          const float r = acc[t] * pr.reward_scales[t];
          total += r;
          bf.episode_sums[(size_t)t * N + env] += r;
        }
      }
      if (bf.extra_reward) total += bf.extra_reward[env];
      if (pr.only_positive_rewards) total = fmaxf(total, 0.0f);
      if (term_on(pr, ELG_REW_TERMINATION)) {
        const float r = acc[ELG_REW_TERMINATION] * pr.reward_scales[ELG_REW_TERMINATION];
        total += r;
        bf.episode_sums[(size_t)ELG_REW_TERMINATION * N + env] += r;
      }
      bf.rew_buf[env] = total;
    }

    if (active && (do_obs || do_hist)) {
      // ---- observation head into shared memory (legged_robot.py:237-244), history (:148-150)
      float* hrow = s_head[slot];
      if (do_obs) {
        hrow[0] = blv.x * pr.obs_scale_lin_vel; hrow[1] = blv.y * pr.obs_scale_lin_vel; hrow[2] = blv.z * pr.obs_scale_lin_vel;
        hrow[3] = bav.x * pr.obs_scale_ang_vel; hrow[4] = bav.y * pr.obs_scale_ang_vel; hrow[5] = bav.z * pr.obs_scale_ang_vel;
        hrow[6] = pg.x; hrow[7] = pg.y; hrow[8] = pg.z;
        hrow[9] = cmd0 * pr.commands_scale[0]; hrow[10] = cmd1 * pr.commands_scale[1]; hrow[11] = cmd2 * pr.commands_scale[2];
      }
      for (int j = 0; j < D; ++j) {
        const size_t e = (size_t)env * D + j;
        const float2 pv = *reinterpret_cast<const float2*>(bf.dof_state + 2 * e);
        const float a = bf.actions[e];
        if (do_obs) {
          hrow[12 + j] = (pv.x - __ldg(bf.default_dof_pos + j)) * pr.obs_scale_dof_pos;
          hrow[12 + D + j] = pv.y * pr.obs_scale_dof_vel;
          hrow[12 + 2 * D + j] = a;
        }
        if (do_hist) {
          bf.last_actions[e] = a;
          bf.last_dof_vel[e] = pv.y;
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
    for (int slot = warp - 1; slot < EPB; slot += kRowWarps) {
      const int env = env0 + slot;
      if (env >= N) break;
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
        bf.obs_buf[(size_t)env * O + k] = finish_obs(s_head[slot][k], u, ns, pr);
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
  // environments per CTA: few for small N (fill all 148 SMs with several CTAs each), many for large N
  if (N <= 12288) {
    elg::elg_step_kernel<8><<<(N + 7) / 8, elg::kStepThreads, 0, st>>>(*dims, *prm, *buf, phase);
  } else if (N <= 32768) {
    elg::elg_step_kernel<16><<<(N + 15) / 16, elg::kStepThreads, 0, st>>>(*dims, *prm, *buf, phase);
  } else {
    elg::elg_step_kernel<32><<<(N + 31) / 32, elg::kStepThreads, 0, st>>>(*dims, *prm, *buf, phase);
  }
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
