// elg_reset.cu -- the sparse, RNG-driven branches of the step as kernels over ALL envs with a per-env predicate,
// so that a step needs no `nonzero()` and no host synchronisation:
//
//   elg_resample_commands   LeggedRobot._post_physics_step_callback's resampling (envs/base/legged_robot.py:389-393)
//                           + _resample_commands (:405-423) for the envs whose episode clock hits the interval
//   elg_reset_envs          LeggedRobot.reset_idx (:162-213): _update_terrain_curriculum (:498-518), _reset_dofs
//                           (:450-465), _reset_root_states (:467-487), _resample_commands, history / timer zeroing,
//                           the sums behind extras["episode"] (as (sum, count) atomics), and the repair of the
//                           observation entries that change when an env resets between rewards and observations
//                           (commands, dof_pos, dof_vel; base velocities and heights stay stale -- App. A-4).
//
// Uniform numbers: column c of a per-env table [N, ELG_RESET_UNIFORMS] supplied by the caller (parity tests feed the host
// path and the kernel the same numbers), or Philox4x32-10 keyed by (seed ^ stream tag) with counter (env, c / 4, step).
// Column layout: [0, D) dof scale | D, D+1 root xy | D+2 .. D+7 root velocity | D+8 cmd x | D+9 cmd y | D+10 heading or yaw |
// D+11 terrain level redraw.  torch_rand_float(lo, hi) = (hi - lo) * u + lo with one rounding per op.
#include <cuda_runtime.h>
#include <stdint.h>

#include "elg_common.cuh"
#include "elg_async.cuh"

namespace elg {

constexpr uint32_t kResetStream = 0x52455345u;   // "RESE": keeps these draws apart from the observation noise stream

struct Uniforms {
  const float* table;   // [N, ELG_RESET_UNIFORMS] or NULL
  uint64_t seed, offset;
  uint32_t env;
  uint4 blk;
  int blk_id;
  __device__ float get(int c) {
    if (table) return table[(size_t)env * ELG_RESET_UNIFORMS + c];
    const int b = c >> 2;
    if (b != blk_id) {
      blk = philox4x32_10(make_uint4(env, (uint32_t)b, (uint32_t)offset, (uint32_t)(offset >> 32)),
                          make_uint2((uint32_t)seed ^ kResetStream, (uint32_t)(seed >> 32)));
      blk_id = b;
    }
    const int w = c & 3;
    return u01(w == 0 ? blk.x : w == 1 ? blk.y : w == 2 ? blk.z : blk.w);
  }
};
__device__ __forceinline__ float rand_range(float lo, float hi, float u) { return add_r(mul_r(sub_r(hi, lo), u), lo); }

// _resample_commands (:405-423) for one env; returns the new (cmd0, cmd1) after the small-command zeroing
__device__ __forceinline__ void resample_one(const ElgResetParams& rp, float* cmd, Uniforms& U, int D) {
  float c0 = rand_range(rp.lin_vel_x[0], rp.lin_vel_x[1], U.get(D + 8));
  float c1 = rand_range(rp.lin_vel_y[0], rp.lin_vel_y[1], U.get(D + 9));
  if (rp.heading_command) cmd[3] = rand_range(rp.heading[0], rp.heading[1], U.get(D + 10));
  else cmd[2] = rand_range(rp.ang_vel_yaw[0], rp.ang_vel_yaw[1], U.get(D + 10));
  const float keep = norm2_t(c0, c1) > 0.2f ? 1.0f : 0.0f;
  cmd[0] = mul_r(c0, keep);
  cmd[1] = mul_r(c1, keep);
}

// the same numbers, one column per lane: lane c holds column c of the env's uniform row (and column c + 32 when the row is longer),
// a warp-uniform column index fetches its value with one shuffle -- one Philox evaluation per lane instead of one per access
struct LaneUniforms { float lo, hi; };
__device__ __forceinline__ float uniform_col(const float* table, uint64_t seed, uint64_t offset, uint32_t env, int c) {
  if (table) return table[(size_t)env * ELG_RESET_UNIFORMS + c];
  const uint4 blk = philox4x32_10_cold(make_uint4(env, (uint32_t)(c >> 2), (uint32_t)offset, (uint32_t)(offset >> 32)),
                                       make_uint2((uint32_t)seed ^ kResetStream, (uint32_t)(seed >> 32)));
  const int w = c & 3;
  return u01(w == 0 ? blk.x : w == 1 ? blk.y : w == 2 ? blk.z : blk.w);
}
__device__ __forceinline__ LaneUniforms lane_uniforms(const float* table, uint64_t seed, uint64_t offset, uint32_t env, int lane, int ncols) {
  LaneUniforms u{0.0f, 0.0f};
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {      // (rolled: one copy of the code)
    const int c = lane + 32 * h;
    const float v = c < ncols ? uniform_col(table, seed, offset, env, c) : 0.0f;
    if (h == 0) u.lo = v;
    else u.hi = v;
    if (ncols <= 32) break;
  }
  return u;
}
__device__ __forceinline__ uint4 noise_block_cold(uint64_t seed, uint64_t offset, uint32_t env, uint32_t lane, uint32_t chunk) {
  return philox4x32_10_cold(make_uint4(env, lane | (chunk << 5), (uint32_t)offset, (uint32_t)(offset >> 32)),
                            make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
__device__ __forceinline__ float uget(const LaneUniforms& u, int c) {   // c: warp-uniform, all lanes call
  return __shfl_sync(0xffffffffu, c < 32 ? u.lo : u.hi, c & 31);   // (c is uniform: every lane offers the same member)
}
// _resample_commands (:405-423) on values: (u_x, u_y, u_third) -> cmd0, cmd1 after the small-command zeroing, heading or yaw rate
__device__ __forceinline__ void resample_vals(const ElgResetParams& rp, float ux, float uy, float ut, float& cmd0, float& cmd1, float& third) {
  const float c0 = rand_range(rp.lin_vel_x[0], rp.lin_vel_x[1], ux);
  const float c1 = rand_range(rp.lin_vel_y[0], rp.lin_vel_y[1], uy);
  third = rp.heading_command ? rand_range(rp.heading[0], rp.heading[1], ut) : rand_range(rp.ang_vel_yaw[0], rp.ang_vel_yaw[1], ut);
  const float keep = norm2_t(c0, c1) > 0.2f ? 1.0f : 0.0f;
  cmd0 = mul_r(c0, keep);
  cmd1 = mul_r(c1, keep);
}

__global__ void __launch_bounds__(128)
elg_resample_kernel(const __grid_constant__ ElgResetParams rp, const int N, const int D, const int C, const int64_t* __restrict__ ep_len,
                    float* __restrict__ commands, const float* __restrict__ uniforms, float* __restrict__ stats_to_zero) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  pdl_launch_dependents();   // programmatic dependent launch: the next kernel's prologue may start; nothing is read before the wait
  pdl_wait();
  if (stats_to_zero && e < ELG_NUM_REWARD_TERMS + 2) stats_to_zero[e] = 0.0f;   // this step's extras["episode"] accumulators
  if (e >= N) return;
  // main / rollout layout: the clock and the random numbers are the MAIN env's; every row of the group ends up with the main's
  // command row (robot_batch_rollout.py:819-838 resamples main envs and copies `commands[main]` to its rollouts)
  const int m = rp.rows_per_main > 0 ? (e / rp.rows_per_main) * rp.rows_per_main : e;
  if ((ep_len[m] + 1) % rp.resample_interval != 0) return;   // the step kernel increments the clock afterwards (:122, :391)
  Uniforms U{uniforms, rp.seed, rp.offset * 2, (uint32_t)m, make_uint4(0, 0, 0, 0), -1};
  float* cmd = commands + (size_t)e * C;
  if (m != e) {   // the column the resampling leaves alone is copied from the main row as well (nobody writes it in this kernel)
    const int keep = rp.heading_command ? 2 : 3;
    if (keep < C) cmd[keep] = commands[(size_t)m * C + keep];
  }
  resample_one(rp, cmd, U, D);
}

// main / rollout layout: does any MAIN env reset in this step?  (reset_idx zeroes the episode sums of the reset rows only then,
// robot_batch_rollout.py:925-931.)  One CTA; the flag is word ELG_NUM_REWARD_TERMS + 1 of the stats vector.
__global__ void __launch_bounds__(1024)
elg_main_reset_flag_kernel(const uint8_t* __restrict__ reset_buf, const int num_main, const int rows_per_main, float* __restrict__ stats) {
  pdl_launch_dependents();
  pdl_wait();
  int any = 0;
  for (int k = threadIdx.x; k < num_main; k += blockDim.x) any |= reset_buf[(size_t)k * rows_per_main];
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) stats[ELG_NUM_REWARD_TERMS + 1] = any ? 1.0f : 0.0f;
}

// One flagged row, handled by one WARP (elg_reset_kernel hands the rows of a 256-env CTA to its warps round-robin): all lanes compute
// the scalar part redundantly from broadcast loads and shuffled uniforms, lane j draws joint j, lane t owns reward term t, and all
// lanes repair the observation row: commands, dof_pos, dof_vel and the height entries, which depend on the new base height
// (compute_observations runs after reset_idx on stale measured_heights, App. A-4), with the very noise samples the step kernel
// attached to those entries.  (Round 1 launched one warp per env: at 65 536 envs that is 65 536 warps to find ~300 resets.)
__device__ __forceinline__ void reset_env(const ElgResetParams& rp, const ElgStepParams& pr, const ElgResetBuffers& rb, const int e, const int lane,
                                          const int N, const int D, const int F, const int C, const int O, const int H) {
  // main / rollout layout (robot_batch_rollout.py:876-940): any reset row gets new joints / root / histories; the terrain
  // curriculum, the command resampling and the extras sums belong to MAIN envs, and a main's new commands go to all its rows
  const int R1 = rp.rows_per_main;
  const int m = R1 > 0 ? (e / R1) * R1 : e;
  const bool is_main = m == e;
  const bool self_reset = rb.reset_buf[e] != 0;
  const bool main_reset = is_main ? self_reset : rb.reset_buf[m] != 0;
  if (!self_reset && !main_reset) return;
  float* rs = rb.root_states + (size_t)e * 13;
  float* cmd = rb.commands + (size_t)e * C;
  float* org = rb.env_origins + (size_t)e * 3;
  float* ds = rb.dof_state + (size_t)e * D * 2;
  if (!self_reset) {
    // a rollout row whose main env resets: only its command row (and the observation entries built from it) change
    const LaneUniforms Um = lane_uniforms(rb.uniforms, rp.seed, rp.offset * 2 + 1, (uint32_t)m, lane, D + 12);
    const int keep = rp.heading_command ? 2 : 3;
    float c0, c1, third, c2 = C > 2 ? cmd[2] : 0.0f;
    const float keep_val = keep < C ? rb.commands[(size_t)m * C + keep] : 0.0f;
    resample_vals(rp, uget(Um, D + 8), uget(Um, D + 9), uget(Um, D + 10), c0, c1, third);
    if (keep == 2 && keep < C) c2 = keep_val;
    if (!rp.heading_command) c2 = third;
    __syncwarp();
    if (lane == 0) {
      if (keep < C) cmd[keep] = keep_val;
      cmd[0] = c0; cmd[1] = c1;
      if (rp.heading_command) cmd[3] = third;
      else cmd[2] = third;
    }
    if (rb.obs_buf && lane >= 9 && lane < 12) {
      const int k = lane;
      float v = (k == 9 ? c0 : k == 10 ? c1 : c2) * pr.commands_scale[k - 9];
      if (pr.noise_mode == ELG_NOISE_TENSOR) v = v + (2.0f * rb.noise_u[(size_t)e * O + k] - 1.0f) * rb.noise_scale_vec[k];
      else if (pr.noise_mode == ELG_NOISE_PHILOX) {
        const int nj = (H + 31) >> 5, hm = (12 + 3 * D + 31) >> 5;
        const bool share = (nj & 7) + hm <= 8;
        const uint4 b = noise_block_cold(pr.noise_seed, pr.noise_offset, e, lane, share ? 1 + (nj >> 3) : 0);
        v = v + sym16(b, share ? (nj & 7) : 0) * rb.noise_scale_vec[k];
      }
      if (pr.clip_observations > 0.0f) v = fminf(fmaxf(v, -pr.clip_observations), pr.clip_observations);
      rb.obs_buf[(size_t)e * O + k] = v;
    }
    return;
  }
  // Everything scalar about the env is computed by ALL lanes from broadcast loads and warp-shuffled random numbers (column c of the
  // env's uniform row lives in lane c), in registers: the loads go out together, nothing is read back from global memory, and lane 0
  // stores the results at the end.  (The first form did the scalar work in lane 0 with read-modify-write chains through global
  // memory -- some twenty dependent round trips: 24.7 us per launch at 65 536 envs whatever the number of resets.)
  const bool zero_sums = R1 <= 0 || rb.stats[ELG_NUM_REWARD_TERMS + 1] != 0.0f;
  // Every global value the rest of the function reads is requested HERE, before the first store: the stores below may alias the
  // loads as far as the compiler knows, so a load written after them waits for its own DRAM round trip -- the height entries alone
  // were six such trips in a row (profiles/README.md r2o: 18 us per launch, 60 % of the samples on the first use of a loaded value).
  constexpr int kMaxNJ = 8, kMaxHM = 4;      // H <= 256, head = 12 + 3 D <= 108 (checked by the host)
  static_assert(ELG_NUM_REWARD_TERMS <= 32, "one lane per reward term");
  const int head = 12 + 3 * D;
  const int nj = (H + 31) >> 5, hm = (head + 31) >> 5;
  const bool repair = rb.obs_buf != nullptr;
  const bool philox = pr.noise_mode == ELG_NOISE_PHILOX, tensor_noise = pr.noise_mode == ELG_NOISE_TENSOR;
  const bool noisy = philox || tensor_noise;
  float hv[kMaxNJ], nsv[kMaxNJ], nuv[kMaxNJ], nsh[kMaxHM], nuh[kMaxHM];
#pragma unroll
  for (int j = 0; j < kMaxNJ; ++j) {
    const int p = lane + 32 * j;
    const bool ok = repair && rb.measured_heights && j < nj && p < H;
    hv[j] = ok ? rb.measured_heights[(size_t)e * H + p] : 0.0f;
    nsv[j] = (ok && noisy) ? rb.noise_scale_vec[head + p] : 0.0f;
    nuv[j] = (ok && tensor_noise) ? rb.noise_u[(size_t)e * O + head + p] : 0.0f;
  }
#pragma unroll
  for (int mm = 0; mm < kMaxHM; ++mm) {
    const int k = lane + 32 * mm;
    const bool ok = repair && mm < hm && k >= 9 && k < 12 + 2 * D;
    nsh[mm] = (ok && noisy) ? rb.noise_scale_vec[k] : 0.0f;
    nuh[mm] = (ok && tensor_noise) ? rb.noise_u[(size_t)e * O + k] : 0.0f;
  }
  const float defpos = lane < D ? rb.default_dof_pos[lane] : 0.0f;
  const bool term_on = lane < ELG_NUM_REWARD_TERMS && ((pr.reward_mask >> lane) & 1u);
  float* const sum_ptr = rb.episode_sums + (size_t)(term_on ? lane : 0) * N + e;
  const float sum_mine = term_on ? *sum_ptr : 0.0f;
  const LaneUniforms U = lane_uniforms(rb.uniforms, rp.seed, rp.offset * 2 + 1, (uint32_t)e, lane, D + 12);
  float o0 = org[0], o1 = org[1], o2 = org[2];
  float cm0 = cmd[0], cm1 = cmd[1], cm2 = C > 2 ? cmd[2] : 0.0f;
  // ---- _update_terrain_curriculum (:498-518): old root position, old commands
  long long lv = 0;
  const bool curr = rp.curriculum && is_main;
  if (curr) {
    const float r0 = rs[0], r1 = rs[1];
    const long long lv0 = rb.terrain_levels[e], ty = rb.terrain_types[e];
    const float dist = norm2_t(sub_r(r0, o0), sub_r(r1, o1));
    const bool up = dist > rp.env_length_half;
    const bool down = (dist < mul_r(mul_r(norm2_t(cm0, cm1), rp.max_episode_length_s), 0.5f)) && !up;
    lv = lv0 + (up ? 1 : 0) - (down ? 1 : 0);
    if (lv >= rp.max_terrain_level) {
      lv = (long long)(uget(U, D + 11) * (float)rp.max_terrain_level);
      if (lv >= rp.max_terrain_level) lv = rp.max_terrain_level - 1;
    } else if (lv < 0) {
      lv = 0;
    }
    const float* to = rb.terrain_origins + ((size_t)lv * rp.terrain_cols + ty) * 3;
    o0 = to[0]; o1 = to[1]; o2 = to[2];
  }
  // ---- _reset_root_states (:467-487)
  float r[13];
#pragma unroll
  for (int k = 0; k < 13; ++k) r[k] = rp.base_init_state[k];
  r[0] = add_r(r[0], o0); r[1] = add_r(r[1], o1); r[2] = add_r(r[2], o2);
  if (rp.custom_origins) {
    r[0] = add_r(r[0], rand_range(-0.5f, 0.5f, uget(U, D)));
    r[1] = add_r(r[1], rand_range(-0.5f, 0.5f, uget(U, D + 1)));
    if (rp.root_z_from_terrain) {   // (robot_batch_rollout.py:1379-1392): .long() truncation, clip to [0, dim - 2], one cell
      const long long cx = (long long)div_r(add_r(r[0], pr.border_size), pr.horizontal_scale);
      const long long cy = (long long)div_r(add_r(r[1], pr.border_size), pr.horizontal_scale);
      const long long px = cx < 0 ? 0 : (cx > pr.hf_rows - 2 ? pr.hf_rows - 2 : cx);
      const long long py = cy < 0 ? 0 : (cy > pr.hf_cols - 2 ? pr.hf_cols - 2 : cy);
      r[2] = add_r(mul_r((float)rb.height_samples[px * pr.hf_cols + py], pr.vertical_scale), rp.base_init_state[2]);
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) r[7 + k] = rand_range(-0.5f, 0.5f, uget(U, D + 2 + k));
  // ---- _resample_commands: main envs draw, rollout rows take their main's draw when it resets too
  const int keep = rp.heading_command ? 2 : 3;
  bool new_cmd = false, copy_keep = false;
  float keep_val = 0.0f, third = 0.0f;
  if (is_main) {
    resample_vals(rp, uget(U, D + 8), uget(U, D + 9), uget(U, D + 10), cm0, cm1, third);
    new_cmd = true;
  } else if (main_reset) {
    const LaneUniforms Um = lane_uniforms(rb.uniforms, rp.seed, rp.offset * 2 + 1, (uint32_t)m, lane, D + 12);
    if (keep < C) { keep_val = rb.commands[(size_t)m * C + keep]; copy_keep = true; }
    resample_vals(rp, uget(Um, D + 8), uget(Um, D + 9), uget(Um, D + 10), cm0, cm1, third);
    new_cmd = true;
  }
  if (copy_keep && keep == 2) cm2 = keep_val;
  if (new_cmd && !rp.heading_command) cm2 = third;
  // ---- _reset_dofs (:450-465); last_dof_vel = new dof_vel (= 0) after the history copy (:149); last_actions already
  // holds `actions` (zeroed by reset_idx, then overwritten by the history copy :148).  D <= 32: dof j is lane j's.
  float newpos = 0.0f;
  if (lane < D) newpos = mul_r(defpos, rand_range(0.5f, 1.5f, U.lo));
  __syncwarp();   // every lane has read the old state: lane 0 may overwrite it
  if (lane < D) {
    ds[2 * lane] = newpos;
    ds[2 * lane + 1] = 0.0f;
    rb.last_dof_vel[(size_t)e * D + lane] = 0.0f;
  }
  if (lane == 0) {
    if (curr) {
      rb.terrain_levels[e] = lv;
      org[0] = o0; org[1] = o1; org[2] = o2;
    }
#pragma unroll
    for (int k = 0; k < 13; ++k) rs[k] = r[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) rb.last_root_vel[(size_t)e * 6 + k] = r[7 + k];              // the history copy after the reset (:150)
    if (copy_keep) cmd[keep] = keep_val;
    if (new_cmd) {
      cmd[0] = cm0; cmd[1] = cm1;
      if (rp.heading_command) cmd[3] = third;
      else cmd[2] = third;
    }
    // ---- timers, episode clock (:191-198)
    for (int f = 0; f < F; ++f) {
      rb.feet_air_time[(size_t)e * F + f] = 0.0f;
      rb.feet_contact_time[(size_t)e * F + f] = 0.0f;
    }
    rb.episode_length_buf[e] = 0;
    if (is_main) {
      atomicAdd(rb.stats + ELG_NUM_REWARD_TERMS, 1.0f);
      if (rb.stats_accum) atomicAdd(rb.stats_accum + ELG_NUM_REWARD_TERMS, 1.0);
    }
  }
  // ---- extras["episode"] (:200-206): (sum, count) over the reset envs, then zero the sums (lane t owns term t)
  if (term_on) {
    if (is_main) {
      atomicAdd(rb.stats + lane, sum_mine);
      if (rb.stats_accum) atomicAdd(rb.stats_accum + lane, (double)sum_mine);
    }
    if (zero_sums) *sum_ptr = 0.0f;
  }
  // ---- observation repair (:234-252 evaluated after the reset)
  if (repair) {
    const bool share = (nj & 7) + hm <= 8;
    float* ob = rb.obs_buf + (size_t)e * O;
    auto finish = [&](int k, float v, float s16, float ns, float nu) {
      if (tensor_noise) v = v + (2.0f * nu - 1.0f) * ns;
      else if (philox) v = v + fmaf(s16, 1.0f / 32768.0f, -1.0f) * ns;
      if (pr.clip_observations > 0.0f) v = fminf(fmaxf(v, -pr.clip_observations), pr.clip_observations);
      ob[k] = v;
    };
    auto sample = [](const uint4& b, int sidx) {
      const uint32_t w = (sidx >> 1) == 0 ? b.x : (sidx >> 1) == 1 ? b.y : (sidx >> 1) == 2 ? b.z : b.w;
      return (float)((w >> (16 * (sidx & 1))) & 0xffffu);
    };
    // head entries k = lane + 32 m: sample (nj % 8) + m of block 1 + nj / 8 when shared, else sample m % 8 of block m / 8.
    // (One Philox block per lane serves every entry of the usual layout -- nj = 6 height rounds + 2 head rounds = 8 samples.)
    uint4 blk = make_uint4(0, 0, 0, 0);
    int blk_id = -1;
    auto block = [&](int id) {
      if (philox && id != blk_id) { blk = noise_block_cold(pr.noise_seed, pr.noise_offset, e, lane, id); blk_id = id; }
    };
#pragma unroll
    for (int mm = 0; mm < kMaxHM; ++mm) {
      if (mm < hm) {      // (warp-uniform)
        const int k = lane + 32 * mm;
        const float pos_k = __shfl_sync(0xffffffffu, newpos, (k - 12) & 31);      // dof k - 12: lane k - 12 drew its new position
        const float def_k = __shfl_sync(0xffffffffu, defpos, (k - 12) & 31);
        block(share ? 1 + (nj >> 3) : (mm >> 3));
        float v;
        bool touch = false;
        if (k >= 9 && k < 12) { v = (k == 9 ? cm0 : k == 10 ? cm1 : cm2) * pr.commands_scale[k - 9]; touch = true; }
        else if (k >= 12 && k < 12 + D) { v = (pos_k - def_k) * pr.obs_scale_dof_pos; touch = true; }
        else if (k >= 12 + D && k < 12 + 2 * D) { v = 0.0f * pr.obs_scale_dof_vel; touch = true; }
        if (touch) finish(k, v, philox ? sample(blk, share ? (nj & 7) + mm : (mm & 7)) : 0.0f, nsh[mm], nuh[mm]);
      }
    }
    // height entries: clip(z - 0.5 - h, -1, 1) * scale with the NEW base height and the stale heights
    if (H > 0 && rb.measured_heights) {
      const float zc = sub_r(r[2], 0.5f);
#pragma unroll
      for (int j = 0; j < kMaxNJ; ++j) {
        if (j < nj) {
          block(1 + (j >> 3));
          const int p = lane + 32 * j;
          if (p < H) {
            const float v = mul_r(fminf(fmaxf(sub_r(zc, hv[j]), -1.0f), 1.0f), pr.obs_scale_height);
            finish(head + p, v, philox ? sample(blk, j & 7) : 0.0f, nsv[j], nuv[j]);
          }
        }
      }
    }
  }
}

// A CTA of eight warps owns 256 consecutive envs: every thread reads the flag of its env (coalesced), ballots + one shared counter
// collect the rows that have work into a shared list, and the warps take the list entries round-robin -- one row per warp at a
// time.  (One warp per 32 envs handling ITS rows one after the other took as long as its unluckiest warp: at 65 536 envs and ~390
// resets some warp always holds three -- 14.3 us per launch against 6.3 us with at most one per warp, profiles/README.md r2q.)
constexpr int kResetThreads = 256;
__global__ void __launch_bounds__(kResetThreads)
elg_reset_kernel(const __grid_constant__ ElgResetParams rp, const __grid_constant__ ElgStepParams pr, const __grid_constant__ ElgResetBuffers rb,
                 const int N, const int D, const int F, const int C, const int O, const int H) {
  __shared__ int s_list[kResetThreads];
  __shared__ int s_n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int mine = blockIdx.x * kResetThreads + threadIdx.x;
  if (threadIdx.x == 0) s_n = 0;
  pdl_launch_dependents();
  pdl_wait();
  bool work = false;
  if (mine < N) {
    work = rb.reset_buf[mine] != 0;
    if (!work && rp.rows_per_main > 0) work = rb.reset_buf[(mine / rp.rows_per_main) * rp.rows_per_main] != 0;   // the main env of this row resets
  }
  __syncthreads();
  const unsigned todo = __ballot_sync(0xffffffffu, work);
  if (todo) {
    int base = 0;
    if (lane == 0) base = atomicAdd(&s_n, __popc(todo));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (work) s_list[base + __popc(todo & ((1u << lane) - 1u))] = mine;
  }
  __syncthreads();
  const int n = s_n;
  for (int i = warp; i < n; i += kResetThreads / 32) {
    reset_env(rp, pr, rb, s_list[i], lane, N, D, F, C, O, H);
    __syncwarp();
  }
}

}  // namespace elg

namespace {
int rfail(int code, const char* msg) { return elg::set_error(code, msg); }

template <typename... Params, typename... Args>
void launch_pdl(void (*k)(Params...), dim3 grid, int threads, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(threads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, k, Params(args)...);
}
}  // namespace

extern "C" {

int elg_sizeof_reset_params(void) { return (int)sizeof(ElgResetParams); }
int elg_sizeof_reset_buffers(void) { return (int)sizeof(ElgResetBuffers); }

int elg_resample_commands(const ElgDims* dims, const ElgResetParams* rp, const int64_t* episode_length_buf, float* commands,
                          const float* uniforms, float* stats_to_zero, void* stream) {
  if (!dims || !rp) return rfail(ELG_ERR_NULL_POINTER, "dims/params is NULL");
  if (rp->resample_interval < 1) return rfail(ELG_ERR_INVALID_ARGUMENT, "resample_interval must be >= 1");
  if (dims->num_dof + 12 > ELG_RESET_UNIFORMS) return rfail(ELG_ERR_UNSUPPORTED, "num_dof + 12 exceeds ELG_RESET_UNIFORMS");
  if (rp->rows_per_main < 0 || (rp->rows_per_main > 0 && dims->num_envs % rp->rows_per_main != 0))
    return rfail(ELG_ERR_INVALID_ARGUMENT, "rows_per_main must divide num_envs");
  if (dims->num_envs == 0) return ELG_OK;
  if (!episode_length_buf || !commands) return rfail(ELG_ERR_NULL_POINTER, "episode_length_buf/commands is NULL");
  launch_pdl(elg::elg_resample_kernel, dim3((dims->num_envs + 127) / 128), 128, (cudaStream_t)stream, *rp, (int)dims->num_envs, (int)dims->num_dof,
             (int)dims->num_commands, episode_length_buf, commands, uniforms, stats_to_zero);
  return elg::check_launch("elg_resample_commands");
}

int elg_reset_envs(const ElgDims* dims, const ElgResetParams* rp, const ElgStepParams* prm, const ElgResetBuffers* buf, void* stream) {
  if (!dims || !rp || !prm || !buf) return rfail(ELG_ERR_NULL_POINTER, "dims/params/buffers is NULL");
  if (dims->num_dof + 12 > ELG_RESET_UNIFORMS) return rfail(ELG_ERR_UNSUPPORTED, "num_dof + 12 exceeds ELG_RESET_UNIFORMS");
  if (dims->num_dof > 32) return rfail(ELG_ERR_UNSUPPORTED, "elg_reset_envs: more than 32 dofs (one lane per joint)");
  if (buf->obs_buf && dims->num_height_points > 256) return rfail(ELG_ERR_UNSUPPORTED, "elg_reset_envs: more than 256 height points");
  if (dims->num_envs == 0) return ELG_OK;
  if (!buf->reset_buf || !buf->root_states || !buf->dof_state || !buf->commands || !buf->env_origins || !buf->default_dof_pos ||
      !buf->last_dof_vel || !buf->last_root_vel || !buf->feet_air_time || !buf->feet_contact_time || !buf->episode_length_buf ||
      !buf->episode_sums || !buf->stats)
    return rfail(ELG_ERR_NULL_POINTER, "a reset buffer is NULL");
  if (rp->curriculum && (!buf->terrain_levels || !buf->terrain_types || !buf->terrain_origins || rp->max_terrain_level < 1 || rp->terrain_cols < 1))
    return rfail(ELG_ERR_INVALID_ARGUMENT, "terrain curriculum needs terrain_levels / types / origins");
  if (rp->root_z_from_terrain && (!buf->height_samples || prm->hf_rows < 2 || prm->hf_cols < 2))
    return rfail(ELG_ERR_INVALID_ARGUMENT, "root_z_from_terrain needs height_samples and the terrain geometry");
  if (buf->obs_buf && prm->noise_mode != ELG_NOISE_OFF && !buf->noise_scale_vec) return rfail(ELG_ERR_NULL_POINTER, "noise_scale_vec is NULL");
  if (buf->obs_buf && prm->noise_mode == ELG_NOISE_TENSOR && !buf->noise_u) return rfail(ELG_ERR_NULL_POINTER, "noise_u is NULL");
  if (rp->rows_per_main < 0 || (rp->rows_per_main > 0 && dims->num_envs % rp->rows_per_main != 0))
    return rfail(ELG_ERR_INVALID_ARGUMENT, "rows_per_main must divide num_envs");
  if (rp->rows_per_main > 0)
    launch_pdl(elg::elg_main_reset_flag_kernel, dim3(1), 1024, (cudaStream_t)stream, buf->reset_buf, (int)(dims->num_envs / rp->rows_per_main),
               (int)rp->rows_per_main, buf->stats);
  launch_pdl(elg::elg_reset_kernel, dim3((dims->num_envs + elg::kResetThreads - 1) / elg::kResetThreads), elg::kResetThreads, (cudaStream_t)stream, *rp, *prm, *buf, (int)dims->num_envs, (int)dims->num_dof,
             (int)dims->num_feet, (int)dims->num_commands, (int)dims->num_obs, (int)dims->num_height_points);
  return elg::check_launch("elg_reset_envs");
}

}  // extern "C"
