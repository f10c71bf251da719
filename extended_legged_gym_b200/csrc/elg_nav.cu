// elg_nav.cu -- navigation command update of the batch-rollout nav task for sm_100a
// (RobotBatchRolloutNav._update_navigation_commands + _check_goal_reached,
// envs/batch_rollout/robot_batch_rollout_nav.py:135-247).  The reference builds the per-env goal tensor with a Python loop
// over every env (twice per callback) and then runs ~40 ATen ops; here one thread does one env: 28 bytes in (position,
// quaternion) + its main's goal, 12 + 12 + 1 bytes out.  Streaming, trivially small -- the point is one launch and no loop.
#include <cuda_runtime.h>
#include <stdint.h>

#include "elg_common.cuh"
#include "elg_async.cuh"

namespace elg {

__global__ void __launch_bounds__(256)
elg_nav_kernel(const int64_t n_envs, const int rows_per_main, const __grid_constant__ ElgNavParams pr, const float* __restrict__ root_states,
               const float* __restrict__ goals, float* __restrict__ commands, float* __restrict__ prev_commands,
               uint8_t* __restrict__ goal_reached, float* __restrict__ distance) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  if (i >= n_envs) return;
  const float* rs = root_states + i * 13;
  const float* g = goals + (i / rows_per_main) * 3;
  const Quat q = {rs[3], rs[4], rs[5], rs[6]};
  const float ex = g[0] - rs[0], ey = g[1] - rs[1], ez = g[2] - rs[2];
  // desired world-frame velocity: kp * error, clipped to max_linear_vel (:158-172)
  Vec3 v = {pr.kp_linear * ex, pr.kp_linear * ey, pr.use_2d_nav ? 0.0f : pr.kp_linear * ez};
  const float mag = pr.use_2d_nav ? norm2_t(v.x, v.y) : norm3_t(v.x, v.y, v.z);
  const float scale = fminf(pr.max_linear_vel / (mag + 1e-8f), 1.0f);
  v.x *= scale; v.y *= scale; v.z *= scale;
  const Vec3 r = quat_rotate_inverse(q, v);   // (:175)
  float ang = 0.0f;
  if (pr.use_2d_nav) {   // yaw rate towards the goal, wrapped through atan2(sin, cos) (:178-194)
    const float yaw = atan2f(2.0f * (q.w * q.z + q.x * q.y), 1.0f - 2.0f * (q.y * q.y + q.z * q.z));
    float d = atan2f(ey, ex) - yaw;
    d = atan2f(sinf(d), cosf(d));
    ang = fminf(fmaxf(pr.kp_angular * d, -pr.max_angular_vel), pr.max_angular_vel);
  }
  float c0 = r.x, c1 = r.y, c2 = ang;
  float* pc = prev_commands + i * 3;
  if (pr.use_prev) {   // (:203-208)
    c0 = pr.smooth * pc[0] + pr.smooth_c * c0;
    c1 = pr.smooth * pc[1] + pr.smooth_c * c1;
    c2 = pr.smooth * pc[2] + pr.smooth_c * c2;
  }
  pc[0] = c0; pc[1] = c1; pc[2] = c2;
  float* cmd = commands + i * pr.num_commands;
  if (pr.zero_reached && goal_reached[i]) {   // flags of the previous check: the whole command row stops (:219-221)
    for (int k = 0; k < pr.num_commands; ++k) cmd[k] = 0.0f;
  } else {
    cmd[0] = c0; cmd[1] = c1; cmd[2] = c2;
  }
  // _check_goal_reached (:224-247): torch.norm over 2 / 3 elements, see norm2_t
  const float dist = pr.use_2d_nav ? norm2_t(ex, ey) : norm3_t(ex, ey, ez);
  goal_reached[i] = dist < pr.tolerance_rad ? 1 : 0;
  if (distance) distance[i] = dist;
}

}  // namespace elg

extern "C" {

int elg_sizeof_nav_params(void) { return (int)sizeof(ElgNavParams); }

int elg_nav_commands(int32_t num_main, int32_t rollouts_per_main, const ElgNavParams* prm, const float* root_states, const float* goal_positions,
                     float* commands, float* prev_commands, uint8_t* goal_reached, float* distance, void* stream) {
  if (!prm) return elg::set_error(ELG_ERR_NULL_POINTER, "nav params is NULL");
  if (num_main < 0 || rollouts_per_main < 0) return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "negative env counts");
  if (prm->num_commands < 3) return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "num_commands < 3");
  if (!root_states || !goal_positions || !commands || !prev_commands || !goal_reached)
    return elg::set_error(ELG_ERR_NULL_POINTER, "nav commands: a pointer is NULL");
  const int64_t n = (int64_t)num_main * (1 + rollouts_per_main);
  if (n == 0) return ELG_OK;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((n + 255) / 256));
  cfg.blockDim = dim3(256);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, elg::elg_nav_kernel, n, (int)(1 + rollouts_per_main), *prm, root_states, goal_positions, commands, prev_commands,
                     goal_reached, distance);
  return elg::check_launch("elg_nav_commands");
}

}  // extern "C"
