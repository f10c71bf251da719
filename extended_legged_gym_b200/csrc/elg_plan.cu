// elg_plan.cu -- kinematic state integration of the planning variant for sm_100a
// (RobotPlanGradSampling._integrate_state_velocities + _sync_integration_to_sim,
// envs/batch_rollout/robot_plan_grad_sampling.py:103-225).  The reference does ~60 indexed ATen ops per call, called once per
// horizon step for every rollout env; here one thread integrates one env and writes the result through to the simulator
// tensors: (6 + D) floats in, 7 + D state + 6 + D velocity + 13 + 2 D + 6 simulator floats out.  Streaming, small.
#include <cuda_runtime.h>
#include <stdint.h>

#include "elg_common.cuh"
#include "elg_async.cuh"

namespace elg {

__global__ void __launch_bounds__(128)
elg_plan_integrate_kernel(const __grid_constant__ ElgPlanParams pr, const __grid_constant__ ElgPlanBuffers bf, const float* __restrict__ state_vels,
                          const int64_t* __restrict__ env_ids, const int64_t rows) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  if (r >= rows) return;
  const int D = pr.num_dof;
  const int64_t e = env_ids ? env_ids[r] : r;
  // state_vels == NULL: write-through only (_sync_integration_to_sim on its own, :197-225) -- the stored velocities, no sub-steps
  const bool integrate = state_vels != nullptr;
  const int nsub = integrate ? pr.n_substeps : 0;
  const float* sv = state_vels + r * (6 + D);
  auto clampf = [](float v, float m) { return fminf(fmaxf(v, -m), m); };
  float* lin = bf.integration_base_lin_vel + e * 3;
  float* ang = bf.integration_base_ang_vel + e * 3;
  float vx, vy, vz, wx, wy, wz;
  if (integrate) {
    vx = clampf(sv[0], pr.max_base_lin_vel), vy = clampf(sv[1], pr.max_base_lin_vel), vz = clampf(sv[2], pr.max_base_lin_vel);
    wx = clampf(sv[3], pr.max_base_ang_vel), wy = clampf(sv[4], pr.max_base_ang_vel), wz = clampf(sv[5], pr.max_base_ang_vel);
  } else {
    vx = lin[0], vy = lin[1], vz = lin[2], wx = ang[0], wy = ang[1], wz = ang[2];
  }
  float* pos = bf.integration_base_pos + e * 3;
  float* quat = bf.integration_base_quat + e * 4;
  float px = pos[0], py = pos[1], pz = pos[2];
  float qx = quat[0], qy = quat[1], qz = quat[2], qw = quat[3];
  const float dt = pr.sub_dt;
  // angle-axis increment of one sub-step (:143-148): the same for every sub-step, the velocities are held constant
  const float wn = norm3_t(wx, wy, wz);
  const float angle = wn * dt;
  const float ax = wx / (wn + 1e-8f), ay = wy / (wn + 1e-8f), az = wz / (wn + 1e-8f);
  // quat_from_angle_axis: normalize(axis) * sin(angle / 2), cos(angle / 2), normalised once more (torch_utils, eps 1e-9)
  const float an = fmaxf(norm3_t(ax, ay, az), 1e-9f);
  const float sh = sinf(angle / 2.0f), ch = cosf(angle / 2.0f);
  float rx = (ax / an) * sh, ry = (ay / an) * sh, rz = (az / an) * sh, rw = ch;
  {
    const float rn = fmaxf(__fsqrt_rn(((rx * rx + ry * ry) + rz * rz) + rw * rw), 1e-9f);
    rx /= rn; ry /= rn; rz /= rn; rw /= rn;
  }
  for (int s = 0; s < nsub; ++s) {
    if (pr.method == 0) {
      px += vx * dt; py += vy * dt; pz += vz * dt;                         // (:139)
    } else {
      px += (((vx + 2.0f * vx) + 2.0f * vx) + vx) * dt / 6.0f;             // (:170-175): k1 = k2 = k3 = k4 = v
      py += (((vy + 2.0f * vy) + 2.0f * vy) + vy) * dt / 6.0f;
      pz += (((vz + 2.0f * vz) + 2.0f * vz) + vz) * dt / 6.0f;
    }
    // quat_mul(q, rot) in the torch_utils factorisation, then renormalise (:151-159)
    const float ww = (qz + qx) * (rx + ry), yy = (qw - qy) * (rw + rz), zz = (qw + qy) * (rw - rz);
    const float xx = ww + yy + zz;
    const float qq = 0.5f * (xx + (qz - qx) * (rx - ry));
    const float nw = qq - ww + (qz - qy) * (ry - rz);
    const float nx = qq - xx + (qx + qw) * (rx + rw);
    const float ny = qq - yy + (qw - qx) * (ry + rz);
    const float nz = qq - zz + (qz + qy) * (rw - rx);
    const float qn = __fsqrt_rn(((nx * nx + ny * ny) + nz * nz) + nw * nw);
    qx = nx / qn; qy = ny / qn; qz = nz / qn; qw = nw / qn;
  }
  if (integrate) {
    pos[0] = px; pos[1] = py; pos[2] = pz;
    quat[0] = qx; quat[1] = qy; quat[2] = qz; quat[3] = qw;
    lin[0] = vx; lin[1] = vy; lin[2] = vz;
    ang[0] = wx; ang[1] = wy; ang[2] = wz;
  }
  float* rs = bf.root_states + e * 13;
  rs[0] = px; rs[1] = py; rs[2] = pz; rs[3] = qx; rs[4] = qy; rs[5] = qz; rs[6] = qw;
  rs[7] = vx; rs[8] = vy; rs[9] = vz; rs[10] = wx; rs[11] = wy; rs[12] = wz;
  const Quat q = {qx, qy, qz, qw};
  const Vec3 bl = quat_rotate_inverse(q, Vec3{vx, vy, vz}), ba = quat_rotate_inverse(q, Vec3{wx, wy, wz});
  float* blv = bf.base_lin_vel + e * 3;
  float* bav = bf.base_ang_vel + e * 3;
  blv[0] = bl.x; blv[1] = bl.y; blv[2] = bl.z;
  bav[0] = ba.x; bav[1] = ba.y; bav[2] = ba.z;
  float* dp = bf.integration_dof_pos + e * D;
  float* dv = bf.integration_dof_vel + e * D;
  float2* ds = reinterpret_cast<float2*>(bf.dof_state) + e * D;
  for (int j = 0; j < D; ++j) {
    float p = dp[j];
    if (!integrate) { ds[j] = make_float2(p, dv[j]); continue; }
    const float jv = clampf(sv[6 + j], pr.max_joint_vel);
    for (int s = 0; s < nsub; ++s) p += jv * dt;                            // (:162, :178)
    if (pr.enforce_joint_limits && bf.dof_pos_limits) p = fminf(fmaxf(p, bf.dof_pos_limits[2 * j]), bf.dof_pos_limits[2 * j + 1]);
    dp[j] = p;
    dv[j] = jv;
    ds[j] = make_float2(p, jv);
  }
}

}  // namespace elg

extern "C" {

int elg_sizeof_plan_params(void) { return (int)sizeof(ElgPlanParams); }
int elg_sizeof_plan_buffers(void) { return (int)sizeof(ElgPlanBuffers); }

int elg_integrate_state_velocities(const ElgPlanParams* prm, const ElgPlanBuffers* buf, const float* state_vels, const int64_t* env_ids,
                                   int64_t num_rows, void* stream) {
  if (!prm || !buf) return elg::set_error(ELG_ERR_NULL_POINTER, "plan params / buffers is NULL");
  if (prm->num_dof < 1 || prm->num_dof > ELG_MAX_DOF) return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "num_dof outside [1, ELG_MAX_DOF]");
  if (prm->method < 0 || prm->method > 1) return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "integration method must be 0 (euler) or 1 (rk4)");
  if (prm->n_substeps < 1) return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "n_substeps < 1");
  if (num_rows < 0) return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "num_rows < 0");
  if (!buf->integration_base_pos || !buf->integration_base_quat || !buf->integration_dof_pos || !buf->integration_base_lin_vel ||
      !buf->integration_base_ang_vel || !buf->integration_dof_vel || !buf->root_states || !buf->dof_state || !buf->base_lin_vel || !buf->base_ang_vel)
    return elg::set_error(ELG_ERR_NULL_POINTER, "state integration: a pointer is NULL");
  if ((reinterpret_cast<uintptr_t>(buf->dof_state) & 7u) != 0) return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "dof_state must be 8-byte aligned");
  if (num_rows == 0) return ELG_OK;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((num_rows + 127) / 128));
  cfg.blockDim = dim3(128);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, elg::elg_plan_integrate_kernel, *prm, *buf, state_vels, env_ids, num_rows);
  return elg::check_launch("elg_integrate_state_velocities");
}

}  // extern "C"
