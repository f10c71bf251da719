// elg_plan.cu -- kinematic state integration of the planning variant for sm_100a
// (RobotPlanGradSampling._integrate_state_velocities + _sync_integration_to_sim,
// envs/batch_rollout/robot_plan_grad_sampling.py:103-225).  The reference does ~60 indexed ATen ops per call, called once per
// horizon step for every rollout env; here one thread integrates one env (the joints: one thread per (env, joint)) and the result
// is written through to the simulator tensors: (6 + D) floats in, 7 + D state + 6 + D velocity + 13 + 2 D + 6 simulator floats
// out, all rows staged through shared memory so that global accesses are row-contiguous.  Streaming, small.
#include <cuda_runtime.h>
#include <stdint.h>

#include "elg_common.cuh"
#include "elg_async.cuh"

namespace elg {

constexpr int kPlanEnvs = 32;       // envs per CTA: one warp integrates, all four warps move rows (the kernel is latency bound:
constexpr int kPlanThreads = 128;   // many small CTAs keep more loads in flight than few large ones: 20.2 -> 10.3 us at 32 768 envs)

// Row-wise copies between global arrays (rows picked by env id) and shared staging, consecutive threads on consecutive words of a
// row: the per-env structs are 3..64 floats, so a thread-per-env access pattern would touch one 32-byte sector per thread and
// instruction (measured: 30.7 us for 32 768 envs); staged this way every sector is written once, whole.
struct RowIter {
  int e, w, qe, qw;
  __device__ RowIter(int W) : e(threadIdx.x / W), w(threadIdx.x % W), qe(kPlanThreads / W), qw(kPlanThreads % W) {}
  __device__ void next(int W) { e += qe; w += qw; if (w >= W) { w -= W; ++e; } }
};
// (the iterator of a row width is built once per thread and copied: its four integer divisions are the expensive part)
__device__ __forceinline__ void gather_rows(float* __restrict__ dst, const float* __restrict__ src, const int64_t* s_id, int n, int W, RowIter it) {
  for (; it.e < n; it.next(W)) dst[it.e * W + it.w] = src[s_id[it.e] * W + it.w];
}
__device__ __forceinline__ void scatter_rows(float* __restrict__ dst, const float* __restrict__ src, const int64_t* s_id, int n, int W, RowIter it) {
  for (; it.e < n; it.next(W)) dst[s_id[it.e] * W + it.w] = src[it.e * W + it.w];
}

__global__ void __launch_bounds__(kPlanThreads)
elg_plan_integrate_kernel(const __grid_constant__ ElgPlanParams pr, const __grid_constant__ ElgPlanBuffers bf, const float* __restrict__ state_vels,
                          const int64_t* __restrict__ env_ids, const int64_t rows) {
  extern __shared__ __align__(16) float sm[];
  __shared__ int64_t s_id[kPlanEnvs];
  const int D = pr.num_dof, SV = 6 + D;
  // shared layout (floats per env): state velocities [6 + D] | root [13] | pos [3] | quat [4] | lin [3] | ang [3] | blv [3] | bav [3] | dof_pos [D] | dof_vel [D]
  float* s_sv = sm;
  float* s_root = s_sv + kPlanEnvs * SV;
  float* s_pos = s_root + kPlanEnvs * 13;
  float* s_quat = s_pos + kPlanEnvs * 3;
  float* s_lin = s_quat + kPlanEnvs * 4;
  float* s_ang = s_lin + kPlanEnvs * 3;
  float* s_blv = s_ang + kPlanEnvs * 3;
  float* s_bav = s_blv + kPlanEnvs * 3;
  float* s_dp = s_bav + kPlanEnvs * 3;
  float* s_dv = s_dp + kPlanEnvs * D;
  const int64_t r0 = (int64_t)blockIdx.x * kPlanEnvs;
  const int n = (int)min((int64_t)kPlanEnvs, rows - r0);
  const int t = threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  if (t < n) s_id[t] = env_ids ? env_ids[r0 + t] : r0 + t;
  __syncthreads();
  const RowIter it3(3), it4(4), it13(13), itD(D), it2D(2 * D);
  // state_vels == NULL: write-through only (_sync_integration_to_sim on its own, :197-225) -- the stored velocities, no sub-steps
  const bool integrate = state_vels != nullptr;
  if (integrate) {
    const float* src = state_vels + r0 * SV;      // the rows of this CTA are contiguous
    for (int i = t; i < n * SV; i += kPlanThreads) s_sv[i] = src[i];
  } else {
    gather_rows(s_lin, bf.integration_base_lin_vel, s_id, n, 3, it3);
    gather_rows(s_ang, bf.integration_base_ang_vel, s_id, n, 3, it3);
    gather_rows(s_dv, bf.integration_dof_vel, s_id, n, D, itD);
  }
  gather_rows(s_pos, bf.integration_base_pos, s_id, n, 3, it3);
  gather_rows(s_quat, bf.integration_base_quat, s_id, n, 4, it4);
  gather_rows(s_dp, bf.integration_dof_pos, s_id, n, D, itD);
  __syncthreads();
  if (t < n) {
    const int nsub = integrate ? pr.n_substeps : 0;
    const float* sv = s_sv + t * SV;
    auto clampf = [](float v, float m) { return fminf(fmaxf(v, -m), m); };
    float vx, vy, vz, wx, wy, wz;
    if (integrate) {
      vx = clampf(sv[0], pr.max_base_lin_vel), vy = clampf(sv[1], pr.max_base_lin_vel), vz = clampf(sv[2], pr.max_base_lin_vel);
      wx = clampf(sv[3], pr.max_base_ang_vel), wy = clampf(sv[4], pr.max_base_ang_vel), wz = clampf(sv[5], pr.max_base_ang_vel);
      s_lin[t * 3] = vx; s_lin[t * 3 + 1] = vy; s_lin[t * 3 + 2] = vz;
      s_ang[t * 3] = wx; s_ang[t * 3 + 1] = wy; s_ang[t * 3 + 2] = wz;
    } else {
      vx = s_lin[t * 3], vy = s_lin[t * 3 + 1], vz = s_lin[t * 3 + 2], wx = s_ang[t * 3], wy = s_ang[t * 3 + 1], wz = s_ang[t * 3 + 2];
    }
    float px = s_pos[t * 3], py = s_pos[t * 3 + 1], pz = s_pos[t * 3 + 2];
    float qx = s_quat[t * 4], qy = s_quat[t * 4 + 1], qz = s_quat[t * 4 + 2], qw = s_quat[t * 4 + 3];
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
    s_pos[t * 3] = px; s_pos[t * 3 + 1] = py; s_pos[t * 3 + 2] = pz;
    s_quat[t * 4] = qx; s_quat[t * 4 + 1] = qy; s_quat[t * 4 + 2] = qz; s_quat[t * 4 + 3] = qw;
    float* rs = s_root + t * 13;
    rs[0] = px; rs[1] = py; rs[2] = pz; rs[3] = qx; rs[4] = qy; rs[5] = qz; rs[6] = qw;
    rs[7] = vx; rs[8] = vy; rs[9] = vz; rs[10] = wx; rs[11] = wy; rs[12] = wz;
    const Quat q = {qx, qy, qz, qw};
    const Vec3 bl = quat_rotate_inverse(q, Vec3{vx, vy, vz}), ba = quat_rotate_inverse(q, Vec3{wx, wy, wz});
    s_blv[t * 3] = bl.x; s_blv[t * 3 + 1] = bl.y; s_blv[t * 3 + 2] = bl.z;
    s_bav[t * 3] = ba.x; s_bav[t * 3 + 1] = ba.y; s_bav[t * 3 + 2] = ba.z;
  }
  if (integrate) {
    // joints: one (env, joint) pair per thread and iteration (:162, :178, :181-186)
    RowIter it = itD;
    for (int i = t; i < n * D; i += kPlanThreads, it.next(D)) {
      const int e = it.e, j = it.w;
      const float jv = fminf(fmaxf(s_sv[e * SV + 6 + j], -pr.max_joint_vel), pr.max_joint_vel);
      float p = s_dp[i];
      for (int s = 0; s < pr.n_substeps; ++s) p += jv * pr.sub_dt;
      if (pr.enforce_joint_limits && bf.dof_pos_limits) p = fminf(fmaxf(p, bf.dof_pos_limits[2 * j]), bf.dof_pos_limits[2 * j + 1]);
      s_dp[i] = p;
      s_dv[i] = jv;
    }
  }
  __syncthreads();
  if (integrate) {
    scatter_rows(bf.integration_base_pos, s_pos, s_id, n, 3, it3);
    scatter_rows(bf.integration_base_quat, s_quat, s_id, n, 4, it4);
    scatter_rows(bf.integration_base_lin_vel, s_lin, s_id, n, 3, it3);
    scatter_rows(bf.integration_base_ang_vel, s_ang, s_id, n, 3, it3);
    scatter_rows(bf.integration_dof_pos, s_dp, s_id, n, D, itD);
    scatter_rows(bf.integration_dof_vel, s_dv, s_id, n, D, itD);
  }
  scatter_rows(bf.root_states, s_root, s_id, n, 13, it13);
  scatter_rows(bf.base_lin_vel, s_blv, s_id, n, 3, it3);
  scatter_rows(bf.base_ang_vel, s_bav, s_id, n, 3, it3);
  {   // dof_state rows: (pos, vel) pairs interleaved, 2 D floats per env
    const int W = 2 * D;
    for (RowIter it = it2D; it.e < n; it.next(W)) {
      const int j = it.w >> 1;
      bf.dof_state[s_id[it.e] * W + it.w] = (it.w & 1) ? s_dv[it.e * D + j] : s_dp[it.e * D + j];
    }
  }
}

static size_t plan_smem_bytes(int D) { return sizeof(float) * kPlanEnvs * (size_t)((6 + D) + 13 + 3 + 4 + 3 + 3 + 3 + 3 + 2 * D); }

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
  const size_t smem = elg::plan_smem_bytes(prm->num_dof);
  static elg::SmemCache smem_cache = {};
  size_t& smem_set = elg::smem_slot(smem_cache);
  if (smem > smem_set) {
    cudaFuncSetAttribute(elg::elg_plan_integrate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    smem_set = smem;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((num_rows + elg::kPlanEnvs - 1) / elg::kPlanEnvs));
  cfg.blockDim = dim3(elg::kPlanThreads);
  cfg.dynamicSmemBytes = smem;
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
