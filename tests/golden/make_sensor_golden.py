"""Generate tests/golden/sensor_envs.npz FROM THE UNMODIFIED REFERENCE (container only):

  confined_*      utils/terrain_confine.py:13-146 convert_2layer_heightfield_to_trimesh on a seeded ground / ceiling pair (ceiling
                  enabled; one case with the slope correction) -- the two-layer terrain mesh of BASELINE config 4
  raydist_*       envs/base/legged_robot_raycast.py:262-297 LeggedRobotRayCast._get_raycast_distances on a synthetic self
  sdfpts_*        envs/batch_rollout/robot_batch_rollout_percept.py:385-441 RobotBatchRolloutPercept._update_sdf_values with a
                  recording MeshSDF stand-in: the query points it builds, and what it stores
  async_*         utils/gait_scheduler.py:104-173 AsyncGaitScheduler and envs/elspider_air/elspider.py:351-363

    python tests/golden/make_sensor_golden.py
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402


def two_layer_fields(seed, rows=28, cols=22):
    rng = np.random.default_rng(seed)
    ground = np.zeros((rows, cols), dtype=np.int16)
    ground[6:12, 4:10] = 40                       # a box
    ground[16:22, 8:18] = (np.arange(10) * 12).astype(np.int16)[None, :]      # a ramp
    ground += rng.integers(-3, 4, size=ground.shape).astype(np.int16)
    ceiling = (ground.astype(np.int32) + 260 + rng.integers(-20, 21, size=ground.shape)).astype(np.int16)
    ceiling[10:16, 6:14] -= 120                   # a hanging block
    return ground, ceiling


def main():
    rh.reference_classes()          # (imports legged_gym.envs first: the package has an import cycle through legged_gym.utils)
    from legged_gym.utils.terrain_confine import convert_2layer_heightfield_to_trimesh
    from legged_gym.envs.base.legged_robot_raycast import LeggedRobotRayCast
    from legged_gym.envs.batch_rollout.robot_batch_rollout_percept import RobotBatchRolloutPercept
    from legged_gym.utils.gait_scheduler import AsyncGaitScheduler, AsyncGaitSchedulerCfg
    from legged_gym.envs.elspider_air.elspider import ElSpider
    out = {}
    for tag, slope in (("confined_a", None), ("confined_b", 0.75)):
        g, c = two_layer_fields(3)
        v, t = convert_2layer_heightfield_to_trimesh(g, c, 0.1, 0.005, slope_threshold=slope, enable_ceiling=True, global_noise=0.0)
        out[f"{tag}__ground"], out[f"{tag}__ceiling"] = g, c
        out[f"{tag}__vertices"], out[f"{tag}__triangles"] = np.asarray(v, dtype=np.float32), np.asarray(t, dtype=np.int32)
    # ---- ray distances
    gen = torch.Generator().manual_seed(1)
    N, R = 7, 13
    hits = torch.randn(N, R, 3, generator=gen) * 4
    found = torch.rand(N, R, generator=gen) > 0.3
    root = torch.randn(N, 13, generator=gen)
    fake = SimpleNamespace(ray_caster=SimpleNamespace(data=SimpleNamespace(ray_hits=hits, ray_hits_found=found), cfg=SimpleNamespace(max_distance=6.0)),
                           root_states=root)
    out["raydist__hits"], out["raydist__found"], out["raydist__root"] = hits.numpy(), found.numpy(), root.numpy()
    out["raydist__normalized"] = LeggedRobotRayCast._get_raycast_distances(fake).numpy()
    out["raydist__raw"] = LeggedRobotRayCast._get_raycast_distances(fake, normalize=False).numpy()
    ids = torch.tensor([5, 0, 2])
    out["raydist__ids"] = ids.numpy()
    out["raydist__normalized_ids"] = LeggedRobotRayCast._get_raycast_distances(fake, env_ids=ids).numpy()
    # ---- SDF query points
    B, bodies = 9, [0, 4, 7]
    offs = [[0.1, -0.05, 0.02], [0.0, 0.0, -0.08]]                 # the third body has no offset configured
    rbs = torch.randn(N * B, 13, generator=gen)
    q = rbs.view(N, B, 13)[:, :, 3:7]
    q /= q.norm(dim=-1, keepdim=True)
    rec = []

    class FakeSDF:
        def query(self, points):
            rec.append(points.clone())
            d = points.norm(dim=-1)
            return d - 1.0, points / d.unsqueeze(-1)

        def nearest_points(self, points):
            s, g = self.query(points)
            rec.pop()
            return points - s.unsqueeze(-1) * g
    env = SimpleNamespace(total_num_envs=N, num_bodies=B, device="cpu", sdf_body_indices=bodies, rigid_body_state=rbs, mesh_sdf=FakeSDF(),
                          cfg=SimpleNamespace(sdf=SimpleNamespace(collision_sphere_pos=offs, compute_nearest_points=True, compute_gradients=True)),
                          sdf_values=torch.zeros(N, 3), sdf_gradients=torch.zeros(N, 3, 3), sdf_nearest_points=torch.zeros(N, 3, 3))
    RobotBatchRolloutPercept._update_sdf_values(env)
    out["sdfpts__rbs"], out["sdfpts__bodies"], out["sdfpts__offsets"] = rbs.numpy(), np.asarray(bodies), np.asarray(offs, dtype=np.float32)
    out["sdfpts__points"] = torch.stack(rec, dim=1).numpy()
    out["sdfpts__values"], out["sdfpts__grad"], out["sdfpts__nearest"] = env.sdf_values.numpy(), env.sdf_gradients.numpy(), env.sdf_nearest_points.numpy()
    # ---- async gait scheduler
    dof_pos = torch.randn(N, 18, generator=gen) * 0.4 + torch.tensor([0.0, 1.0, 1.0] * 6)
    foot_pos = torch.randn(N, 6, 3, generator=gen) * 0.1
    sch = AsyncGaitScheduler(None, None, None, None, None, dof_pos, None, foot_pos, None, N, "cpu", AsyncGaitSchedulerCfg())
    out["async__dof_pos"], out["async__foot_pos"] = dof_pos.numpy(), foot_pos.numpy()
    out["async__dof_align"], out["async__dof_nominal_pos"], out["async__foot_z_align"] = (sch.reward_dof_align().numpy(), sch.reward_dof_nominal_pos().numpy(),
                                                                                        sch.reward_foot_z_align().numpy())
    scales = SimpleNamespace(dof_align=1.0, dof_nominal_pos=[0.0, 0.2], reward_foot_z_align=[0.0, 0.6])
    for stage in (0, 1):
        e = SimpleNamespace(cfg=SimpleNamespace(rewards=SimpleNamespace(async_gait_scheduler=scales)), async_gait_scheduler=sch, reward_scales_stage=stage)
        out[f"async__combined_stage{stage}"] = ElSpider._reward_async_gait_scheduler(e).numpy()
    path = os.path.join(HERE, "sensor_envs.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB", {k: v.shape for k, v in out.items() if "confined" in k})


if __name__ == "__main__":
    main()
