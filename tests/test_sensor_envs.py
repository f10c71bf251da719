"""Sensor-carrying env classes and the meshes of BASELINE config 4 (missing 1, 2, 3, 5, 6, 7).
not-gpu: the observation glue restated in oracle/sensor_oracle.py and utils/gait_scheduler.py against tests/golden/sensor_envs.npz
         (generated from the unmodified reference by tests/golden/make_sensor_golden.py); OBJ round trip of the two-layer mesh.
gpu:     ray cast + SDF on the reference's two-layer (ground + ceiling) confined mesh and on the same mesh re-read from OBJ, against
         the float64 brute force; LeggedRobotRayCast / RobotBatchRolloutPercept / LeggedRobotDepth / ElSpider / Go2 end to end."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import common  # noqa: E402
from oracle import mesh_oracle as mo, sensor_oracle as so  # noqa: E402
from extended_legged_gym_b200 import synthetic  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sensor_envs.npz")
DEV = "cuda:0"
RTOL, ATOL = 1e-5, 1e-6
Z = np.load(GOLDEN)
T = lambda k: torch.from_numpy(np.asarray(Z[k]))


def write_obj(path, v, t):
    with open(path, "w") as fh:
        for p in v:
            fh.write(f"v {float(p[0])!r} {float(p[1])!r} {float(p[2])!r}\n")
        for f in t:
            fh.write(f"f {int(f[0]) + 1} {int(f[1]) + 1} {int(f[2]) + 1}\n")


# ---------------------------------------------------------------------------------------------- CPU
def test_raycast_distance_oracle_matches_reference_fixture():
    hits, found, root = T("raydist__hits"), T("raydist__found"), T("raydist__root")
    assert torch.equal(so.raycast_distances(hits, found, root[:, :3], 6.0), T("raydist__normalized"))
    assert torch.equal(so.raycast_distances(hits, found, root[:, :3], 6.0, normalize=False), T("raydist__raw"))
    ids = T("raydist__ids")
    assert torch.equal(so.raycast_distances(hits[ids], found[ids], root[ids, :3], 6.0), T("raydist__normalized_ids"))


def test_sdf_query_point_oracle_matches_reference_fixture():
    offs = [torch.from_numpy(o) for o in Z["sdfpts__offsets"]] + [None]
    pts = so.sdf_query_points(T("sdfpts__rbs"), 9, Z["sdfpts__bodies"].tolist(), None)
    got = []
    for k, b in enumerate(Z["sdfpts__bodies"].tolist()):
        got.append(so.sdf_query_points(T("sdfpts__rbs"), 9, [b], None if offs[k] is None else [offs[k]])[:, 0])
    got = torch.stack(got, dim=1)
    assert torch.equal(got, T("sdfpts__points"))
    assert torch.equal(pts[:, 2], T("sdfpts__points")[:, 2])          # the body without an offset is its own position
    d = got.norm(dim=-1)
    assert torch.allclose(so.nearest_points(got, d - 1.0, got / d.unsqueeze(-1)), T("sdfpts__nearest"), rtol=0, atol=0)


def test_async_gait_scheduler_matches_reference_fixture():
    from extended_legged_gym_b200.utils.gait_scheduler import AsyncGaitScheduler, AsyncGaitSchedulerCfg
    n = Z["async__dof_pos"].shape[0]
    s = AsyncGaitScheduler(None, None, None, None, None, T("async__dof_pos"), None, T("async__foot_pos"), None, n, "cpu", AsyncGaitSchedulerCfg())
    assert torch.equal(s.reward_dof_align(), T("async__dof_align"))
    assert torch.equal(s.reward_dof_nominal_pos(), T("async__dof_nominal_pos"))
    assert torch.equal(s.reward_foot_z_align(), T("async__foot_z_align"))


@pytest.mark.parametrize("tag", ["confined_a", "confined_b"])
def test_two_layer_mesh_obj_round_trip_and_oracle_known_answers(tag, tmp_path):
    from extended_legged_gym_b200.utils.ray_caster import load_obj
    v, t = Z[f"{tag}__vertices"], Z[f"{tag}__triangles"]
    path = str(tmp_path / "confined.obj")
    write_obj(path, v, t)
    v2, t2 = load_obj(path)
    assert np.array_equal(v2, v) and np.array_equal(t2, t)
    # between the layers: up hits the ceiling, down hits the ground (both face orientations count)
    o = np.array([[1.0, 1.0, 0.6], [1.0, 1.0, 0.6]], np.float32)
    d = np.array([[0, 0, 1.0], [0, 0, -1.0]], np.float32)
    hits, found, dist, _ = mo.raycast_mesh(o, d, 10.0, v, t)
    assert found.all() and hits[0, 2] > 0.62 and hits[1, 2] < 0.1          # (under the hanging block the ceiling sits at ~0.7 m)
    # polygon faces and negative indices are part of the OBJ subset the loader reads
    with open(path, "w") as fh:
        fh.write("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nf 1/1/1 2/2/2 3/3/3 4/4/4\nf -4 -3 -2\n")
    v3, t3 = load_obj(path)
    assert t3.tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 2]] and v3.shape == (4, 3)


# ---------------------------------------------------------------------------------------------- GPU
def rays_between_layers(v, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = v.min(0), v.max(0)
    o = np.stack([rng.uniform(lo[0] + 0.05, hi[0] - 0.05, n), rng.uniform(lo[1] + 0.05, hi[1] - 0.05, n), rng.uniform(0.35, 0.9, n)], axis=1).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["confined_a", "confined_b"])
@pytest.mark.parametrize("via_obj", [False, True])
def test_ray_and_sdf_on_two_layer_confined_mesh(tag, via_obj, tmp_path):
    from extended_legged_gym_b200.utils.mesh_sdf import MeshSDF, MeshSDFCfg
    from extended_legged_gym_b200.utils.ray_caster import Mesh, load_obj, raycast_mesh
    v, t = Z[f"{tag}__vertices"], Z[f"{tag}__triangles"]
    if via_obj:
        path = str(tmp_path / "confined.obj")
        write_obj(path, v, t)
        sdf = MeshSDF(MeshSDFCfg(mesh_paths=[path], max_distance=1.5, enable_caching=False), device=DEV)
        v, t = load_obj(path)
        mesh = Mesh(v, t, DEV)
    else:
        mesh = Mesh(v, t, DEV)
        sdf = MeshSDF(MeshSDFCfg(vertices=v, triangles=t, max_distance=1.5, enable_caching=False), device=DEV)
    o, d = rays_between_layers(v, 4000, 5)
    for max_dist in (0.4, 3.0):
        hits, found, dist = raycast_mesh(torch.from_numpy(o).to(DEV), torch.from_numpy(d).to(DEV), max_dist, mesh, return_distance=True)
        wh, wf, wd, _ = mo.raycast_mesh(o, d, max_dist, v, t)
        assert np.array_equal(found.cpu().numpy(), wf) and np.array_equal(dist.cpu().numpy(), wd)
        assert np.allclose(hits.cpu().numpy(), wh, rtol=RTOL, atol=ATOL)
    assert 0.3 < wf.mean() < 1.0      # floor and ceiling over a 2.7 m x 2.1 m patch, open to the sides
    pts = o[:1500]
    s, g = sdf.query(torch.from_numpy(pts).to(DEV))
    ws, wg, _, _ = mo.sdf_query(pts, 1.5, v, t)
    assert np.allclose(np.abs(s.cpu().numpy()), np.abs(ws), rtol=RTOL, atol=ATOL)          # unsigned distance to a non-closed two-layer mesh
    same = np.sign(s.cpu().numpy()) == np.sign(ws)
    assert same.mean() > 0.999, f"{(~same).sum()} signs differ"
    assert np.allclose(g.cpu().numpy()[same], wg[same], rtol=1e-4, atol=1e-5)


def small_terrain_env(cls, cfg_mod, n, **cfg_over):
    from extended_legged_gym_b200.sim_backend import SyntheticSim
    cfg, spec, st = common.make_case_state("anymal_c_rough", n, seed=2)
    for k, val in cfg_over.items():
        obj, _, leaf = k.rpartition(".")
        tgt = cfg
        for part in obj.split("."):
            tgt = getattr(tgt, part)
        setattr(tgt, leaf, val)
    cfg_mod(cfg)
    hf = synthetic.make_height_field(rows=120, cols=120, border=20, tile=40, seed=1)
    cfg.terrain.border_size = 2.0
    cfg.terrain.num_rows, cfg.terrain.num_cols, cfg.terrain.terrain_length, cfg.terrain.terrain_width = 2, 2, 4.0, 4.0
    cfg.terrain.curriculum = False
    st["root_states"][:, 0:2] = torch.rand(st["root_states"].shape[0], 2, generator=torch.Generator().manual_seed(3)) * 6 + 1
    st["root_states"][:, 2] = 0.6
    B = spec.num_bodies
    st["rigid_body_state"].view(-1, B, 13)[:, :, 0:3] = st["root_states"][:, None, 0:3] + torch.randn(st["root_states"].shape[0], B, 3, generator=torch.Generator().manual_seed(4)) * 0.2
    env = cls(cfg, None, SyntheticSim(cfg, cfg.env.num_envs if cls.__name__.startswith("RobotBatch") else n, DEV, spec=spec, height_samples=hf, state=st) if False else
              SyntheticSim(cfg, st["root_states"].shape[0], DEV, spec=spec, height_samples=hf, state=st), DEV, True)
    env.set_env_state(st)
    v, t = synthetic.heightfield_to_trimesh(hf, cfg.terrain.horizontal_scale, cfg.terrain.vertical_scale, cfg.terrain.border_size)
    return env, st, v, t


@pytest.mark.gpu
def test_legged_robot_raycast_observations_in_place():
    from extended_legged_gym_b200.envs import LeggedRobotRayCast
    n, rays = 64, 32

    def mod(cfg):
        cfg.env.num_envs = n
        cfg.raycaster.enable_raycast, cfg.raycaster.ray_pattern, cfg.raycaster.num_rays = True, "cone", rays
        cfg.raycaster.max_distance, cfg.raycaster.offset_pos, cfg.raycaster.ray_angle = 3.0, [0.3, 0.0, 0.0], 70
        cfg.env.num_observations = 235 + rays
    env, st, v, t = small_terrain_env(LeggedRobotRayCast, mod, n)
    env.noise_u = torch.rand(n, env.num_obs).to(DEV)
    obs, _, _, _, _ = env.step(torch.zeros(n, 12, device=DEV))
    torch.cuda.synchronize()
    data = env.ray_caster.data
    # hits against the brute force on the same rays
    o, d = mo.sensor_rays(env.ray_caster._pattern_origins.cpu().numpy(), env.ray_caster._pattern_directions.cpu().numpy(),
                          st["root_states"][:, :3].numpy(), st["root_states"][:, 3:7].numpy(), False)
    wh, wf, _, _ = mo.raycast_mesh(o.reshape(-1, 3), d.reshape(-1, 3), 3.0, v, t)
    assert np.array_equal(data.ray_hits_found.cpu().numpy().reshape(-1), wf)
    assert np.allclose(data.ray_hits.cpu().numpy().reshape(-1, 3), wh, rtol=1e-4, atol=1e-5)
    assert 0.2 < wf.mean() < 1.0
    want = so.raycast_distances(data.ray_hits.cpu(), data.ray_hits_found.cpu(), st["root_states"][:, :3], 3.0)
    assert torch.allclose(obs[:, 235:].cpu(), want, rtol=RTOL, atol=ATOL)
    assert torch.allclose(env._get_raycast_distances().cpu(), want, rtol=RTOL, atol=ATOL)
    assert env.raycast_distances.data_ptr() == env.obs_buf[:, 235:].data_ptr()
    # the columns in front are the base class's observation row
    from extended_legged_gym_b200.envs import LeggedRobot
    base, _, _, _ = small_terrain_env(LeggedRobot, lambda c: setattr(c.env, "num_envs", n), n)
    base.noise_u = env.noise_u[:, :235].contiguous()
    bobs, _, _, _, _ = base.step(torch.zeros(n, 12, device=DEV))
    assert torch.allclose(obs[:, :235], bobs, rtol=RTOL, atol=ATOL)


@pytest.mark.gpu
def test_rollout_percept_sdf_values_one_launch():
    from extended_legged_gym_b200.envs import RobotBatchRolloutPercept
    mains, rollouts, rays = 6, 7, 16
    n = mains * (1 + rollouts)
    bodies = ["base", "LF_FOOT", "RH_FOOT"]
    offs = [[0.1, -0.05, 0.02], [0.0, 0.0, -0.03]]

    def mod(cfg):
        cfg.env.num_envs, cfg.env.rollout_envs = mains, rollouts
        cfg.raycaster.enable_raycast, cfg.raycaster.ray_pattern = True, "spherical"
        cfg.raycaster.spherical_num_azimuth, cfg.raycaster.spherical_num_elevation, cfg.raycaster.max_distance = 4, 4, 2.5
        cfg.sdf = type("sdf", (), dict(enable_sdf=True, mesh_paths=[], max_distance=1.0, enable_caching=True, update_freq=1, query_bodies=bodies,
                                       collision_sphere_radius=[], collision_sphere_pos=offs, compute_gradients=True, compute_nearest_points=True,
                                       include_in_obs=True))
        cfg.env.num_observations = 235 + rays + len(bodies)
    env, st, v, t = small_terrain_env(RobotBatchRolloutPercept, mod, n)
    env.add_noise = False
    env.post_physics_step()
    torch.cuda.synchronize()
    idx = [env.sim.spec.body_names.index(b) for b in bodies]
    pts = so.sdf_query_points(st["rigid_body_state"], env.num_bodies, idx, [torch.tensor(o) for o in offs] + [None]) if False else None
    per_body = []
    for k, b in enumerate(idx):
        per_body.append(so.sdf_query_points(st["rigid_body_state"], env.num_bodies, [b], [torch.tensor(offs[k])] if k < len(offs) else None)[:, 0])
    pts = torch.stack(per_body, dim=1)
    assert torch.allclose(env.sdf_query_points.cpu(), pts, rtol=RTOL, atol=ATOL)
    ws, wg, _, _ = mo.sdf_query(env.sdf_query_points.cpu().numpy().reshape(-1, 3), 1.0, v, t)
    got_s = env.sdf_values.cpu().numpy().reshape(-1)
    assert np.allclose(got_s, ws, rtol=RTOL, atol=ATOL)
    assert (np.abs(ws) < 1.0).mean() > 0.3
    assert np.allclose(env.sdf_gradients.cpu().numpy().reshape(-1, 3), wg, rtol=1e-4, atol=1e-5)
    near = so.nearest_points(env.sdf_query_points.cpu(), env.sdf_values.cpu(), env.sdf_gradients.cpu())
    assert torch.allclose(env.sdf_nearest_points.cpu(), near, rtol=RTOL, atol=ATOL)
    assert torch.equal(env.obs_buf[:, 235 + rays:], env.sdf_values) and env.sdf_values.data_ptr() == env.obs_buf[:, 235 + rays:].data_ptr()
    want = so.raycast_distances(env.ray_caster.data.ray_hits.cpu(), env.ray_caster.data.ray_hits_found.cpu(), st["root_states"][:, :3], 2.5)
    assert torch.allclose(env.obs_buf[:, 235:235 + rays].cpu(), want, rtol=RTOL, atol=ATOL)
    # the rollout-mode step refreshes both sensors as well (_post_physics_step_callback_rollout)
    env.sdf_values.zero_()
    env.post_physics_step_rollout()
    assert np.allclose(env.sdf_values.cpu().numpy().reshape(-1), ws, rtol=RTOL, atol=ATOL)


@pytest.mark.gpu
def test_legged_robot_depth_owns_a_camera_with_update_interval():
    from extended_legged_gym_b200.envs import LeggedRobotDepth
    from extended_legged_gym_b200.utils.depth_camera import DepthCameraWarp
    n = 16

    def mod(cfg):
        cfg.env.num_envs = n
        cfg.depth.camera_type, cfg.depth.update_interval = "Warp", 2
    env, st, v, t = small_terrain_env(LeggedRobotDepth, mod, n)
    assert env.is_depth_enabled()
    env.step(torch.zeros(n, 12, device=DEV))                 # counter 0: the camera updates
    first = env.get_depth_images().clone()
    ref = DepthCameraWarp(env.cfg.depth, DEV, n, v, t)
    ref.update(env.dt, env.root_states[:, :3], env.root_states[:, 3:7])
    ref.update_depth_buffer(None, env.episode_length_buf)
    assert torch.equal(first, ref.depth_buffer)
    assert first.shape == (n, 2, 28, 56) and float(first.abs().max()) <= 0.6 and float(first.std()) > 0   # (bicubic taps overshoot +-0.5)
    env.root_states[:, 0] += 0.3
    env.step(torch.zeros(n, 12, device=DEV))                 # counter 1: skipped
    assert torch.equal(env.get_depth_images(), first)
    env.step(torch.zeros(n, 12, device=DEV))                 # counter 2: updated from the moved robots
    assert not torch.equal(env.get_depth_images(), first)
    assert env.get_depth_observation().shape == (n, 28, 56)


@pytest.mark.gpu
def test_elspider_async_gait_scheduler_term_and_go2_class():
    from extended_legged_gym_b200.envs import TASKS, ElSpider, Go2
    from extended_legged_gym_b200.sim_backend import SyntheticSim
    assert TASKS["go2_rough"][0] is Go2
    n = Z["async__dof_pos"].shape[0]
    cfg, spec, st = common.make_case_state("elspider_air_rough", n, seed=0)
    cfg.env.num_envs = n
    cfg.rewards.scales.async_gait_scheduler = -0.3
    cfg.rewards.async_gait_scheduler = type("a", (), dict(dof_align=1.0, dof_nominal_pos=[0.0, 0.2], reward_foot_z_align=[0.0, 0.6]))
    st["dof_state"].view(n, 18, 2)[..., 0] = T("async__dof_pos")
    env = ElSpider(cfg, None, SyntheticSim(cfg, n, DEV, spec=spec, height_samples=synthetic.make_height_field(seed=0), state=st), DEV, True)
    env.set_env_state(st)
    assert "async_gait_scheduler" in env._python_terms
    env.async_gait_scheduler.foot_pos = T("async__foot_pos").to(DEV)
    for stage in (0, 1):
        env.reward_scales_stage = stage
        got = env._reward_async_gait_scheduler().cpu()
        assert torch.allclose(got, T(f"async__combined_stage{stage}"), rtol=RTOL, atol=ATOL)
    # the term takes part in the step: rew_buf differs from the run without it by scale * dt * value
    env.reward_scales_stage = cfg.rewards.reward_min_stage
    env.add_noise = False
    env.post_physics_step()
    assert float(env.episode_sums["async_gait_scheduler"].abs().sum()) > 0
    # Go2 = the hooks of Anymal on the go2 asset (gait clock in the kernel)
    cfg2, spec2, st2 = common.make_case_state("go2_rough", 32, seed=0)
    cfg2.env.num_envs = 32
    g = Go2(cfg2, None, SyntheticSim(cfg2, 32, DEV, spec=spec2, height_samples=synthetic.make_height_field(seed=0), state=st2), DEV, True)
    g.set_env_state(st2)
    g0 = g.gait_idx.clone()
    g.post_physics_step()
    assert g.gait_cfg.period == 0.6 and torch.allclose(g.gait_idx, torch.remainder(g0 + g.dt / 0.6, 1.0), atol=1e-6)
