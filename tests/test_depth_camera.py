"""Warp depth camera.
not-gpu: the oracle restatement against the fixture generated from the unmodified reference class.
gpu: the fused elg_camera_pose / elg_depth_camera path against the same fixture and against the oracle at a
larger size -- camera poses and depth frames within 1e-5 rel / 1e-6 abs, hit / miss pixels identical."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_depth_golden as mg  # noqa: E402
from oracle import mesh_oracle as mo  # noqa: E402
from oracle.depth_oracle import DepthOracle  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "depth_camera.npz")
DEV = "cuda:0"
RTOL, ATOL = 1e-5, 1e-6


def steps_of(name):
    z = np.load(GOLDEN)
    out = []
    for s in range(3):
        out.append({k: torch.from_numpy(z[f"{name}__s{s}__{k}"]) for k in ("pos", "quat", "ep", "u", "camera_pos", "camera_rot", "depth_buffer", "obs")})
    return torch.from_numpy(z[f"{name}__ray_directions"]), out


@pytest.mark.parametrize("name", list(mg.CASES))
def test_oracle_matches_reference_fixture(name):
    c = mg.CASES[name]
    v, t, _ = mo.heightfield_mesh(40, 40, seed=3)
    ora = DepthOracle(mg.make_cfg(c), c["n"], v, t)
    dirs, steps = steps_of(name)
    assert torch.equal(ora.ray_directions[0], dirs)
    for s in steps:
        ora.update(s["pos"], s["quat"])
        assert torch.equal(ora.camera_pos, s["camera_pos"]) and torch.equal(ora.camera_rot, s["camera_rot"])
        ora.update_depth_buffer(s["ep"], s["u"])
        assert torch.equal(ora.depth_buffer, s["depth_buffer"])
        assert torch.equal(ora.depth_buffer[:, -2], s["obs"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(mg.CASES))
def test_depth_camera_matches_reference_fixture(name):
    from extended_legged_gym_b200.utils.depth_camera import DepthCameraWarp
    c = mg.CASES[name]
    v, t, _ = mo.heightfield_mesh(40, 40, seed=3)
    cam = DepthCameraWarp(mg.make_cfg(c), DEV, c["n"], v, t)
    dirs, steps = steps_of(name)
    assert torch.equal(cam.ray_directions[0].cpu(), dirs)
    for i, s in enumerate(steps):
        cam.update(0.02, s["pos"].to(DEV), s["quat"].to(DEV))
        cam.noise_u = s["u"].to(DEV)
        cam.update_depth_buffer(None, s["ep"].to(DEV))
        torch.cuda.synchronize()
        assert torch.allclose(cam.camera_pos.cpu(), s["camera_pos"], rtol=RTOL, atol=ATOL)
        assert torch.allclose(cam.camera_rot.cpu(), s["camera_rot"], rtol=RTOL, atol=ATOL)
        assert torch.allclose(cam.depth_buffer.cpu(), s["depth_buffer"], rtol=RTOL, atol=2e-6), f"step {i}"
        assert torch.allclose(cam.get_depth_observation().cpu(), s["obs"], rtol=RTOL, atol=2e-6)


@pytest.mark.gpu
def test_depth_camera_matches_oracle_default_config():
    """cfg.depth defaults (legged_robot_config.py:105-125): 60x30 rays, resized to 56x28, hfov 100, far 2, 2 frames."""
    from extended_legged_gym_b200.utils.depth_camera import DepthCameraFake, DepthCameraWarp
    c = dict(original=(60, 30), resized=(56, 28), far_clip=2.0, near_clip=0.0, dis_noise=0.0, buffer_len=2, n=6, seed=4)
    v, t, _ = mo.heightfield_mesh(60, 60, seed=9)
    cfg = mg.make_cfg(c)
    cam = DepthCameraWarp(cfg, DEV, c["n"], v, t)
    cam.raw_depth = torch.zeros(c["n"], 30, 60, device=DEV)
    ora = DepthOracle(cfg, c["n"], v, t)
    for step in range(3):
        pos, quat = mg.poses(c["n"], 40 + step)
        pos = pos * 1.5
        ep = torch.tensor([0, 1, 2, 50, 1, 9]) + step
        ora.update(pos, quat)
        ora.update_depth_buffer(ep)
        cam.update(0.02, pos.to(DEV), quat.to(DEV))
        cam.update_depth_buffer(None, ep.to(DEV))
        torch.cuda.synchronize()
        raw = cam.raw_depth.cpu()
        assert torch.equal(raw == -2.0, ora.raw_depth == -2.0), "hit / miss pixels differ"
        assert torch.allclose(raw, ora.raw_depth, rtol=RTOL, atol=ATOL)
        assert torch.allclose(cam.depth_buffer.cpu(), ora.depth_buffer, rtol=RTOL, atol=2e-6), f"step {step}"
    assert 0.05 < float((ora.raw_depth > -2.0).float().mean()) < 1.0
    fake = DepthCameraFake(cfg, DEV, 3)
    fake.update(0.02, None, None)
    fake.update_depth_buffer(None, None)
    assert fake.depth_buffer.shape == (3, 2, 28, 56) and bool((fake.get_depth_observation() == -0.5).all())


@pytest.mark.gpu
def test_depth_camera_partial_pose_update_and_ring_buffer():
    from extended_legged_gym_b200.utils.depth_camera import DepthCameraWarp
    c = dict(original=(16, 8), resized=(16, 8), far_clip=4.0, near_clip=0.0, dis_noise=0.0, buffer_len=4, n=8, seed=0)
    v, t, _ = mo.heightfield_mesh(40, 40, seed=3)
    cam = DepthCameraWarp(mg.make_cfg(c), DEV, 8, v, t)
    pos, quat = mg.poses(8, 3)
    cam.update(0.02, pos.to(DEV), quat.to(DEV))
    p0 = cam.camera_pos.clone()
    cam.update(0.02, (pos + 1.0).to(DEV), quat.to(DEV), env_ids=torch.tensor([2, 5], device=DEV))
    moved = (cam.camera_pos != p0).any(dim=1).cpu().tolist()
    assert moved == [False, False, True, False, False, True, False, False]
    cam.update(0.02, pos.to(DEV), quat.to(DEV))
    frames = []
    for k in range(5):
        cam.update(0.02, (pos + 0.05 * k).to(DEV), quat.to(DEV))
        cam.update_depth_buffer(None, torch.full((8,), 1 + k, device=DEV))
        frames.append(cam.depth_buffer[:, -1].clone())
        if k == 0:
            assert all(torch.equal(cam.depth_buffer[:, j], frames[0]) for j in range(4))      # initialised with the first frame
    assert all(torch.equal(cam.depth_buffer[:, j], frames[1 + j]) for j in range(4))          # then a sliding window
