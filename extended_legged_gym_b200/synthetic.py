"""Synthetic PhysX state for the hot path (SURVEY.md §8d "Configs as concrete inputs").

PhysX is outside the scope of this build; the per-step path is exercised on state
tensors with the shapes, dtypes and layouts Isaac Gym hands to
``LeggedRobot._init_buffers`` (legged_gym/legged_gym/envs/base/legged_robot.py:564-584):

  root_states      [N, 13]    pos3, quat xyzw 4, lin vel 3, ang vel 3
  dof_state        [N*D, 2]   (pos, vel) interleaved -> dof_pos/dof_vel are stride-2 views
  contact_forces   [N*B, 3]
  rigid_body_state [N*B, 13]

plus the env-owned history buffers.  Everything is generated on CPU from a seeded
``torch.Generator`` so the oracle and the CUDA path consume identical bytes.
"""
import math
from typing import Dict

import numpy as np
import torch


def make_height_field(rows: int = 900, cols: int = 900, border: int = 250, tile: int = 50, seed: int = 0) -> torch.Tensor:
    """int16 [rows, cols] height samples: zero border, seeded 8x8 tiles of slope / stairs / noise.

    Shape contract of ``Terrain.heightsamples`` (legged_gym/legged_gym/utils/terrain.py:54-61):
    default cfg = 8x8 tiles of 5 m at 0.1 m/px + 2 x 25 m border = 900 x 900.
    """
    rng = np.random.default_rng(seed)
    hf = np.zeros((rows, cols), dtype=np.int16)
    inner_r, inner_c = rows - 2 * border, cols - 2 * border
    if inner_r <= 0 or inner_c <= 0:
        return torch.from_numpy(hf)
    nr, nc = max(inner_r // tile, 1), max(inner_c // tile, 1)
    for i in range(nr):
        for j in range(nc):
            r0, c0 = border + i * tile, border + j * tile
            r1 = border + inner_r if i == nr - 1 else r0 + tile
            c1 = border + inner_c if j == nc - 1 else c0 + tile
            h, w = r1 - r0, c1 - c0
            kind = (i * nc + j) % 3
            if kind == 0:      # pyramid slope
                yy, xx = np.mgrid[0:h, 0:w]
                d = np.minimum(np.minimum(yy, h - 1 - yy), np.minimum(xx, w - 1 - xx))
                patch = d * rng.integers(2, 9)
            elif kind == 1:    # concentric stairs
                yy, xx = np.mgrid[0:h, 0:w]
                d = np.minimum(np.minimum(yy, h - 1 - yy), np.minimum(xx, w - 1 - xx))
                patch = (d // 4) * rng.integers(10, 40)
            else:              # uniform noise +-10 units
                patch = rng.integers(-10, 11, size=(h, w))
            hf[r0:r1, c0:c1] = patch.astype(np.int16)
    hf[border:rows - border, border:cols - border] += rng.integers(-2, 3, size=(inner_r, inner_c)).astype(np.int16)
    return torch.from_numpy(hf)


def _quat_mul_xyzw(a, b):
    x1, y1, z1, w1 = a.unbind(-1)
    x2, y2, z2, w2 = b.unbind(-1)
    return torch.stack([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                        w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                        w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
                        w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2], dim=-1)


def make_state(num_envs: int, num_dof: int, num_bodies: int, feet_indices, penalised_indices, termination_indices,
               default_dof_pos, foot_offsets=None, num_commands: int = 4, seed: int = 0,
               xy_range=(2.0, 38.0), max_episode_length: int = 1000, clip_actions: float = 100.0) -> Dict[str, torch.Tensor]:
    """Distributions of SURVEY.md §8d.  All tensors CPU, fp32 unless noted."""
    g = torch.Generator().manual_seed(seed)
    N, D, B = num_envs, num_dof, num_bodies
    F = len(feet_indices)

    def randn(*s):
        return torch.randn(*s, generator=g, dtype=torch.float32)

    def rand(*s):
        return torch.rand(*s, generator=g, dtype=torch.float32)

    st: Dict[str, torch.Tensor] = {}
    root = torch.zeros(N, 13)
    root[:, 0:2] = xy_range[0] + (xy_range[1] - xy_range[0]) * rand(N, 2)
    root[:, 2] = 0.55 + 0.05 * randn(N)
    yaw = (2 * rand(N) - 1) * math.pi
    rp = 0.1 * randn(N, 2)
    zero = torch.zeros(N)
    q_yaw = torch.stack([zero, zero, torch.sin(yaw / 2), torch.cos(yaw / 2)], -1)
    q_roll = torch.stack([torch.sin(rp[:, 0] / 2), zero, zero, torch.cos(rp[:, 0] / 2)], -1)
    q_pitch = torch.stack([zero, torch.sin(rp[:, 1] / 2), zero, torch.cos(rp[:, 1] / 2)], -1)
    q = _quat_mul_xyzw(q_yaw, _quat_mul_xyzw(q_pitch, q_roll))
    root[:, 3:7] = q / q.norm(dim=-1, keepdim=True)
    root[:, 7:13] = 0.5 * randn(N, 6)
    st["root_states"] = root

    q0 = torch.as_tensor(default_dof_pos, dtype=torch.float32).view(1, D)
    dof_state = torch.zeros(N * D, 2)
    dof_state.view(N, D, 2)[..., 0] = q0 + 0.2 * randn(N, D)
    dof_state.view(N, D, 2)[..., 1] = randn(N, D)
    st["dof_state"] = dof_state

    st["actions"] = randn(N, D).clamp_(-clip_actions, clip_actions)
    st["last_actions"] = st["actions"] + 0.1 * randn(N, D)
    st["last_dof_vel"] = dof_state.view(N, D, 2)[..., 1] + 0.1 * randn(N, D)
    st["last_root_vel"] = root[:, 7:13] + 0.1 * randn(N, 6)
    st["base_lin_acc"] = randn(N, 3)
    st["base_ang_acc"] = randn(N, 3)

    cf = torch.zeros(N, B, 3)
    foot_contact = rand(N, F) < 0.5
    ff = torch.zeros(N, F, 3)
    ff[..., 2] = 200.0 * rand(N, F)
    ff[..., :2] = 10.0 * randn(N, F, 2)
    # a few "stumbles": large horizontal force relative to vertical
    stumble = rand(N, F) < 0.03
    ff[..., :2] = torch.where(stumble.unsqueeze(-1), ff[..., :2] * 80.0, ff[..., :2])
    cf[:, feet_indices] = ff * foot_contact.unsqueeze(-1)
    if len(penalised_indices):
        P = len(penalised_indices)
        cf[:, penalised_indices] = 5.0 * randn(N, P, 3) * (rand(N, P) < 0.02).unsqueeze(-1)
    if len(termination_indices):
        T = len(termination_indices)
        cf[:, termination_indices] = 10.0 * randn(N, T, 3) * (rand(N, T) < 0.005).unsqueeze(-1)
    st["contact_forces"] = cf.view(N * B, 3).contiguous()

    rbs = torch.zeros(N, B, 13)
    rbs[..., 0:3] = root[:, None, 0:3] + 0.3 * randn(N, B, 3)
    rbs[..., 3:7] = root[:, None, 3:7]
    rbs[..., 7:13] = 0.5 * randn(N, B, 6)
    if foot_offsets is None:
        foot_offsets = [(0.4 * (1 if i % 2 == 0 else -1), 0.25 * (1 if i < F // 2 else -1), -0.5) for i in range(F)]
    offs = torch.tensor(foot_offsets, dtype=torch.float32)
    rbs[:, feet_indices, 0:2] = root[:, None, 0:2] + offs[None, :, 0:2]
    rbs[:, feet_indices, 2] = 0.2 * rand(N, F)
    rbs[:, feet_indices, 7:10] = 0.5 * randn(N, F, 3)
    st["rigid_body_state"] = rbs.view(N * B, 13).contiguous()

    st["episode_length_buf"] = torch.randint(0, max_episode_length + 2, (N,), generator=g, dtype=torch.int64)
    cmd = 2 * rand(N, num_commands) - 1
    if num_commands > 3:
        cmd[:, 3] *= 3.14
    small = cmd[:, :2].norm(dim=1) <= 0.2
    cmd[small, :2] = 0.0
    st["commands"] = cmd
    st["feet_air_time"] = 0.6 * rand(N, F) * (~foot_contact)
    st["feet_contact_time"] = 0.6 * rand(N, F) * foot_contact
    st["last_contacts"] = rand(N, F) < 0.5
    st["gait_idx"] = rand(N)
    st["noise_u"] = rand(N, 1)  # placeholder; per-config noise tensors are drawn by the caller at [N, O]
    return st


def clone_state(st: Dict[str, torch.Tensor], device=None) -> Dict[str, torch.Tensor]:
    return {k: (v.clone() if device is None else v.to(device).clone()) for k, v in st.items()}


def heightfield_to_trimesh(height_samples, horizontal_scale: float = 0.1, vertical_scale: float = 0.005, border_size: float = 25.0):
    """Triangle mesh of an int16 height field with the vertex / triangle layout of
    ``isaacgym.terrain_utils.convert_heightfield_to_trimesh`` without slope correction (what ``Terrain`` hands to
    ``gym.add_triangle_mesh``, legged_gym/legged_gym/utils/terrain.py:76-80): vertex (i, j) at
    (i * hscale - border, j * hscale - border, h * vscale), two triangles per cell -> 900 x 900 samples give
    810 000 vertices and 1 616 402 triangles."""
    hf = np.asarray(height_samples.cpu() if torch.is_tensor(height_samples) else height_samples)
    rows, cols = hf.shape
    ii, jj = np.meshgrid(np.arange(rows, dtype=np.float32), np.arange(cols, dtype=np.float32), indexing="ij")
    v = np.stack([ii * np.float32(horizontal_scale) - np.float32(border_size), jj * np.float32(horizontal_scale) - np.float32(border_size),
                  hf.astype(np.float32) * np.float32(vertical_scale)], axis=-1).reshape(-1, 3).astype(np.float32)
    idx = np.arange(rows * cols, dtype=np.int32).reshape(rows, cols)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, 1:].ravel()
    t = np.empty((2 * (rows - 1) * (cols - 1), 3), dtype=np.int32)
    t[0::2] = np.stack([a, d, c], axis=1)
    t[1::2] = np.stack([a, b, d], axis=1)
    return v, t
