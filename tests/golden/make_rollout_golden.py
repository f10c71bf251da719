"""Generate tests/golden/rollout_clone.npz from the UNMODIFIED reference (container only):
RobotBatchRollout._init_env_indices / _sync_main_to_rollout / _cache_main_env_states / _restore_main_env_states
(envs/batch_rollout/robot_batch_rollout.py) bound to a synthetic ``self``.

    python tests/golden/make_rollout_golden.py
"""
import os
import sys
from types import SimpleNamespace
from unittest import mock

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness, rollout_oracle as ro  # noqa: E402


def reference_run(num_main, rollouts, seed, drift):
    ref_harness.install()
    from legged_gym.envs.batch_rollout.robot_batch_rollout import RobotBatchRollout as Ref
    o = ro.make_rollout_state(num_main, rollouts, seed=seed)
    o.gym, o.sim = mock.MagicMock(), None
    o.t_main = o.t_rollout = 0.0
    o.cfg = SimpleNamespace(domain_rand=SimpleNamespace(rollout_envs_sync_pos_drift=drift))
    inputs = ro.snapshot(o)
    Ref._init_env_indices(o)
    stages = {}
    torch.manual_seed(1000 + seed)
    u = torch.rand_like(o.base_pos[o.rollout_env_indices])       # the sample _sync_main_to_rollout draws next
    torch.manual_seed(1000 + seed)
    Ref._sync_main_to_rollout(o)
    stages["sync"] = ro.snapshot(o)
    Ref._cache_main_env_states(o)
    # perturb everything, then restore the main rows
    g = torch.Generator().manual_seed(77 + seed)
    for k in ro.STATE_KEYS:
        t = getattr(o, k)
        if t.dtype == torch.bool:
            t ^= torch.rand(t.shape, generator=g) > 0.5
        else:
            t += torch.randn(t.shape, generator=g)
    stages["perturbed"] = ro.snapshot(o)
    Ref._restore_main_env_states(o)
    stages["restore"] = ro.snapshot(o)
    idx = dict(main_env_indices=o.main_env_indices, rollout_env_indices=o.rollout_env_indices, rollout_to_main_map=o.rollout_to_main_map)
    return inputs, u, stages, idx


def main():
    out = {}
    for tag, (m, r, seed, drift) in {"a": (3, 5, 0, 0.0), "b": (4, 7, 1, 0.05)}.items():
        inputs, u, stages, idx = reference_run(m, r, seed, drift)
        out[f"{tag}__meta"] = np.array([m, r, seed], dtype=np.int64)
        out[f"{tag}__drift"] = np.array([drift], dtype=np.float32)
        out[f"{tag}__drift_u"] = u.numpy()
        for k, v in inputs.items():
            out[f"{tag}__in__{k}"] = v.numpy()
        for st, d in stages.items():
            for k, v in d.items():
                out[f"{tag}__{st}__{k}"] = v.numpy()
        for k, v in idx.items():
            out[f"{tag}__idx__{k}"] = v.numpy()
    path = os.path.join(ROOT, "tests", "golden", "rollout_clone.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
