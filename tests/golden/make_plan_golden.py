"""Generate tests/golden/plan_integration.npz from the UNMODIFIED reference (container only):
RobotPlanGradSampling._integrate_state_velocities + _sync_integration_to_sim
(envs/batch_rollout/robot_plan_grad_sampling.py:103-225) bound to a synthetic ``self``; Euler with sub-stepping on an index
subset with joint limits, and the "rk4" form on all envs.

    python tests/golden/make_plan_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import plan_oracle as po, ref_harness  # noqa: E402

CASES = {"a": dict(n=24, d=18, seed=0, method="euler", enforce=True, max_step=0.01, dt=0.02, subset=True),
         "b": dict(n=16, d=12, seed=1, method="rk4", enforce=False, max_step=0.05, dt=0.02, subset=False)}


def main():
    ref_harness.install()
    from legged_gym.envs.batch_rollout.robot_plan_grad_sampling import RobotPlanGradSampling as Ref
    out = {}
    for tag, c in CASES.items():
        o = po.make_state(c["n"], c["d"], c["seed"], c["method"], c["enforce"], c["max_step"])
        g = torch.Generator().manual_seed(50 + c["seed"])
        idx = torch.arange(1, c["n"], 3) if c["subset"] else torch.arange(c["n"])
        sv = torch.randn(len(idx), 6 + c["d"], generator=g) * torch.tensor([2.5] * 3 + [1.5] * 3 + [6.0] * c["d"])
        sv[0, 3:6] = 0.0          # zero angular velocity: the 1e-8 guard of the axis
        for k, v in po.snapshot(o).items():
            out[f"{tag}__in__{k}"] = v.numpy()
        out[f"{tag}__in__dof_state"] = o.dof_state.clone().numpy()
        out[f"{tag}__state_vels"], out[f"{tag}__env_ids"] = sv.numpy(), idx.numpy()
        Ref._integrate_state_velocities(o, sv, c["dt"], idx)
        Ref._sync_integration_to_sim(o, idx)
        for k, v in po.snapshot(o).items():
            out[f"{tag}__out__{k}"] = v.numpy()
    path = os.path.join(ROOT, "tests", "golden", "plan_integration.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
