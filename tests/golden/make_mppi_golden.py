"""Generate tests/golden/mppi.npz from the UNMODIFIED in-tree MPPI update of the reference
(legged_gym/tests/score_sampling/cmp_mppi_wbfo.py:188-260 OptimizationComparison.mppi_optimization, one iteration),
bound to a synthetic ``self`` whose sampler returns fixed samples and whose cost function returns fixed step rewards.

    python tests/golden/make_mppi_golden.py
"""
import importlib.util
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402


def reference_update(step_rewards, samples, temp):
    """step_rewards [S, T], samples [S, K, D] -> updated mean trajectory [K, D] through the reference method."""
    ref_harness.install()
    from legged_gym.envs.base.legged_robot import LeggedRobot  # noqa: F401
    path = os.path.join(ref_harness.REFERENCE_ROOT, "legged_gym", "legged_gym", "tests", "score_sampling", "cmp_mppi_wbfo.py")
    spec = importlib.util.spec_from_file_location("cmp_mppi_wbfo_ref", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    S, K, D = samples.shape
    me = SimpleNamespace(num_samples=S, temp_sample=temp, device="cpu", sample_trajectory=lambda mean: samples.clone(),
                         node2dense=lambda nodes: nodes)

    def cost(dense_with_time):
        return step_rewards if dense_with_time.shape[0] == S else torch.zeros(1, step_rewards.shape[1])
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        _, hist = mod.OptimizationComparison.mppi_optimization(me, torch.zeros(K, D), cost, num_iterations=1)
    return hist["mean_traj"][1]


def main():
    g = torch.Generator().manual_seed(0)
    out = {}
    for tag, (M, S, T, K, D, temp) in {"a": (3, 127, 16, 5, 12, 0.05), "b": (2, 64, 20, 5, 12, 0.1)}.items():
        r = torch.randn(M, S, T, generator=g) * 0.3 + torch.randn(M, S, 1, generator=g)
        u = torch.randn(M, S, K, D, generator=g)
        trajs = torch.stack([reference_update(r[m], u[m], temp) for m in range(M)])
        out.update({f"{tag}__rewards": r.numpy(), f"{tag}__samples": u.numpy(), f"{tag}__temp": np.array([temp], np.float32),
                    f"{tag}__mean_traj": trajs.numpy()})
    path = os.path.join(ROOT, "tests", "golden", "mppi.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
