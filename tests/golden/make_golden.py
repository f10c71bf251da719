"""Generate the golden fixtures under tests/golden/ FROM THE UNMODIFIED REFERENCE.

Runs in the build container only (needs /root/reference; the third-party binaries the reference
imports are stubbed by oracle/ref_harness.py, its own LeggedRobot methods run unmodified on
torch-CPU).  For every case of tests/common.CASES: seeded synthetic PhysX state (N envs), then
STEPS x [ _compute_torques -> post_physics_step() ] of the reference, with the global CPU RNG
seeded per step.  Inputs, per-step RNG seeds and all resulting tensors are stored in one .npz per
case so that the GPU box (which has no /root/reference) can check both the oracle port and the
CUDA path against the reference's own numbers.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_harness as rh  # noqa: E402
import common  # noqa: E402
from extended_legged_gym_b200 import synthetic  # noqa: E402
from extended_legged_gym_b200.utils.helpers import class_to_dict, update_class_from_dict  # noqa: E402

N_ENVS, STEPS = 96, 3
HF_SEED = 0


def reference_cfg_for(case):
    """Reference config instance carrying exactly the values of our case config."""
    cfg_cls, _, ref_name = common.CASES[case]
    rc = rh.reference_classes()
    base = ref_name or {"a1_all_terms": "A1RoughCfg", "go2_all_terms_heading": "Go2RoughCfg", "elspider_all_terms": "ElSpiderAirRoughCfg",
                        "elspider_air_rough": "ElSpiderAirRoughCfg", "a1_control_V": "A1RoughCfg", "go2_control_T": "Go2RoughCfg"}[case]
    ref_cfg = rc[base]()
    update_class_from_dict(ref_cfg, class_to_dict(cfg_cls()))
    return ref_cfg


def reference_env_for(case, spec, st, hf):
    """the reference's own class for the case (LeggedRobot, or ElSpider for the hexapod cases) on the synthetic state"""
    cls = rh.reference_classes()[common.REF_ENV_CLASS[case]] if case in common.REF_ENV_CLASS else None
    env = rh.make_reference_env(reference_cfg_for(case), spec, st, hf, cls=cls)
    if cls is not None:
        # ElSpider.post_physics_step (elspider.py:334-338) also advances a GaitScheduler object its skipped __init__ would have
        # created; the scheduler does not feed any enabled term of these cases, the harness supplies an inert one
        from types import SimpleNamespace
        env.gait_scheduler = SimpleNamespace(step=lambda *a, **k: None)
    return env


def height_field():
    return synthetic.make_height_field(seed=HF_SEED)


def run_reference(case, n_envs=N_ENVS, steps=STEPS, seed=0, adversarial=True):
    cfg, spec, st = common.make_case_state(case, n_envs, seed=seed, adversarial=adversarial)
    inputs = {k: v.clone() for k, v in st.items()}
    hf = height_field()
    env = reference_env_for(case, spec, st, hf)
    g = torch.Generator().manual_seed(1000 + seed)
    out = {}
    for s in range(steps):
        noise_u = torch.rand(n_envs, env.num_obs, generator=g)
        torch.manual_seed(5000 + 17 * s + seed)
        # the reference draws its observation noise with torch.rand_like; feed it the same numbers
        # the CUDA path gets by making rand_like return the prepared tensor for this step
        orig = torch.rand_like
        torch.rand_like = lambda t, *a, **k: noise_u.clone() if t.shape == noise_u.shape else orig(t, *a, **k)
        try:
            env.torques = env._compute_torques(env.actions).view(env.torques.shape)
            env.post_physics_step()
        finally:
            torch.rand_like = orig
        snap = common.snapshot(env)
        if torch.is_tensor(env.measured_heights):
            pass
        for k, v in snap.items():
            out[f"s{s}__{k}"] = v.numpy()
        out[f"s{s}__noise_u"] = noise_u.numpy()
        for k, v in env.extras.get("episode", {}).items():
            out[f"s{s}__extras__{k}"] = np.asarray(float(v))
    return inputs, out


def main():
    only = sys.argv[1:]
    for case in common.CASES:
        if only and case not in only:
            continue
        inputs, out = run_reference(case)
        blob = {f"in__{k}": v.numpy() for k, v in inputs.items()}
        blob.update(out)
        blob["meta__n_envs"] = np.asarray(N_ENVS)
        blob["meta__steps"] = np.asarray(STEPS)
        path = os.path.join(HERE, f"{case}.npz")
        np.savez_compressed(path, **blob)
        print(f"{case}: {len(blob)} arrays -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
