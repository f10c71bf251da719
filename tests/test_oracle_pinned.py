"""Pin the oracle (oracle/legged_oracle.py) before anything trusts it:
  1. against the committed fixtures generated from the UNMODIFIED reference (tests/golden/*.npz);
  2. against the reference itself, imported live through oracle/ref_harness.py, when
     /root/reference is present (this container), at a larger N and other seeds.
Same torch ops in the same order on the same CPU => the comparison is bit-exact."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import common  # noqa: E402
from oracle import ref_harness  # noqa: E402
from oracle.legged_oracle import LeggedOracle  # noqa: E402
from extended_legged_gym_b200 import synthetic  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(case):
    z = np.load(os.path.join(GOLDEN, f"{case}.npz"))
    inputs = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in__")}
    steps = int(z["meta__steps"])
    outs = [{k.split("__", 1)[1]: torch.from_numpy(np.asarray(z[k])) for k in z.files if k.startswith(f"s{s}__")}
            for s in range(steps)]
    return inputs, outs


def step_seed(s, seed=0):
    return 5000 + 17 * s + seed


@pytest.mark.parametrize("case", list(common.CASES))
def test_oracle_matches_golden(case):
    inputs, outs = load_golden(case)
    cfg_cls, spec_fn, _ = common.CASES[case]
    ora = LeggedOracle(cfg_cls(), spec_fn(), {k: v.clone() for k, v in inputs.items()}, synthetic.make_height_field(seed=0))
    for s, want in enumerate(outs):
        torch.manual_seed(step_seed(s))
        ora.torques = ora.compute_torques(ora.actions).view(ora.torques.shape)
        ora.post_physics_step(noise_u=want["noise_u"])
        got = common.snapshot(ora)
        for k, w in want.items():
            if k == "noise_u":
                continue
            if k.startswith("extras__"):
                g = float(ora.extras["episode"][k[8:]])
                assert g == float(w), f"{case} step {s} {k}: {g} vs {float(w)}"
                continue
            assert torch.equal(got[k], w.to(got[k].dtype)), f"{case} step {s}: {k} differs from the reference fixture"
    # the fixtures exercise the sparse paths too
    assert any(bool(o["reset_buf"].any()) for o in outs)


def test_golden_covers_every_reward_term():
    seen = set()
    for case in common.CASES:
        _, outs = load_golden(case)
        seen |= {k[4:] for k in outs[0] if k.startswith("sum_")}
    from extended_legged_gym_b200 import _lib
    missing = set(_lib.REWARD_TERMS) - seen - {"gait_scheduler"}
    assert not missing, f"no fixture exercises: {sorted(missing)}"


@pytest.mark.reference
@pytest.mark.skipif(not ref_harness.available(), reason="/root/reference not present")
@pytest.mark.parametrize("case", list(common.CASES))
def test_oracle_matches_live_reference(case):
    sys.path.insert(0, GOLDEN)
    import make_golden
    n, seed = 512, 3
    cfg, spec, st = common.make_case_state(case, n, seed=seed, adversarial=True)
    hf = synthetic.make_height_field(seed=0)
    env = make_golden.reference_env_for(case, spec, {k: v.clone() for k, v in st.items()}, hf)
    ora = LeggedOracle(cfg, spec, {k: v.clone() for k, v in st.items()}, hf)
    for s in range(2):
        torch.manual_seed(step_seed(s, seed))
        env.torques = env._compute_torques(env.actions).view(env.torques.shape)
        env.post_physics_step()
        torch.manual_seed(step_seed(s, seed))
        ora.torques = ora.compute_torques(ora.actions).view(ora.torques.shape)
        ora.post_physics_step()
        a, b = common.snapshot(env), common.snapshot(ora)
        assert set(a) == set(b)
        for k in a:
            assert torch.equal(a[k], b[k]), f"{case} step {s}: {k}"


@pytest.mark.reference
@pytest.mark.skipif(not ref_harness.available(), reason="/root/reference not present")
def test_configs_match_reference():
    from extended_legged_gym_b200.utils.helpers import class_to_dict
    rc = ref_harness.reference_classes()
    from legged_gym.utils.helpers import class_to_dict as ref_c2d
    for case, (cfg_cls, _, ref_name) in common.CASES.items():
        if ref_name:
            assert class_to_dict(cfg_cls()) == ref_c2d(rc[ref_name]()), case
    from extended_legged_gym_b200.envs import ElSpiderAirRoughCfg
    mine, ref = class_to_dict(ElSpiderAirRoughCfg()), ref_c2d(rc["ElSpiderAirRoughCfg"]())
    for d in (mine, ref):      # machine-local path of the reference config / a key only this repo's terrain class has
        d["terrain"].pop("terrain_file", None)
        d["terrain"].pop("confined_terrain_proportions", None)
    assert mine == ref, "elspider_air_rough"
    from extended_legged_gym_b200.envs.base.legged_robot_config import LeggedRobotCfg
    assert class_to_dict(LeggedRobotCfg()) == ref_c2d(rc["LeggedRobotCfg"]())
