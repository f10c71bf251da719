"""Generate tests/golden/dial_mpc_terms.npz FROM THE UNMODIFIED REFERENCE (container only): the DIAL-MPC reward terms of
``AnymalCTrajGradSampling`` / ``Go2TrajGradSampling`` (envs/anymal_c/batch_rollout/anymal_c_traj_grad_sampling.py:148-356,
envs/go2/batch_rollout/go2_traj_grad_sampling.py) evaluated as bound methods on a bare object that carries seeded tensors
(tests/test_robot_rollout_classes.py:_dial_fake_self builds the same object for this repo's classes).

    python tests/golden/make_dial_golden.py
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
import test_robot_rollout_classes as T  # noqa: E402
from extended_legged_gym_b200.envs.anymal_c.batch_rollout.anymal_c_traj_grad_sampling import DialMpcRewardMixin  # noqa: E402 (tables only)


def main():
    rh.install()
    from legged_gym.envs.anymal_c.batch_rollout.anymal_c_traj_grad_sampling import AnymalCTrajGradSampling as RefA
    from legged_gym.envs.go2.batch_rollout.go2_traj_grad_sampling import Go2TrajGradSampling as RefG
    blob = {}
    for robot, Ref in (("anymal_c", RefA), ("go2", RefG)):
        for heading in (False, True):
            for gait in T.DIAL_GAITS:
                ref = T._dial_fake_self(Ref, T.DIAL_N, 5, heading, torch.long)
                ref._gait = gait
                ref._gait_phase = {k: (torch.zeros(4) if k == "stand" else torch.tensor(v)) for k, v in DialMpcRewardMixin.GAIT_PHASES.items()}
                ref._gait_params = {k: torch.tensor(v) for k, v in DialMpcRewardMixin.GAIT_PARAMS.items()}
                for name in T.DIAL_TERMS:
                    if hasattr(ref, "_reward_" + name):
                        blob[f"{robot}__{int(heading)}__{gait}__{name}"] = getattr(ref, "_reward_" + name)().float().numpy()
                blob[f"{robot}__{int(heading)}__{gait}__feet_air_time"] = ref.feet_air_time.numpy()
                blob[f"{robot}__{int(heading)}__{gait}__last_contacts"] = ref.last_contacts.numpy()
    path = os.path.join(HERE, "dial_mpc_terms.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB", len(blob), "arrays")


if __name__ == "__main__":
    main()
