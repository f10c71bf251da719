"""Unitree A1 rough-terrain config (legged_gym/legged_gym/envs/a1/a1_config.py:33-84)."""
from ..base.legged_robot_config import LeggedRobotCfg, LeggedRobotCfgPPO


def _unitree_angles(hip, thigh_front, thigh_rear, calf):
    out = {}
    for leg in ("FL", "RL", "FR", "RR"):
        out[f"{leg}_hip_joint"] = hip if leg[1] == "L" else -hip
        out[f"{leg}_thigh_joint"] = thigh_front if leg[0] == "F" else thigh_rear
        out[f"{leg}_calf_joint"] = calf
    return out


class A1RoughCfg(LeggedRobotCfg):
    class init_state(LeggedRobotCfg.init_state):
        pos = [0.0, 0.0, 0.42]
        default_joint_angles = _unitree_angles(0.1, 0.8, 1.0, -1.5)

    class control(LeggedRobotCfg.control):
        control_type = "P"
        stiffness = {"joint": 20.0}
        damping = {"joint": 0.5}
        action_scale = 0.25
        decimation = 4

    class asset(LeggedRobotCfg.asset):
        file = "{LEGGED_GYM_ROOT_DIR}/resources/robots/a1/urdf/a1.urdf"
        name = "a1"
        foot_name = "foot"
        penalize_contacts_on = ["thigh", "calf"]
        terminate_after_contacts_on = ["base"]
        self_collisions = 1

    class rewards(LeggedRobotCfg.rewards):
        soft_dof_pos_limit = 0.9
        base_height_target = 0.25

        class scales(LeggedRobotCfg.rewards.scales):
            torques = -0.0002
            dof_pos_limits = -10.0


class A1RoughCfgPPO(LeggedRobotCfgPPO):
    class algorithm(LeggedRobotCfgPPO.algorithm):
        entropy_coef = 0.01

    class runner(LeggedRobotCfgPPO.runner):
        run_name = ""
        experiment_name = "rough_a1"
