"""Unitree Go2 rough-terrain config with heading commands
(legged_gym/legged_gym/envs/go2/flat/go2_rough_config.py)."""
from ..base.legged_robot_config import LeggedRobotCfg, LeggedRobotCfgPPO
from ..a1.a1_config import _unitree_angles


class Go2RoughCfg(LeggedRobotCfg):
    class env(LeggedRobotCfg.env):
        num_envs = 4096
        num_observations = 235
        num_privileged_obs = None
        num_actions = 12

    class terrain(LeggedRobotCfg.terrain):
        curriculum = True
        mesh_type = "trimesh"
        measure_heights = True

    class commands(LeggedRobotCfg.commands):
        curriculum = False
        max_curriculum = 1.0
        num_commands = 4
        resampling_time = 10.0
        heading_command = True

        class ranges:
            lin_vel_x = [-1.0, 1.0]
            lin_vel_y = [-1.0, 1.0]
            ang_vel_yaw = [-1, 1]
            heading = [-3.14, 3.14]

    class init_state(LeggedRobotCfg.init_state):
        pos = [0.0, 0.0, 0.33]
        default_joint_angles = _unitree_angles(0.1, 0.8, 0.8, -1.5)

    class control(LeggedRobotCfg.control):
        stiffness = {"joint": 30.0}
        damping = {"joint": 0.8}
        action_scale = 0.3
        decimation = 4
        use_actuator_network = False

    class asset(LeggedRobotCfg.asset):
        file = "{LEGGED_GYM_ROOT_DIR}/resources/robots/go2/urdf/go2_description.urdf"
        name = "go2"
        foot_name = "foot"
        penalize_contacts_on = ["thigh", "calf"]
        terminate_after_contacts_on = ["base", "Head_upper"]
        self_collisions = 1
        flip_visual_attachments = True
        fix_base_link = False

    class rewards(LeggedRobotCfg.rewards):
        soft_dof_pos_limit = 0.9
        max_contact_force = 350.0
        base_height_target = 0.25

        class scales(LeggedRobotCfg.rewards.scales):
            orientation = -0.5
            action_rate = -0.001

    class domain_rand(LeggedRobotCfg.domain_rand):
        randomize_friction = True
        friction_range = [0.5, 1.25]
        randomize_base_mass = True


class Go2RoughCfgPPO(LeggedRobotCfgPPO):
    class runner(LeggedRobotCfgPPO.runner):
        run_name = ""
        experiment_name = "rough_go2"
