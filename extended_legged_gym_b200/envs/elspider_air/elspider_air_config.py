"""ElSpider Air (hexapod: 6 legs x (HAA, HFE, KFE) = 18 DOF) rough-terrain config.  Attribute names and values follow
legged_gym/legged_gym/envs/elspider_air/mixed_terrains/elspider_air_rough_config.py:36-170 of the reference
(``tests/test_oracle_pinned.py::test_configs_match_reference`` compares the two dictionaries in the build container)."""
from ..base.legged_robot_config import LeggedRobotCfg, LeggedRobotCfgPPO

_JOINT_ANGLE = {"HAA": 0.0, "HFE": 0.6, "KFE": 0.6}          # default angle per joint type [rad]
_JOINT_KP, _JOINT_KD = 80.0, 2.0                             # the ANYdrive PD gains
_ROOT = "{LEGGED_GYM_ROOT_DIR}/resources"


def _per_joint(value_by_type, legs=("LB", "LF", "LM", "RB", "RF", "RM")):
    return {f"{leg}_{joint}": value for joint, value in value_by_type.items() for leg in legs}


class ElSpiderAirRoughCfg(LeggedRobotCfg):
    class env(LeggedRobotCfg.env):
        num_actions = 18
        num_observations = 12 + 3 * 18 + 187
        num_envs = 4096

    class init_state(LeggedRobotCfg.init_state):
        default_joint_angles = _per_joint(_JOINT_ANGLE)
        pos = [0.0, 0.0, 0.4]

    class control(LeggedRobotCfg.control):
        use_actuator_network = True
        actuator_net_file = _ROOT + "/actuator_nets/anydrive_v3_lstm.pt"
        action_scale = 0.5
        decimation = 4
        stiffness = {joint: _JOINT_KP for joint in _JOINT_ANGLE}
        damping = {joint: _JOINT_KD for joint in _JOINT_ANGLE}

    class asset(LeggedRobotCfg.asset):
        name = "elspider_air"
        file = _ROOT + "/robots/el_mini/urdf/el_mini.urdf"
        foot_name = "FOOT"
        terminate_after_contacts_on = ["trunk"]
        penalize_contacts_on = ["THIGH", "HIP"]
        flip_visual_attachments = False
        self_collisions = 0

    class terrain(LeggedRobotCfg.terrain):
        mesh_type = "trimesh"
        measure_heights = True
        curriculum = True
        max_init_terrain_level = 0
        border_size = 100
        num_rows, num_cols = 10, 10
        terrain_length, terrain_width = 8.0, 8.0
        terrain_proportions = [0.1, 0.1, 0.3, 0.3, 0.2]      # smooth slope, rough slope, stairs up, stairs down, discrete

    class domain_rand(LeggedRobotCfg.domain_rand):
        added_mass_range = [-5.0, 5.0]
        randomize_base_mass = True

    class rewards(LeggedRobotCfg.rewards):
        only_positive_rewards = True
        max_contact_force = 500.0
        base_height_target = 0.25
        # two reward stages: a list-valued scale holds one value per stage; the stage advances when the mean episode reward
        # passes the threshold (LeggedRobotRewMixin.update_reward_scales)
        multi_stage_rewards = True
        reward_min_stage, reward_max_stage = 0, 1
        reward_stage_threshold = 5.0

        class scales(LeggedRobotCfg.rewards.scales):
            # tracking
            tracking_lin_vel, tracking_ang_vel = 1.0, 0.5
            # base motion and posture
            lin_vel_z, ang_vel_xy, orientation = -2.0, -0.05, -0.3
            base_height = [-2.0, -4.0]
            # joints
            torques, dof_acc, dof_pos_limits, action_rate = -0.00001, -2.5e-8, -1.0, -0.001
            # feet and contacts
            feet_air_time, collision, gait_2_step = 1.0, -1.0, -5.0
            feet_slip = [-0.0, -0.4]
            # present but switched off
            termination = dof_vel = feet_stumble = stand_still = -0.0

        class async_gait_scheduler:       # weights of the async_gait_scheduler term (disabled in scales)
            reward_foot_z_align = [0.2, 0.05]
            dof_nominal_pos = [0.1, 0.2]
            dof_align = 0.5

        class raibert_planner:            # read by the FootTrackElSpider variant only
            planner_type = 0
            foot_pos_track, base_quat_track, base_pos_track = 0.3, 0.5, 1.0


class ElSpiderAirRoughCfgPPO(LeggedRobotCfgPPO):
    class runner(LeggedRobotCfgPPO.runner):
        experiment_name = "rough_elspider_air"
        run_name = ""
