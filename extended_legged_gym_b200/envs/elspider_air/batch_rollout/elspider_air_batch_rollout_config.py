"""Config of ``ElSpiderAirBatchRollout`` (values of envs/elspider_air/batch_rollout/elspider_air_batch_rollout_config.py:41-300 in
/root/reference/legged_gym/legged_gym; compared with the reference's class by tests/test_robot_rollout_classes.py)."""
from ...batch_rollout.robot_batch_rollout_config import RobotBatchRolloutPerceptCfg, RobotBatchRolloutCfgPPO
from ....utils.gait_scheduler import AsyncGaitSchedulerCfg


class ElSpiderAirBatchRolloutCfg(RobotBatchRolloutPerceptCfg):
    class gait_scheduler:
        period = 1.4
        duty = 0.5
        foot_phases = [0.0, 0.5, 0.0, 0.5, 0.0, 0.5]
        dt = 0.005
        swing_height = 0.07
        track_sigma = 0.25

    class async_gait_scheduler(AsyncGaitSchedulerCfg):
        dof_nominal_pos = [0.0, 1.0, 1.0] * 6

    class env(RobotBatchRolloutPerceptCfg.env):
        num_envs = 32
        rollout_envs = 0
        num_observations = 66
        num_actions = 18
        episode_length_s = 20

    class terrain(RobotBatchRolloutPerceptCfg.terrain):
        mesh_type = "confined_trimesh"
        measure_heights = False
        curriculum = True
        max_init_terrain_level = 2
        terrain_length = 6.0
        terrain_width = 6.0
        num_rows = 2
        num_cols = 1
        difficulty_scale = 0.6
        terrain_proportions = [0.2, 0.2, 0.3, 0.2, 0.1]
        confined_terrain_proportions = [0.0, 1.0, 0.0, 0.0]
        use_terrain_obj = False
        terrain_file = "resources/terrains/confined/confined_terrain.obj"

    class raycaster(RobotBatchRolloutPerceptCfg.raycaster):
        enable_raycast = False
        ray_pattern = "spherical2"
        num_rays = 10
        ray_angle = 30.0
        terrain_file = None
        max_distance = 10.0
        attach_yaw_only = False
        offset_pos = [0.0, 0.0, 0.0]
        spherical_num_azimuth = 16
        spherical_num_elevation = 8
        spherical2_num_points = 128
        spherical2_polar_axis = [0.0, 0.0, 1.0]

    class sdf(RobotBatchRolloutPerceptCfg.sdf):
        enable_sdf = False
        mesh_paths = []
        max_distance = 10.0
        enable_caching = True
        update_freq = 5
        query_bodies = ["trunk", "RF_SHANK", "RM_SHANK", "RB_SHANK", "LF_SHANK", "LM_SHANK", "LB_SHANK"]
        compute_gradients = True
        compute_nearest_points = True
        include_in_obs = True

    class commands(RobotBatchRolloutPerceptCfg.commands):
        curriculum = False
        max_curriculum = 1.0
        num_commands = 4
        resampling_time = 10.0
        heading_command = False

        class ranges(RobotBatchRolloutPerceptCfg.commands.ranges):
            lin_vel_x = [-1.5, 1.5]
            lin_vel_y = [-0.6, 0.6]
            ang_vel_yaw = [-0.6, 0.6]
            heading = [-3.14, 3.14]

    class init_state(RobotBatchRolloutPerceptCfg.init_state):
        pos = [0.0, 0.0, 0.32]
        rot = [0.0, 0.0, 0.0, 1.0]
        default_joint_angles = {f"{leg}_{j}": v for j, v in (("HAA", 0.0), ("HFE", 0.6), ("KFE", 0.6))
                                for leg in ("RF", "RM", "RB", "LF", "LM", "LB")}

    class control(RobotBatchRolloutPerceptCfg.control):
        stiffness = {'HAA': 80.0, 'HFE': 80.0, 'KFE': 80.0}
        damping = {'HAA': 2.0, 'HFE': 2.0, 'KFE': 2.0}
        action_scale = 0.2
        decimation = 4
        use_actuator_network = False
        actuator_net_file = "{LEGGED_GYM_ROOT_DIR}/resources/actuator_nets/anydrive_v3_lstm.pt"

    class asset(RobotBatchRolloutPerceptCfg.asset):
        file = "{LEGGED_GYM_ROOT_DIR}/resources/robots/el_mini/urdf/el_mini_collsp.urdf"
        name = "elspider"
        foot_name = "FOOT"
        penalize_contacts_on = ["base", "HIP", "THIGH", "SHANK"]
        terminate_after_contacts_on = []
        self_collisions = 0
        flip_visual_attachments = False

    class rewards(RobotBatchRolloutPerceptCfg.rewards):
        max_contact_force = 500.0
        base_height_target = 0.34
        only_positive_rewards = False
        multi_stage_rewards = True
        tracking_sigma = 0.25

        class scales:
            termination = -0.0
            tracking_lin_vel = 3.0
            tracking_ang_vel = 0.5
            lin_vel_z = -2.0
            ang_vel_xy = -0.05
            orientation = -5.0
            torques = -0.00001
            dof_vel = -0.0
            dof_acc = -0.5e-8
            base_height = -8.0
            feet_slip = [-0.0, -0.4]
            feet_air_time = 0.8
            collision = -0.05
            feet_stumble = -0.4
            feet_stumble_liftup = 1.0
            action_rate = -0.001
            stand_still = -0.0
            dof_pos_limits = -1.0
            gait_2_step = -1.0

        class async_gait_scheduler:
            dof_align = 1.0
            dof_nominal_pos = [0.05, 0.2]
            reward_foot_z_align = [0.1, 0.6]

    class domain_rand(RobotBatchRolloutPerceptCfg.domain_rand):
        randomize_base_mass = True
        added_mass_range = [-5.0, 5.0]
        rollout_envs_sync_pos_drift = 0.0


class ElSpiderAirBatchRolloutCfgPPO(RobotBatchRolloutCfgPPO):
    class policy(RobotBatchRolloutCfgPPO.policy):
        actor_hidden_dims = [128, 64, 32]
        critic_hidden_dims = [128, 64, 32]
        activation = 'elu'

    class algorithm(RobotBatchRolloutCfgPPO.algorithm):
        entropy_coef = 0.01

    class runner(RobotBatchRolloutCfgPPO.runner):
        run_name = ''
        experiment_name = 'elspider_air_batch_rollout'
        load_run = -1
        max_iterations = 3000
        multi_stage_rewards = True
