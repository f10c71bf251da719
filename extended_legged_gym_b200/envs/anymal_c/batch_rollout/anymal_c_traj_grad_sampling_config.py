"""Config of ``AnymalCTrajGradSampling`` (values of envs/anymal_c/batch_rollout/anymal_c_traj_grad_sampling_config.py:41-308 in
/root/reference/legged_gym/legged_gym; tests/test_robot_rollout_classes.py compares the blocks the per-step path reads with the
reference's class in the build container).  ``trajectory_opt`` carries the reference's values; the update rules other than "mppi"
and the spline interpolation belong to the external ``traj_sampling`` package: the in-tree MPPI update with
linear node interpolation is what runs here whatever ``update_method`` / ``interp_method`` say.  The reinforcement-learning warm
start (``rl_warmstart``: a policy checkpoint of the author's machine) is not part of the per-step path."""
from ...batch_rollout.robot_traj_grad_sampling_config import RobotTrajGradSamplingCfg, RobotTrajGradSamplingCfgPPO


class AnymalCTrajGradSamplingCfg(RobotTrajGradSamplingCfg):
    class env(RobotTrajGradSamplingCfg.env):
        num_envs = 1
        rollout_envs = 1
        env_spacing = 0.4
        num_observations = 48
        num_actions = 12
        episode_length_s = 20

    class trajectory_opt(RobotTrajGradSamplingCfg.trajectory_opt):
        enable_traj_opt = True
        num_diffuse_steps = 1
        num_diffuse_steps_init = 6
        num_samples = 127
        temp_sample = 0.1
        horizon_samples = 16
        horizon_nodes = 4
        horizon_diffuse_factor = 0.9
        traj_diffuse_factor = 0.5
        noise_scaling = 1.5
        update_method = "avwbfo"
        gamma = 1.00
        interp_method = "spline"
        compute_predictions = True

    class terrain(RobotTrajGradSamplingCfg.terrain):
        use_terrain_obj = False
        mesh_type = "plane"
        measure_heights = False
        curriculum = False
        terrain_proportions = [0.1, 0.1, 0.35, 0.25, 0.2]

    class commands(RobotTrajGradSamplingCfg.commands):
        curriculum = False
        max_curriculum = 1.0
        num_commands = 4
        resampling_time = 4.0
        heading_command = False

        class ranges(RobotTrajGradSamplingCfg.commands.ranges):
            lin_vel_x = [-1.5, 1.5]
            lin_vel_y = [-1.0, 1.0]
            ang_vel_yaw = [-1.0, 1.0]
            heading = [-3.14, 3.14]

    class init_state(RobotTrajGradSamplingCfg.init_state):
        pos = [0.0, 0.0, 0.5]
        rot = [0.0, 0.0, 0.0, 1.0]
        default_joint_angles = {
            'LF_HAA': 0.0, 'LF_HFE': 0.4, 'LF_KFE': -1.1,
            'RF_HAA': 0.0, 'RF_HFE': 0.4, 'RF_KFE': -1.1,
            'LH_HAA': 0.0, 'LH_HFE': -0.4, 'LH_KFE': 1.1,
            'RH_HAA': 0.0, 'RH_HFE': -0.4, 'RH_KFE': 1.1,
        }

    class control(RobotTrajGradSamplingCfg.control):
        control_type = 'P'
        jointpos_action_normalization = False
        stiffness = {'HAA': 80.0, 'HFE': 80.0, 'KFE': 80.0}
        damping = {'HAA': 2.0, 'HFE': 2.0, 'KFE': 2.0}
        action_scale = 0.5
        decimation = 4
        use_actuator_network = False
        actuator_net_file = "{LEGGED_GYM_ROOT_DIR}/resources/actuator_nets/anydrive_v3_lstm.pt"

    class asset(RobotTrajGradSamplingCfg.asset):
        file = "{LEGGED_GYM_ROOT_DIR}/resources/robots/anymal_c/urdf/anymal_c_boldshankcoll.urdf"
        name = "anymal_c"
        foot_name = "FOOT"
        penalize_contacts_on = ["SHANK", "THIGH", "base"]
        terminate_after_contacts_on = []
        self_collisions = 1

    class rewards(RobotTrajGradSamplingCfg.rewards):
        max_contact_force = 500.0
        base_height_target = 0.5
        only_positive_rewards = False
        multi_stage_rewards = False
        reward_stage_threshold = 6.0
        reward_min_stage = 0
        reward_max_stage = 1
        tracking_sigma = 0.25

        class scales:
            termination = -0.0
            tracking_lin_vel = 5.0
            tracking_ang_vel = 0.5
            lin_vel_z = -1.0
            ang_vel_xy = -0.5
            orientation = -2.0
            torques = -0.00001
            dof_vel = -0.0
            dof_acc = -2.5e-7
            feet_air_time = 1.0
            collision = -2
            feet_stumble = -0.0
            action_rate = -0.001
            stand_still = -0.0

    class gait_scheduler:
        period = 1.0
        duty = 0.5
        foot_phases = [0.0, 0.5, 0.0, 0.5]
        dt = 0.02
        swing_height = 0.1
        track_sigma = 0.25

    class domain_rand(RobotTrajGradSamplingCfg.domain_rand):
        randomize_base_mass = True
        added_mass_range = [-5.0, 5.0]


class AnymalCTrajGradSamplingCfgPPO(RobotTrajGradSamplingCfgPPO):
    class policy(RobotTrajGradSamplingCfgPPO.policy):
        actor_hidden_dims = [128, 64, 32]
        critic_hidden_dims = [128, 64, 32]
        activation = 'elu'

    class algorithm(RobotTrajGradSamplingCfgPPO.algorithm):
        entropy_coef = 0.01

    class runner(RobotTrajGradSamplingCfgPPO.runner):
        run_name = ''
        experiment_name = 'anymal_c_traj_grad_sampling'
        load_run = -1
        max_iterations = 3000
