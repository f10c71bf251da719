"""Config of ``Go2TrajGradSampling`` (values of envs/go2/batch_rollout/go2_traj_grad_sampling_config.py:36-300 in
/root/reference/legged_gym/legged_gym; compared with the reference's class by tests/test_robot_rollout_classes.py).  Its reward set IS the
DIAL-MPC set (``gaits, upright, yaw, vel, ang_vel, height``): subclass terms evaluated with torch next to the (empty) kernel registry."""
from ...batch_rollout.robot_traj_grad_sampling_config import RobotTrajGradSamplingCfg, RobotTrajGradSamplingCfgPPO


class Go2TrajGradSamplingCfg(RobotTrajGradSamplingCfg):
    class env(RobotTrajGradSamplingCfg.env):
        num_envs = 1
        rollout_envs = 256
        env_spacing = 2.0
        num_observations = 48
        num_actions = 12
        episode_length_s = 20

    class trajectory_opt(RobotTrajGradSamplingCfg.trajectory_opt):
        enable_traj_opt = True
        num_diffuse_steps = 2
        num_diffuse_steps_init = 6
        num_samples = 255
        temp_sample = 0.05
        horizon_samples = 16
        horizon_nodes = 4
        horizon_diffuse_factor = 0.9
        traj_diffuse_factor = 0.5
        noise_scaling = 0.8
        update_method = "mppi"
        gamma = 1.00
        interp_method = "spline"
        compute_predictions = False

    class terrain(RobotTrajGradSamplingCfg.terrain):
        use_terrain_obj = False
        mesh_type = "plane"
        measure_heights = False
        curriculum = False

    class commands(RobotTrajGradSamplingCfg.commands):
        curriculum = False
        max_curriculum = 1.0
        num_commands = 4
        resampling_time = 4.0
        heading_command = False

        class ranges(RobotTrajGradSamplingCfg.commands.ranges):
            lin_vel_x = [-0.8, 0.8]
            lin_vel_y = [-0.6, 0.6]
            ang_vel_yaw = [-0.8, 0.8]
            heading = [-3.14, 3.14]

    class init_state(RobotTrajGradSamplingCfg.init_state):
        pos = [0.0, 0.0, 0.27]
        rot = [0.0, 0.0, 0.0, 1.0]
        default_joint_angles = {
            'FL_hip_joint': 0.0, 'FL_thigh_joint': 0.9, 'FL_calf_joint': -1.8,
            'FR_hip_joint': 0.0, 'FR_thigh_joint': 0.9, 'FR_calf_joint': -1.8,
            'RL_hip_joint': 0.0, 'RL_thigh_joint': 0.9, 'RL_calf_joint': -1.8,
            'RR_hip_joint': 0.0, 'RR_thigh_joint': 0.9, 'RR_calf_joint': -1.8,
        }

    class control(RobotTrajGradSamplingCfg.control):
        control_type = 'P'
        jointpos_action_normalization = False
        stiffness = {'joint': 40.0}
        damping = {'joint': 1.0}
        action_scale = 1.0
        decimation = 4
        use_actuator_network = False
        actuator_net_file = "{LEGGED_GYM_ROOT_DIR}/resources/actuator_nets/go2_actuator_net.pt"

    class asset(RobotTrajGradSamplingCfg.asset):
        file = "{LEGGED_GYM_ROOT_DIR}/resources/robots/go2/urdf/go2_description.urdf"
        name = "go2"
        foot_name = "foot"
        penalize_contacts_on = ["thigh", "calf"]
        terminate_after_contacts_on = ["base"]
        self_collisions = 1

    class rewards(RobotTrajGradSamplingCfg.rewards):
        max_contact_force = 350.0
        base_height_target = 0.30
        only_positive_rewards = False
        multi_stage_rewards = False
        reward_stage_threshold = 5.0
        reward_min_stage = 0
        reward_max_stage = 1
        tracking_sigma = 0.25

        class scales:
            termination = -0.0
            gaits = 0.1
            air_time = 0.0
            pos = 0.0
            upright = 0.5
            yaw = 0.3
            vel = 1.0
            ang_vel = 0.3
            height = 10.0
            energy = 0.0
            alive = 0.0

    class gait_scheduler:
        period = 1.0
        duty = 0.5
        foot_phases = [0.0, 0.5, 0.0, 0.5]
        dt = 0.02
        swing_height = 0.1
        track_sigma = 0.25

    class domain_rand(RobotTrajGradSamplingCfg.domain_rand):
        randomize_base_mass = False
        added_mass_range = [-1.0, 1.0]
        randomize_friction = False
        randomize_restitution = False


class Go2TrajGradSamplingCfgPPO(RobotTrajGradSamplingCfgPPO):
    class policy(RobotTrajGradSamplingCfgPPO.policy):
        actor_hidden_dims = [128, 64, 32]
        critic_hidden_dims = [128, 64, 32]
        activation = 'elu'

    class algorithm(RobotTrajGradSamplingCfgPPO.algorithm):
        entropy_coef = 0.01
        learning_rate = 3e-4
        num_learning_epochs = 8
        mini_batch_size = 4096

    class runner(RobotTrajGradSamplingCfgPPO.runner):
        run_name = ''
        experiment_name = 'go2_traj_grad_sampling'
        load_run = -1
        max_iterations = 2000
        multi_stage_rewards = False
        save_interval = 100
