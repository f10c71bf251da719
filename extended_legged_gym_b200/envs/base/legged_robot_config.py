"""``LeggedRobotCfg`` / ``LeggedRobotCfgPPO`` -- the configuration API of the hot path.

Attribute names, nesting and default values are the reference's (legged_gym/legged_gym/envs/base/legged_robot_config.py:34-316):
user configs subclass these nested classes by name, so the drop-in keeps that contract, and
``tests/test_oracle_pinned.py::test_configs_match_reference`` compares the resulting dictionaries in the build container.
The sections are ordered by who reads them: the per-step kernels first, then the perception utilities, then what is only
carried for the simulator and the caller.
"""
from .base_config import BaseConfig


def _decimetres(lo: int, hi: int):
    """[lo/10, ..., hi/10]: the height-scan grid axes, e.g. (-8, 8) -> -0.8 ... 0.8"""
    return [k / 10.0 for k in range(lo, hi + 1)]


class LeggedRobotCfg(BaseConfig):
    # ------------------------------------------------------------------ read by the step / reset kernels
    class sim:
        dt = 0.005                     # physics step; the control step is decimation x this
        substeps = 1
        up_axis = 1                    # 0: y, 1: z
        gravity = [0.0, 0.0, -9.81]

        class physx:                   # solver settings, handed to the simulator unchanged
            solver_type, num_threads = 1, 10
            num_position_iterations, num_velocity_iterations = 4, 0
            contact_offset, rest_offset = 0.01, 0.0
            bounce_threshold_velocity, max_depenetration_velocity = 0.5, 1.0
            max_gpu_contact_pairs = 2 ** 23
            default_buffer_size_multiplier = 5
            contact_collection = 2

    class control:
        decimation = 4
        action_scale = 0.5
        control_type = "P"             # P: position targets, V: velocity targets, T: torques
        damping = {"joint_a": 1.0, "joint_b": 1.5}
        stiffness = {"joint_a": 10.0, "joint_b": 15.0}

    class commands:
        num_commands = 4               # lin_vel_x, lin_vel_y, ang_vel_yaw, heading
        heading_command = False        # True: the yaw rate follows the heading error
        resampling_time = 10.0
        curriculum, max_curriculum = False, 1.0

        class ranges:
            heading = [-3.14, 3.14]
            ang_vel_yaw = [-1, 1]
            lin_vel_y = [-1.0, 1.0]
            lin_vel_x = [-1.0, 1.0]

    class normalization:
        clip_actions = clip_observations = 100.0

        class obs_scales:
            height_measurements = 5.0
            dof_pos, dof_vel = 1.0, 0.05
            lin_vel, ang_vel = 2.0, 0.25

    class noise:
        noise_level = 1.0
        add_noise = True

        class noise_scales:
            height_measurements = 0.1
            gravity = 0.05
            lin_vel, ang_vel = 0.1, 0.2
            dof_pos, dof_vel = 0.01, 1.5

    class rewards:
        tracking_sigma = 0.25
        base_height_target = 1.0
        max_contact_force = 100.0
        soft_dof_pos_limit = soft_dof_vel_limit = soft_torque_limit = 1.0
        only_positive_rewards = True
        # reward stages: list-valued scales hold one value per stage
        multi_stage_rewards = False
        reward_min_stage = reward_max_stage = 0
        reward_stage_threshold = 6.0

        class scales:
            tracking_lin_vel, tracking_ang_vel = 1.0, 0.5
            lin_vel_z, ang_vel_xy = -2.0, -0.05
            torques, dof_acc, action_rate = -0.00001, -2.5e-7, -0.01
            feet_air_time, collision = 1.0, -1.0
            # registered, weight zero
            termination = orientation = dof_vel = base_height = feet_stumble = stand_still = -0.0

    class terrain:
        mesh_type = "trimesh"          # none | plane | heightfield | trimesh | confined_trimesh
        horizontal_scale, vertical_scale = 0.1, 0.005      # [m] per height-sample cell, [m] per int16 unit
        border_size = 25               # [m]
        measure_heights = True
        measured_points_x = _decimetres(-8, 8)             # 17 columns
        measured_points_y = _decimetres(-5, 5)             # 11 rows -> 187 scan points
        curriculum = True
        max_init_terrain_level = 5
        num_rows = num_cols = 8
        terrain_length = terrain_width = 5.0
        terrain_proportions = [0.1, 0.1, 0.35, 0.25, 0.2]
        confined_terrain_proportions = [0.25, 0.5, 0.75, 1.0]
        slope_treshold = 0.75
        static_friction = dynamic_friction = 1.0
        restitution = 0.0
        selected, terrain_kwargs = False, None
        use_terrain_obj, terrain_file = False, None

    class env:
        num_envs = 4096
        num_actions = 12
        num_observations = 235
        num_privileged_obs = None      # None -> step() returns None and the runner reuses obs
        episode_length_s = 20
        send_timeouts = True
        env_spacing = 3.0

    class init_state:
        default_joint_angles = {"joint_a": 0.0, "joint_b": 0.0}
        pos, rot = [0.0, 0.0, 1.0], [0.0, 0.0, 0.0, 1.0]   # rot: xyzw
        lin_vel, ang_vel = [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]

    class domain_rand:
        push_robots, push_interval_s, max_push_vel_xy = True, 15, 1.0
        randomize_friction, friction_range = True, [0.5, 1.25]
        randomize_base_mass, added_mass_range = False, [-1.0, 1.0]

    # ------------------------------------------------------------------ read by the perception utilities
    class raycaster:
        enable_raycast = False
        ray_pattern = "cone"           # single | grid | cone | spherical | spherical2
        num_rays, ray_angle, max_distance = 32, 60, 10.0
        spherical_num_azimuth, spherical_num_elevation = 8, 4
        spherical2_num_points, spherical2_polar_axis = 32, [0.0, 0.0, 1.0]
        offset_pos, attach_yaw_only = [0.5, 0.0, 0.0], False
        terrain_file = None

    class depth:
        camera_type = "Warp"           # None | "IsaacGym" | "Warp" | "Fake"  ("Warp" -> the B200 ray caster here)
        original, resized = (60, 30), (56, 28)
        horizontal_fov = 100
        near_clip, far_clip = 0, 2
        buffer_len, update_interval = 2, 1
        position, angle = [0.5, 0, 0.03], [30, 30]
        dis_noise, scale, invert = 0.0, 1, True

    # ------------------------------------------------------------------ carried for the simulator
    class asset:
        name, file = "legged_robot", ""
        foot_name = "None"
        penalize_contacts_on, terminate_after_contacts_on = [], []
        self_collisions = 0
        fix_base_link = disable_gravity = False
        collapse_fixed_joints = replace_cylinder_with_capsule = flip_visual_attachments = True
        default_dof_drive_mode = 3
        density, armature, thickness = 0.001, 0.0, 0.01
        angular_damping = linear_damping = 0.0
        max_angular_velocity = max_linear_velocity = 1000.0

    class obstacle_gen:
        enable_obstacles = False
        min_obstacles, max_obstacles = 5, 15
        cluster_probability = 0.3
        spawn_height_range, spawn_radius_range = [0.3, 1.0], [1.5, 6.0]
        stone_density_range = [800, 2000]
        stone_friction_range, stone_restitution_range = [0.3, 0.9], [0.1, 0.4]

    class viewer:
        ref_env = 0
        pos, lookat = [10, 0, 6], [11.0, 5, 3.0]


class LeggedRobotCfgPPO(BaseConfig):
    """Carried for the caller (rsl_rl ``OnPolicyRunner``); nothing on the hot path reads it."""
    runner_class_name = "OnPolicyRunner"
    seed = 1

    class runner:
        policy_class_name, algorithm_class_name = "ActorCritic", "PPO"
        num_steps_per_env, max_iterations, save_interval = 24, 1500, 50
        experiment_name, run_name = "test", ""
        resume, resume_path = False, None
        load_run = checkpoint = -1
        multi_stage_rewards = False

    class algorithm:
        learning_rate, schedule, desired_kl = 1.0e-3, "adaptive", 0.01
        gamma, lam = 0.99, 0.95
        clip_param, entropy_coef = 0.2, 0.01
        value_loss_coef, use_clipped_value_loss = 1.0, True
        num_learning_epochs, num_mini_batches = 5, 4
        max_grad_norm = 1.0

    class policy:
        activation = "elu"
        init_noise_std = 1.0
        actor_hidden_dims = [512, 256, 128]
        critic_hidden_dims = [512, 256, 128]
