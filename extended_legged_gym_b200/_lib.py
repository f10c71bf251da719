"""ctypes binding of ``include/elg_b200.h`` -- the thin layer between the Python host classes and
the hand-written sm_100a kernels.  No CPU fallback: if the shared library is missing it is built
(nvcc), and if that is impossible loading raises.
"""
import ctypes as C
import os
from typing import Optional

from . import build as _build

MAX_DOF, MAX_FEET, MAX_PENALISED, MAX_TERMINATION = 32, 8, 16, 8

REWARD_TERMS = [
    "action_rate", "ang_vel_xy", "base_foot_height", "base_height", "collision", "dof_acc", "dof_pos_limits", "dof_vel",
    "dof_vel_limits", "feet_air_time", "feet_contact_forces", "feet_slip", "feet_stumble", "feet_stumble_liftup",
    "four_footup", "gait_2_step", "gait_scheduler", "jump_air", "lin_vel_z", "orientation", "stand_still", "termination",
    "torque_limits", "torques", "tracking_ang_vel", "tracking_lin_vel"]
NUM_REWARD_TERMS = len(REWARD_TERMS)
TERM_ID = {n: i for i, n in enumerate(REWARD_TERMS)}
assert REWARD_TERMS == sorted(REWARD_TERMS)

CONTROL_TYPES = {"P": 0, "V": 1, "T": 2}
NOISE_OFF, NOISE_TENSOR, NOISE_PHILOX = 0, 1, 2
PHASE_DERIVE, PHASE_TERMINATION, PHASE_REWARD, PHASE_OBS, PHASE_HISTORY = 1, 2, 4, 8, 16
PHASE_PRE = PHASE_DERIVE | PHASE_TERMINATION | PHASE_REWARD
PHASE_POST = PHASE_OBS | PHASE_HISTORY
PHASE_FUSED = PHASE_PRE | PHASE_POST


class ElgDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("num_envs", "num_dof", "num_bodies", "num_feet", "num_penalised", "num_termination",
                                          "num_height_points", "num_obs", "num_commands")] + [
        ("feet_idx", C.c_int32 * MAX_FEET), ("penalised_idx", C.c_int32 * MAX_PENALISED),
        ("termination_idx", C.c_int32 * MAX_TERMINATION)]


class ElgStepParams(C.Structure):
    _fields_ = [
        ("dt", C.c_float), ("sim_dt", C.c_float), ("acc_ema", C.c_float), ("acc_ema_c", C.c_float),
        ("max_episode_length", C.c_int64),
        ("control_type", C.c_int32), ("action_scale", C.c_float),
        ("heading_command", C.c_int32), ("measure_heights", C.c_int32), ("terrain_is_plane", C.c_int32),
        ("only_positive_rewards", C.c_int32), ("noise_mode", C.c_int32), ("rollout_mode", C.c_int32), ("clip_observations", C.c_float),
        ("gravity_vec", C.c_float * 3),
        ("obs_scale_lin_vel", C.c_float), ("obs_scale_ang_vel", C.c_float), ("obs_scale_dof_pos", C.c_float),
        ("obs_scale_dof_vel", C.c_float), ("obs_scale_height", C.c_float),
        ("commands_scale", C.c_float * 3),
        ("border_size", C.c_float), ("horizontal_scale", C.c_float), ("vertical_scale", C.c_float),
        ("hf_rows", C.c_int32), ("hf_cols", C.c_int32), ("height_points_env_stride", C.c_int32),
        ("reward_mask", C.c_uint32), ("reward_scales", C.c_float * NUM_REWARD_TERMS),
        ("tracking_sigma", C.c_float), ("base_height_target", C.c_float), ("max_contact_force", C.c_float),
        ("soft_dof_vel_limit", C.c_float), ("soft_torque_limit", C.c_float), ("speed_min", C.c_float),
        ("stand_still_threshold", C.c_float),
        ("gait_increment", C.c_float), ("gait_swing_height", C.c_float), ("gait_foot_phases", C.c_float * MAX_FEET),
        ("gait_2_step_hexapod", C.c_int32), ("terminate_upside_down", C.c_int32),
        ("rows_per_main", C.c_int32), ("rollout_rew_stride", C.c_int32),
        ("height_obs_bound", C.c_float), ("reserved1", C.c_float),
        ("noise_seed", C.c_uint64), ("noise_offset", C.c_uint64)]


_BUF_FIELDS = [
    "root_states", "dof_state", "contact_forces", "rigid_body_state", "actions", "torques", "default_dof_pos",
    "dof_pos_limits", "dof_vel_limits", "torque_limits", "height_samples", "height_field_min", "height_points", "noise_scale_vec", "noise_u",
    "extra_reward",
    "last_actions", "last_dof_vel", "last_root_vel", "base_lin_acc", "base_ang_acc", "commands", "feet_air_time",
    "feet_contact_time", "last_contacts", "episode_length_buf", "episode_sums", "gait_idx", "gait_prev_foot_z",
    "base_lin_vel", "base_ang_vel", "projected_gravity", "foot_positions", "foot_velocities", "measured_heights",
    "reset_buf", "time_out_buf", "rew_buf", "obs_buf", "dof_consts", "step_counter", "rollout_rew_out"]


class ElgStepBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _BUF_FIELDS]


MAX_CLONE_FIELDS = 24
CLONE_SYNC, CLONE_CACHE, CLONE_RESTORE = 0, 1, 2
# actuator-network weight blob (ELG_ACTNET_* in include/elg_b200.h)
ACTNET_IN_SCALE, ACTNET_OUT_SCALE, ACTNET_W_IH0 = 0, 2, 4
ACTNET_W_HH0 = ACTNET_W_IH0 + 64
ACTNET_B_IH0 = ACTNET_W_HH0 + 256
ACTNET_B_HH0 = ACTNET_B_IH0 + 32
ACTNET_W_IH1 = ACTNET_B_HH0 + 32
ACTNET_W_HH1 = ACTNET_W_IH1 + 256
ACTNET_B_IH1 = ACTNET_W_HH1 + 256
ACTNET_B_HH1 = ACTNET_B_IH1 + 32
ACTNET_W_LIN = ACTNET_B_HH1 + 32
ACTNET_B_LIN = ACTNET_W_LIN + 8
ACTNET_WORDS = ACTNET_B_LIN + 4


class ElgNavParams(C.Structure):
    _fields_ = [("use_2d_nav", C.c_int32), ("use_prev", C.c_int32), ("num_commands", C.c_int32), ("zero_reached", C.c_int32),
                ("kp_linear", C.c_float), ("kp_angular", C.c_float), ("max_linear_vel", C.c_float), ("max_angular_vel", C.c_float),
                ("smooth", C.c_float), ("smooth_c", C.c_float), ("tolerance_rad", C.c_float)]


class ElgPlanParams(C.Structure):
    _fields_ = [("num_dof", C.c_int32), ("method", C.c_int32), ("n_substeps", C.c_int32), ("enforce_joint_limits", C.c_int32),
                ("sub_dt", C.c_float), ("max_base_lin_vel", C.c_float), ("max_base_ang_vel", C.c_float), ("max_joint_vel", C.c_float)]


class ElgPlanBuffers(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("integration_base_pos", "integration_base_quat", "integration_dof_pos", "integration_base_lin_vel",
                                          "integration_base_ang_vel", "integration_dof_vel", "dof_pos_limits", "root_states", "dof_state",
                                          "base_lin_vel", "base_ang_vel")]


class ElgCloneField(C.Structure):
    _fields_ = [("base", C.c_void_p), ("cache", C.c_void_p), ("row_bytes", C.c_int32), ("reserved", C.c_int32)]


class ElgCloneTable(C.Structure):
    _fields_ = [("num_fields", C.c_int32), ("num_main", C.c_int32), ("rollouts_per_main", C.c_int32), ("drift_field", C.c_int32),
                ("fields", ElgCloneField * MAX_CLONE_FIELDS)]


class ElgCamParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("out_width", C.c_int32), ("out_height", C.c_int32), ("buffer_len", C.c_int32),
                ("resize", C.c_int32), ("max_taps", C.c_int32), ("near_clip", C.c_float), ("far_clip", C.c_float), ("noise_scale", C.c_float)]


RESET_UNIFORMS = 48


class ElgResetParams(C.Structure):
    _fields_ = [("lin_vel_x", C.c_float * 2), ("lin_vel_y", C.c_float * 2), ("ang_vel_yaw", C.c_float * 2), ("heading", C.c_float * 2),
                ("heading_command", C.c_int32), ("resample_interval", C.c_int32), ("base_init_state", C.c_float * 13),
                ("custom_origins", C.c_int32), ("curriculum", C.c_int32), ("env_length_half", C.c_float), ("max_episode_length_s", C.c_float),
                ("max_terrain_level", C.c_int32), ("terrain_cols", C.c_int32), ("seed", C.c_uint64), ("offset", C.c_uint64),
                ("rows_per_main", C.c_int32), ("root_z_from_terrain", C.c_int32)]


_RESET_FIELDS = ["reset_buf", "root_states", "dof_state", "commands", "env_origins", "terrain_levels", "terrain_types", "terrain_origins",
                 "default_dof_pos", "last_dof_vel", "last_root_vel", "feet_air_time", "feet_contact_time", "episode_length_buf", "episode_sums",
                 "stats", "stats_accum", "obs_buf", "measured_heights", "noise_scale_vec", "noise_u", "uniforms", "height_samples"]


class ElgResetBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _RESET_FIELDS]


class ElgError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None


def lib_path() -> str:
    return _build.LIB_PATH


def load() -> C.CDLL:
    """Load (building first if needed) the CUDA library; raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    # build() returns at once when the source fingerprint matches the stamp next to the library, so a stale .so is never
    # loaded after a csrc edit
    path = _build.build(force=os.environ.get("ELG_REBUILD") == "1")
    if os.environ.get("ELG_LIB_PATH"):       # diagnostics only (A/B runs against another build of the same ABI, scripts/)
        path = os.environ["ELG_LIB_PATH"]
    lib = C.CDLL(path)
    lib.elg_last_error.restype = C.c_char_p
    lib.elg_reward_term_name.restype = C.c_char_p
    lib.elg_reward_term_name.argtypes = [C.c_int]
    for fn, st in (("elg_sizeof_dims", ElgDims), ("elg_sizeof_step_params", ElgStepParams), ("elg_sizeof_step_buffers", ElgStepBuffers),
                   ("elg_sizeof_clone_table", ElgCloneTable), ("elg_sizeof_cam_params", ElgCamParams), ("elg_sizeof_reset_params", ElgResetParams),
                   ("elg_sizeof_reset_buffers", ElgResetBuffers), ("elg_sizeof_nav_params", ElgNavParams),
                   ("elg_sizeof_plan_params", ElgPlanParams), ("elg_sizeof_plan_buffers", ElgPlanBuffers)):
        got = getattr(lib, fn)()
        if got != C.sizeof(st) and not os.environ.get("ELG_LIB_PATH"):
            raise ElgError(f"ABI mismatch: {fn}() = {got}, python mirror = {C.sizeof(st)}")
    for i, name in enumerate(REWARD_TERMS):
        if lib.elg_reward_term_name(i).decode() != name:
            raise ElgError(f"reward registry mismatch at id {i}")
    vp, i64 = C.c_void_p, C.c_int64
    lib.elg_compute_torques.argtypes = [C.POINTER(ElgDims), C.POINTER(ElgStepParams)] + [vp] * 9 + [i64, vp]
    if hasattr(lib, "elg_rollout_actions"):
        lib.elg_rollout_actions.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, i64, C.c_float, vp, vp, vp, vp]
    lib.elg_post_physics_step.argtypes = [C.POINTER(ElgDims), C.POINTER(ElgStepParams), C.POINTER(ElgStepBuffers), C.c_uint32, vp]
    lib.elg_set_step_tuning.argtypes = [C.c_int] * 4
    lib.elg_get_heights.argtypes = [C.POINTER(ElgDims), C.POINTER(ElgStepParams)] + [vp] * 5 + [vp]
    lib.elg_clone_rows.argtypes = [C.POINTER(ElgCloneTable), C.c_int, C.c_float, vp, C.c_uint64, C.c_uint64, vp]
    lib.elg_set_step_debug.argtypes = [vp]
    if hasattr(lib, "elg_probe_empty"):
        lib.elg_probe_empty.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vp]
        lib.elg_probe_roundtrip.argtypes = [vp, vp, i64, i64, C.c_int, C.c_int, vp]
    lib.elg_stage_block.argtypes = [vp, vp, i64, C.c_int, C.c_int, vp]
    lib.elg_set_clone_tuning.argtypes = [C.c_int]
    lib.elg_nav_commands.argtypes = [C.c_int32, C.c_int32, C.POINTER(ElgNavParams)] + [vp] * 7
    lib.elg_integrate_state_velocities.argtypes = [C.POINTER(ElgPlanParams), C.POINTER(ElgPlanBuffers), vp, vp, i64, vp]
    lib.elg_normalizer_scratch_bytes.restype = i64
    lib.elg_normalizer_scratch_bytes.argtypes = [i64, C.c_int32]
    lib.elg_set_normalizer_tuning.argtypes = [C.c_int]
    lib.elg_normalize_observations.argtypes = [i64, C.c_int32] + [vp] * 5 + [C.c_float, i64, C.c_int32] + [vp] * 7
    if lib.elg_actuator_net_words() != ACTNET_WORDS:
        raise ElgError(f"ABI mismatch: elg_actuator_net_words() = {lib.elg_actuator_net_words()}, python mirror = {ACTNET_WORDS}")
    if hasattr(lib, "elg_set_actuator_tuning"):
        lib.elg_set_actuator_tuning.argtypes = [C.c_int]
    lib.elg_actuator_net_torques.argtypes = [C.POINTER(ElgDims), vp, C.c_float] + [vp] * 7
    lib.elg_actuator_net_bind.argtypes = [vp, vp]
    lib.elg_mesh_create.argtypes = [vp, C.c_int32, vp, C.c_int32, C.POINTER(vp)]
    lib.elg_mesh_create_ex.argtypes = [vp, C.c_int32, vp, C.c_int32, C.c_int32, C.POINTER(vp)]
    lib.elg_mesh_free.argtypes = [vp]
    if hasattr(lib, "elg_mesh_grid_info"):
        lib.elg_mesh_grid_info.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        lib.elg_set_mesh_tuning.argtypes = [C.c_int]
    lib.elg_mesh_info.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_float)]
    lib.elg_raycast.argtypes = [vp, vp, vp, i64, C.c_float, vp, vp, vp, vp, vp]
    lib.elg_raycast_sensor.argtypes = [vp, vp, vp, C.c_int32, vp, vp, vp, i64, C.c_int, C.c_float, vp, vp, vp]
    if hasattr(lib, "elg_raycast_sensor_obs"):
        lib.elg_raycast_sensor_obs.argtypes = [vp, vp, vp, C.c_int32, vp, vp, vp, i64, C.c_int, C.c_float, vp, vp, vp, C.c_int32, C.c_int32, vp, i64, vp]
    lib.elg_camera_pose.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp, vp]
    lib.elg_depth_camera.argtypes = [vp, C.POINTER(ElgCamParams)] + [vp] * 9 + [i64, vp, vp, vp]
    lib.elg_sdf_query.argtypes = [vp, vp, i64, C.c_float, C.c_float, vp, vp, vp, vp, vp]
    if hasattr(lib, "elg_sdf_query_bodies"):
        lib.elg_sdf_query_bodies.argtypes = [vp, vp, C.c_int32, vp, vp, C.c_int32, vp, i64, C.c_float, C.c_float, vp, i64, vp, vp, vp, vp]
    lib.elg_mesh_mean_edge.argtypes = [vp]
    lib.elg_mesh_mean_edge.restype = C.c_double
    lib.elg_mppi_costs.argtypes = [vp, i64, i64, C.c_int32, vp, vp]
    lib.elg_mppi_partials.argtypes = [vp, i64, C.c_int32, C.c_int32, C.c_int32, vp, C.c_int32, C.c_float, vp, vp]
    lib.elg_mppi_finish.argtypes = [vp, i64, C.c_int32, vp, vp]
    if hasattr(lib, "elg_mppi_update"):
        lib.elg_mppi_partials_ranked.argtypes = [vp, i64, C.c_int32, C.c_int32, C.c_int32, vp, C.c_int32, C.c_float, vp, vp]
        lib.elg_comm_unique_id.argtypes = [vp]
        lib.elg_comm_init.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp)]
        lib.elg_comm_destroy.argtypes = [vp]
        lib.elg_comm_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.elg_episode_stats_allreduce.argtypes = [vp, C.c_int32, vp, vp]
        lib.elg_comm_warmup.argtypes = [vp, vp, C.c_int32, vp]
        lib.elg_mppi_update.argtypes = [vp, vp, i64, C.c_int32, C.c_int32, C.c_int32, C.c_float, vp, vp, vp, vp, vp]
    lib.elg_resample_commands.argtypes = [C.POINTER(ElgDims), C.POINTER(ElgResetParams), vp, vp, vp, vp, vp]
    lib.elg_reset_envs.argtypes = [C.POINTER(ElgDims), C.POINTER(ElgResetParams), C.POINTER(ElgStepParams), C.POINTER(ElgResetBuffers), vp]
    lib.elg_prepare_height_field.argtypes = [vp, C.c_int32, C.c_int32, C.c_float, vp, vp]
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().elg_last_error().decode()
        if rc == -1 and "controller type" in msg:
            raise NameError(msg)          # same exception type as the reference (legged_robot.py:447)
        raise ElgError(f"{what or 'elg call'} failed ({rc}): {msg}")


def ptr(t) -> Optional[int]:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
