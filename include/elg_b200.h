/*
 * elg_b200.h -- C ABI of the B200-native per-step hot path of extended_legged_gym.
 *
 * The reference (MasterYip/extended_legged_gym) is pure Python and has no FFI layer of its
 * own: its boundary is the Python class API.  This header is the
 * boundary a maintainer binds from Python with ctypes: every entry
 * point replaces one group of reference methods, cited below as file:line relative to
 * legged_gym/legged_gym/ in the reference tree.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types cross the boundary.
 *   - every pointer is a DEVICE pointer borrowed from the caller (a torch tensor or a
 *     PhysX-owned tensor); the library never allocates, frees or synchronises, except the
 *     opaque mesh handle (elg_mesh_*) which owns its BVH.
 *   - `stream` is a cudaStream_t passed as void*; kernels are enqueued on it and are CUDA
 *     graph capturable (no host reads).
 *   - return value: 0 on success, negative ElgStatus on error; elg_last_error() describes it.
 *   - dtypes: fp32; int64 episode_length_buf; int16 height_samples; uint8 for torch.bool.
 *   - quaternions are xyzw (Isaac Gym convention).
 */
#ifndef ELG_B200_H
#define ELG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ELG_ABI_VERSION 4

#define ELG_MAX_DOF 32
#define ELG_MAX_FEET 8
#define ELG_MAX_PENALISED 16
#define ELG_MAX_TERMINATION 8

typedef enum ElgStatus {
  ELG_OK = 0,
  ELG_ERR_INVALID_ARGUMENT = -1,
  ELG_ERR_UNSUPPORTED = -2,
  ELG_ERR_CUDA = -3,
  ELG_ERR_NULL_POINTER = -4
} ElgStatus;

/* Reward registry: ids are the ALPHABETICAL order of the reference's `_reward_*` names, which is
 * the order `class_to_dict(cfg.rewards.scales)` yields and therefore the fp32 summation order of
 * compute_reward (utils/helpers.py:43-58, envs/base/legged_robot.py:215-232,
 * envs/base/legged_robot_rew_mixin.py:41-234, envs/anymal_c/anymal.py:112-114). */
typedef enum ElgRewardTerm {
  ELG_REW_ACTION_RATE = 0,
  ELG_REW_ANG_VEL_XY,
  ELG_REW_BASE_FOOT_HEIGHT,
  ELG_REW_BASE_HEIGHT,
  ELG_REW_COLLISION,
  ELG_REW_DOF_ACC,
  ELG_REW_DOF_POS_LIMITS,
  ELG_REW_DOF_VEL,
  ELG_REW_DOF_VEL_LIMITS,
  ELG_REW_FEET_AIR_TIME,
  ELG_REW_FEET_CONTACT_FORCES,
  ELG_REW_FEET_SLIP,
  ELG_REW_FEET_STUMBLE,
  ELG_REW_FEET_STUMBLE_LIFTUP,
  ELG_REW_FOUR_FOOTUP,
  ELG_REW_GAIT_2_STEP,
  ELG_REW_GAIT_SCHEDULER,
  ELG_REW_JUMP_AIR,
  ELG_REW_LIN_VEL_Z,
  ELG_REW_ORIENTATION,
  ELG_REW_STAND_STILL,
  ELG_REW_TERMINATION, /* added after the only-positive clip (legged_robot.py:228-232) */
  ELG_REW_TORQUE_LIMITS,
  ELG_REW_TORQUES,
  ELG_REW_TRACKING_ANG_VEL,
  ELG_REW_TRACKING_LIN_VEL,
  ELG_NUM_REWARD_TERMS
} ElgRewardTerm;

/* returns the reference name of a term ("action_rate", ...) or NULL */
const char* elg_reward_term_name(int term);

typedef enum ElgControlType { ELG_CONTROL_P = 0, ELG_CONTROL_V = 1, ELG_CONTROL_T = 2 } ElgControlType;

typedef enum ElgNoiseMode {
  ELG_NOISE_OFF = 0,    /* cfg.noise.add_noise == False */
  ELG_NOISE_TENSOR = 1, /* uniform [0,1) samples supplied by the caller (parity mode, = torch.rand_like) */
  ELG_NOISE_PHILOX = 2  /* generated in-kernel: Philox4x32-10 keyed by seed, counter (env, lane | block << 5, step); 16-bit samples */
} ElgNoiseMode;

/* Sections of post_physics_step (legged_robot.py:113-150), OR-ed into `phase`.  The reset path
 * (reset_idx, host side) sits between ELG_PHASE_PRE and ELG_PHASE_POST. */
#define ELG_PHASE_DERIVE 1u       /* episode counter, base-frame state, feet gather, heading cmd, heights (:122-139) */
#define ELG_PHASE_TERMINATION 2u  /* check_termination (:155-160) */
#define ELG_PHASE_REWARD 4u       /* compute_reward + registry (:215-232) */
#define ELG_PHASE_OBS 8u          /* compute_observations (+noise, +clip) (:234-252) */
#define ELG_PHASE_HISTORY 16u     /* last_actions / last_dof_vel / last_root_vel (:148-150) */
#define ELG_PHASE_PRE (ELG_PHASE_DERIVE | ELG_PHASE_TERMINATION | ELG_PHASE_REWARD)
#define ELG_PHASE_POST (ELG_PHASE_OBS | ELG_PHASE_HISTORY)
#define ELG_PHASE_FUSED (ELG_PHASE_PRE | ELG_PHASE_POST)

typedef struct ElgDims {
  int32_t num_envs;          /* N */
  int32_t num_dof;           /* D  (== num_actions) */
  int32_t num_bodies;        /* B */
  int32_t num_feet;          /* F */
  int32_t num_penalised;     /* P */
  int32_t num_termination;   /* T */
  int32_t num_height_points; /* H (0 when cfg.terrain.measure_heights is False) */
  int32_t num_obs;           /* O */
  int32_t num_commands;      /* C */
  int32_t feet_idx[ELG_MAX_FEET];
  int32_t penalised_idx[ELG_MAX_PENALISED];
  int32_t termination_idx[ELG_MAX_TERMINATION];
} ElgDims;

/* cfg scalars baked by the host (legged_robot.py:_parse_cfg :847-860, _prepare_reward_function :649-674) */
typedef struct ElgStepParams {
  float dt;                  /* control.decimation * sim.dt */
  float sim_dt;
  float acc_ema;             /* 0.9 (legged_robot.py:85) */
  float acc_ema_c;           /* fp32(1 - acc_ema), evaluated in double like the Python expression */
  int64_t max_episode_length; /* floor(np.ceil(episode_length_s / dt)); time_out = ep_len > this */
  int32_t control_type;      /* ElgControlType */
  float action_scale;
  int32_t heading_command;
  int32_t measure_heights;
  int32_t terrain_is_plane;  /* mesh_type == 'plane' -> heights are zeros (legged_robot.py:913-914) */
  int32_t only_positive_rewards;
  int32_t noise_mode;        /* ElgNoiseMode */
  int32_t rollout_mode;      /* post_physics_step_rollout (batch_rollout/robot_batch_rollout.py:763-817): with ELG_PHASE_DERIVE no
                                episode counter, heading command or height scan (measured_heights is an input); rewards do not
                                accumulate episode sums (compute_reward_rollout :969-985) */
  float clip_observations;   /* <= 0: no clip; > 0: step()'s clip fused (legged_robot.py:107-108) */
  float gravity_vec[3];      /* normalised gravity, (0,0,-1) */
  float obs_scale_lin_vel, obs_scale_ang_vel, obs_scale_dof_pos, obs_scale_dof_vel, obs_scale_height;
  float commands_scale[3];
  /* terrain (legged_robot.py:925-938) */
  float border_size, horizontal_scale, vertical_scale;
  int32_t hf_rows, hf_cols;
  int32_t height_points_env_stride; /* 0: one [H,3] grid shared by all envs; else elements between envs */
  /* rewards */
  uint32_t reward_mask;                    /* bit t set <=> term t has a non-zero scale */
  float reward_scales[ELG_NUM_REWARD_TERMS]; /* fp32(scale * dt) */
  float tracking_sigma, base_height_target, max_contact_force;
  float soft_dof_vel_limit, soft_torque_limit, speed_min;
  float stand_still_threshold;             /* speed_min in the base mixin (:221) */
  /* gait scheduler (utils/gait_scheduler.py:63-81, anymal.py:60-66) */
  float gait_increment;      /* fp32(dt / period) */
  float gait_swing_height;
  float gait_foot_phases[ELG_MAX_FEET];
  /* hexapod class ElSpider (envs/elspider_air/elspider.py): */
  int32_t gait_2_step_hexapod;     /* gait_2_step over the tripods (0,1,5) / (2,3,4) of six feet (:365-408) instead of the quadruped pairs */
  int32_t terminate_upside_down;   /* reset |= projected_gravity.z > 0: 1 = every row (elspider.py:340-345), 2 = main rows of the main / rollout layout only (anymal_c_batch_rollout.py:192-199) */
  /* main / rollout env layout (envs/batch_rollout/robot_batch_rollout.py:119-164): 0 = flat env list; R1 = 1 + rollouts per main:
     row r is a main env iff r % R1 == 0.  check_termination (:857-866) ORs time-outs into the main rows only. */
  int32_t rows_per_main;
  /* rollout mode only: elements between the reward slots of consecutive rollout envs in ElgStepBuffers.rollout_rew_out (= horizon) */
  int32_t rollout_rew_stride;
  /* upper bound of |height observation| = obs_scale_height + max height-entry noise scale, or 0 when unknown: lets the kernel drop
     step()'s clip on those entries when it cannot change a value */
  float height_obs_bound;
  float reserved1;
  /* in-kernel noise */
  uint64_t noise_seed;
  uint64_t noise_offset;     /* step counter, so successive steps draw fresh numbers */
} ElgStepParams;

typedef struct ElgStepBuffers {
  /* ---- PhysX-owned state, read only (legged_robot.py:575-584) ---- */
  const float* root_states;      /* [N,13] */
  const float* dof_state;        /* [N*D,2] */
  const float* contact_forces;   /* [N*B,3] */
  const float* rigid_body_state; /* [N*B,13] */
  /* ---- per-step inputs ---- */
  const float* actions;          /* [N,D] */
  const float* torques;          /* [N,D] (output of elg_compute_torques) */
  const float* default_dof_pos;  /* [D] */
  const float* dof_pos_limits;   /* [D,2] */
  const float* dof_vel_limits;   /* [D] */
  const float* torque_limits;    /* [D] */
  const int16_t* height_samples; /* [rows, cols] or NULL */
  const float* height_field_min; /* [rows, cols] fp32 from elg_prepare_height_field, or NULL (kernel gathers the 3 int16 cells) */
  const float* height_points;    /* [H,3] (or [N,H,3] with height_points_env_stride) or NULL */
  const float* noise_scale_vec;  /* [O] or NULL */
  const float* noise_u;          /* [N,O] uniform samples for ELG_NOISE_TENSOR, else NULL */
  const float* extra_reward;     /* [N] pre-scaled sum of user-defined Python terms, added before the clip; or NULL */
  /* ---- env-owned state, read + written ---- */
  float* last_actions;           /* [N,D] */
  float* last_dof_vel;           /* [N,D] */
  float* last_root_vel;          /* [N,6] */
  float* base_lin_acc;           /* [N,3] */
  float* base_ang_acc;           /* [N,3] */
  float* commands;               /* [N,C] */
  float* feet_air_time;          /* [N,F] */
  float* feet_contact_time;      /* [N,F] */
  uint8_t* last_contacts;        /* [N,F] bool */
  int64_t* episode_length_buf;   /* [N] */
  float* episode_sums;           /* [ELG_NUM_REWARD_TERMS, N]; row t only touched if term t enabled */
  float* gait_idx;               /* [N] or NULL */
  float* gait_prev_foot_z;       /* [N,F] foot z handed to GaitScheduler.step last step, or NULL */
  /* ---- outputs ---- */
  float* base_lin_vel;           /* [N,3] */
  float* base_ang_vel;           /* [N,3] */
  float* projected_gravity;      /* [N,3] */
  float* foot_positions;         /* [N,F,3] */
  float* foot_velocities;        /* [N,F,3] */
  float* measured_heights;       /* [N,H] or NULL */
  uint8_t* reset_buf;            /* [N] bool */
  uint8_t* time_out_buf;         /* [N] bool */
  float* rew_buf;                /* [N] */
  float* obs_buf;                /* [N,O] */
  /* ---- optional extras ---- */
  const float* dof_consts;       /* optional [5 D]: default_dof_pos | dof_pos_limits (2 D) | dof_vel_limits | torque_limits back to back
                                    (the same values as the four arrays above): one bulk copy per CTA instead of four */
  const uint64_t* step_counter;  /* device word added to noise_offset, or NULL: lets a CUDA graph that is replayed many times draw fresh
                                    in-kernel noise (the host-side noise_offset is frozen into a captured launch) */
  float* rollout_rew_out;        /* rollout mode with rows_per_main > 0, or NULL: the reward of rollout env (main k, rollout r) is ALSO
                                    written to rollout_rew_out[(k * (rows_per_main - 1) + r) * rollout_rew_stride] -- column i of the
                                    [num_rollout_envs, horizon] reward table of rollout_batch (robot_traj_grad_sampling.py:262-266) */
} ElgStepBuffers;

/* ABI self-description so the Python mirror structs can be checked at load time */
int elg_abi_version(void);
int elg_sizeof_dims(void);
int elg_sizeof_step_params(void);
int elg_sizeof_step_buffers(void);
const char* elg_last_error(void);

/* LeggedRobot._compute_torques (envs/base/legged_robot.py:425-448; rollout twin
 * envs/batch_rollout/robot_batch_rollout.py:1016-1037).  dof_state is the interleaved [N*D,2]
 * PhysX tensor (dof_pos/dof_vel are its stride-2 views).  env_ids == NULL: all N envs;
 * otherwise num_ids rows selected by int64 index (rollout variant). */
int elg_compute_torques(const ElgDims* dims, const ElgStepParams* prm, const float* actions, const float* dof_state,
                        const float* last_dof_vel, const float* p_gains, const float* d_gains, const float* torque_limits,
                        const float* default_dof_pos, float* torques, const int64_t* env_ids, int64_t num_ids, void* stream);

/* RobotBatchRollout.step_rollout's action hand-over (envs/batch_rollout/robot_batch_rollout.py:643-656, with the joint-target
 * denormalisation of robot_traj_grad_sampling.py:326-345 when joint_lower / joint_range are given): rollout env (main k, rollout r)
 * takes rollout_actions[k * R + r] -- optionally lower + (clamp(a, -1, 1) + 1) * range / 2 -- clipped to +- clip_actions, written to
 * row k * (1 + R) + 1 + r of `actions`; main rows are left alone.  One launch instead of clip + index_put (+ the copy that makes a
 * strided slice of the action plan contiguous). */
int elg_rollout_actions(const float* rollout_actions /*[M*R, A], rows src_row_stride floats apart*/, int32_t num_main, int32_t rollouts_per_main,
                        int32_t num_actions, int64_t src_row_stride /* 0: dense (= num_actions); e.g. horizon * A for step i of an [M*R, horizon, A] plan */,
                        float clip_actions, const float* joint_lower, const float* joint_range, float* actions, void* stream);

/* LeggedRobot.post_physics_step body (envs/base/legged_robot.py:122-150) with
 * _post_physics_step_callback's heading + heights (:394-401), check_termination (:155-160),
 * compute_reward (:215-232) + the _reward_* registry, compute_observations (:234-252) and the
 * history copies (:148-150), as ONE kernel.  `phase` is an OR of ELG_PHASE_* sections. */
int elg_post_physics_step(const ElgDims* dims, const ElgStepParams* prm, const ElgStepBuffers* buf, uint32_t phase, void* stream);

/* Launch-geometry override of elg_post_physics_step for benchmarking sweeps (no reference counterpart):
 * envs per chunk (multiple of 4; <= 32 for the generic kernel, <= 28 for the lean one), CTAs per SM (generic kernel only);
 * disable_bulk bit 0 forces the element-wise staging path of the generic kernel instead of TMA bulk copies, bit 1 disables the
 * lean kernel (elg_step_fast.cu) so that every call takes the generic one.  threads_per_cta is a diagnostic word for the lean
 * kernel: 16 launches it without programmatic dependent launch.  envs_per_chunk == 0 restores the built-in geometry. */
int elg_set_step_tuning(int envs_per_chunk, int threads_per_cta, int ctas_per_sm, int disable_bulk);

/* Diagnostic (no reference counterpart): when set to a device buffer of >= 64 int64, CTA 0 of every following
 * elg_post_physics_step launch records clock64() at its stage boundaries there; NULL switches it off. */
int elg_set_step_debug(long long* device_stamps);

/* Launch-floor probes (no reference counterpart; csrc/elg_probe.cu): an empty kernel and a pure bulk-copy round trip with the
 * step kernel's launch shape, so that the fixed cost of one launch in a PDL chain / CUDA graph and the cost of moving the
 * step's bytes with no arithmetic can be measured on the same box as the step itself (profiles/). */
int elg_probe_empty(int grid, int threads, int smem_bytes, int pdl, void* stream);
int elg_probe_roundtrip(const void* src, void* dst, int64_t bytes_in_per_cta, int64_t bytes_out_per_cta, int grid, int pdl, void* stream);

/* Probe (csrc/elg_stage.cu; no reference counterpart): SM-issued block copy between pinned (mapped) host memory and device memory,
 * measured against the copy engine for the end-to-end step's packed transfers (no faster: the PCIe link sets the rate).  dst / src /
 * bytes: multiples of 16.  mode 0: 16-byte load / store kernel; mode 1: TMA bulk pieces through shared memory.  grid <= 0: default. */
int elg_stage_block(void* dst, const void* src, int64_t bytes, int mode, int grid, void* stream);

/* LeggedRobot._get_heights (envs/base/legged_robot.py:900-938), standalone. cells_out (optional,
 * int32 [N,H,2]) receives the clipped (px, py) terrain cell of every point for index parity tests. */
int elg_get_heights(const ElgDims* dims, const ElgStepParams* prm, const float* root_states, const int16_t* height_samples,
                    const float* height_points, float* measured_heights, int32_t* cells_out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Main -> rollout state clone (envs/batch_rollout/robot_batch_rollout.py:1447-1535 _sync_main_to_rollout,
 * :1537-1583 _cache_main_env_states, :1585-1640 _restore_main_env_states; env layout _init_env_indices :119-164:
 * main env k is row k * (1 + rollouts_per_main) of every per-env tensor, its rollout envs are the rows behind it).
 * A field is one per-env tensor with a fixed number of bytes per env row; aliasing views of one tensor (base_pos /
 * base_quat of root_states, dof_pos / dof_vel of dof_state) are passed once, as the underlying tensor. */
#define ELG_MAX_CLONE_FIELDS 24
typedef enum ElgCloneMode {
  ELG_CLONE_SYNC = 0,   /* main row -> each of its rollout rows (+ optional position drift) */
  ELG_CLONE_CACHE = 1,  /* main rows -> cache tensors [num_main, row] (fields with cache == NULL are skipped) */
  ELG_CLONE_RESTORE = 2 /* cache tensors -> main rows */
} ElgCloneMode;
typedef struct ElgCloneField {
  void* base;        /* [num_main * (1 + rollouts_per_main), row] */
  void* cache;       /* [num_main, row] or NULL */
  int32_t row_bytes;
  int32_t reserved;
} ElgCloneField;
typedef struct ElgCloneTable {
  int32_t num_fields, num_main, rollouts_per_main;
  int32_t drift_field; /* index of the field whose first three floats are the base position (root_states), or -1 */
  ElgCloneField fields[ELG_MAX_CLONE_FIELDS];
} ElgCloneTable;
int elg_sizeof_clone_table(void);
/* drift: cfg.domain_rand.rollout_envs_sync_pos_drift (<= 0: none); drift_u: [num_main * rollouts_per_main, 3] uniform
 * samples (= torch.rand_like(base_pos[rollout_env_indices]), robot_batch_rollout.py:1493-1497) or NULL for in-kernel
 * Philox4x32-10 keyed by seed with counter (sample index, offset). */
int elg_clone_rows(const ElgCloneTable* table, int mode, float drift, const float* drift_u, uint64_t seed, uint64_t offset, void* stream);
/* Diagnostic (no reference counterpart): disable_bulk != 0 makes ELG_CLONE_SYNC write with per-thread 16-byte stores instead
 * of TMA bulk stores from replicated shared-memory tiles (A/B measurements); 0 restores the default. */
int elg_set_clone_tuning(int disable_bulk);

/* ---------------------------------------------------------------------------------------------------------------
 * Triangle-mesh queries (the reference delegates these to warp-lang 1.7: wp.Mesh / wp.mesh_query_ray /
 * wp.mesh_query_point_sign_normal).  ElgMesh is the only object the library owns: an opaque handle holding the
 * device-resident 4-wide BVH of one static mesh, the counterpart of utils/ray_caster.py:29-42 convert_to_warp_mesh. */
typedef struct ElgMesh ElgMesh;
/* vertices [V,3] fp32 and triangles [M,3] int32 are HOST pointers (the reference also builds from numpy arrays); the
 * BVH is built on the host once and uploaded to the current CUDA device. */
int elg_mesh_create(const float* vertices, int32_t num_vertices, const int32_t* triangles, int32_t num_triangles, ElgMesh** out);
/* The same with the leaf size of the BVH chosen by the caller: leaf_triangles = 0 (default: 3, best for ray casts on B200) or 1 ... 4
 * (4 is best for closest-point / SDF queries, where a leaf's triangles are all evaluated anyway).  Results do not depend on it. */
int elg_mesh_create_ex(const float* vertices, int32_t num_vertices, const int32_t* triangles, int32_t num_triangles, int32_t leaf_triangles,
                       ElgMesh** out);
int elg_mesh_free(ElgMesh* mesh);
int elg_mesh_info(const ElgMesh* mesh, int32_t* num_triangles, int32_t* num_nodes, float* bounds6);
/* Height-field-derived meshes (vertices on a regular xy grid, one or more layers, two triangles per cell -- the terrain meshes of
 * utils/terrain.py:76-80 and the two-layer confined terrain of utils/terrain_confine.py:13-146 without slope correction) also get a
 * grid accelerator at creation: rays walk the cells front to back instead of the BVH and run the same exact triangle tests -- identical
 * hit flags, distances and triangle ids.  Measured slower than the 4-wide BVH on B200 (it cannot skip empty air), so it is opt-in:
 * elg_set_mesh_tuning(1) selects it (A/B runs, tests), 0 (default) the BVH.  elg_mesh_grid_info reports it (layers == 0: none). */
int elg_mesh_grid_info(const ElgMesh* mesh, int32_t* layers, int32_t* nx, int32_t* ny);
int elg_set_mesh_tuning(int use_grid);

/* raycast_mesh + raycast_mesh_kernel (utils/ray_caster.py:45-167): closest hit with t in [0, max_dist) against both
 * face orientations; ray_hits = origin + t * direction, or the end point origin + max_dist * direction on a miss;
 * hits_found is torch.bool storage.  hit_distance (t, or max_dist on a miss) and hit_triangle (caller's triangle id, -1 on
 * a miss) are optional extras.  All ray buffers are device pointers, [num_rays, 3] fp32 contiguous. */
int elg_raycast(const ElgMesh* mesh, const float* ray_origins, const float* ray_directions, int64_t num_rays, float max_dist,
                float* ray_hits, uint8_t* hits_found, float* hit_distance, int32_t* hit_triangle, void* stream);

/* RayCaster._update_ray_casting (utils/ray_caster.py:558-594) fused with the ray cast.  pattern_origins / _directions
 * [num_rays, 3]: one sensor's pattern in the sensor frame (RayCaster.ray_origins[0] / ray_directions[0]); sensor_pos
 * [*, 3], sensor_quat [*, 4] xyzw (RayCaster.data.pos / .rot); env_ids: the num_sensors rows to update (int64) or NULL
 * for rows 0..num_sensors-1; yaw_only = cfg.attach_yaw_only.  ray_hits [*, num_rays, 3] and hits_found [*, num_rays]
 * are written at the selected rows only. */
int elg_raycast_sensor(const ElgMesh* mesh, const float* pattern_origins, const float* pattern_directions, int32_t num_rays,
                       const float* sensor_pos, const float* sensor_quat, const int64_t* env_ids, int64_t num_sensors, int yaw_only,
                       float max_dist, float* ray_hits, uint8_t* hits_found, void* stream);

/* The same launch with LeggedRobotRayCast._get_raycast_distances (envs/base/legged_robot_raycast.py:262-297) fused in:
 * distances[e, r] = |hit - dist_origins[e]| measured from the ROBOT BASE (dist_origins = root_states, dist_origin_stride = 13), and
 * with normalize != 0 the observation (1 - clamp(d / max_dist, 0, 1)) * found -- written where compute_observations (:232-260) would
 * concatenate it, so the ray observations never exist as a separate tensor. */
int elg_raycast_sensor_obs(const ElgMesh* mesh, const float* pattern_origins, const float* pattern_directions, int32_t num_rays,
                           const float* sensor_pos, const float* sensor_quat, const int64_t* env_ids, int64_t num_sensors, int yaw_only,
                           float max_dist, float* ray_hits, uint8_t* hits_found, const float* dist_origins, int32_t dist_origin_stride,
                           int32_t normalize, float* distances, int64_t distances_row_stride /* elements between env rows: num_rays for a dense
                           [N, num_rays] table, num_obs when the rows are the trailing columns of obs_buf */, void* stream);

/* DepthCameraWarp (utils/depth_camera.py:256-571) + DepthCameraBase.process_depth_image (:84-138). */
typedef struct ElgCamParams {
  int32_t width, height;         /* cfg.depth.original */
  int32_t out_width, out_height; /* cfg.depth.resized */
  int32_t buffer_len;
  int32_t resize;                /* 1: sizes differ, apply the separable tap tables (antialiased bicubic of torchvision's Resize) */
  int32_t max_taps;              /* taps per output pixel in the tables */
  float near_clip, far_clip;
  float noise_scale;             /* fp32(cfg.depth.dis_noise * 2); 0 or noise_u == NULL: no noise */
} ElgCamParams;
int elg_sizeof_cam_params(void);
/* DepthCameraWarp.update (:501-566): camera_pos = sensor_pos + quat_apply(sensor_quat, offset_pos); camera_rot =
 * quat_mul(sensor_quat, offset_quat) with the reference's literal argument order.  offset_* are HOST pointers. */
int elg_camera_pose(const float* sensor_pos, const float* sensor_quat, const int64_t* env_ids, int64_t num, const float* offset_pos3,
                    const float* offset_quat4, float* camera_pos, float* camera_rot, void* stream);
/* DepthCameraWarp.update_depth_buffer (:402-499), one launch: cast the [height * width] ray grid of every camera with
 * max_dist = far_clip, depth = -|hit - camera_pos| or -far_clip, + noise_scale * (noise_u[env] - 0.5), clip to
 * [-far, -near], resize, normalise to [-0.5, 0.5], then init (episode_length_buf <= 1) or shift-append the env's ring
 * buffer depth_buffer [num_envs, buffer_len, out_height, out_width].  raw_depth (optional) receives the unprocessed
 * [num_envs, height, width] image.  resize_*_start [out], resize_*_weights [out, max_taps]. */
int elg_depth_camera(const ElgMesh* mesh, const ElgCamParams* cam, const float* ray_directions, const float* camera_pos, const float* camera_rot,
                     const int64_t* episode_length_buf, const float* noise_u, const int32_t* resize_x_start, const float* resize_x_weights,
                     const int32_t* resize_y_start, const float* resize_y_weights, int64_t num_envs, float* depth_buffer, float* raw_depth,
                     void* stream);

/* MeshSDF.query + query_sdf_kernel (utils/mesh_sdf.py:38-116, :230-314): signed distance to the closest point of the mesh
 * within max_distance and its unit gradient.  Sign and closest face follow wp.mesh_query_point_sign_normal restated
 * order-independently: among the faces within d_min + epsilon * (mean edge length) of the point, the one whose unit normal
 * is most aligned with the offset decides (lowest triangle id on ties); sign = sign(n . (p - c)).  gradient = (p - c) / d
 * flipped inside, or the face normal when d <= 1e-6; nothing within max_distance: sdf = max_distance, gradient 0.
 * closest_points [n,3] / closest_face [n] are optional (nearest_points without the second query of :316-336). */
int elg_sdf_query(const ElgMesh* mesh, const float* points, int64_t num_points, float max_distance, float epsilon, float* sdf, float* grad,
                  float* closest_points, int32_t* closest_face, void* stream);
double elg_mesh_mean_edge(const ElgMesh* mesh);
/* RobotBatchRolloutPercept._update_sdf_values (envs/batch_rollout/robot_batch_rollout_percept.py:385-441) as ONE launch over (env,
 * query body) instead of a Python loop of gather + quat_rotate + one (or two) Warp round trips per body: the query point of body k is
 * rigid_body_state[env, body_indices[k], 0:3] + quat_rotate(its quaternion, sphere_offsets[k]) (HOST arrays; an offset whose first
 * component is NaN means "no offset", :401-414); sdf[env * sdf_row_stride + k], grad / nearest_points / query_points [N, K, 3]
 * (optional) are written for the selected rows only.  nearest = p - sdf * grad (utils/mesh_sdf.py:316-336) from the same traversal. */
int elg_sdf_query_bodies(const ElgMesh* mesh, const float* rigid_body_state /*[N*B,13]*/, int32_t num_bodies, const int32_t* body_indices,
                         const float* sphere_offsets, int32_t num_query_bodies, const int64_t* env_ids, int64_t num_rows, float max_distance,
                         float epsilon, float* sdf, int64_t sdf_row_stride, float* grad, float* nearest_points, float* query_points, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Navigation command update of the batch-rollout nav task (envs/batch_rollout/robot_batch_rollout_nav.py:135-247:
 * _update_navigation_commands followed by _check_goal_reached, both called from _post_physics_step_callback and its
 * rollout twin).  Env i steers towards goal_positions[i / (1 + rollouts_per_main)]: proportional world-frame velocity
 * clipped to max_linear_vel, rotated into the base frame; yaw rate towards the goal (2-D mode); exponential smoothing
 * with prev_commands when use_prev; commands[:, 0:3] overwritten, rows of envs whose goal_reached flag was set by the
 * PREVIOUS check are zeroed entirely; then goal_reached = distance < tolerance_rad.  One launch instead of two Python
 * loops over every env. */
typedef struct ElgNavParams {
  int32_t use_2d_nav;
  int32_t use_prev;          /* 0: prev_commands holds nothing yet (reference: prev_commands is None) */
  int32_t num_commands;      /* row length of commands */
  int32_t zero_reached;      /* 0: goal_reached holds nothing yet (reference: goal_reached is None) -> no rows are zeroed */
  float kp_linear, kp_angular, max_linear_vel, max_angular_vel;
  float smooth, smooth_c;    /* cmd_smooth_factor and fp32(1 - cmd_smooth_factor) */
  float tolerance_rad;
} ElgNavParams;
int elg_sizeof_nav_params(void);
int elg_nav_commands(int32_t num_main, int32_t rollouts_per_main, const ElgNavParams* prm, const float* root_states /*[N,13]*/,
                     const float* goal_positions /*[num_main,3]*/, float* commands /*[N,C]*/, float* prev_commands /*[N,3]*/,
                     uint8_t* goal_reached /*[N] bool*/, float* distance /*[N] or NULL*/, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Kinematic state integration of the planning variant (envs/batch_rollout/robot_plan_grad_sampling.py:103-195
 * _integrate_state_velocities followed by :197-225 _sync_integration_to_sim -- the reference always calls the pair): for the
 * selected envs clamp the commanded state velocities (3 linear + 3 angular + D joint), integrate base position, base
 * orientation (angle-axis increment, quat_mul, renormalise) and joint positions over n_substeps of sub_dt (Euler, or the
 * reference's "rk4" form), optionally clamp the joints to dof_pos_limits, store the velocities, and write the result
 * through to root_states / dof_state / base_lin_vel / base_ang_vel.  env_ids == NULL: all num_envs rows, state_vels row r
 * belongs to env r; otherwise state_vels row r belongs to env env_ids[r]. */
typedef struct ElgPlanParams {
  int32_t num_dof;
  int32_t method;            /* 0 euler, 1 rk4 */
  int32_t n_substeps;        /* ceil(dt / min(dt, max_integration_step)) */
  int32_t enforce_joint_limits;
  float sub_dt;              /* dt / n_substeps */
  float max_base_lin_vel, max_base_ang_vel, max_joint_vel;
} ElgPlanParams;
typedef struct ElgPlanBuffers {
  float* integration_base_pos;      /* [N,3] */
  float* integration_base_quat;     /* [N,4] */
  float* integration_dof_pos;       /* [N,D] */
  float* integration_base_lin_vel;  /* [N,3] */
  float* integration_base_ang_vel;  /* [N,3] */
  float* integration_dof_vel;       /* [N,D] */
  const float* dof_pos_limits;      /* [D,2] or NULL */
  float* root_states;               /* [N,13] */
  float* dof_state;                 /* [N*D,2] */
  float* base_lin_vel;              /* [N,3] */
  float* base_ang_vel;              /* [N,3] */
} ElgPlanBuffers;
int elg_sizeof_plan_params(void);
int elg_sizeof_plan_buffers(void);
/* state_vels == NULL: no integration, only the write-through of the stored integration_* state (_sync_integration_to_sim alone). */
int elg_integrate_state_velocities(const ElgPlanParams* prm, const ElgPlanBuffers* buf, const float* state_vels /*[rows, 6 + D]*/,
                                   const int64_t* env_ids, int64_t num_rows, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Caller-side fusion of the step outputs: rsl_rl EmpiricalNormalization.forward
 * (rsl_rl/modules/normalizer.py:43-75) with the normalised rows written directly to their destination -- e.g. the rollout-storage
 * slot observations[step] (rsl_rl/storage/rollout_storage.py:95) -- and, optionally, the reward / done columns copied to theirs
 * (:99-100).  training != 0: the batch statistics update mean / var / std / count first (skipped on the device, without a host
 * read, once count >= until; until < 0: never stops), exactly the reference's update rule:
 *   count += N; rate = N / count; d = mean_x - mean; mean += rate d; var += rate (var_x - var + d (mean_x - mean)); std = sqrt(var)
 * then out = (x - mean) / (std + eps).  mean / var / std are the [1, O] buffers of the module, count its int64 scalar (all on
 * the device).  `scratch` (elg_normalizer_scratch_bytes(N, O) bytes, 16-byte aligned, caller-owned, ZEROED once before its
 * first use -- its header holds the completion tickets, which every call leaves at zero again) is only needed when
 * training.  At most 1920 columns.  out may alias x; out == NULL updates the statistics only (EmpiricalNormalization.update). */
int64_t elg_normalizer_scratch_bytes(int64_t num_rows, int32_t num_cols);
int elg_set_normalizer_tuning(int mode);   /* bits 0-1: 0 = column-parallel single launch (thread-block cluster over the rows) up to
                                             65 536 rows, else two launches (default); 1 = always two launches; 2 = row-parallel single
                                             launch with a grid-wide hand-over (A/B, tests); bits 2-4: force the cluster size
                                             (1 -> 1 CTA, 2 -> 2, 3 -> 4, 4 -> 8) */
int elg_normalize_observations(int64_t num_rows, int32_t num_cols, const float* x, float* mean, float* var, float* std, int64_t* count, float eps,
                               int64_t until, int32_t training, float* out, void* scratch, const float* rew /*[N] or NULL*/,
                               float* rew_out, const uint8_t* dones /*[N] or NULL*/, uint8_t* dones_out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Actuator-network torques (envs/anymal_c/anymal.py:93-105, the default torque path of the anymal_c_* configs:
 * control.use_actuator_network, mixed_terrains/anymal_c_rough_config.py:68-69).  The TorchScript module
 * resources/actuator_nets/anydrive_v3_lstm.pt is `LSTMsea`: x * in_scale -> 2-layer LSTM(input 2, hidden 8, batch_first, one
 * time step) -> Linear(8, 1) -> * out_scale, applied to every (env, dof) row with
 *   x = (actions * action_scale + default_dof_pos - dof_pos, dof_vel),   gates in torch order (i, f, g, o).
 * `weights` is a device blob of ELG_ACTNET_WORDS floats in the layout below (host side: Anymal._pack_actuator_net);
 * hidden / cell are the [2, N*D, 8] state tensors (sea_hidden_state / sea_cell_state), updated in place. */
#define ELG_ACTNET_HIDDEN 8
#define ELG_ACTNET_GATES 32
enum {
  ELG_ACTNET_IN_SCALE = 0,                                   /* [2]                      */
  ELG_ACTNET_OUT_SCALE = 2,                                  /* [1] (+ 1 pad)            */
  ELG_ACTNET_W_IH0 = 4,                                      /* [32, 2]  weight_ih_l0    */
  ELG_ACTNET_W_HH0 = ELG_ACTNET_W_IH0 + 64,                  /* [32, 8]  weight_hh_l0    */
  ELG_ACTNET_B_IH0 = ELG_ACTNET_W_HH0 + 256,                 /* [32]                     */
  ELG_ACTNET_B_HH0 = ELG_ACTNET_B_IH0 + 32,                  /* [32]                     */
  ELG_ACTNET_W_IH1 = ELG_ACTNET_B_HH0 + 32,                  /* [32, 8]  weight_ih_l1    */
  ELG_ACTNET_W_HH1 = ELG_ACTNET_W_IH1 + 256,                 /* [32, 8]  weight_hh_l1    */
  ELG_ACTNET_B_IH1 = ELG_ACTNET_W_HH1 + 256,                 /* [32]                     */
  ELG_ACTNET_B_HH1 = ELG_ACTNET_B_IH1 + 32,                  /* [32]                     */
  ELG_ACTNET_W_LIN = ELG_ACTNET_B_HH1 + 32,                  /* [8]      linear.weight   */
  ELG_ACTNET_B_LIN = ELG_ACTNET_W_LIN + 8,                   /* [1] (+ 3 pad)            */
  ELG_ACTNET_WORDS = ELG_ACTNET_B_LIN + 4
};
int elg_actuator_net_words(void);
/* Kernel selection (no reference counterpart; every form returns bit-identical results): 0 (default) = one thread per (env, dof) row,
 * shared-memory weights (staged before the grid-dependency wait when the blob is the bound one); 2 = the same, staged after the wait;
 * 5 = the weights as constant-bank operands (bound blob only); 1 = eight lanes per row, one hidden unit each; 3 / 4 = four warps per
 * 32 rows, two hidden units per warp (constant bank / shared memory).  B200, 4096 envs x 12 dofs: DESIGN.md 4.2b. */
int elg_set_actuator_tuning(int mode);
/* Optional: declare the blob's contents FROZEN and copy them into the device's constant bank (stream-ordered).  Calls of
 * elg_actuator_net_torques that pass the SAME `weights` pointer afterwards may read the blob before the grid-dependency wait of a
 * programmatic dependent launch (i.e. while the preceding kernel of the stream is still running), and forms 3 / 5 take the weights
 * as constant operands; call it again after changing the blob's contents, or with NULL to unbind.  One bound blob per device. */
int elg_actuator_net_bind(const float* weights, void* stream);
int elg_actuator_net_torques(const ElgDims* dims, const float* weights, float action_scale, const float* actions, const float* dof_state,
                             const float* default_dof_pos, float* hidden, float* cell, float* torques, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * MPPI cost-weighted control update, batched over the main envs (in-tree statement: legged_gym/tests/score_sampling/
 * cmp_mppi_wbfo.py:216-233; the production optimiser is the external traj_sampling package, call sites
 * envs/batch_rollout/robot_traj_grad_sampling.py:62-69, :222-280).  Split in three so that the sample dimension can be
 * sharded across GPUs: costs are all-gathered, partial sums all-reduced (host side: utils/mppi.py). */
int elg_mppi_costs(const float* rewards /*[M,S,T]*/, int64_t num_main, int64_t num_samples, int32_t horizon, float* costs /*[M,S]*/, void* stream);
int elg_mppi_partials(const float* costs_all /*[M,S_total]*/, int64_t num_main, int32_t samples_total, int32_t first_local_sample,
                      int32_t samples_local, const float* samples /*[M,S_local,traj_size]*/, int32_t traj_size, float temperature,
                      float* partial /*[M, 1 + traj_size]: sum_e, sum_e * sample*/, void* stream);
int elg_mppi_finish(const float* partial, int64_t num_main, int32_t traj_size, float* mean_traj /*[M,traj_size]*/, void* stream);
/* the same stage over the rank-major cost layout an all-gather leaves: costs_ranked [num_ranks][M][samples_local] */
int elg_mppi_partials_ranked(const float* costs_ranked, int64_t num_main, int32_t num_ranks, int32_t rank, int32_t samples_local,
                             const float* samples, int32_t traj_size, float temperature, float* partial, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Collectives on the compute stream (csrc/elg_nccl.cu).  ElgComm wraps one ncclComm_t per rank; NCCL is
 * resolved with dlopen at the first call (the copy torch has mapped when there is one), so a single-GPU process never needs
 * it.  elg_comm_unique_id fills 128 bytes (ncclUniqueId) on rank 0; the host distributes them and every rank calls
 * elg_comm_init.  All calls enqueue on `stream`, are CUDA-graph capturable and never synchronise the host.
 *   elg_episode_stats_allreduce  in-place sum of `n` doubles: the (per-term sums, count, ...) vector elg_reset_envs accumulates
 *                                (ElgResetBuffers.stats_accum) -> extras["episode"] means of legged_robot.py:200-206 over ALL envs
 *   elg_mppi_update              cmp_mppi_wbfo.py:216-233 with the rollouts of every main env sharded over the ranks (equal
 *                                shares, rank order): costs -> all-gather -> weights + partial sums -> all-reduce -> mean trajectories.
 *                                costs_ranked [world, M, samples_local] and partial [M, 1 + traj_size] are caller scratch.
 *                                comm == NULL: single rank, no collective. */
typedef struct ElgComm ElgComm;
int elg_comm_unique_id(void* out128);
int elg_comm_init(const void* unique_id128, int rank, int world, ElgComm** out);
int elg_comm_destroy(ElgComm* comm);
int elg_comm_info(const ElgComm* comm, int* rank, int* world, int* nccl_version);
int elg_comm_warmup(ElgComm* comm, void* scratch /* >= 8 * world * n bytes */, int32_t n, void* stream);   /* first use of every collective kind, eagerly */
int elg_episode_stats_allreduce(double* stats, int32_t n, ElgComm* comm, void* stream);
int elg_mppi_update(const float* rewards /*[M,S_local,T]*/, const float* samples /*[M,S_local,traj_size]*/, int64_t num_main, int32_t samples_local,
                    int32_t horizon, int32_t traj_size, float temperature, float* costs_ranked, float* partial, float* mean_traj /*[M,traj_size]*/,
                    ElgComm* comm, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Sparse RNG-driven branches as predicated kernels over all envs: no nonzero(), no host synchronisation.
 * elg_resample_commands: envs/base/legged_robot.py:389-393 + _resample_commands :405-423 for envs with
 * (episode_length_buf + 1) % resample_interval == 0 (call BEFORE elg_post_physics_step, which increments the clock).
 * elg_reset_envs: reset_idx :162-213 for envs with reset_buf set (call AFTER the fused step): terrain curriculum :498-518,
 * _reset_dofs :450-465, _reset_root_states :467-487, _resample_commands, histories / timers / clock, (sum, count) of the
 * episode sums for extras["episode"] into stats (atomics; zeroed by elg_resample_commands when it is given the pointer, else by
 * the caller), episode sums zeroed,
 * and the command / dof_pos / dof_vel observation entries recomputed (obs_buf may be NULL).
 * update_command_curriculum (:520-531) stays on the host (it edits Python-side ranges once per max_episode_length steps). */
#define ELG_RESET_UNIFORMS 48   /* columns of the optional per-env uniform table: see csrc/elg_reset.cu */
typedef struct ElgResetParams {
  float lin_vel_x[2], lin_vel_y[2], ang_vel_yaw[2], heading[2];   /* command_ranges */
  int32_t heading_command;
  int32_t resample_interval;     /* int(cfg.commands.resampling_time / dt) */
  float base_init_state[13];
  int32_t custom_origins;
  int32_t curriculum;            /* cfg.terrain.curriculum (and init_done) */
  float env_length_half;         /* terrain.env_length / 2 */
  float max_episode_length_s;
  int32_t max_terrain_level;
  int32_t terrain_cols;          /* second dim of terrain_origins */
  uint64_t seed, offset;         /* Philox key / step counter when no uniform table is given */
  /* main / rollout env layout (envs/batch_rollout/robot_batch_rollout.py): 0 = flat env list; R1 = 1 + rollouts per main.
   * Commands are resampled for MAIN envs only (clock / reset flag of the main row) and copied to every row of the group
   * (:819-838, :900-915); the terrain curriculum and the extras["episode"] sums cover main rows only (:884-887, :925-931);
   * episode sums of the reset rows are zeroed only when at least one main env resets in the step (:925-931). */
  int32_t rows_per_main;
  /* RobotBatchRollout._reset_root_states (:1366-1404): with custom origins on a heightfield / trimesh terrain the new base height is
   * height_samples[cell(x, y)] * vertical_scale + base_init_state[2] (single cell, no min-of-3; terrain geometry from ElgStepParams) */
  int32_t root_z_from_terrain;
} ElgResetParams;
typedef struct ElgResetBuffers {
  const uint8_t* reset_buf;      /* [N] bool */
  float* root_states;            /* [N,13] */
  float* dof_state;              /* [N*D,2] */
  float* commands;               /* [N,C] */
  float* env_origins;            /* [N,3] */
  int64_t* terrain_levels;       /* [N] or NULL */
  const int64_t* terrain_types;  /* [N] or NULL */
  const float* terrain_origins;  /* [rows, cols, 3] or NULL */
  const float* default_dof_pos;  /* [D] */
  float* last_dof_vel;           /* [N,D] */
  float* last_root_vel;          /* [N,6] */
  float* feet_air_time;          /* [N,F] */
  float* feet_contact_time;      /* [N,F] */
  int64_t* episode_length_buf;   /* [N] */
  float* episode_sums;           /* [ELG_NUM_REWARD_TERMS, N] */
  float* stats;                  /* [ELG_NUM_REWARD_TERMS + 2]: per-term sums, count; word [NUM + 1] is scratch (main-reset flag) */
  double* stats_accum;           /* [ELG_NUM_REWARD_TERMS + 1] or NULL: the same (sum, count) added on top of what is there -- the
                                    running totals a sharded run all-reduces once per K steps (elg_episode_stats_allreduce) */
  float* obs_buf;                /* [N,O] or NULL */
  const float* measured_heights; /* [N,H] (stale heights the repaired height observations are built from) or NULL */
  const float* noise_scale_vec;  /* [O] */
  const float* noise_u;          /* [N,O] (ELG_NOISE_TENSOR) or NULL */
  const float* uniforms;         /* [N, ELG_RESET_UNIFORMS] or NULL (Philox) */
  const int16_t* height_samples; /* [rows, cols]; only read with root_z_from_terrain */
} ElgResetBuffers;
int elg_sizeof_reset_params(void);
int elg_sizeof_reset_buffers(void);
int elg_resample_commands(const ElgDims* dims, const ElgResetParams* rp, const int64_t* episode_length_buf, float* commands,
                          const float* uniforms, float* stats_to_zero /* [ELG_NUM_REWARD_TERMS + 2] or NULL */, void* stream);
int elg_reset_envs(const ElgDims* dims, const ElgResetParams* rp, const ElgStepParams* prm, const ElgResetBuffers* buf, void* stream);

/* Init-time helper for _get_heights (envs/base/legged_robot.py:932-938): out[i][j] =
 * fp32(min(hs[i][j], hs[i+1][j], hs[i][j+1])) * vertical_scale for i <= rows-2, j <= cols-2 (0 elsewhere) -- the
 * value the reference computes per height point, tabulated once per (static) terrain so the step kernel gathers one
 * float instead of three int16.  Bit-identical to the per-point evaluation. */
int elg_prepare_height_field(const int16_t* height_samples, int32_t rows, int32_t cols, float vertical_scale, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ELG_B200_H */
