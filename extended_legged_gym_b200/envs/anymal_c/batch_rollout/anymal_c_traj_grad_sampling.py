"""``AnymalCTrajGradSampling`` -- the ANYmal-C task class of the sampling-based trajectory optimiser
(envs/anymal_c/batch_rollout/anymal_c_traj_grad_sampling.py:17-381 in /root/reference/legged_gym/legged_gym; BASELINE config 5 runs
this robot): ``RobotTrajGradSampling`` plus

  gait scheduler      :60-75, :324-331   ``cfg.gait_scheduler`` on the env clock (``GaitClockMixin``)
  DIAL-MPC reward set :112-290, :344-356 ``gaits, air_time, pos, upright, yaw, vel, ang_vel, height, energy, alive, no_fly`` -- the terms
                                         of the reference's dial-mpc configs (anymal_c_dialmpc_flat_config.py).  They are subclass terms in
                                         the sense of SURVEY section 8b: torch expressions on the device, evaluated next to the kernel's
                                         built-ins (``_python_terms``; main steps and, after the DERIVE section, rollout steps).  The
                                         class's DEFAULT config (anymal_c_traj_grad_sampling_config.py:212-229) enables stock terms only,
                                         so its horizon loop stays one CUDA graph of kernel launches.

``tests/test_robot_rollout_classes.py`` compares every term with the unmodified reference method on the same tensors (CPU, bit for bit).
"""
import torch

from ....utils.math_utils import quat_apply, quat_apply_yaw
from ...batch_rollout.robot_traj_grad_sampling import RobotTrajGradSampling
from .anymal_c_batch_rollout import GaitClockMixin


class DialMpcRewardMixin:
    # foot order of the asset; per gait: phase offsets and (duty ratio, cadence [Hz], swing amplitude [m])  (:42-57)
    GAIT_PHASES = {"stand": [0.0, 0.0, 0.0, 0.0], "walk": [0.0, 0.5, 0.75, 0.25], "trot": [0.0, 0.5, 0.5, 0.0],
                   "canter": [0.0, 0.33, 0.33, 0.66], "gallop": [0.0, 0.05, 0.4, 0.35]}
    GAIT_PARAMS = {"stand": [1.0, 1.0, 0.0], "walk": [0.75, 1.0, 0.08], "trot": [0.45, 2.0, 0.08],
                   "canter": [0.4, 4.0, 0.06], "gallop": [0.3, 3.5, 0.10]}

    def _init_dial_mpc(self):
        self._gait = "trot"
        self._gait_phase = {k: torch.tensor(v, device=self.device) for k, v in self.GAIT_PHASES.items()}
        self._gait_params = {k: torch.tensor(v, device=self.device) for k, v in self.GAIT_PARAMS.items()}

    def get_foot_step(self, duty_ratio, cadence, amplitude, phases, time):
        """target foot heights of a gait at ``time`` (:114-144): zero in stance, a half sine of ``amplitude`` over the swing"""
        gait_phase = torch.fmod(time * cadence + phases, 1.0)
        swing = ~(gait_phase < duty_ratio)
        heights = torch.zeros_like(gait_phase)
        heights[swing] = amplitude * torch.sin((gait_phase[swing] - duty_ratio) / (1.0 - duty_ratio) * torch.pi)
        return heights

    def _reward_gaits(self):      # (:148-168) the clock of the ROLLOUT envs on every row: the main rows' value is not used
        duty_ratio, cadence, amplitude = self._gait_params[self._gait]
        target = self.get_foot_step(duty_ratio, cadence, amplitude, self._gait_phase[self._gait], self.t_rollout)
        target = target.unsqueeze(0).repeat(self.total_num_envs, 1)
        return -torch.sum(((target - self.foot_positions[:, :, 2]) / 0.05) ** 2, dim=1)

    def _reward_air_time(self):   # (:170-190) feet_air_time's bookkeeping with a 0.1 s threshold; writes the timers and last_contacts
        contact = self.contact_forces[:, self.feet_indices, 2] > 1.0
        contact_filt = torch.logical_or(contact, self.last_contacts)
        self.last_contacts.copy_(contact)        # (in place: the kernels and the clone tables hold this tensor's pointer)
        first_contact = (self.feet_air_time > 0.0) * contact_filt
        self.feet_air_time += self.dt
        rew = torch.sum((self.feet_air_time - 0.1) * first_contact, dim=1)
        self.feet_air_time *= ~contact_filt
        return rew

    def _reward_pos(self):        # (:192-210)
        pos = self.root_states[:, :3]
        elapsed = torch.ones_like(self.commands[:, 0]) * self.t_main
        target = self.commands[:, :3] * self.dt * elapsed.unsqueeze(1)
        head = torch.tensor([0.285, 0.0, 0.0], device=self.device)
        head_pos = pos + quat_apply(self.base_quat, head.repeat(pos.shape[0], 1)) * 0.285
        return -torch.sum((head_pos - target) ** 2, dim=1)

    def _reward_upright(self):    # (:212-222)
        up = torch.zeros_like(self.projected_gravity)
        up[:, 2] = -1.0
        return -torch.sum((self.projected_gravity - up) ** 2, dim=1)

    def _reward_yaw(self):        # (:224-240)
        fwd = quat_apply_yaw(self.base_quat, torch.tensor([1.0, 0.0, 0.0], device=self.device).repeat(self.total_num_envs, 1))
        yaw = torch.atan2(fwd[:, 1], fwd[:, 0])
        diff = yaw - (self.commands[:, 3] if self.cfg.commands.heading_command else 0.0)
        return -torch.square(torch.atan2(torch.sin(diff), torch.cos(diff)))

    def _reward_vel(self):        # (:242-250)
        return -torch.sum((self.base_lin_vel[:, :2] - self.commands[:, :2]) ** 2, dim=1)

    def _reward_ang_vel(self):    # (:252-260)
        return -torch.square(self.base_ang_vel[:, 2] - self.commands[:, 2])

    def _reward_height(self):     # (:262-273)
        return -torch.square(self.root_states[:, 2] - self.cfg.rewards.base_height_target)

    def _reward_energy(self):     # (:275-284) positive mechanical power, normalised by 160 W
        power = torch.clamp(self.torques * self.dof_vel, min=0.0) / 160.0
        return -torch.sum(power ** 2, dim=1)

    def _reward_alive(self):      # (:286-288) 1 - reset_buf (the reference's expression accepts integer flags only; flags are bool here)
        return 1.0 - self.reset_buf.to(torch.float)

    def _reward_no_fly(self):     # (:344-352) at least one foot on the ground
        contacts = self.contact_forces[:, self.feet_indices, 2] > 0.1
        return 1.0 * (torch.sum(1.0 * contacts, dim=1) >= 1)


class AnymalCTrajGradSampling(GaitClockMixin, DialMpcRewardMixin, RobotTrajGradSampling):
    def __init__(self, cfg, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True):
        super().__init__(cfg, sim_params, physics_engine, sim_device, headless)
        self._init_dial_mpc()
        self._init_gait_clock()
