"""``Anymal`` -- ANYmal-C task class of the reference (envs/anymal_c/anymal.py:48-108 in
/root/reference/legged_gym/legged_gym): ``LeggedRobot`` + the actuator-network torque path + the gait scheduler.

  _compute_torques()   :93-105  -> elg_actuator_net_torques (one thread per (env, dof): LSTMsea end to end) when
                                  ``cfg.control.use_actuator_network``; PD controller otherwise (base class)
  _init_buffers()      :85-91   sea_input / sea_hidden_state / sea_cell_state (+ the per-env views), same shapes
  reset_idx()          :79-83   clears the network state of the reset envs
  gait scheduler       :60-77, :107-113, utils/gait_scheduler.py:63-81 -> evaluated inside the fused step kernel

The network parameters come from the TorchScript file the config names (``actuator_net_file``; ``{LEGGED_GYM_ROOT_DIR}`` is
taken from the environment variable of that name or from ``cfg.control.legged_gym_root_dir``) or from
``cfg.control.actuator_net_weights`` (a dict / ``.npz`` with the module's ``state_dict`` + ``in_scale`` + ``out_scale``).
There is no fallback: if neither can be loaded the constructor raises.
"""
import ctypes as C
import os
from types import SimpleNamespace

import numpy as np
import torch

from ... import _lib
from ..base.legged_robot import LeggedRobot
from ..base.legged_robot_rew_mixin import _stock

_WEIGHT_LAYOUT = (("in_scale", _lib.ACTNET_IN_SCALE, 2), ("out_scale", _lib.ACTNET_OUT_SCALE, 1),
                  ("lstm.weight_ih_l0", _lib.ACTNET_W_IH0, 64), ("lstm.weight_hh_l0", _lib.ACTNET_W_HH0, 256),
                  ("lstm.bias_ih_l0", _lib.ACTNET_B_IH0, 32), ("lstm.bias_hh_l0", _lib.ACTNET_B_HH0, 32),
                  ("lstm.weight_ih_l1", _lib.ACTNET_W_IH1, 256), ("lstm.weight_hh_l1", _lib.ACTNET_W_HH1, 256),
                  ("lstm.bias_ih_l1", _lib.ACTNET_B_IH1, 32), ("lstm.bias_hh_l1", _lib.ACTNET_B_HH1, 32),
                  ("linear.weight", _lib.ACTNET_W_LIN, 8), ("linear.bias", _lib.ACTNET_B_LIN, 1))


def load_actuator_net_weights(cfg_control):
    """dict name -> float32 CPU tensor with the LSTMsea parameters, from cfg.control (see the module docstring)."""
    w = getattr(cfg_control, "actuator_net_weights", None)
    if w is not None:
        if isinstance(w, (str, os.PathLike)):
            z = np.load(w)
            w = {k: z[k] for k in z.files}
        return {name: torch.as_tensor(np.asarray(w[name]), dtype=torch.float).reshape(-1) for name, _, _ in _WEIGHT_LAYOUT}
    root = os.environ.get("LEGGED_GYM_ROOT_DIR", getattr(cfg_control, "legged_gym_root_dir", ""))
    path = cfg_control.actuator_net_file.format(LEGGED_GYM_ROOT_DIR=root)
    if not os.path.exists(path):
        raise FileNotFoundError(f"actuator network '{path}' not found: set LEGGED_GYM_ROOT_DIR (or cfg.control.legged_gym_root_dir) to the "
                                "legged_gym checkout that holds resources/actuator_nets/, or provide cfg.control.actuator_net_weights")
    net = torch.jit.load(path, map_location="cpu")     # anymal.py:56-58
    sd = {k: v.detach().float() for k, v in net.state_dict().items()}
    sd["in_scale"], sd["out_scale"] = net.in_scale.detach().float(), net.out_scale.detach().float()
    return {name: sd[name].reshape(-1) for name, _, _ in _WEIGHT_LAYOUT}


class Anymal(LeggedRobot):
    def __init__(self, cfg, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True):
        super().__init__(cfg, sim_params, physics_engine, sim_device, headless)
        self.actuator_net_blob = None
        if self.cfg.control.use_actuator_network:
            self.actuator_net_blob = self._pack_actuator_net(load_actuator_net_weights(self.cfg.control))
            # the parameters never change after this: keep them in the constant bank (immediate operands of the kernel's FMAs)
            with torch.cuda.device(self.device):
                _lib.check(self._lib.elg_actuator_net_bind(self.actuator_net_blob.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream),
                           "elg_actuator_net_bind")
        # gait scheduler (anymal.py:60-77): period 0.6 s, trot phases, 0.15 m swing height -- state lives in two env buffers,
        # the update and the foot-z tracking reward run inside the step kernel
        self.gait_cfg = SimpleNamespace(dt=self.dt, period=0.6, foot_phases=[0.0, 0.5, 0.5, 0.0], swing_height=0.15)
        self.gait_idx = torch.zeros(self.num_envs, device=self.device)
        self.gait_prev_foot_z = torch.zeros(self.num_envs, len(self.feet_indices), device=self.device)
        self._params_dirty = True

    def _pack_actuator_net(self, weights):
        blob = torch.zeros(_lib.ACTNET_WORDS, dtype=torch.float)
        for name, off, n in _WEIGHT_LAYOUT:
            t = weights[name]
            if t.numel() != n:
                raise _lib.ElgError(f"actuator network: '{name}' has {t.numel()} elements, LSTMsea(2 -> 8 x 2 layers -> 1) needs {n}")
            blob[off:off + n] = t
        return blob.to(self.device)

    def _init_buffers(self):
        super()._init_buffers()
        N, A = self.num_envs, self.num_actions
        z = dict(device=self.device, requires_grad=False)
        self.sea_input = torch.zeros(N * A, 1, 2, **z)          # kept for attribute parity; the kernel forms the inputs in registers
        self.sea_hidden_state = torch.zeros(2, N * A, 8, **z)
        self.sea_cell_state = torch.zeros(2, N * A, 8, **z)
        self.sea_hidden_state_per_env = self.sea_hidden_state.view(2, N, A, 8)
        self.sea_cell_state_per_env = self.sea_cell_state.view(2, N, A, 8)

    def _compute_torques(self, actions):
        if not self.cfg.control.use_actuator_network:
            return super()._compute_torques(actions)
        self._sync_native()
        if not (actions.is_cuda and actions.is_contiguous() and actions.dtype == torch.float):
            actions = actions.to(self.device, torch.float).contiguous()
        out = self.torques
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self._lib.elg_actuator_net_torques(C.byref(self._dims), self.actuator_net_blob.data_ptr(), float(self.cfg.control.action_scale),
                                                actions.data_ptr(), self.dof_state.data_ptr(), self.default_dof_pos.data_ptr(),
                                                self.sea_hidden_state.data_ptr(), self.sea_cell_state.data_ptr(), out.data_ptr(), stream)
        _lib.check(rc, "elg_actuator_net_torques")
        return out

    def reset_idx(self, env_ids):
        super().reset_idx(env_ids)
        if len(env_ids) == 0:
            return
        # additionally empty the actuator network state (anymal.py:79-83)
        self.sea_hidden_state_per_env[:, env_ids] = 0.0
        self.sea_cell_state_per_env[:, env_ids] = 0.0

    def _launch_reset(self):
        """fused reset path: the same clearing, predicated on the reset flags (no index list, no host sync)"""
        super()._launch_reset()
        m = self._reset_bool.view(1, self.num_envs, 1, 1)
        self.sea_hidden_state_per_env.masked_fill_(m, 0.0)
        self.sea_cell_state_per_env.masked_fill_(m, 0.0)

    @_stock
    def _reward_gait_scheduler(self):
        """Foot-height tracking of the gait scheduler (anymal.py:112-114 -> utils/gait_scheduler.py:74-81): torch statement
        for callers of the registry; inside step() the fused kernel evaluates the same expression (ELG_REW_GAIT_SCHEDULER)
        on the feet and phases GaitScheduler.step stored after the previous step."""
        rew = torch.zeros(self.num_envs, device=self.device)
        for i, phase in enumerate(self.gait_cfg.foot_phases):
            ph = torch.remainder(self.gait_idx + phase, 1.0)
            target = torch.where(ph < 0.5, self.gait_cfg.swing_height * torch.sin(2 * torch.pi * ph), torch.zeros_like(ph))
            rew += torch.square(target - self.gait_prev_foot_z[:, i])
        return rew
