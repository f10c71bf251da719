"""``BaseTask``: buffer allocation and the VecEnv-facing accessors.

Follows the contract of the reference's ``BaseTask`` (envs/base/base_task.py:41-119): buffer
names, dtypes (``reset_buf`` starts as int64 ones, ``episode_length_buf`` int64, ``time_out_buf``
bool) and ``reset()`` = ``reset_idx(all)`` + one zero-action ``step``.  Viewer / rendering are
outside the hot path and are not provided.
"""
import torch


class BaseTask:
    def __init__(self, cfg, sim_params, physics_engine, sim_device, headless):
        self.sim_params = sim_params
        self.physics_engine = physics_engine
        self.sim_device = sim_device
        self.headless = headless
        self.device = sim_device if isinstance(sim_device, str) else str(sim_device)
        if not self.device.startswith("cuda"):
            raise RuntimeError("extended_legged_gym_b200 runs the per-step path on a B200 only: sim_device must be "
                               f"a cuda device, got '{self.device}' (there is no CPU fallback)")
        self.num_envs = cfg.env.num_envs
        self.num_obs = cfg.env.num_observations
        self.num_privileged_obs = cfg.env.num_privileged_obs
        self.num_actions = cfg.env.num_actions

        f32 = dict(device=self.device, dtype=torch.float)
        self.obs_buf = torch.zeros(self.num_envs, self.num_obs, **f32)
        self.rew_buf = torch.zeros(self.num_envs, **f32)
        self.reset_buf = torch.ones(self.num_envs, device=self.device, dtype=torch.long)
        self.episode_length_buf = torch.zeros(self.num_envs, device=self.device, dtype=torch.long)
        self.time_out_buf = torch.zeros(self.num_envs, device=self.device, dtype=torch.bool)
        self.privileged_obs_buf = None if self.num_privileged_obs is None else \
            torch.zeros(self.num_envs, self.num_privileged_obs, **f32)
        self.extras = {}
        self.create_sim()
        self.enable_viewer_sync = True
        self.viewer = None

    def create_sim(self):
        raise NotImplementedError

    def get_observations(self):
        return self.obs_buf

    def get_privileged_observations(self):
        return self.privileged_obs_buf

    def reset_idx(self, env_ids):
        raise NotImplementedError

    def reset(self):
        self.reset_idx(torch.arange(self.num_envs, device=self.device))
        obs, privileged_obs, _, _, _ = self.step(torch.zeros(self.num_envs, self.num_actions, device=self.device, requires_grad=False))
        return obs, privileged_obs

    def step(self, actions):
        raise NotImplementedError

    def render(self, sync_frame_time=True):
        return None
