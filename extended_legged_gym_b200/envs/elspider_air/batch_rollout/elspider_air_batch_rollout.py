"""``ElSpiderAirBatchRollout`` -- the hexapod's main / rollout task class of the reference
(envs/elspider_air/batch_rollout/elspider_air_batch_rollout.py:46-230 in /root/reference/legged_gym/legged_gym):
``RobotBatchRolloutPercept`` with ``ElSpider``'s hooks -- 18 DOF / 6 feet, tripod ``_reward_gait_2_step`` (:201-230), the actuator
network on every row (``_compute_torques(actions, env_ids=None)`` :152-168, state cleared for reset rows :137-150), the async-scheduler
posture term (:178-190) -- and two rules of its own:

  check_termination  :170-176  EVERY upside-down row is reset, rollout rows included (``ElgStepParams.terminate_upside_down = 1``; the
                               ANYmal / Go2 classes flag main rows only)
  gait scheduler     :64-78, :132-135  ``cfg.gait_scheduler``; advanced by ``post_physics_step_rollout`` ONLY (incrementally, no clock
                               argument): a main step leaves the phase and the stored feet untouched

The hexapod runs through the generic step kernel in both modes (the lean kernel is the 12-DOF / 4-foot layout).  Oracle:
``RobotBatchRolloutOracle(upside_down_rows="all", gait_period=None)``, bit-identical to the unmodified reference class on
tests/golden/rollout_step_anymal.npz (tag d); on the B200 the class's main step reproduces that fixture
(tests/test_robot_rollout_classes.py).  Its rollout-mode step has NOT been run on hardware yet (the round's GPU budget ended there):
the generic kernel's rollout mode is exercised with the 12-DOF robots only."""
from types import SimpleNamespace

from ...batch_rollout.robot_batch_rollout_percept import RobotBatchRolloutPercept
from ..elspider import ElSpider


class ElSpiderAirBatchRollout(ElSpider, RobotBatchRolloutPercept):
    def __init__(self, cfg, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True):
        super().__init__(cfg, sim_params, physics_engine, sim_device, headless)
        g = getattr(self.cfg, "gait_scheduler", None)
        if g is not None:
            self.gait_cfg = SimpleNamespace(dt=g.dt, period=g.period, foot_phases=list(g.foot_phases), swing_height=g.swing_height)
            self._params_dirty = True

    def _compute_torques(self, actions, env_ids=None):
        torques = super()._compute_torques(actions)
        return torques if env_ids is None else torques[env_ids]

    def post_physics_step(self):
        # the step kernel advances the gait phase and stores the feet heights on every launch; this class's scheduler is not
        # stepped by the main step (only post_physics_step_rollout calls it, :132-135): put both back
        keep = None if self.gait_idx is None else (self.gait_idx.clone(), self.gait_prev_foot_z.clone())
        super().post_physics_step()
        if keep is not None:
            self.gait_idx.copy_(keep[0])
            self.gait_prev_foot_z.copy_(keep[1])
