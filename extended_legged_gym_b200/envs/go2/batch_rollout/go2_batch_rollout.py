"""``Go2BatchRollout`` -- the Unitree Go2 main / rollout task class of the reference (envs/go2/batch_rollout/go2_batch_rollout.py:49-230
in /root/reference/legged_gym/legged_gym).  Its body is ``AnymalCBatchRollout``'s, hook for hook (the two reference files differ in
the class name and the default config only): optional actuator-network torques on every row, the gait scheduler on the env clock,
upside-down MAIN robots reset (:193-200), the async-scheduler posture term."""
from ...anymal_c.batch_rollout.anymal_c_batch_rollout import AnymalCBatchRollout


class Go2BatchRollout(AnymalCBatchRollout):
    pass
