"""``Go2`` -- task class of the reference's Unitree Go2 (envs/go2/go2.py:47-111 in /root/reference/legged_gym/legged_gym): the same hooks
as ``Anymal`` -- optional actuator-network torque path (:52-55, :84-98; off in go2_rough_config.py:90), gait scheduler with period 0.6 s,
trot phases and 0.15 m swing height stepped after every env step (:58-75, :108-111), ``_reward_gait_scheduler`` (:113-115), and the
network-state clearing in ``reset_idx`` (:77-82).  All of it is ``Anymal``'s implementation here: the gait clock and the foot-height
tracking reward run inside the step kernel."""
from ..anymal_c.anymal import Anymal


class Go2(Anymal):
    pass
