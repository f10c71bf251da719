"""``Go2TrajGradSampling`` -- the Unitree Go2 task class of the sampling-based trajectory optimiser
(envs/go2/batch_rollout/go2_traj_grad_sampling.py:33-330 in /root/reference/legged_gym/legged_gym): ``RobotTrajGradSampling`` plus the
DIAL-MPC reward set and its gait tables -- the reference file carries the same terms as ``AnymalCTrajGradSampling`` minus ``no_fly``
(``alive`` as ``1 - reset_buf.long()``, :198), without a gait scheduler object.  The class's default config enables six of them and
no stock term: the registry launch then only adds the Python-side sum, clips and stores."""
from ...anymal_c.batch_rollout.anymal_c_traj_grad_sampling import DialMpcRewardMixin
from ...batch_rollout.robot_traj_grad_sampling import RobotTrajGradSampling


class Go2TrajGradSampling(DialMpcRewardMixin, RobotTrajGradSampling):
    def __init__(self, cfg, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True):
        super().__init__(cfg, sim_params, physics_engine, sim_device, headless)
        self._init_dial_mpc()
