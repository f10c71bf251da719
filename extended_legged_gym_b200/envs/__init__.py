"""Environment classes and configs on the hot path (mirrors ``legged_gym.envs`` for the BASELINE configs)."""
from .base.legged_robot_config import LeggedRobotCfg, LeggedRobotCfgPPO
from .base.legged_robot import LeggedRobot
from .base.legged_robot_raycast import LeggedRobotRayCast
from .base.legged_robot_depthcam import LeggedRobotDepth
from .anymal_c.anymal_c_config import AnymalCRoughCfg, AnymalCRoughCfgPPO, AnymalCFlatCfg, AnymalCFlatCfgPPO
from .anymal_c.anymal import Anymal
from .a1.a1_config import A1RoughCfg, A1RoughCfgPPO
from .go2.go2_config import Go2RoughCfg, Go2RoughCfgPPO
from .go2.go2 import Go2
from .elspider_air.elspider_air_config import ElSpiderAirRoughCfg, ElSpiderAirRoughCfgPPO
from .elspider_air.elspider import ElSpider
from .batch_rollout.robot_batch_rollout import RobotBatchRollout
from .batch_rollout.robot_batch_rollout_config import RobotBatchRolloutCfg, RobotBatchRolloutCfgPPO, RobotBatchRolloutPerceptCfg
from .batch_rollout.robot_batch_rollout_percept import RobotBatchRolloutPercept
from .batch_rollout.robot_traj_grad_sampling import RobotTrajGradSampling
from .anymal_c.batch_rollout.anymal_c_batch_rollout import AnymalCBatchRollout
from .anymal_c.batch_rollout.anymal_c_batch_rollout_config import AnymalCBatchRolloutCfg, AnymalCBatchRolloutCfgPPO
from .anymal_c.batch_rollout.anymal_c_traj_grad_sampling import AnymalCTrajGradSampling
from .anymal_c.batch_rollout.anymal_c_traj_grad_sampling_config import AnymalCTrajGradSamplingCfg, AnymalCTrajGradSamplingCfgPPO
from .go2.batch_rollout.go2_batch_rollout import Go2BatchRollout
from .go2.batch_rollout.go2_batch_rollout_config import Go2BatchRolloutCfg, Go2BatchRolloutCfgPPO
from .go2.batch_rollout.go2_traj_grad_sampling import Go2TrajGradSampling
from .elspider_air.batch_rollout.elspider_air_batch_rollout import ElSpiderAirBatchRollout
from .elspider_air.batch_rollout.elspider_air_batch_rollout_config import ElSpiderAirBatchRolloutCfg, ElSpiderAirBatchRolloutCfgPPO
from .go2.batch_rollout.go2_traj_grad_sampling_config import Go2TrajGradSamplingCfg, Go2TrajGradSamplingCfgPPO
from .batch_rollout.robot_traj_grad_sampling_config import RobotTrajGradSamplingCfg, RobotTrajGradSamplingCfgPPO
from .batch_rollout.robot_batch_rollout_nav import RobotBatchRolloutNav
from .batch_rollout.robot_batch_rollout_nav_config import RobotBatchRolloutNavCfg, RobotBatchRolloutNavCfgPPO
from .batch_rollout.robot_plan_grad_sampling import KinematicStateIntegration, RobotPlanGradSampling
from .batch_rollout.robot_plan_grad_sampling_config import RobotPlanGradSamplingCfg, RobotPlanGradSamplingCfgPPO

TASKS = {
    "anymal_c_rough": (Anymal, AnymalCRoughCfg, AnymalCRoughCfgPPO),      # legged_gym/envs/__init__.py registers Anymal for both
    "anymal_c_flat": (Anymal, AnymalCFlatCfg, AnymalCFlatCfgPPO),
    "a1": (LeggedRobot, A1RoughCfg, A1RoughCfgPPO),
    "go2_rough": (Go2, Go2RoughCfg, Go2RoughCfgPPO),                      # legged_gym/envs/__init__.py:66
    "elspider_air_rough": (ElSpider, ElSpiderAirRoughCfg, ElSpiderAirRoughCfgPPO),
    "anymal_c_batch_rollout": (AnymalCBatchRollout, AnymalCBatchRolloutCfg, AnymalCBatchRolloutCfgPPO),   # legged_gym/envs/__init__.py
    "go2_batch_rollout": (Go2BatchRollout, Go2BatchRolloutCfg, Go2BatchRolloutCfgPPO),
    "anymal_c_traj_grad_sampling": (AnymalCTrajGradSampling, AnymalCTrajGradSamplingCfg, AnymalCTrajGradSamplingCfgPPO),
    "go2_traj_grad_sampling": (Go2TrajGradSampling, Go2TrajGradSamplingCfg, Go2TrajGradSamplingCfgPPO),
    "elspider_air_batch_rollout": (ElSpiderAirBatchRollout, ElSpiderAirBatchRolloutCfg, ElSpiderAirBatchRolloutCfgPPO),
}
