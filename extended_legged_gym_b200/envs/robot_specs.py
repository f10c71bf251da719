"""Asset facts the hot path needs, without PhysX.

In the reference these numbers come out of Isaac Gym's URDF import inside
``LeggedRobot._create_envs`` / ``_process_dof_props``
(legged_gym/legged_gym/envs/base/legged_robot.py:344-372, 725-815): rigid-body
and DOF names (Isaac Gym orders siblings alphabetically and collapses fixed
joints unless ``dont_collapse``), and the URDF position / velocity / effort
limits.  PhysX is outside the scope of this build, so the same facts are
tabulated here from the URDFs under ``legged_gym/resources/robots`` (SURVEY.md
App. C).  Body *indices* are then derived from the cfg's name filters exactly the
way the reference does (substring match, cfg list order).
"""
from dataclasses import dataclass, field
from typing import Dict, List, Tuple


@dataclass
class RobotSpec:
    name: str
    body_names: List[str]
    dof_names: List[str]
    # URDF <limit lower upper velocity effort>, per DOF
    dof_lower: List[float]
    dof_upper: List[float]
    dof_velocity: List[float]
    dof_effort: List[float]
    # nominal foot offsets in the base frame, used only by the synthetic-state generator
    foot_offsets: List[Tuple[float, float, float]] = field(default_factory=list)

    @property
    def num_dof(self) -> int:
        return len(self.dof_names)

    @property
    def num_bodies(self) -> int:
        return len(self.body_names)

    def indices_matching(self, patterns) -> List[int]:
        """Same selection rule as the reference: for each pattern in cfg order, every body whose
        name contains it, in body order (legged_robot.py:764-770, 801-815)."""
        if isinstance(patterns, str):
            patterns = [patterns]
        out: List[int] = []
        for pat in patterns:
            out.extend(i for i, n in enumerate(self.body_names) if pat in n)
        return out


def _quadruped(name, legs, segs, dof_suffix, lower, upper, vel, eff, extra_bodies=(), foot_xy=(0.35, 0.2), foot_z=-0.5):
    bodies = ["base"]
    for leg in legs:
        bodies.extend(f"{leg}_{s}" for s in segs)
    bodies.extend(extra_bodies)
    bodies = ["base"] + sorted(bodies[1:], key=lambda n: (n.split("_")[0], 0))  # stable: groups alphabetical by prefix
    dofs, lo, up, ve, ef = [], [], [], [], []
    for leg in sorted(legs):
        for k, suf in enumerate(dof_suffix):
            dofs.append(f"{leg}_{suf}")
            lo.append(lower[k] if not callable(lower[k]) else lower[k](leg))
            up.append(upper[k] if not callable(upper[k]) else upper[k](leg))
            ve.append(vel[k])
            ef.append(eff[k])
    offs = []
    for leg in sorted(legs):
        sx = 1.0 if leg[-2] == "F" or leg[0:2] in ("LF", "RF") else -1.0
        sy = 1.0 if leg[0] == "L" or leg[1] == "L" else -1.0
        offs.append((sx * foot_xy[0], sy * foot_xy[1], foot_z))
    return RobotSpec(name, bodies, dofs, lo, up, ve, ef, offs)


def anymal_c() -> RobotSpec:
    # anymal_c.urdf: revolute joints carry effort 80 / velocity 20 and no position range (-> 0, 0)
    return _quadruped("anymal_c", ["LF", "LH", "RF", "RH"], ["HIP", "THIGH", "SHANK", "FOOT"],
                      ["HAA", "HFE", "KFE"], [0.0] * 3, [0.0] * 3, [20.0] * 3, [80.0] * 3,
                      foot_xy=(0.44, 0.26), foot_z=-0.55)


def a1() -> RobotSpec:
    return _quadruped("a1", ["FL", "FR", "RL", "RR"], ["hip", "thigh", "calf", "foot"],
                      ["hip_joint", "thigh_joint", "calf_joint"],
                      [-0.802851455917, -1.0471975512, -2.69653369433],
                      [0.802851455917, 4.18879020479, -0.916297857297],
                      [52.4, 28.6, 28.6], [20.0, 55.0, 55.0], foot_xy=(0.18, 0.13), foot_z=-0.32)


def go2() -> RobotSpec:
    rear = lambda leg: leg[0] == "R"
    spec = _quadruped("go2", ["FL", "FR", "RL", "RR"], ["hip", "thigh", "calf", "foot"],
                      ["hip_joint", "thigh_joint", "calf_joint"],
                      [-1.0472, lambda l: -0.5236 if rear(l) else -1.5708, -2.7227],
                      [1.0472, lambda l: 4.5379 if rear(l) else 3.4907, -0.83776],
                      [30.1, 30.1, 15.70], [23.7, 23.7, 45.43],
                      extra_bodies=("Head_upper", "Head_lower"), foot_xy=(0.19, 0.14), foot_z=-0.3)
    return spec


def elspider_air() -> RobotSpec:
    """el_mini.urdf (resources/robots/el_mini/urdf): hexapod, 6 x (HAA, HFE, KFE); the fixed base -> trunk joint keeps ``trunk`` as
    the body the config terminates on.  Legs in alphabetical order LB LF LM RB RF RM (the foot order ElSpider._reward_gait_2_step
    documents, elspider.py:366)."""
    legs = ["LB", "LF", "LM", "RB", "RF", "RM"]
    bodies = ["base", "trunk"] + [f"{leg}_{seg}" for leg in legs for seg in ("HIP", "THIGH", "SHANK", "FOOT")]
    dofs = [f"{leg}_{j}" for leg in legs for j in ("HAA", "HFE", "KFE")]
    lo, up = [-0.785, -0.5233, -0.6978] * 6, [0.785, 3.14, 3.925] * 6
    x = {"F": 0.3, "M": 0.0, "B": -0.3}
    offs = [(x[leg[1]], (0.35 if leg[0] == "L" else -0.35) * (1.2 if leg[1] == "M" else 1.0), -0.25) for leg in legs]
    return RobotSpec("elspider_air", bodies, dofs, lo, up, [21.0] * 18, [33.5] * 18, offs)


ROBOTS: Dict[str, callable] = {"anymal_c": anymal_c, "a1": a1, "go2": go2, "elspider_air": elspider_air, "el_mini": elspider_air,
                               "elspider": elspider_air}      # (asset.name of the hexapod's main / rollout configs)


def get_robot_spec(name: str) -> RobotSpec:
    for key, fn in ROBOTS.items():
        if name.startswith(key) or key in name:
            return fn()
    raise KeyError(f"no RobotSpec tabulated for asset '{name}'")
