"""Generate tests/golden/ray_patterns.npz from the UNMODIFIED reference (container only):
RayCasterPatternCfg.create_pattern (utils/ray_caster.py:205-363) for every pattern type, on the CPU.

    python tests/golden/make_mesh_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

CASES = {
    "single": dict(pattern_type="SINGLE_RAY", single_ray_direction=[0.0, 0.0, -1.0]),
    "grid": dict(pattern_type="GRID", grid_dims=(4, 7), grid_width=1.5, grid_height=0.8),
    "cone": dict(pattern_type="CONE", cone_num_rays=12, cone_angle=25.0),
    "spherical": dict(pattern_type="SPHERICAL", spherical_num_azimuth=9, spherical_num_elevation=5),
    "spherical2": dict(pattern_type="SPHERICAL2", spherical2_num_points=40),
    "spherical2_axis": dict(pattern_type="SPHERICAL2", spherical2_num_points=24, spherical2_polar_axis=[1.0, 0.0, 0.0],
                            ellipsoid_axes=[1.0, 0.6, 0.4]),
}


def main():
    ref_harness.install()
    from legged_gym.envs.base.legged_robot import LeggedRobot  # noqa: F401  (resolves the package's import cycle first)
    from legged_gym.utils.ray_caster import PatternType, RayCasterPatternCfg
    out = {}
    for name, kw in CASES.items():
        kw = dict(kw)
        kw["pattern_type"] = getattr(PatternType, kw["pattern_type"])
        o, d = RayCasterPatternCfg(**kw).create_pattern("cpu")
        out[name + "__origins"] = o.numpy()
        out[name + "__directions"] = d.numpy()
    path = os.path.join(ROOT, "tests", "golden", "ray_patterns.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
