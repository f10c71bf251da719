"""``LeggedRobotDepth`` -- LeggedRobotRayCast that owns a depth camera, updated every ``cfg.depth.update_interval`` steps.

Mirrors envs/base/legged_robot_depthcam.py:4-192 of the reference: camera construction from the terrain mesh (:18-108; the
"IsaacGym" camera type needs the simulator's renderer and is not available here), the decimated update after the step
(:110-130: camera pose from the root state, then ``update_depth_buffer``), and the accessors (:154-192).  One update = two
launches (``elg_camera_pose`` + the fused ``elg_depth_camera``: ray grid, BVH walk, depth post-processing, resize, ring buffer).
"""
from ...utils.depth_camera import DepthCameraFake, DepthCameraWarp
from .legged_robot_raycast import LeggedRobotRayCast


class LeggedRobotDepth(LeggedRobotRayCast):
    def __init__(self, cfg, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True):
        self.depth_camera = None
        self.depth_update_counter = 0
        super().__init__(cfg, sim_params, physics_engine, sim_device, headless)
        self._create_depth_camera()

    def _create_depth_camera(self):
        """_create_envs' camera part (:24-100)"""
        kind = self.cfg.depth.camera_type
        if kind is None:
            return
        if kind == "Warp":
            v, t = self._terrain_mesh()
            self.depth_camera = DepthCameraWarp(cfg=self.cfg.depth, device=self.device, num_envs=self.num_envs,
                                                terrain_vertices=v.cpu().numpy(), terrain_triangles=t.cpu().numpy())
        elif kind == "Fake":
            self.depth_camera = DepthCameraFake(cfg=self.cfg.depth, device=self.device, num_envs=self.num_envs)
        elif kind == "IsaacGym":
            raise NotImplementedError("camera_type 'IsaacGym' renders through the simulator (outside this build): use 'Warp' (the B200 ray caster) or 'Fake'")
        else:
            print(f"Warning: Unknown camera type '{kind}'. Depth camera disabled.")

    def post_physics_step(self):
        super().post_physics_step()
        if self.cfg.depth.camera_type is not None and self.depth_camera is not None:
            if self.depth_update_counter % self.cfg.depth.update_interval == 0:
                if self.cfg.depth.camera_type == "Warp":
                    self.depth_camera.update(dt=self.dt, sensor_pos=self.root_states[:, :3], sensor_rot=self.root_states[:, 3:7])
                self.depth_camera.update_depth_buffer(None, self.episode_length_buf)
            self.depth_update_counter += 1

    def get_depth_images(self):
        if self.cfg.depth.camera_type is not None and self.depth_camera is not None:
            return self.depth_camera.get_depth_buffer()
        return None

    def get_depth_observation(self):
        if self.cfg.depth.camera_type is not None and self.depth_camera is not None:
            return self.depth_camera.get_depth_observation()
        return None

    def is_depth_enabled(self):
        return self.cfg.depth.camera_type is not None and self.depth_camera is not None and self.depth_camera.is_enabled()
