"""TEST INFRASTRUCTURE ONLY (oracle) -- the checker, never the product path.

torch-CPU restatements of the observation glue of the sensor-carrying env classes of the reference
(/root/reference/legged_gym/legged_gym):

  raycast_distances   envs/base/legged_robot_raycast.py:262-297  LeggedRobotRayCast._get_raycast_distances
  sdf_query_points    envs/batch_rollout/robot_batch_rollout_percept.py:385-441  the point construction of _update_sdf_values
  nearest_points      utils/mesh_sdf.py:316-336  MeshSDF.nearest_points

``tests/test_sensor_envs.py`` pins them to the unmodified reference methods bound to a synthetic ``self`` (container only) and to
``tests/golden/sensor_envs.npz``.  The mesh queries underneath (Warp in the reference) are the brute force of
``oracle/mesh_oracle.py`` -- parity unpinned, see its header.
"""
import torch

from . import torch_utils as tu


def raycast_distances(ray_hits, ray_hits_found, root_pos, max_distance, normalize=True):
    """(:278-297) distances are measured from the robot base, not from the ray origin; a miss reads 0 after normalisation."""
    distances = torch.norm(ray_hits - root_pos.unsqueeze(1), dim=2)
    if not normalize:
        return distances
    nd = 1.0 - torch.clamp(distances / max_distance, 0.0, 1.0)
    nd = nd * ray_hits_found.float()
    return nd.reshape(nd.shape[0], -1)


def sdf_query_points(rigid_body_state, num_bodies, body_indices, sphere_offsets=None):
    """(:396-414) world positions of the query spheres: body position + quat_rotate(body quaternion, sphere offset).
    rigid_body_state [N * B, 13]; returns [N, len(body_indices), 3]."""
    rbs = rigid_body_state.view(-1, num_bodies, 13)
    out = []
    for k, b in enumerate(body_indices):
        pos, quat = rbs[:, b, 0:3], rbs[:, b, 3:7]
        if sphere_offsets is not None:
            off = sphere_offsets[k].unsqueeze(0).expand(pos.shape[0], 3)
            pos = pos + tu.quat_rotate(quat, off)
        out.append(pos)
    return torch.stack(out, dim=1)


def nearest_points(points, sdf, grad):
    """MeshSDF.nearest_points (mesh_sdf.py:316-336): p - sdf * gradient"""
    return points - sdf.unsqueeze(-1) * grad
