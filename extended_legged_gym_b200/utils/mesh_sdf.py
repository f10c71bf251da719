"""Signed distance queries against a triangle mesh -- host side of the B200 path.

Same public surface as the reference module (utils/mesh_sdf.py in /root/reference/legged_gym/legged_gym):
``MeshSDFCfg``, ``MeshSDFData``, ``MeshSDF.query`` (:230-314), ``MeshSDF.nearest_points`` (:316-336),
``MeshSDF.clear_cache``.  The reference ships every query to the host, through a Warp kernel and back (:268-293);
here the points stay on the GPU and ``query`` is one launch of ``elg_sdf_query`` on the mesh's BVH.
"""
import os
from dataclasses import dataclass, field
from typing import List, Tuple

import torch

from .. import _lib
from .ray_caster import Mesh, load_obj


@dataclass
class MeshSDFCfg:
    mesh_paths: List[str] = field(default_factory=list)
    vertices: torch.Tensor = None
    triangles: torch.Tensor = None
    default_sdf_value: float = 1000.0
    max_distance: float = 100.0
    enable_caching: bool = False


@dataclass
class MeshSDFData:
    sdf_values: torch.Tensor = None
    sdf_gradients: torch.Tensor = None


class MeshSDF:
    EPSILON = 1.0e-3        # the literal the reference passes to the kernel (:284)

    def __init__(self, cfg: MeshSDFCfg, device: str = "cuda:0"):
        self.cfg, self.device = cfg, device
        self._cache = {}
        self.meshes = {}
        self._data = MeshSDFData()
        self._is_initialized = False
        self._initialize()

    def _initialize(self):
        if self._is_initialized:
            return
        if self.cfg.mesh_paths:
            for path in self.cfg.mesh_paths:
                if not os.path.isfile(path):
                    print(f"Failed to load mesh {path}: file not found")
                    continue
                v, f = load_obj(path)
                self.meshes[path] = Mesh(v, f, self.device, leaf_triangles=4)
        elif self.cfg.vertices is not None and self.cfg.triangles is not None:
            self.meshes["custom_mesh"] = Mesh(self.cfg.vertices, self.cfg.triangles, self.device, leaf_triangles=4)
        else:
            raise ValueError("No mesh paths or vertices/triangles provided for SDF calculation.")
        if not self.meshes:
            raise RuntimeError("No meshes were successfully loaded or created.")
        self._is_initialized = True

    def _run(self, points, want_closest):
        rank = points.dim()
        if rank not in (2, 3):
            raise ValueError(f"Expected points to have rank 2 or 3, got {rank}")
        if not points.is_cuda:
            raise _lib.ElgError("MeshSDF has no CPU path: points must be a CUDA tensor")
        flat = points.reshape(-1, 3)
        flat = flat if (flat.dtype == torch.float and flat.is_contiguous()) else flat.to(torch.float).contiguous()
        n = flat.shape[0]
        sdf = torch.empty(n, dtype=torch.float, device=flat.device)
        grad = torch.empty(n, 3, dtype=torch.float, device=flat.device)
        closest = torch.empty(n, 3, dtype=torch.float, device=flat.device) if want_closest else None
        mesh = next(iter(self.meshes.values()))            # only the first mesh is queried (:279)
        stream = torch.cuda.current_stream(flat.device).cuda_stream
        rc = _lib.load().elg_sdf_query(mesh.id, flat.data_ptr(), n, float(self.cfg.max_distance), self.EPSILON, sdf.data_ptr(),
                                       grad.data_ptr(), _lib.ptr(closest), None, stream)
        _lib.check(rc, "elg_sdf_query")
        if rank == 3:
            b, m = points.shape[0], points.shape[1]
            sdf, grad = sdf.reshape(b, m), grad.reshape(b, m, 3)
            closest = closest.reshape(b, m, 3) if closest is not None else None
        return sdf, grad, closest

    def query(self, points: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        key = None
        if self.cfg.enable_caching:                        # byte-string cache of the reference (:259-263): forces a D2H copy
            key = points.detach().reshape(-1, 3).cpu().numpy().tobytes()
            if key in self._cache:
                return self._cache[key]
        sdf, grad, _ = self._run(points, False)
        self._data.sdf_values, self._data.sdf_gradients = sdf, grad
        if key is not None:
            self._cache[key] = (sdf, grad)
        return sdf, grad

    def nearest_points(self, query_points: torch.Tensor) -> torch.Tensor:
        """p - sdf * grad (:316-336), from ONE query."""
        sdf, grad = self.query(query_points)
        return query_points - sdf.unsqueeze(-1) * grad

    def closest_points(self, query_points: torch.Tensor) -> torch.Tensor:
        """The closest surface points themselves, as found by the traversal (no reconstruction error)."""
        return self._run(query_points, True)[2]

    def clear_cache(self):
        self._cache.clear()

    @property
    def data(self) -> MeshSDFData:
        return self._data
