"""Mesh signed distance.
not-gpu: oracle known answers from the reference's demo (tests/mesh_sdf/test_mesh_sdf.py:22-94: unit icosphere, origin
'should be close to -1.0', inside / outside counts).  gpu: elg_sdf_query through MeshSDF against the float64 brute
force -- signs and closest faces identical, distances / gradients to fp32 rounding."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import mesh_oracle as mo  # noqa: E402

DEV = "cuda:0"
RTOL, ATOL = 1e-5, 1e-6


def test_oracle_icosphere_known_answers():
    v, t = mo.icosphere(3)
    sdf, grad, closest, face = mo.sdf_query(np.array([[0, 0, 0], [0, 0, 2.0], [0, 0, 0.5], [5, 5, 5]], np.float32), 3.0, v, t)
    assert -1.0 < sdf[0] < -0.97                                   # faces of an inscribed icosphere lie inside the unit sphere
    assert abs(sdf[1] - 1.0) < 2e-3 and np.allclose(grad[1], [0, 0, 1], atol=2e-2)
    assert -0.5 <= sdf[2] < -0.49 and grad[2, 2] > 0.99            # inside: gradient still points outward
    assert sdf[3] == np.float32(3.0) and not grad[3].any() and face[3] == -1      # nothing within max_distance
    rng = np.random.default_rng(0)
    p = rng.uniform(-1.5, 1.5, size=(2000, 3)).astype(np.float32)
    s, g, _, _ = mo.sdf_query(p, 10.0, v, t)
    r = np.linalg.norm(p, axis=1)
    clear = np.abs(r - 1.0) > 0.02
    assert np.array_equal((s < 0)[clear], (r < 1.0)[clear])
    assert np.allclose(np.abs(s), np.abs(r - 1.0), atol=0.02)
    assert np.allclose(np.linalg.norm(g, axis=1), 1.0, atol=1e-5)


def sdf_check(v, t, pts, max_distance, enable_caching=False):
    from extended_legged_gym_b200.utils.mesh_sdf import MeshSDF, MeshSDFCfg
    m = MeshSDF(MeshSDFCfg(vertices=torch.from_numpy(v), triangles=torch.from_numpy(t), max_distance=max_distance, enable_caching=enable_caching), DEV)
    P = torch.from_numpy(pts).to(DEV)
    sdf, grad = m.query(P)
    closest = m.closest_points(P)
    torch.cuda.synchronize()
    ws, wg, wc, wf = mo.sdf_query(pts.reshape(-1, 3), max_distance, v, t)
    s, g, c = sdf.cpu().numpy().reshape(-1), grad.cpu().numpy().reshape(-1, 3), closest.cpu().numpy().reshape(-1, 3)
    assert np.array_equal(np.signbit(s), np.signbit(ws)), f"{int((np.signbit(s) != np.signbit(ws)).sum())} signs differ"
    assert np.allclose(s, ws, rtol=RTOL, atol=ATOL)
    assert np.allclose(g, wg, rtol=1e-5, atol=2e-6)
    assert np.allclose(c, wc, rtol=RTOL, atol=ATOL)
    return m, sdf, grad, ws


@pytest.mark.gpu
def test_sdf_matches_bruteforce_icosphere():
    v, t = mo.icosphere(3)
    rng = np.random.default_rng(1)
    pts = rng.uniform(-1.6, 1.6, size=(5000, 3)).astype(np.float32)
    pts[:5] = [[0, 0, 0], [0, 0, 1], [1, 0, 0], [0, 0, 2.0], [0.3, 0.3, 0.3]]
    m, sdf, grad, ws = sdf_check(v, t, pts, 100.0)
    assert -1.0 < float(sdf[0]) < -0.97
    # rank-3 input, nearest_points = p - sdf * grad lies on the surface
    P3 = torch.from_numpy(pts[:600].reshape(6, 100, 3)).to(DEV)
    s3, g3 = m.query(P3)
    assert s3.shape == (6, 100) and g3.shape == (6, 100, 3)
    assert torch.equal(s3.reshape(-1), sdf[:600])
    near = m.nearest_points(P3)
    again, _ = m.query(near)
    assert float(again.abs().max()) < 1e-4      # fp32 reconstruction p - sdf * grad


@pytest.mark.gpu
def test_sdf_matches_bruteforce_box_edges_and_misses():
    """closest features on edges / corners of a box: the sign comes from the most aligned face"""
    v, t = mo.box_mesh((-1, -1, -1), (1, 1, 1))
    rng = np.random.default_rng(2)
    pts = rng.uniform(-2.5, 2.5, size=(4000, 3)).astype(np.float32)
    pts[:6] = [[2, 2, 2], [1.5, 1.5, 0], [0, 0, 0], [0.999, 0.999, 0.999], [1, 1, 1], [0, 0, 1]]
    m, sdf, grad, ws = sdf_check(v, t, pts, 1.2)
    assert float(sdf[2]) == -1.0 and float(sdf[0]) == pytest.approx(1.2)      # centre: -1; far corner: clamped to max_distance
    inside = (np.abs(pts) < 0.98).all(axis=1)
    outside = (np.abs(pts) > 1.02).any(axis=1) & (ws < 1.2)
    assert (ws[inside] < 0).all() and (ws[outside] > 0).all()


@pytest.mark.gpu
def test_sdf_matches_bruteforce_on_open_terrain_and_cache():
    v, t, _ = mo.heightfield_mesh(40, 40, seed=5, origin=(20.0, 20.0))
    rng = np.random.default_rng(3)
    xy = rng.uniform(19.8, 24.1, size=(3000, 2))
    z = rng.uniform(-0.3, 0.9, size=(3000, 1))
    pts = np.concatenate([xy, z], axis=1).astype(np.float32)
    m, sdf, grad, ws = sdf_check(v, t, pts, 0.5, enable_caching=True)
    assert 0.1 < (ws < 0).mean() < 0.9 and (ws == np.float32(0.5)).any()
    s2, g2 = m.query(torch.from_numpy(pts).to(DEV))              # served from the byte-string cache
    assert s2 is sdf and len(m._cache) == 1
    m.clear_cache()
    assert not m._cache
