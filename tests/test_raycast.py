"""Ray casting.
not-gpu: oracle known answers from the reference's demo scripts, ray-pattern generator against the fixture generated
from the unmodified reference, ABI argument checks.  gpu: elg_raycast / elg_raycast_sensor against the float64
brute-force oracle -- hit flags and triangle-level decisions bit-exact, hit points / distances to fp32 rounding."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import mesh_oracle as mo  # noqa: E402
from extended_legged_gym_b200 import _lib  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda:0"
RTOL, ATOL = 1e-5, 1e-6


# ---------------------------------------------------------------------------------------------- CPU
def test_oracle_box_known_answer():
    """reference tests/ray_cast/test_ray_caster.py:134-151: one ray from (0,0,5) straight down at a box whose top
    face is at z = 2 must report a hit on that face."""
    v, t = mo.box_mesh()
    hits, found, dist, tri = mo.raycast_mesh(np.array([[0, 0, 5.0]], np.float32), np.array([[0, 0, -1.0]], np.float32), 100.0, v, t)
    assert found[0] and abs(hits[0, 2] - 2.0) < 1e-6 and abs(dist[0] - 3.0) < 1e-6 and tri[0] in (2, 3)
    # too short a ray: miss, hit point is the end point (ray_caster.py:88-92)
    hits, found, dist, _ = mo.raycast_mesh(np.array([[0, 0, 5.0]], np.float32), np.array([[0, 0, -1.0]], np.float32), 2.5, v, t)
    assert not found[0] and np.allclose(hits[0], [0, 0, 2.5]) and dist[0] == np.float32(2.5)
    # from inside the box: the far wall is hit (both face orientations count)
    hits, found, _, _ = mo.raycast_mesh(np.array([[0, 0, 1.5]], np.float32), np.array([[1.0, 0, 0]], np.float32), 10.0, v, t)
    assert found[0] and abs(hits[0, 0] - 1.0) < 1e-6


def test_ray_patterns_match_reference_fixture():
    from extended_legged_gym_b200.utils.ray_caster import PatternType, RayCasterPatternCfg
    sys.path.insert(0, GOLD)
    import make_mesh_golden as mg
    z = np.load(os.path.join(GOLD, "ray_patterns.npz"))
    for name, kw in mg.CASES.items():
        kw = dict(kw)
        kw["pattern_type"] = getattr(PatternType, kw["pattern_type"])
        o, d = RayCasterPatternCfg(**kw).create_pattern("cpu")
        assert np.array_equal(o.numpy(), z[name + "__origins"]), name
        assert np.allclose(d.numpy(), z[name + "__directions"], rtol=1e-6, atol=1e-7), name
        assert d.dtype == torch.float32


def test_raycast_abi_argument_checks():
    lib = _lib.load()
    assert lib.elg_raycast(None, None, None, 1, 1.0, None, None, None, None, None) == -4
    assert b"Mesh cannot be None" in lib.elg_last_error()
    h = C.c_void_p()
    v = np.zeros((3, 3), np.float32)
    t = np.array([[0, 1, 5]], np.int32)
    assert lib.elg_mesh_create(v.ctypes.data, 3, t.ctypes.data, 1, C.byref(h)) == -1      # index out of range
    assert lib.elg_mesh_create(None, 3, t.ctypes.data, 1, C.byref(h)) == -4
    assert lib.elg_mesh_free(None) == 0
    from extended_legged_gym_b200.utils.ray_caster import raycast_mesh
    with pytest.raises(ValueError):
        raycast_mesh(torch.zeros(2, 3), torch.zeros(2, 3), 1.0, None)


# ---------------------------------------------------------------------------------------------- GPU
def camera_rays(rng, n, lo, hi, zmin, zmax, down=False):
    o = np.stack([rng.uniform(lo[0], hi[0], n), rng.uniform(lo[1], hi[1], n), rng.uniform(zmin, zmax, n)], axis=1).astype(np.float32)
    if down:
        d = np.tile(np.array([[0, 0, -1.0]], np.float32), (n, 1))
    else:
        d = rng.normal(size=(n, 3)).astype(np.float32)
        d[:, 2] = -np.abs(d[:, 2]) * 0.7
        d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d.astype(np.float32)


def check_against_oracle(mesh, v, t, o, d, max_dist):
    from extended_legged_gym_b200.utils.ray_caster import raycast_mesh
    hits, found, dist = raycast_mesh(torch.from_numpy(o).to(DEV), torch.from_numpy(d).to(DEV), max_dist, mesh, return_distance=True)
    torch.cuda.synchronize()
    wh, wf, wd, _ = mo.raycast_mesh(o, d, max_dist, v, t)
    assert np.array_equal(found.cpu().numpy(), wf), f"{int((found.cpu().numpy() != wf).sum())} hit flags differ"
    assert np.array_equal(dist.cpu().numpy(), wd), "hit distances differ (expected bit-exact: same fp64 chain)"
    assert np.allclose(hits.cpu().numpy(), wh, rtol=RTOL, atol=ATOL)
    return wf


@pytest.mark.gpu
def test_box_known_answer_gpu():
    from extended_legged_gym_b200.utils.ray_caster import convert_to_warp_mesh, raycast_mesh
    v, t = mo.box_mesh()
    mesh = convert_to_warp_mesh(v, t, device="cuda")
    o = torch.tensor([[[0.0, 0.0, 5.0], [0.0, 0.0, 1.5], [5.0, 5.0, 5.0]]], device=DEV)
    d = torch.tensor([[[0.0, 0.0, -1.0], [1.0, 0.0, 0.0], [0.0, 0.0, -1.0]]], device=DEV)
    hits, found = raycast_mesh(o, d, 100.0, mesh)
    assert hits.shape == (1, 3, 3) and found.shape == (1, 3) and found.dtype == torch.bool
    assert found.cpu().tolist() == [[True, True, False]]
    assert torch.allclose(hits[0, 0].cpu(), torch.tensor([0.0, 0.0, 2.0])) and torch.allclose(hits[0, 1].cpu(), torch.tensor([1.0, 0.0, 1.5]))
    assert torch.allclose(hits[0, 2].cpu(), torch.tensor([5.0, 5.0, -95.0]))          # miss: end point at max_dist


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols,origin", [(24, 31, (0.0, 0.0)), (40, 40, (30.0, -12.5))])
def test_raycast_matches_bruteforce_on_terrain(rows, cols, origin):
    from extended_legged_gym_b200.utils.ray_caster import Mesh
    v, t, _ = mo.heightfield_mesh(rows, cols, seed=rows, origin=origin)
    mesh = Mesh(v, t, DEV)
    rng = np.random.default_rng(rows)
    lo, hi = v.min(0), v.max(0)
    for down, max_dist in ((True, 5.0), (False, 2.0), (False, 10.0)):
        o, d = camera_rays(rng, 3000, lo - 0.3, hi + 0.3, hi[2] + 0.05, hi[2] + 1.0, down)
        wf = check_against_oracle(mesh, v, t, o, d, max_dist)
        assert 0.05 < wf.mean() < 1.0 or down
    # grazing and degenerate rays: along grid lines, exactly on vertices, zero-length max_dist
    gx = (np.arange(0, rows, 3) * 0.1 + origin[0]).astype(np.float32)
    o = np.stack([gx, np.full_like(gx, origin[1] + 0.1 * 5), np.full_like(gx, 3.0)], axis=1)
    d = np.tile(np.array([[0, 0, -1.0]], np.float32), (len(gx), 1))
    check_against_oracle(mesh, v, t, o, d, 10.0)
    d2 = np.tile(np.array([[0, 1.0, 0]], np.float32), (len(gx), 1))
    o2 = o.copy()
    o2[:, 2] = 0.1
    check_against_oracle(mesh, v, t, o2, d2, 10.0)
    check_against_oracle(mesh, v, t, o, d, 0.0)


@pytest.mark.gpu
def test_raycast_on_closed_mesh_inside_and_outside():
    from extended_legged_gym_b200.utils.ray_caster import Mesh
    v, t = mo.icosphere(3)
    mesh = Mesh(v, t, DEV)
    rng = np.random.default_rng(3)
    o = rng.uniform(-2, 2, size=(4000, 3)).astype(np.float32)
    d = rng.normal(size=(4000, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    wf = check_against_oracle(mesh, v, t, o, d.astype(np.float32), 3.0)
    inside = np.linalg.norm(o, axis=1) < 0.97
    assert wf[inside].all()                      # a ray that starts inside a closed surface always hits it


@pytest.mark.gpu
@pytest.mark.parametrize("yaw_only", [True, False])
def test_ray_caster_sensor_matches_oracle(yaw_only):
    from extended_legged_gym_b200.utils.ray_caster import PatternType, RayCaster, RayCasterCfg, RayCasterPatternCfg
    v, t, _ = mo.heightfield_mesh(50, 50, seed=7)
    n_env = 37
    cfg = RayCasterCfg(pattern_cfg=RayCasterPatternCfg(pattern_type=PatternType.SPHERICAL2, spherical2_num_points=48),
                       vertices=torch.from_numpy(v), triangles=torch.from_numpy(t), max_distance=3.0, attach_yaw_only=yaw_only,
                       offset_pos=[0.1, 0.0, 0.05])
    rc = RayCaster(cfg, n_env, DEV)
    assert rc.num_rays == 48 and rc.ray_origins.shape == (n_env, 48, 3)
    g = torch.Generator().manual_seed(0)
    pos = torch.stack([torch.rand(n_env, generator=g) * 4.5 + 0.2, torch.rand(n_env, generator=g) * 4.5 + 0.2,
                       torch.rand(n_env, generator=g) * 0.5 + 0.6], dim=1)
    quat = torch.randn(n_env, 4, generator=g)
    quat = quat / quat.norm(dim=1, keepdim=True)
    rc.update(0.02, pos.to(DEV), quat.to(DEV))
    torch.cuda.synchronize()
    o, d = mo.sensor_rays(rc._pattern_origins.cpu().numpy(), rc._pattern_directions.cpu().numpy(), pos.numpy(), quat.numpy(), yaw_only)
    wh, wf, _, _ = mo.raycast_mesh(o.reshape(-1, 3), d.reshape(-1, 3), 3.0, v, t)
    assert np.array_equal(rc.data.ray_hits_found.cpu().numpy().reshape(-1), wf)
    assert np.allclose(rc.data.ray_hits.cpu().numpy().reshape(-1, 3), wh, rtol=RTOL, atol=ATOL)
    assert torch.equal(rc.data.pos.cpu(), pos) and torch.equal(rc.data.rot.cpu(), quat)
    # partial update: only the listed envs move
    before = rc.data.ray_hits.clone()
    ids = torch.tensor([3, 11, 20], device=DEV)
    pos2 = pos.clone()
    pos2[:, 2] += 0.3
    rc.update(0.02, pos2.to(DEV), quat.to(DEV), env_ids=ids)
    torch.cuda.synchronize()
    changed = (rc.data.ray_hits != before).any(dim=2).any(dim=1).cpu()
    assert set(changed.nonzero().flatten().tolist()) <= {3, 11, 20} and bool(changed[3])


@pytest.mark.gpu
def test_raycast_large_terrain_properties():
    """250 k-triangle terrain, 400 k downward rays: every ray over the terrain hits, the hit height lies inside the
    height range of the terrain, and a random subset agrees with the brute force."""
    from extended_legged_gym_b200.utils.ray_caster import Mesh, raycast_mesh
    v, t, _ = mo.heightfield_mesh(354, 354, seed=1)
    mesh = Mesh(v, t, DEV)
    assert mesh.num_triangles == 2 * 353 * 353
    n = 400_000
    g = torch.Generator(device=DEV).manual_seed(0)
    xy = torch.rand(n, 2, device=DEV, generator=g) * 35.0 + 0.1
    o = torch.cat([xy, torch.full((n, 1), 2.0, device=DEV)], dim=1)
    d = torch.tensor([0.0, 0.0, -1.0], device=DEV).repeat(n, 1)
    hits, found = raycast_mesh(o, d, 5.0, mesh)
    assert bool(found.all())
    assert float(hits[:, 2].min()) >= float(v[:, 2].min()) - 1e-6 and float(hits[:, 2].max()) <= float(v[:, 2].max()) + 1e-6
    assert torch.equal(hits[:, :2], o[:, :2])
    sel = torch.randperm(n, generator=torch.Generator().manual_seed(1))[:300]
    # brute force only against the triangles near each ray would need the BVH; use the full mesh on a small subset
    wh, wf, _, _ = mo.raycast_mesh(o[sel].cpu().numpy(), d[sel].cpu().numpy(), 5.0, v, t, chunk=8)
    assert wf.all() and np.allclose(hits[sel].cpu().numpy(), wh, rtol=RTOL, atol=ATOL)


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["terrain", "terrain_offset", "confined_two_layer"])
def test_grid_walk_is_bit_identical_to_the_bvh_walk(which):
    """height-field-derived meshes carry a regular-grid accelerator (opt-in, elg_set_mesh_tuning(1): measured slower than the BVH):
    same hit flags, same distances, same triangle ids as the default BVH walk, and both equal the float64 brute force"""
    from extended_legged_gym_b200 import _lib, synthetic
    from extended_legged_gym_b200.utils.ray_caster import Mesh
    lib = _lib.load()
    if which == "confined_two_layer":
        z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sensor_envs.npz"))
        v, t = z["confined_a__vertices"], z["confined_a__triangles"]
        layers = 2
    else:
        hf = synthetic.make_height_field(rows=90, cols=70, border=10, tile=25, seed=4)
        v, t = synthetic.heightfield_to_trimesh(hf, 0.1, 0.005, 1.0 if which == "terrain" else -37.3)
        layers = 1
    mesh = Mesh(v, t, DEV)
    assert mesh.grid[0] == layers, f"grid accelerator not detected: {mesh.grid}"
    rng = np.random.default_rng(11)
    lo, hi = v.min(0), v.max(0)
    n = 6000
    o = np.stack([rng.uniform(lo[0] - 0.5, hi[0] + 0.5, n), rng.uniform(lo[1] - 0.5, hi[1] + 0.5, n), rng.uniform(lo[2] - 0.2, hi[2] + 0.8, n)], axis=1).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    # special rays: straight down / up, along grid lines (x and y), exactly through vertices, horizontal, starting outside the grid
    k = n // 6
    d[:k] = [0, 0, -1]
    d[k:k + 200] = [0, 0, 1]
    xs, ys = np.unique(v[:, 0]), np.unique(v[:, 1])
    o[k + 200:k + 400, 0] = rng.choice(xs, 200)              # on a grid line x = const ...
    d[k + 200:k + 300] = [0, 1, 0]                           # ... travelling along it
    d[k + 300:k + 400] = [0, 0, -1]                          # ... or dropping onto it
    o[k + 400:k + 600, 0], o[k + 400:k + 600, 1] = rng.choice(xs, 200), rng.choice(ys, 200)      # exactly above vertices
    d[k + 400:k + 600] = [0, 0, -1]
    o[k + 600:k + 800, 1] = rng.choice(ys, 200)
    d[k + 600:k + 800] = [1, 0, 0]
    d[k + 800:k + 1000, 2] = 0.0
    d[k + 800:k + 1000] /= np.linalg.norm(d[k + 800:k + 1000], axis=1, keepdims=True) + 1e-12
    O, D = torch.from_numpy(o).to(DEV), torch.from_numpy(d.astype(np.float32)).to(DEV)
    for max_dist in (0.7, 3.0, 25.0):
        out = {}
        for mode in (0, 1):
            lib.elg_set_mesh_tuning(mode)
            hits = torch.empty(n, 3, device=DEV)
            found = torch.empty(n, dtype=torch.bool, device=DEV)
            dist = torch.empty(n, device=DEV)
            tri = torch.empty(n, dtype=torch.int32, device=DEV)
            _lib.check(lib.elg_raycast(mesh.id, O.data_ptr(), D.data_ptr(), n, max_dist, hits.data_ptr(), found.data_ptr(), dist.data_ptr(),
                                       tri.data_ptr(), None))
            torch.cuda.synchronize()
            out[mode] = (hits.cpu(), found.cpu(), dist.cpu(), tri.cpu())
        lib.elg_set_mesh_tuning(0)
        for a, b, name in zip(out[0], out[1], ("hits", "found", "distance", "triangle")):
            assert torch.equal(a, b), f"{which} max_dist={max_dist}: grid walk and BVH walk differ in {name} ({int((a != b).sum())} entries)"
        wh, wf, wd, wt = mo.raycast_mesh(o, d.astype(np.float32), max_dist, v, t)
        assert np.array_equal(out[0][1].numpy(), wf) and np.array_equal(out[0][2].numpy(), wd)
        assert 0.02 < wf.mean() < 0.99
