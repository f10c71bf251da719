"""TEST INFRASTRUCTURE ONLY (oracle) -- the checker, never the product path.  PARITY UNPINNED for the Warp parts.

float64 brute-force restatement of the mesh queries the reference delegates to warp-lang==1.7.0
(legged_gym/setup.py:17; the wheel is absent from /root/reference and from this image, so its kernels cannot be run):

  raycast_mesh        utils/ray_caster.py:45-167   wp.mesh_query_ray: closest hit with t in [0, max_dist), both face
                                                    orientations; miss -> end point origin + max_dist * direction
  sdf_query           utils/mesh_sdf.py:38-116     wp.mesh_query_point_sign_normal + mesh_eval_position
  ray caster / depth camera host arithmetic        utils/ray_caster.py:558-594, utils/depth_camera.py:402-566 (these are
                                                    torch code of the reference itself and are restated op by op)

Every float64 expression below is a chain of individually rounded elementwise numpy operations, in the order the
CUDA kernels use (csrc/elg_mesh.cu), so kernel and oracle agree to the last bit of the fp32 results.  The informal
known answers of the reference's demo scripts (tests/ray_cast/test_ray_caster.py:134-151 box mesh; tests/mesh_sdf/
test_mesh_sdf.py:42-47 icosphere) are pinned in tests/test_mesh_oracle.py.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.
"""
import numpy as np


def raycast_mesh(origins, directions, max_dist, vertices, triangles, chunk=256):
    """origins, directions: [R,3] float32; vertices [V,3] float32; triangles [M,3] int.
    Returns hits [R,3] float32, found [R] bool, t [R] float32 (max_dist on a miss), tri [R] int (-1 on a miss)."""
    o32 = np.ascontiguousarray(origins, dtype=np.float32).reshape(-1, 3)
    d32 = np.ascontiguousarray(directions, dtype=np.float32).reshape(-1, 3)
    V = np.asarray(vertices, dtype=np.float32).astype(np.float64)
    T = np.asarray(triangles).astype(np.int64)
    v0, v1, v2 = V[T[:, 0]], V[T[:, 1]], V[T[:, 2]]
    e1, e2 = v1 - v0, v2 - v0
    R = o32.shape[0]
    t_best = np.full(R, np.float64(np.float32(max_dist)))
    tri = np.full(R, -1, dtype=np.int64)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        for a in range(0, R, chunk):
            o = o32[a:a + chunk].astype(np.float64)[:, None, :]
            d = d32[a:a + chunk].astype(np.float64)[:, None, :]
            dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]
            e1x, e1y, e1z = e1[None, :, 0], e1[None, :, 1], e1[None, :, 2]
            e2x, e2y, e2z = e2[None, :, 0], e2[None, :, 1], e2[None, :, 2]
            px = dy * e2z - dz * e2y
            py = dz * e2x - dx * e2z
            pz = dx * e2y - dy * e2x
            det = (e1x * px + e1y * py) + e1z * pz
            inv = 1.0 / det
            sx, sy, sz = o[..., 0] - v0[None, :, 0], o[..., 1] - v0[None, :, 1], o[..., 2] - v0[None, :, 2]
            u = ((sx * px + sy * py) + sz * pz) * inv
            qx = sy * e1z - sz * e1y
            qy = sz * e1x - sx * e1z
            qz = sx * e1y - sy * e1x
            v = ((dx * qx + dy * qy) + dz * qz) * inv
            t = ((e2x * qx + e2y * qy) + e2z * qz) * inv
            ok = (det != 0.0) & (u >= 0.0) & (u <= 1.0) & (v >= 0.0) & ((u + v) <= 1.0) & (t >= 0.0) & (t < t_best[a:a + chunk, None])
            tt = np.where(ok, t, np.inf)
            k = np.argmin(tt, axis=1)
            tb = tt[np.arange(tt.shape[0]), k]
            hit = np.isfinite(tb)
            t_best[a:a + chunk] = np.where(hit, tb, t_best[a:a + chunk])
            tri[a:a + chunk] = np.where(hit, k, -1)
    found = tri >= 0
    t32 = t_best.astype(np.float32)
    hits = o32 + t32[:, None] * d32          # fp32 multiply, fp32 add (ray_caster.py:86)
    return hits.astype(np.float32), found, t32, tri


# ----------------------------------------------------------------------------------------------
# isaacgym.torch_utils.quat_apply / math_utils.quat_apply_yaw in fp32 numpy, one rounding per torch op
# ----------------------------------------------------------------------------------------------
def _cross(a, b):
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1).astype(np.float32)


def quat_apply(q, b):
    q = np.asarray(q, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    xyz, w = q[..., :3], q[..., 3:4]
    t = (_cross(xyz, b) * np.float32(2.0)).astype(np.float32)
    return ((b + w * t).astype(np.float32) + _cross(xyz, t)).astype(np.float32)


def quat_apply_yaw(q, b):
    q = np.array(q, dtype=np.float32, copy=True)
    q[..., :2] = 0.0
    n = np.sqrt((q[..., 2] * q[..., 2] + q[..., 3] * q[..., 3]).astype(np.float32)).astype(np.float32)   # x, y are 0
    q = (q / np.maximum(n, np.float32(1e-9))[..., None]).astype(np.float32)
    return quat_apply(q, b)


def sensor_rays(pattern_origins, pattern_dirs, sensor_pos, sensor_quat, yaw_only):
    """RayCaster._update_ray_casting (utils/ray_caster.py:566-579): world rays of every (sensor, pattern ray)."""
    N, n = sensor_pos.shape[0], pattern_dirs.shape[0]
    q = np.repeat(np.asarray(sensor_quat, np.float32)[:, None, :], n, axis=1)
    po = np.broadcast_to(np.asarray(pattern_origins, np.float32)[None], (N, n, 3))
    pd = np.broadcast_to(np.asarray(pattern_dirs, np.float32)[None], (N, n, 3))
    rot = quat_apply_yaw if yaw_only else quat_apply
    o = (rot(q, po) + np.asarray(sensor_pos, np.float32)[:, None, :]).astype(np.float32)
    return o, rot(q, pd)


# ----------------------------------------------------------------------------------------------
# fixture meshes
# ----------------------------------------------------------------------------------------------
def box_mesh(lo=(-1.0, -1.0, 1.0), hi=(1.0, 1.0, 2.0)):
    """8-vertex box like the one in the reference's tests/ray_cast/test_ray_caster.py:96-130."""
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    v = np.array([[x0, y0, z0], [x1, y0, z0], [x1, y1, z0], [x0, y1, z0], [x0, y0, z1], [x1, y0, z1], [x1, y1, z1], [x0, y1, z1]], np.float32)
    # outward-facing winding: bottom, top, front (y0), right (x1), back (y1), left (x0)
    t = np.array([[0, 2, 1], [0, 3, 2], [4, 5, 6], [4, 6, 7], [0, 1, 5], [0, 5, 4], [1, 2, 6], [1, 6, 5], [2, 3, 7], [2, 7, 6],
                  [3, 0, 4], [3, 4, 7]], np.int32)
    return v, t


def heightfield_mesh(rows, cols, hscale=0.1, vscale=0.005, seed=0, origin=(0.0, 0.0)):
    """Grid mesh with the triangulation of isaacgym.terrain_utils.convert_heightfield_to_trimesh (two triangles per
    cell, vertex (i, j) at (i * hscale, j * hscale, h * vscale)) over a seeded random int16 field of slopes and steps."""
    rng = np.random.default_rng(seed)
    ii, jj = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    h = (40 * np.sin(ii / 7.0) * np.cos(jj / 5.0) + rng.integers(-6, 7, size=(rows, cols)) + 30 * ((ii // 8 + jj // 8) % 2)).astype(np.int16)
    v = np.stack([ii * hscale + origin[0], jj * hscale + origin[1], h * vscale], axis=-1).reshape(-1, 3).astype(np.float32)
    idx = (ii * cols + jj)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, 1:].ravel()
    t = np.concatenate([np.stack([a, d, c], axis=1), np.stack([a, b, d], axis=1)], axis=0).astype(np.int32)
    return v, t, h


def icosphere(subdivisions=2, radius=1.0):
    """Unit icosphere (tests/mesh_sdf/test_mesh_sdf.py:22-30 uses trimesh.creation.icosphere): outward-facing."""
    phi = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, phi, 0), (1, phi, 0), (-1, -phi, 0), (1, -phi, 0), (0, -1, phi), (0, 1, phi), (0, -1, -phi), (0, 1, -phi),
         (phi, 0, -1), (phi, 0, 1), (-phi, 0, -1), (-phi, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdivisions):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[key] = len(v) - 1
            return cache[key]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return (np.array(v) * radius).astype(np.float32), np.array(f, np.int32)


# ----------------------------------------------------------------------------------------------
# signed distance (utils/mesh_sdf.py:38-116 on top of wp.mesh_query_point_sign_normal) -- PARITY UNPINNED
# ----------------------------------------------------------------------------------------------
def _dot(a, b):
    return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]


def closest_on_triangles(p, a, b, c):
    """Ericson 5.1.5, vectorised: p [R,1,3], a/b/c [1,M,3] float64 -> closest points [R,M,3], squared distances [R,M]."""
    ab, ac, ap = b - a, c - a, p - a
    d1, d2 = _dot(ab, ap), _dot(ac, ap)
    bp = p - b
    d3, d4 = _dot(ab, bp), _dot(ac, bp)
    cp = p - c
    d5, d6 = _dot(ab, cp), _dot(ac, cp)
    vc = d1 * d4 - d3 * d2
    vb = d5 * d2 - d1 * d6
    va = d3 * d6 - d5 * d4
    with np.errstate(divide="ignore", invalid="ignore"):
        v_ab = (d1 / (d1 - d3))[..., None]
        w_ac = (d2 / (d2 - d6))[..., None]
        w_bc = ((d4 - d3) / ((d4 - d3) + (d5 - d6)))[..., None]
        denom = 1.0 / ((va + vb) + vc)
    v_in, w_in = (vb * denom)[..., None], (vc * denom)[..., None]
    conds = [(d1 <= 0) & (d2 <= 0), (d3 >= 0) & (d4 <= d3), (vc <= 0) & (d1 >= 0) & (d3 <= 0), (d6 >= 0) & (d5 <= d6),
             (vb <= 0) & (d2 >= 0) & (d6 <= 0), (va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0)]
    cands = [a + 0 * p, b + 0 * p, a + v_ab * ab, c + 0 * p, a + w_ac * ac, b + w_bc * (c - b)]
    q = (a + ab * v_in) + ac * w_in
    for cond, cand in zip(reversed(conds), reversed(cands)):       # the first matching region wins
        q = np.where(cond[..., None], cand, q)
    off = p - q
    return q, _dot(off, off)


def sdf_query(points, max_distance, vertices, triangles, epsilon=1.0e-3, chunk=128):
    """Returns sdf [R] f32, grad [R,3] f32, closest [R,3] f32, face [R] (-1: nothing within max_distance)."""
    P = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3).astype(np.float64)
    V = np.asarray(vertices, dtype=np.float32).astype(np.float64)
    T = np.asarray(triangles).astype(np.int64)
    a, b, c = V[T[:, 0]][None], V[T[:, 1]][None], V[T[:, 2]][None]
    e1, e2 = (b - a)[0], (c - a)[0]
    fn = np.stack([e1[:, 1] * e2[:, 2] - e1[:, 2] * e2[:, 1], e1[:, 2] * e2[:, 0] - e1[:, 0] * e2[:, 2],
                   e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]], axis=1)
    fl = np.sqrt(_dot(fn, fn))
    edges = np.concatenate([np.linalg.norm((b - a)[0], axis=1), np.linalg.norm((c - b)[0], axis=1), np.linalg.norm((a - c)[0], axis=1)])
    eps_abs = np.float64(np.float32(epsilon)) * (edges.reshape(3, -1).T.sum() / (3.0 * len(T)))
    D = np.float64(np.float32(max_distance))
    R = P.shape[0]
    sdf = np.empty(R, np.float32)
    grad = np.zeros((R, 3), np.float32)
    closest = np.zeros((R, 3), np.float32)
    face = np.full(R, -1, np.int64)
    for s in range(0, R, chunk):
        p = P[s:s + chunk, None, :]
        q, d2 = closest_on_triangles(p, a, b, c)
        # zero-area triangles (the slope correction of the terrain converters collapses some) can fall through every vertex / edge
        # region into the interior formula, whose 1 / (va + vb + vc) is then 1 / 0: such an evaluation takes no part (their
        # edges belong to non-degenerate neighbours as well); the kernel's `d2 <= best` comparison drops the NaN the same way
        d2 = np.where(np.isnan(d2), np.inf, d2)
        best = np.minimum(d2.min(axis=1), D * D)
        hit = d2.min(axis=1) <= D * D
        lim = np.sqrt(best) + eps_abs
        cand = d2 <= (lim * lim)[:, None]
        off = p - q
        dt = _dot(fn[None], off)
        with np.errstate(divide="ignore", invalid="ignore"):
            score = np.where(fl[None] > 0, np.abs(dt) / fl[None], 0.0)
        score = np.where(cand, score, -1.0)
        k = np.argmax(score, axis=1)                                 # first maximum = lowest triangle id
        ar = np.arange(len(k))
        dist = np.sqrt(d2[ar, k])
        sign = np.where(dt[ar, k] < 0, -1.0, 1.0)
        with np.errstate(divide="ignore", invalid="ignore"):
            g_off = off[ar, k] / dist[:, None]
            g_n = fn[k] * np.where(fl[k] > 0, 1.0 / fl[k], 0.0)[:, None]
        g = np.where((dist > 1.0e-6)[:, None], g_off, g_n)
        sdf[s:s + chunk] = np.where(hit, (dist * sign), D).astype(np.float32)
        grad[s:s + chunk] = np.where(hit[:, None], g * sign[:, None], 0.0).astype(np.float32)
        closest[s:s + chunk] = np.where(hit[:, None], q[ar, k], 0.0).astype(np.float32)
        face[s:s + chunk] = np.where(hit, k, -1)
    return sdf, grad, closest, face


# ----------------------------------------------------------------------------------------------
# the sign rule Warp documents for mesh_query_point_sign_normal, restated for COMPARISON with the rule above
# (tests/test_sdf_sign_rule.py): angle-weighted pseudo-normal of the closest feature (Baerentzen & Aanaes 2005) --
# face interior: the face normal; edge: sum of the unit normals of the faces sharing it; vertex: sum over the incident
# faces of (interior angle at the vertex) x (unit normal).  Not used by the product path or by the parity tests.
# ----------------------------------------------------------------------------------------------
def torus_mesh(major=1.0, minor=0.4, nu=24, nv=12):
    """closed manifold torus around the z axis: its inner ring is made of saddle vertices and saddle edges"""
    u = np.arange(nu) * (2 * np.pi / nu)
    w = np.arange(nv) * (2 * np.pi / nv)
    uu, ww = np.meshgrid(u, w, indexing="ij")
    v = np.stack([(major + minor * np.cos(ww)) * np.cos(uu), (major + minor * np.cos(ww)) * np.sin(uu), minor * np.sin(ww)], axis=-1)
    idx = np.arange(nu * nv).reshape(nu, nv)
    a, b = idx, np.roll(idx, -1, axis=0)
    c, d = np.roll(idx, -1, axis=1), np.roll(np.roll(idx, -1, axis=0), -1, axis=1)
    t = np.concatenate([np.stack([a.ravel(), b.ravel(), d.ravel()], axis=1), np.stack([a.ravel(), d.ravel(), c.ravel()], axis=1)])
    return v.reshape(-1, 3).astype(np.float32), t.astype(np.int32)


def points_inside_by_parity(points, vertices, triangles, direction=(0.3713, 0.5127, 0.7741)):
    """ground truth for a CLOSED mesh: a point is inside iff a ray from it crosses the surface an odd number of times"""
    P = np.asarray(points, np.float64).reshape(-1, 3)
    V = np.asarray(vertices, np.float32).astype(np.float64)
    T = np.asarray(triangles).astype(np.int64)
    d = np.asarray(direction, np.float64)
    d = d / np.linalg.norm(d)
    a, e1, e2 = V[T[:, 0]], V[T[:, 1]] - V[T[:, 0]], V[T[:, 2]] - V[T[:, 0]]
    pv = np.cross(d, e2)
    det = (e1 * pv).sum(-1)
    inside = np.zeros(len(P), bool)
    for i, p in enumerate(P):
        s = p - a
        with np.errstate(divide="ignore", invalid="ignore"):
            u = (s * pv).sum(-1) / det
            q = np.cross(s, e1)
            v = (q @ d) / det
            t = (e2 * q).sum(-1) / det
        inside[i] = (np.count_nonzero((u >= 0) & (v >= 0) & (u + v <= 1) & (t > 0)) & 1) == 1
    return inside


def sdf_sign_pseudonormal(points, vertices, triangles, tol=1.0e-9):
    """sign (+1 outside / -1 inside) by the angle-weighted pseudo-normal of the closest feature; exhaustive closest-point search"""
    P = np.asarray(points, np.float32).reshape(-1, 3).astype(np.float64)
    V = np.asarray(vertices, np.float32).astype(np.float64)
    T = np.asarray(triangles).astype(np.int64)
    # weld vertices that share a position (fixture meshes may duplicate them)
    _, weld = np.unique(np.round(V, 9), axis=0, return_inverse=True)
    W = weld.reshape(-1)[T]
    a, b, c = V[T[:, 0]], V[T[:, 1]], V[T[:, 2]]
    fn = np.cross(b - a, c - a)
    fl = np.linalg.norm(fn, axis=1)
    un = fn / np.where(fl > 0, fl, 1.0)[:, None]

    def angle(p, q, r):      # interior angle at p
        x, y = q - p, r - p
        cs = (x * y).sum(-1) / np.maximum(np.linalg.norm(x, axis=1) * np.linalg.norm(y, axis=1), 1e-300)
        return np.arccos(np.clip(cs, -1.0, 1.0))
    ang = np.stack([angle(a, b, c), angle(b, c, a), angle(c, a, b)], axis=1)
    nvert = int(W.max()) + 1
    vert_n = np.zeros((nvert, 3))
    for k in range(3):
        np.add.at(vert_n, W[:, k], ang[:, k:k + 1] * un)
    edge_n = {}
    for f in range(len(T)):
        for i, j in ((0, 1), (1, 2), (2, 0)):
            key = (min(W[f, i], W[f, j]), max(W[f, i], W[f, j]))
            edge_n[key] = edge_n.get(key, 0.0) + un[f]
    sign = np.ones(len(P))
    A, B, C = a[None], b[None], c[None]
    for s in range(0, len(P), 64):
        p = P[s:s + 64, None, :]
        q, d2 = closest_on_triangles(p, A, B, C)
        d2 = np.where(np.isnan(d2), np.inf, d2)
        k = np.argmin(d2, axis=1)
        for r, f in enumerate(k):
            qq = q[r, f]
            corners = (a[f], b[f], c[f])
            near = [np.linalg.norm(qq - x) <= tol * (1.0 + np.linalg.norm(x)) for x in corners]
            if any(near):
                n = vert_n[W[f, int(np.argmax(near))]]
            else:
                n = un[f]
                for i, j in ((0, 1), (1, 2), (2, 0)):
                    x, y = corners[i], corners[j]
                    e = y - x
                    tt = np.dot(qq - x, e) / np.dot(e, e)
                    if np.linalg.norm(x + tt * e - qq) <= tol * (1.0 + np.linalg.norm(qq)):
                        n = edge_n[(min(W[f, i], W[f, j]), max(W[f, i], W[f, j]))]
                        break
            sign[s + r] = -1.0 if np.dot(P[s + r] - qq, n) < 0 else 1.0
    return sign
