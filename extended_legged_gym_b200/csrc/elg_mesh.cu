// elg_mesh.cu -- triangle-mesh BVH, ray casting, depth camera and signed-distance queries for sm_100a.
//
// Replaces what the reference delegates to NVIDIA Warp 1.7 through a GPU -> CPU numpy -> Warp -> CPU -> GPU round
// trip every call (paths relative to legged_gym/legged_gym/ in the reference):
//   utils/ray_caster.py:45-92   raycast_mesh_kernel   (wp.mesh_query_ray)            -> elg_raycast
//   utils/ray_caster.py:29-42   convert_to_warp_mesh  (wp.Mesh: BVH build)           -> elg_mesh_create / elg_mesh_free
//   utils/depth_camera.py:402-499 DepthCameraWarp.update_depth_buffer                 -> elg_depth_camera (fused)
//   utils/mesh_sdf.py:38-116    query_sdf_kernel (wp.mesh_query_point_sign_normal)    -> elg_sdf_query
//
// Data structure: a 4-wide BVH.  The host builds a binary tree top-down with binned SAH (16 bins, leaves of <= 3
// triangles -- the terrain is static, the build runs once at init), collapses it to 4 children per node and uploads
//   nodes   128 bytes each: child boxes as SoA float4 rows (lo.x[4], lo.y[4], lo.z[4], hi.x[4], hi.y[4], hi.z[4]),
//           int4 child codes (>= 0 inner node, < 0 leaf: ~(first << 2 | count - 1), INT_MIN empty)
//   tris    48 bytes each, in leaf order: v0, v1, v2 as fp32 + the caller's triangle id
// so one node visit is eight 16-byte loads from one 128-byte line and a leaf is a contiguous run.  The default
// terrain (1.6 M triangles) is 77 MB of triangles + 17 MB of nodes: resident in the 126 MB L2 while rays stay local.
//
// Numerics: box tests are fp32 and conservative (boxes padded at build time, far plane scaled by 1 + 2^-21); the
// triangle tests run in fp64 with individually rounded operations in the same order as the numpy oracle
// (oracle/mesh_oracle.py), so hit distances agree with the float64 brute force to the last fp32 bit instead of
// carrying the ~4e-6 m cancellation error of an fp32 test at terrain coordinates of tens of metres.
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "elg_common.cuh"
#include "elg_async.cuh"

struct ElgMesh {
  int32_t num_vertices, num_triangles, num_nodes;
  float4* nodes;      // device, 8 float4 per node
  float4* tris;       // device, 3 float4 per triangle (leaf order)
  float lo[3], hi[3]; // scene bounds
  double avg_edge;    // mean edge length (mesh_query_point_sign_normal's epsilon is relative to it)
  int device;
  // regular-grid accelerator (height-field-derived meshes, one or more layers over the same xy grid): see GridView
  int grid_layers, grid_nx, grid_ny;
  float grid_x0, grid_y0, grid_dx, grid_dy, grid_pad;
  float2* grid_cellz;   // device [layers][nx-1][ny-1]: (min, max) height of the cell's corners
  int2* grid_celltri;   // device, same shape: the cell's two triangles as positions in `tris`
};

namespace elg {

constexpr int kStack = 48;
constexpr int kEmpty = INT32_MIN;

struct Hit {
  double t;
  int tri;
};

// ---------------------------------------------------------------------------------------------
// fp64 Moeller-Trumbore, two-sided, every operation rounded on its own (matches oracle/mesh_oracle.py)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }

__device__ __forceinline__ bool ray_triangle(const double ox, const double oy, const double oz, const double dx, const double dy,
                                             const double dz, const float4 a, const float4 b, const float4 c, double& t_out) {
  const double v0x = a.x, v0y = a.y, v0z = a.z, v1x = a.w, v1y = b.x, v1z = b.y, v2x = b.z, v2y = b.w, v2z = c.x;
  const double e1x = dsub(v1x, v0x), e1y = dsub(v1y, v0y), e1z = dsub(v1z, v0z);
  const double e2x = dsub(v2x, v0x), e2y = dsub(v2y, v0y), e2z = dsub(v2z, v0z);
  const double px = dsub(dmul(dy, e2z), dmul(dz, e2y));
  const double py = dsub(dmul(dz, e2x), dmul(dx, e2z));
  const double pz = dsub(dmul(dx, e2y), dmul(dy, e2x));
  const double det = dadd(dadd(dmul(e1x, px), dmul(e1y, py)), dmul(e1z, pz));
  if (det == 0.0) return false;
  const double inv = 1.0 / det;
  const double sx = dsub(ox, v0x), sy = dsub(oy, v0y), sz = dsub(oz, v0z);
  const double u = dmul(dadd(dadd(dmul(sx, px), dmul(sy, py)), dmul(sz, pz)), inv);
  if (!(u >= 0.0 && u <= 1.0)) return false;
  const double qx = dsub(dmul(sy, e1z), dmul(sz, e1y));
  const double qy = dsub(dmul(sz, e1x), dmul(sx, e1z));
  const double qz = dsub(dmul(sx, e1y), dmul(sy, e1x));
  const double v = dmul(dadd(dadd(dmul(dx, qx), dmul(dy, qy)), dmul(dz, qz)), inv);
  if (!(v >= 0.0 && dadd(u, v) <= 1.0)) return false;
  const double t = dmul(dadd(dadd(dmul(e2x, qx), dmul(e2y, qy)), dmul(e2z, qz)), inv);
  if (!(t >= 0.0)) return false;
  t_out = t;
  return true;
}

// closest hit of one ray with t in [0, max_t); returns false on a miss
__device__ __forceinline__ bool trace(const float4* __restrict__ nodes, const float4* __restrict__ tris, const float ox, const float oy,
                                      const float oz, const float dx, const float dy, const float dz, const float max_t, Hit& hit) {
  const float ix = 1.0f / dx, iy = 1.0f / dy, iz = 1.0f / dz;
  const double Ox = ox, Oy = oy, Oz = oz, Dx = dx, Dy = dy, Dz = dz;
  double t_best = (double)max_t;
  float t_cull = max_t;
  int best_tri = -1;
  // one 8-byte stack entry per node: (child code, entry distance) -- one local store per push, one load per pop
  int2 stack[kStack];
  int sp = 0;
  stack[sp++] = make_int2(0, __float_as_int(0.0f));
  // Warp-coherent "while-while" traversal: every ray still visits its nodes and triangles in exactly its own stack order (so
  // hits are what the one-loop form returns, bit for bit), but the warp alternates between a NODE phase, in which the lanes
  // that are still searching expand inner nodes until they pop a leaf, and a LEAF phase, in which all lanes holding a leaf run
  // the fp64 triangle tests together.  A ballot per node step ends the node phase when no lane searches any more.
  // (One loop with "leaf or node" per iteration kept 17 of 32 lanes busy on the depth-camera workload.  Leaving the node
  // phase earlier -- while 1/16 ... all of the lanes still search -- was measured and is monotonically slower:
  // depth camera 5.21 Grays/s with this rule, 5.14 / 4.81 / 4.66 / 4.58 / 4.45 at 2 / 4 / 8 / 16 / 32 thirty-seconds.)
  // (Round 2 also measured a STREAMING form for the depth camera -- one warp works through a block of pixels and a lane whose ray is
  //  finished takes the next pixel when fewer than k lanes are still busy: bit-identical images, but 4.1-5.1 / 2.3-3.2 Grays/s
  //  (far clip 2 / 10 m) for k = 1 ... 33 against 5.9 / 4.3 for one trace() per pixel round: new rays start at the root while the
  //  others are deep in the tree, and the per-lane ray state kept across the loop costs registers.  profiles/r3i_depth_refill.txt.)
  // The slab planes of two children at a time go through the packed fp32 pipe (FADD2 / FMUL2: one instruction, two individually
  // rounded results -- the same values as the scalar form): 24 packed instead of 48 scalar instructions per node.
  const f32x2 nox = pack2(-ox, -ox), noy = pack2(-oy, -oy), noz = pack2(-oz, -oz);
  const f32x2 ix2 = pack2(ix, ix), iy2 = pack2(iy, iy), iz2 = pack2(iz, iz);
  const unsigned wmask = __activemask();
  int pending = 0;   // leaf code waiting for its triangle tests (leaf codes are negative)
  for (;;) {
    for (;;) {
      const bool searching = pending == 0 && sp > 0;
      if (!__any_sync(wmask, searching)) break;
      if (!searching) continue;
      --sp;
      const int2 top = stack[sp];
      const int code = top.x;
      if (__int_as_float(top.y) > t_cull) continue;
      if (code < 0) {
        pending = code;
        continue;
      }
      const ulonglong2* np = reinterpret_cast<const ulonglong2*>(nodes + 8 * (size_t)code);
      const ulonglong2 lx = __ldg(np), ly = __ldg(np + 1), lz = __ldg(np + 2), hx = __ldg(np + 3), hy = __ldg(np + 4), hz = __ldg(np + 5);
      const int4 ch = __ldg(reinterpret_cast<const int4*>(np + 6));
      float ax[4], bx[4], ay[4], by[4], az[4], bz[4];
      // (lo - o) * inv per plane, children (0, 1) and (2, 3) packed; a - b == a + (-b) exactly
      unpack2(mul2(add2(lx.x, nox), ix2), ax[0], ax[1]); unpack2(mul2(add2(lx.y, nox), ix2), ax[2], ax[3]);
      unpack2(mul2(add2(hx.x, nox), ix2), bx[0], bx[1]); unpack2(mul2(add2(hx.y, nox), ix2), bx[2], bx[3]);
      unpack2(mul2(add2(ly.x, noy), iy2), ay[0], ay[1]); unpack2(mul2(add2(ly.y, noy), iy2), ay[2], ay[3]);
      unpack2(mul2(add2(hy.x, noy), iy2), by[0], by[1]); unpack2(mul2(add2(hy.y, noy), iy2), by[2], by[3]);
      unpack2(mul2(add2(lz.x, noz), iz2), az[0], az[1]); unpack2(mul2(add2(lz.y, noz), iz2), az[2], az[3]);
      unpack2(mul2(add2(hz.x, noz), iz2), bz[0], bz[1]); unpack2(mul2(add2(hz.y, noz), iz2), bz[2], bz[3]);
      float tn[4];
      int cc[4] = {ch.x, ch.y, ch.z, ch.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        // slab test; fminf / fmaxf drop the NaN of 0 * inf when the origin lies on a slab plane of an axis-parallel ray
        const float t0 = fmaxf(fmaxf(fminf(ax[c], bx[c]), fminf(ay[c], by[c])), fmaxf(fminf(az[c], bz[c]), 0.0f));
        const float t1 = fminf(fminf(fmaxf(ax[c], bx[c]), fmaxf(ay[c], by[c])), fminf(fmaxf(az[c], bz[c]), t_cull)) * 1.0000005f;
        tn[c] = (cc[c] != kEmpty && t0 <= t1) ? t0 : FLT_MAX;
      }
      // sort the four children by entry distance (5-comparator network), push far to near.  (Measured in SASS: 32-bit keys --
      // distance bits with the child slot in the two low mantissa bits -- sorted with integer min / max need 10 instructions for
      // the network but a slot -> child select and a branch per push: 165 instructions per node step against 156 for this form.)
#define CSWAP(i, j)                                   \
  if (tn[i] > tn[j]) {                                \
    const float tt = tn[i]; tn[i] = tn[j]; tn[j] = tt; \
    const int ct = cc[i]; cc[i] = cc[j]; cc[j] = ct;  \
  }
      CSWAP(0, 1) CSWAP(2, 3) CSWAP(0, 2) CSWAP(1, 3) CSWAP(1, 2)
#undef CSWAP
#pragma unroll
      for (int c = 3; c >= 0; --c)
        if (tn[c] != FLT_MAX && sp < kStack) stack[sp++] = make_int2(cc[c], __float_as_int(tn[c]));
    }
    if (!__any_sync(wmask, pending != 0 || sp > 0)) break;
    if (pending != 0) {   // leaf
      const int first = (~pending) >> 2, count = ((~pending) & 3) + 1;
      pending = 0;
      for (int i = 0; i < count; ++i) {
        const float4* tp = tris + 3 * (size_t)(first + i);
        const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
        double t;
        if (ray_triangle(Ox, Oy, Oz, Dx, Dy, Dz, a, b, c, t)) {
          const int id = __float_as_int(c.y);
          if (t < t_best || (t == t_best && best_tri >= 0 && id < best_tri)) {   // ties: the lowest triangle id (what argmin picks)
            t_best = t;
            best_tri = id;
            t_cull = __double2float_ru(t);
          }
        }
      }
    }
  }
  hit.t = t_best;
  hit.tri = best_tri;
  return best_tri >= 0;
}

// ---------------------------------------------------------------------------------------------
// Regular-grid fast path.  The terrain meshes of the BASELINE configs are triangulated height fields (utils/terrain.py:76-80;
// the confined terrain is two of them, ground and ceiling, over the same grid: utils/terrain_confine.py:13-146): vertex (i, j) of
// a layer sits at (x_i, y_j), every triangle lies inside one cell.  For such a mesh a ray needs no tree: it walks the slabs of
// its major horizontal axis front to back, in each slab the one or two cells its (padded) footprint touches, skips a cell when
// the ray's height interval over the slab misses the cell's [min, max], and otherwise runs the SAME fp64 ray / triangle test on
// the cell's two triangles as the BVH path does.  Slabs are ordered by entry distance, so the walk stops at the first slab that
// starts behind the best hit.  Every candidate the exact test could accept is visited (footprints are padded by more than the
// fp32 error of the slab arithmetic), the accepted minimum is taken over the same triangle tests: distances and hit flags are
// bit-identical to the BVH path (tests/test_raycast.py compares them).
// ---------------------------------------------------------------------------------------------
struct GridView {
  int layers, nx, ny;          // vertices per axis; layers == 0: no grid, use the BVH
  float x0, y0, dx, dy, pad;
  const float2* cellz;
  const int2* celltri;
};

__device__ __forceinline__ bool trace_grid(const GridView g, const float4* __restrict__ tris, const float ox, const float oy, const float oz,
                                           const float dx, const float dy, const float dz, const float max_t, Hit& hit) {
  const double Ox = ox, Oy = oy, Oz = oz, Dx = dx, Dy = dy, Dz = dz;
  double t_best = (double)max_t;
  int best_tri = -1;
  const int cx = g.nx - 1, cy = g.ny - 1;
  auto test_cell = [&](const int i, const int j, const float zlo, const float zhi, const bool zcull) {
    for (int l = 0; l < g.layers; ++l) {
      const size_t c = ((size_t)l * cx + i) * cy + j;
      if (zcull) {
        const float2 z = __ldg(g.cellz + c);
        if (zhi < z.x || zlo > z.y) continue;
      }
      const int2 tt = __ldg(g.celltri + c);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float4* tp = tris + 3 * (size_t)(k == 0 ? tt.x : tt.y);
        const float4 a = __ldg(tp), b = __ldg(tp + 1), cc = __ldg(tp + 2);
        double t;
        if (ray_triangle(Ox, Oy, Oz, Dx, Dy, Dz, a, b, cc, t)) {
          const int id = __float_as_int(cc.y);
          if (t < t_best || (t == t_best && best_tri >= 0 && id < best_tri)) {
            t_best = t;
            best_tri = id;
          }
        }
      }
    }
  };
  const bool xmajor = fabsf(dx) >= fabsf(dy);
  const float ou = xmajor ? ox : oy, ov = xmajor ? oy : ox, du = xmajor ? dx : dy, dv = xmajor ? dy : dx;
  const float U0 = xmajor ? g.x0 : g.y0, V0 = xmajor ? g.y0 : g.x0, DU = xmajor ? g.dx : g.dy, DV = xmajor ? g.dy : g.dx;
  const int nu = xmajor ? cx : cy, nv = xmajor ? cy : cx;
  const float iDU = 1.0f / DU, iDV = 1.0f / DV, pad = g.pad;
  auto cell_range = [](const float lo, const float hi, const float base, const float inv, const int n, int& a, int& b) {
    // cells [a, b] whose extent meets [lo, hi]; empty (a > b) when the interval lies outside the grid
    const float fa = floorf((lo - base) * inv), fb = floorf((hi - base) * inv);
    a = fa < 0.0f ? 0 : (fa > (float)(n - 1) ? n : (int)fa);
    b = fb < 0.0f ? -1 : (fb > (float)(n - 1) ? n - 1 : (int)fb);
  };
  if (!(fabsf(du) * max_t > 4.0f * pad)) {
    // (near-)vertical ray: its whole footprint is a padded point -- at most 2 x 2 cells, no slab walk
    const float ext = fabsf(du) * max_t + pad;
    int ia, ib, ja, jb;
    cell_range(ou - ext, ou + ext, U0, iDU, nu, ia, ib);
    cell_range(ov - ext, ov + ext, V0, iDV, nv, ja, jb);
    for (int i = ia; i <= ib; ++i)
      for (int j = ja; j <= jb; ++j) test_cell(xmajor ? i : j, xmajor ? j : i, 0.0f, 0.0f, false);
  } else {
    const float inv_du = 1.0f / du;
    const float u_end = ou + du * max_t;
    int ia, ib;
    cell_range(fminf(ou, u_end) - pad, fmaxf(ou, u_end) + pad, U0, iDU, nu, ia, ib);
    const int step = du > 0.0f ? 1 : -1;
    const float zpad = pad * (1.0f + fabsf(dz * inv_du)), vpad = pad * (1.0f + fabsf(dv * inv_du));
    for (int i = du > 0.0f ? ia : ib, left = ib - ia + 1; left > 0; --left, i += step) {
      const float ua = U0 + (float)i * DU - pad, ub = U0 + (float)(i + 1) * DU + pad;
      float t0 = (ua - ou) * inv_du, t1 = (ub - ou) * inv_du;
      if (t0 > t1) { const float tt = t0; t0 = t1; t1 = tt; }
      t0 = fmaxf(t0, 0.0f);
      t1 = fminf(t1, max_t);
      if (t0 > t1) continue;
      if ((double)t0 > t_best) break;          // every later slab starts behind the best hit
      const float va = ov + t0 * dv, vb = ov + t1 * dv;
      int ja, jb;
      cell_range(fminf(va, vb) - vpad, fmaxf(va, vb) + vpad, V0, iDV, nv, ja, jb);
      const float za = oz + t0 * dz, zb = oz + t1 * dz;
      const float zlo = fminf(za, zb) - zpad, zhi = fmaxf(za, zb) + zpad;
      for (int j = ja; j <= jb; ++j) test_cell(xmajor ? i : j, xmajor ? j : i, zlo, zhi, true);
    }
  }
  hit.t = t_best;
  hit.tri = best_tri;
  return best_tri >= 0;
}

// the mesh query every ray kernel calls: grid walk for height-field-derived meshes, BVH otherwise -- a compile-time choice, so that
// neither instantiation carries the other's registers (the BVH stack, the slab state)
template <bool kGrid>
__device__ __forceinline__ bool trace_any(const GridView g, const float4* __restrict__ nodes, const float4* __restrict__ tris, const float ox,
                                          const float oy, const float oz, const float dx, const float dy, const float dz, const float max_t,
                                          Hit& hit) {
  if (kGrid) return trace_grid(g, tris, ox, oy, oz, dx, dy, dz, max_t, hit);
  return trace(nodes, tris, ox, oy, oz, dx, dy, dz, max_t, hit);
}

// ---------------------------------------------------------------------------------------------
// raycast_mesh (utils/ray_caster.py:45-92): one thread per ray
// ---------------------------------------------------------------------------------------------
template <bool kGrid>
__global__ void __launch_bounds__(128)
elg_raycast_kernel(const GridView gv, const float4* __restrict__ nodes, const float4* __restrict__ tris, const float* __restrict__ origins,
                   const float* __restrict__ dirs, const long long n, const float max_dist, float* __restrict__ hits,
                   uint8_t* __restrict__ found, float* __restrict__ dist_out, int* __restrict__ tri_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float ox = origins[3 * i], oy = origins[3 * i + 1], oz = origins[3 * i + 2];
  const float dx = dirs[3 * i], dy = dirs[3 * i + 1], dz = dirs[3 * i + 2];
  Hit h;
  const bool ok = trace_any<kGrid>(gv, nodes, tris, ox, oy, oz, dx, dy, dz, max_dist, h);
  const float t = ok ? (float)h.t : max_dist;   // miss: the ray end point (ray_caster.py:88-92)
  hits[3 * i] = add_r(ox, mul_r(t, dx));
  hits[3 * i + 1] = add_r(oy, mul_r(t, dy));
  hits[3 * i + 2] = add_r(oz, mul_r(t, dz));
  found[i] = ok ? 1 : 0;
  if (dist_out) dist_out[i] = t;
  if (tri_out) tri_out[i] = ok ? h.tri : -1;
}

// isaacgym.torch_utils.quat_apply(q, b) = b + w t + xyz x t, t = 2 (xyz x b), every torch op rounded on its own
__device__ __forceinline__ void quat_apply_r(const float qx, const float qy, const float qz, const float qw, const float bx,
                                             const float by, const float bz, float& rx, float& ry, float& rz) {
  const float tx = mul_r(sub_r(mul_r(qy, bz), mul_r(qz, by)), 2.0f);
  const float ty = mul_r(sub_r(mul_r(qz, bx), mul_r(qx, bz)), 2.0f);
  const float tz = mul_r(sub_r(mul_r(qx, by), mul_r(qy, bx)), 2.0f);
  rx = add_r(add_r(bx, mul_r(qw, tx)), sub_r(mul_r(qy, tz), mul_r(qz, ty)));
  ry = add_r(add_r(by, mul_r(qw, ty)), sub_r(mul_r(qz, tx), mul_r(qx, tz)));
  rz = add_r(add_r(bz, mul_r(qw, tz)), sub_r(mul_r(qx, ty), mul_r(qy, tx)));
}

// ---------------------------------------------------------------------------------------------
// RayCaster._update_ray_casting (utils/ray_caster.py:558-594) fused with the ray cast: the world ray of
// (sensor, pattern ray) is built in registers -- quat_apply or quat_apply_yaw (utils/math_utils.py:40-44) of the
// pattern origin / direction by the sensor quaternion, + sensor position -- instead of materialising [N, n, 3]
// origin and direction tensors.  One thread per (sensor, ray); rays of one sensor share a warp where n >= 32.
// ---------------------------------------------------------------------------------------------
template <bool kGrid>
__global__ void __launch_bounds__(128)
elg_raycast_sensor_kernel(const GridView gv, const float4* __restrict__ nodes, const float4* __restrict__ tris, const float* __restrict__ pat_o,
                          const float* __restrict__ pat_d, const int n_rays, const float* __restrict__ pos,
                          const float* __restrict__ quat, const int64_t* __restrict__ env_ids, const long long n_sensors,
                          const int yaw_only, const float max_dist, float* __restrict__ hits, uint8_t* __restrict__ found,
                          const float* __restrict__ dist_org, const int dist_stride, const int normalize, float* __restrict__ dist_out,
                          const long long dist_out_stride) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sensors * n_rays) return;
  const long long s = i / n_rays;
  const int r = (int)(i - s * n_rays);
  const long long e = env_ids ? env_ids[s] : s;
  float qx = quat[4 * e], qy = quat[4 * e + 1], qz = quat[4 * e + 2], qw = quat[4 * e + 3];
  if (yaw_only) {   // normalize((0, 0, qz, qw)), clamp(min=1e-9)
    float nrm = __fsqrt_rn(add_r(mul_r(qz, qz), mul_r(qw, qw)));
    nrm = fmaxf(nrm, 1e-9f);
    qx = 0.0f;
    qy = 0.0f;
    qz = div_r(qz, nrm);
    qw = div_r(qw, nrm);
  }
  float ox, oy, oz, dx, dy, dz;
  quat_apply_r(qx, qy, qz, qw, pat_o[3 * r], pat_o[3 * r + 1], pat_o[3 * r + 2], ox, oy, oz);
  ox = add_r(ox, pos[3 * e]);
  oy = add_r(oy, pos[3 * e + 1]);
  oz = add_r(oz, pos[3 * e + 2]);
  quat_apply_r(qx, qy, qz, qw, pat_d[3 * r], pat_d[3 * r + 1], pat_d[3 * r + 2], dx, dy, dz);
  Hit h;
  const bool ok = trace_any<kGrid>(gv, nodes, tris, ox, oy, oz, dx, dy, dz, max_dist, h);
  const float t = ok ? (float)h.t : max_dist;
  const long long o = (e * n_rays + r) * 3;
  const float hx = add_r(ox, mul_r(t, dx)), hy = add_r(oy, mul_r(t, dy)), hz = add_r(oz, mul_r(t, dz));
  hits[o] = hx;
  hits[o + 1] = hy;
  hits[o + 2] = hz;
  found[e * n_rays + r] = ok ? 1 : 0;
  if (dist_out) {
    // LeggedRobotRayCast._get_raycast_distances (envs/base/legged_robot_raycast.py:262-297): the distance is measured from the
    // ROBOT BASE (root_states[:, :3]), not from the ray origin; observation = (1 - clamp(d / max, 0, 1)) * found
    const float* org = dist_org + (size_t)e * dist_stride;
    float d = norm3_t(sub_r(hx, org[0]), sub_r(hy, org[1]), sub_r(hz, org[2]));
    if (normalize) d = mul_r(sub_r(1.0f, fminf(fmaxf(div_r(d, max_dist), 0.0f), 1.0f)), ok ? 1.0f : 0.0f);
    dist_out[e * dist_out_stride + r] = d;
  }
}

// isaacgym.torch_utils.quat_mul(a, b) (xyzw), one rounding per torch op
__device__ __forceinline__ void quat_mul_r(const float x1, const float y1, const float z1, const float w1, const float x2, const float y2,
                                           const float z2, const float w2, float& x, float& y, float& z, float& w) {
  const float ww = mul_r(add_r(z1, x1), add_r(x2, y2));
  const float yy = mul_r(sub_r(w1, y1), add_r(w2, z2));
  const float zz = mul_r(add_r(w1, y1), sub_r(w2, z2));
  const float xx = add_r(add_r(ww, yy), zz);
  const float qq = mul_r(0.5f, add_r(xx, mul_r(sub_r(z1, x1), sub_r(x2, y2))));
  w = add_r(sub_r(qq, ww), mul_r(sub_r(z1, y1), sub_r(y2, z2)));
  x = add_r(sub_r(qq, xx), mul_r(add_r(x1, w1), add_r(x2, w2)));
  y = add_r(sub_r(qq, yy), mul_r(sub_r(w1, x1), add_r(y2, z2)));
  z = add_r(sub_r(qq, zz), mul_r(add_r(z1, y1), sub_r(w2, x2)));
}

// DepthCameraWarp.update (utils/depth_camera.py:501-566): camera pose = base pose (x) mounting offset.
// offset_quat is the 4-vector the reference builds ([w, x, y, z] order) and then feeds to the xyzw quat_mul -- the
// arithmetic is replicated literally.
__global__ void __launch_bounds__(128)
elg_camera_pose_kernel(const float* __restrict__ pos, const float* __restrict__ quat, const int64_t* __restrict__ env_ids, const long long n,
                       const float ox, const float oy, const float oz, const float q0, const float q1, const float q2, const float q3,
                       float* __restrict__ cam_pos, float* __restrict__ cam_rot) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long e = env_ids ? env_ids[i] : i;
  const float qx = quat[4 * e], qy = quat[4 * e + 1], qz = quat[4 * e + 2], qw = quat[4 * e + 3];
  float wx, wy, wz;
  quat_apply_r(qx, qy, qz, qw, ox, oy, oz, wx, wy, wz);
  cam_pos[3 * e] = add_r(pos[3 * e], wx);
  cam_pos[3 * e + 1] = add_r(pos[3 * e + 1], wy);
  cam_pos[3 * e + 2] = add_r(pos[3 * e + 2], wz);
  float x, y, z, w;
  quat_mul_r(qx, qy, qz, qw, q0, q1, q2, q3, x, y, z, w);
  cam_rot[4 * e] = x; cam_rot[4 * e + 1] = y; cam_rot[4 * e + 2] = z; cam_rot[4 * e + 3] = w;
}

// ---------------------------------------------------------------------------------------------
// DepthCameraWarp.update_depth_buffer (utils/depth_camera.py:402-499) + DepthCameraBase.process_depth_image (:84-138)
// as ONE kernel, one CTA per camera: the ray of every pixel is rotated into the world in registers, cast, turned into
// -distance (or -far_clip), the image gets its scalar noise, is clipped, resized (separable antialiased bicubic, the
// taps tabulated by the host from torchvision itself), normalised and pushed into the env's frame ring buffer -- the
// reference's per-env Python loop (:486-499).  28 bytes per camera in, 4 bytes per output pixel out.
// ---------------------------------------------------------------------------------------------
// (register cap: 72 registers / 3 CTAs per SM 5.74 Grays/s, 64 / 4 CTAs 5.94-6.01, 48 / 5 CTAs 5.44 -- far clip 2 m, r2v / r3q)
template <bool kGrid>
__global__ void __launch_bounds__(256, 4)
elg_depth_camera_kernel(const GridView gv, const float4* __restrict__ nodes, const float4* __restrict__ tris, const __grid_constant__ ElgCamParams cp,
                        const float* __restrict__ ray_dirs, const float* __restrict__ cam_pos, const float* __restrict__ cam_rot,
                        const int64_t* __restrict__ ep_len, const float* __restrict__ noise_u, const int32_t* __restrict__ rx_start,
                        const float* __restrict__ rx_w, const int32_t* __restrict__ ry_start, const float* __restrict__ ry_w,
                        float* __restrict__ depth_buffer, float* __restrict__ raw_depth) {
  extern __shared__ float s_img[];   // [h * w] image, then [h * out_w] horizontally resized rows
  const int e = blockIdx.x;
  const int W = cp.width, Hh = cp.height, OW = cp.out_width, OH = cp.out_height;
  const int npx = W * Hh;
  const float px = cam_pos[3 * e], py = cam_pos[3 * e + 1], pz = cam_pos[3 * e + 2];
  const float qx = cam_rot[4 * e], qy = cam_rot[4 * e + 1], qz = cam_rot[4 * e + 2], qw = cam_rot[4 * e + 3];
  float noise = 0.0f;
  if (cp.noise_scale != 0.0f && noise_u) noise = mul_r(cp.noise_scale, sub_r(noise_u[e], 0.5f));
  for (int r = threadIdx.x; r < npx; r += blockDim.x) {
    float dx, dy, dz;
    quat_apply_r(qx, qy, qz, qw, ray_dirs[3 * r], ray_dirs[3 * r + 1], ray_dirs[3 * r + 2], dx, dy, dz);
    Hit h;
    const bool ok = trace_any<kGrid>(gv, nodes, tris, px, py, pz, dx, dy, dz, cp.far_clip, h);
    float d = -cp.far_clip;
    if (ok) {
      const float t = (float)h.t;
      const float hx = add_r(px, mul_r(t, dx)), hy = add_r(py, mul_r(t, dy)), hz = add_r(pz, mul_r(t, dz));
      d = -norm3_t(sub_r(hx, px), sub_r(hy, py), sub_r(hz, pz));     // torch.norm(hits - camera_pos) (:464)
    }
    if (raw_depth) raw_depth[(size_t)e * npx + r] = d;
    d = add_r(d, noise);
    d = fminf(fmaxf(d, -cp.far_clip), -cp.near_clip);
    s_img[r] = d;
  }
  __syncthreads();
  const float* src = s_img;
  if (cp.resize) {
    float* tmp = s_img + max(npx, OW * OH);   // [Hh, OW]
    for (int i = threadIdx.x; i < Hh * OW; i += blockDim.x) {
      const int y = i / OW, xo = i - y * OW;
      const int x0 = rx_start[xo];
      float acc = 0.0f;
      for (int k = 0; k < cp.max_taps; ++k) {
        const float wgt = rx_w[xo * cp.max_taps + k];
        if (wgt != 0.0f) acc = add_r(acc, mul_r(wgt, s_img[y * W + min(x0 + k, W - 1)]));
      }
      tmp[i] = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < OH * OW; i += blockDim.x) {
      const int yo = i / OW, xo = i - yo * OW;
      const int y0 = ry_start[yo];
      float acc = 0.0f;
      for (int k = 0; k < cp.max_taps; ++k) {
        const float wgt = ry_w[yo * cp.max_taps + k];
        if (wgt != 0.0f) acc = add_r(acc, mul_r(wgt, tmp[min(y0 + k, Hh - 1) * OW + xo]));
      }
      s_img[i] = acc;                  // the original image is dead: reuse its head for the final frame
    }
    __syncthreads();
  }
  // normalise (:56-69) and push into the ring buffer (:483-499)
  const int opx = OW * OH;
  const bool init = ep_len[e] <= 1;
  const float range = sub_r(cp.far_clip, cp.near_clip);
  float* buf = depth_buffer + (size_t)e * cp.buffer_len * opx;
  for (int i = threadIdx.x; i < opx; i += blockDim.x) {
    const float v = sub_r(div_r(sub_r(mul_r(src[i], -1.0f), cp.near_clip), range), 0.5f);
    if (init) {
      for (int k = 0; k < cp.buffer_len; ++k) buf[(size_t)k * opx + i] = v;
    } else {
      for (int k = 0; k + 1 < cp.buffer_len; ++k) buf[(size_t)k * opx + i] = buf[(size_t)(k + 1) * opx + i];
      buf[(size_t)(cp.buffer_len - 1) * opx + i] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// signed distance: closest point on a triangle (Ericson, Real-Time Collision Detection 5.1.5) in fp64 with
// individually rounded operations, same expression order as oracle/mesh_oracle.py
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double ddot(double ax, double ay, double az, double bx, double by, double bz) {
  return dadd(dadd(dmul(ax, bx), dmul(ay, by)), dmul(az, bz));
}
struct Closest {
  double cx, cy, cz;   // closest point
  double d2;           // squared distance
};
__device__ __forceinline__ Closest closest_on_triangle(const double px, const double py, const double pz, const float4 A, const float4 Bv,
                                                       const float4 Cv) {
  const double ax = A.x, ay = A.y, az = A.z, bx = A.w, by = Bv.x, bz = Bv.y, cx = Bv.z, cy = Bv.w, cz = Cv.x;
  const double abx = dsub(bx, ax), aby = dsub(by, ay), abz = dsub(bz, az);
  const double acx = dsub(cx, ax), acy = dsub(cy, ay), acz = dsub(cz, az);
  const double apx = dsub(px, ax), apy = dsub(py, ay), apz = dsub(pz, az);
  const double d1 = ddot(abx, aby, abz, apx, apy, apz), d2 = ddot(acx, acy, acz, apx, apy, apz);
  const double bpx = dsub(px, bx), bpy = dsub(py, by), bpz = dsub(pz, bz);
  const double d3 = ddot(abx, aby, abz, bpx, bpy, bpz), d4 = ddot(acx, acy, acz, bpx, bpy, bpz);
  const double cpx = dsub(px, cx), cpy = dsub(py, cy), cpz = dsub(pz, cz);
  const double d5 = ddot(abx, aby, abz, cpx, cpy, cpz), d6 = ddot(acx, acy, acz, cpx, cpy, cpz);
  const double vc = dsub(dmul(d1, d4), dmul(d3, d2));
  const double vb = dsub(dmul(d5, d2), dmul(d1, d6));
  const double va = dsub(dmul(d3, d6), dmul(d5, d4));
  // (Measured: forming every region's candidate and selecting -- no seven-way branch -- is SLOWER: 223 vs 288 Mpoints/s; the four
  //  fp64 divisions per triangle cost more than the divergence, and most of the idle lanes of this kernel sit in the traversal.)
  double qx, qy, qz;
  if (d1 <= 0.0 && d2 <= 0.0) {                                   // vertex A
    qx = ax; qy = ay; qz = az;
  } else if (d3 >= 0.0 && d4 <= d3) {                             // vertex B
    qx = bx; qy = by; qz = bz;
  } else if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) {               // edge AB
    const double v = d1 / dsub(d1, d3);
    qx = dadd(ax, dmul(v, abx)); qy = dadd(ay, dmul(v, aby)); qz = dadd(az, dmul(v, abz));
  } else if (d6 >= 0.0 && d5 <= d6) {                             // vertex C
    qx = cx; qy = cy; qz = cz;
  } else if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) {               // edge AC
    const double w = d2 / dsub(d2, d6);
    qx = dadd(ax, dmul(w, acx)); qy = dadd(ay, dmul(w, acy)); qz = dadd(az, dmul(w, acz));
  } else if (va <= 0.0 && dsub(d4, d3) >= 0.0 && dsub(d5, d6) >= 0.0) {   // edge BC
    const double w = dsub(d4, d3) / dadd(dsub(d4, d3), dsub(d5, d6));
    qx = dadd(bx, dmul(w, dsub(cx, bx))); qy = dadd(by, dmul(w, dsub(cy, by))); qz = dadd(bz, dmul(w, dsub(cz, bz)));
  } else {                                                        // interior
    const double denom = 1.0 / dadd(dadd(va, vb), vc);
    const double v = dmul(vb, denom), w = dmul(vc, denom);
    qx = dadd(dadd(ax, dmul(abx, v)), dmul(acx, w));
    qy = dadd(dadd(ay, dmul(aby, v)), dmul(acy, w));
    qz = dadd(dadd(az, dmul(abz, v)), dmul(acz, w));
  }
  Closest r;
  r.cx = qx; r.cy = qy; r.cz = qz;
  const double ox = dsub(px, qx), oy = dsub(py, qy), oz = dsub(pz, qz);
  r.d2 = ddot(ox, oy, oz, ox, oy, oz);
  return r;
}

// Visits every triangle whose padded leaf box lies within sqrt(r2) of p, nearest boxes first.  F(tri float4 x3) may
// shrink r2 (returns the new squared radius, or a negative value to keep it).
template <typename F>
__device__ __forceinline__ void visit_near(const float4* __restrict__ nodes, const float4* __restrict__ tris, const float px, const float py,
                                           const float pz, float r2, F&& f) {
  int2 stack[kStack];      // (child code, lower bound of the squared distance): one 8-byte local store per push
  int sp = 0;
  stack[sp++] = make_int2(0, __float_as_int(0.0f));
  // warp-coherent while-while, as in trace(): node phase until no lane searches, then all pending leaves together.  The box
  // distances of two children at a time go through the packed fp32 pipe (same individually rounded values as the scalar form).
  const f32x2 pxp = pack2(px, px), pyp = pack2(py, py), pzp = pack2(pz, pz);
  const f32x2 pxn = pack2(-px, -px), pyn = pack2(-py, -py), pzn = pack2(-pz, -pz);
  const f32x2 neg1 = pack2(-1.0f, -1.0f);
  const unsigned wmask = __activemask();
  int pending = 0;
  for (;;) {
    for (;;) {
      const bool searching = pending == 0 && sp > 0;
      if (!__any_sync(wmask, searching)) break;
      if (!searching) continue;
      --sp;
      const int2 top = stack[sp];
      const int code = top.x;
      if (__int_as_float(top.y) > r2) continue;
      if (code < 0) {
        pending = code;
        continue;
      }
      const ulonglong2* np = reinterpret_cast<const ulonglong2*>(nodes + 8 * (size_t)code);
      const ulonglong2 lx = __ldg(np), ly = __ldg(np + 1), lz = __ldg(np + 2), hx = __ldg(np + 3), hy = __ldg(np + 4), hz = __ldg(np + 5);
      const int4 ch = __ldg(reinterpret_cast<const int4*>(np + 6));
      float dn[4];
      int cc[4] = {ch.x, ch.y, ch.z, ch.w};
      // lo - p and p - hi per axis (p - hi == (-1) * hi + p would fuse: it is formed as p + (-hi) with an exact negation)
      float lxm[4], hxm[4], lym[4], hym[4], lzm[4], hzm[4];
      unpack2(add2(lx.x, pxn), lxm[0], lxm[1]); unpack2(add2(lx.y, pxn), lxm[2], lxm[3]);
      unpack2(add2(mul2(hx.x, neg1), pxp), hxm[0], hxm[1]); unpack2(add2(mul2(hx.y, neg1), pxp), hxm[2], hxm[3]);
      unpack2(add2(ly.x, pyn), lym[0], lym[1]); unpack2(add2(ly.y, pyn), lym[2], lym[3]);
      unpack2(add2(mul2(hy.x, neg1), pyp), hym[0], hym[1]); unpack2(add2(mul2(hy.y, neg1), pyp), hym[2], hym[3]);
      unpack2(add2(lz.x, pzn), lzm[0], lzm[1]); unpack2(add2(lz.y, pzn), lzm[2], lzm[3]);
      unpack2(add2(mul2(hz.x, neg1), pzp), hzm[0], hzm[1]); unpack2(add2(mul2(hz.y, neg1), pzp), hzm[2], hzm[3]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float ex = fmaxf(fmaxf(lxm[c], hxm[c]), 0.0f);
        const float ey = fmaxf(fmaxf(lym[c], hym[c]), 0.0f);
        const float ez = fmaxf(fmaxf(lzm[c], hzm[c]), 0.0f);
        const float d = (ex * ex + ey * ey + ez * ez) * 0.999999f;     // lower bound of the squared distance to the box
        dn[c] = (cc[c] != kEmpty && d <= r2) ? d : FLT_MAX;
      }
#define CSWAP(i, j)                                   \
  if (dn[i] > dn[j]) {                                \
    const float tt = dn[i]; dn[i] = dn[j]; dn[j] = tt; \
    const int ct = cc[i]; cc[i] = cc[j]; cc[j] = ct;  \
  }
      CSWAP(0, 1) CSWAP(2, 3) CSWAP(0, 2) CSWAP(1, 3) CSWAP(1, 2)
#undef CSWAP
#pragma unroll
      for (int c = 3; c >= 0; --c)
        if (dn[c] != FLT_MAX && sp < kStack) stack[sp++] = make_int2(cc[c], __float_as_int(dn[c]));
    }
    if (!__any_sync(wmask, pending != 0 || sp > 0)) break;
    if (pending != 0) {
      const int first = (~pending) >> 2, count = ((~pending) & 3) + 1;
      pending = 0;
      for (int i = 0; i < count; ++i) {
        const float4* tp = tris + 3 * (size_t)(first + i);
        const float nr2 = f(__ldg(tp), __ldg(tp + 1), __ldg(tp + 2));
        if (nr2 >= 0.0f) r2 = nr2;
      }
    }
  }
}

// query_sdf_kernel (utils/mesh_sdf.py:38-116) on top of an order-independent statement of
// wp.mesh_query_point_sign_normal: pass 1 finds the exact minimum distance d_min; pass 2 looks at every face within
// d_min + epsilon * mean_edge and takes the one whose unit normal is most aligned with the offset (|n . (p - c)|
// largest, lowest triangle id on ties) -- the face that decides the sign at edges and vertices.
struct SdfResult { float sdf, gx, gy, gz, cx, cy, cz; int face; };

__device__ __forceinline__ SdfResult sdf_at(const float4* __restrict__ nodes, const float4* __restrict__ tris, const float pxf, const float pyf,
                                            const float pzf, const float max_distance, const double eps_abs) {
  const double px = pxf, py = pyf, pz = pzf;
  const double D = (double)max_distance;
  // pass 1: exact minimum squared distance
  double best = dmul(D, D);
  bool any = false;
  visit_near(nodes, tris, pxf, pyf, pzf, __double2float_ru(best) * 1.000001f, [&](const float4 a, const float4 b, const float4 c) -> float {
    const Closest q = closest_on_triangle(px, py, pz, a, b, c);
    if (q.d2 <= best) {
      best = q.d2;
      any = true;
      return __double2float_ru(best) * 1.000001f;
    }
    return -1.0f;
  });
  if (!any) {   // nothing within max_distance (mesh_sdf.py:112-115)
    return SdfResult{max_distance, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, -1};
  }
  // pass 2: the deciding face among the near-ties
  const double dmin = sqrt(best);
  const double lim = dadd(dmin, eps_abs);
  const double lim2 = dmul(lim, lim);
  double score = -1.0, sdot = 0.0, bx = 0.0, by = 0.0, bz = 0.0, bd2 = 0.0, nx = 0.0, ny = 0.0, nz = 0.0;
  int btri = 0x7fffffff;
  visit_near(nodes, tris, pxf, pyf, pzf, __double2float_ru(lim2) * 1.000001f, [&](const float4 a, const float4 b, const float4 c) -> float {
    const Closest q = closest_on_triangle(px, py, pz, a, b, c);
    if (q.d2 <= lim2) {
      const double e1x = dsub((double)a.w, (double)a.x), e1y = dsub((double)b.x, (double)a.y), e1z = dsub((double)b.y, (double)a.z);
      const double e2x = dsub((double)b.z, (double)a.x), e2y = dsub((double)b.w, (double)a.y), e2z = dsub((double)c.x, (double)a.z);
      const double fx = dsub(dmul(e1y, e2z), dmul(e1z, e2y)), fy = dsub(dmul(e1z, e2x), dmul(e1x, e2z)), fz = dsub(dmul(e1x, e2y), dmul(e1y, e2x));
      const double fl = sqrt(ddot(fx, fy, fz, fx, fy, fz));
      const double ox = dsub(px, q.cx), oy = dsub(py, q.cy), oz = dsub(pz, q.cz);
      const double dt = ddot(fx, fy, fz, ox, oy, oz);
      const double sc = fl > 0.0 ? fabs(dt) / fl : 0.0;
      const int tri = __float_as_int(c.y);
      if (sc > score || (sc == score && tri < btri)) {
        score = sc; sdot = dt; btri = tri;
        bx = q.cx; by = q.cy; bz = q.cz; bd2 = q.d2;
        nx = fx; ny = fy; nz = fz;
      }
    }
    return -1.0f;
  });
  const double dist = sqrt(bd2);
  const double sign = sdot < 0.0 ? -1.0 : 1.0;
  double gx, gy, gz;
  if (dist > 1.0e-6) {       // offset direction, flipped inside (mesh_sdf.py:89-94)
    gx = dsub(px, bx) / dist; gy = dsub(py, by) / dist; gz = dsub(pz, bz) / dist;
  } else {                   // on the surface: the face normal (:95-107)
    const double fl = sqrt(ddot(nx, ny, nz, nx, ny, nz));
    const double il = fl > 0.0 ? 1.0 / fl : 0.0;
    gx = dmul(nx, il); gy = dmul(ny, il); gz = dmul(nz, il);
  }
  return SdfResult{(float)dmul(dist, sign), (float)dmul(gx, sign), (float)dmul(gy, sign), (float)dmul(gz, sign), (float)bx, (float)by, (float)bz, btri};
}


__global__ void __launch_bounds__(128)
elg_sdf_kernel(const float4* __restrict__ nodes, const float4* __restrict__ tris, const float* __restrict__ points, const long long n,
               const float max_distance, const double eps_abs, float* __restrict__ sdf, float* __restrict__ grad,
               float* __restrict__ closest, int* __restrict__ face) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const SdfResult r = sdf_at(nodes, tris, points[3 * i], points[3 * i + 1], points[3 * i + 2], max_distance, eps_abs);
  sdf[i] = r.sdf;
  grad[3 * i] = r.gx; grad[3 * i + 1] = r.gy; grad[3 * i + 2] = r.gz;
  if (closest) { closest[3 * i] = r.cx; closest[3 * i + 1] = r.cy; closest[3 * i + 2] = r.cz; }
  if (face) face[i] = r.face;
}

// RobotBatchRolloutPercept._update_sdf_values (envs/batch_rollout/robot_batch_rollout_percept.py:385-441) as ONE launch over
// (env, query body): the query point is the body position + quat_rotate(body quaternion, collision-sphere offset) (:396-414, one
// rounding per torch op like isaacgym.torch_utils.quat_rotate), the signed distance / gradient come from the same traversal,
// and nearest = p - sdf * gradient (utils/mesh_sdf.py:316-336) without the second query the reference runs for it.
constexpr int kMaxSdfBodies = 16;
struct SdfBodies {
  int num_bodies, count;
  int body_idx[kMaxSdfBodies];
  float offset[kMaxSdfBodies][3];
  int has_offset[kMaxSdfBodies];
};
__global__ void __launch_bounds__(128)
elg_sdf_bodies_kernel(const float4* __restrict__ nodes, const float4* __restrict__ tris, const float* __restrict__ rbs,
                      const __grid_constant__ SdfBodies sb, const int64_t* __restrict__ env_ids, const long long n_rows,
                      const float max_distance, const double eps_abs, float* __restrict__ sdf, const long long sdf_row_stride,
                      float* __restrict__ grad, float* __restrict__ nearest, float* __restrict__ points_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * sb.count) return;
  const long long row = i / sb.count;
  const int k = (int)(i - row * sb.count);
  const long long e = env_ids ? env_ids[row] : row;
  const float* st = rbs + ((size_t)e * sb.num_bodies + sb.body_idx[k]) * 13;
  float px = st[0], py = st[1], pz = st[2];
  if (sb.has_offset[k]) {
    // quat_rotate(q, v) = v (2 w^2 - 1) + 2 w (q_xyz x v) + 2 q_xyz (q_xyz . v)   (dot through bmm: sequential sum)
    const float qx = st[3], qy = st[4], qz = st[5], qw = st[6];
    const float vx = sb.offset[k][0], vy = sb.offset[k][1], vz = sb.offset[k][2];
    const float s = sub_r(mul_r(2.0f, mul_r(qw, qw)), 1.0f);
    const float cx = sub_r(mul_r(qy, vz), mul_r(qz, vy)), cy = sub_r(mul_r(qz, vx), mul_r(qx, vz)), cz = sub_r(mul_r(qx, vy), mul_r(qy, vx));
    const float d = add_r(add_r(mul_r(qx, vx), mul_r(qy, vy)), mul_r(qz, vz));
    const float w2 = mul_r(qw, 2.0f);
    px = add_r(px, add_r(add_r(mul_r(vx, s), mul_r(cx, w2)), mul_r(mul_r(qx, d), 2.0f)));
    py = add_r(py, add_r(add_r(mul_r(vy, s), mul_r(cy, w2)), mul_r(mul_r(qy, d), 2.0f)));
    pz = add_r(pz, add_r(add_r(mul_r(vz, s), mul_r(cz, w2)), mul_r(mul_r(qz, d), 2.0f)));
  }
  const SdfResult r = sdf_at(nodes, tris, px, py, pz, max_distance, eps_abs);
  const long long o = e * sb.count + k;
  sdf[e * sdf_row_stride + k] = r.sdf;
  if (grad) { grad[3 * o] = r.gx; grad[3 * o + 1] = r.gy; grad[3 * o + 2] = r.gz; }
  if (nearest) {
    nearest[3 * o] = sub_r(px, mul_r(r.sdf, r.gx));
    nearest[3 * o + 1] = sub_r(py, mul_r(r.sdf, r.gy));
    nearest[3 * o + 2] = sub_r(pz, mul_r(r.sdf, r.gz));
  }
  if (points_out) { points_out[3 * o] = px; points_out[3 * o + 1] = py; points_out[3 * o + 2] = pz; }
}

// ---------------------------------------------------------------------------------------------
// host: binned-SAH binary build, collapse to 4-wide, upload
// ---------------------------------------------------------------------------------------------
struct Box {
  float lo[3], hi[3];
  void reset() { lo[0] = lo[1] = lo[2] = FLT_MAX; hi[0] = hi[1] = hi[2] = -FLT_MAX; }
  void grow(const float* p) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); } }
  void grow(const Box& b) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); } }
  double area() const {
    const double dx = (double)hi[0] - lo[0], dy = (double)hi[1] - lo[1], dz = (double)hi[2] - lo[2];
    return (dx < 0 || dy < 0 || dz < 0) ? 0.0 : 2.0 * (dx * dy + dy * dz + dz * dx);
  }
};
struct BNode {
  Box box;
  int left, right;   // children (inner) or -1
  int first, count;  // triangle range (leaf)
};

struct Builder {
  const float* V;
  const int32_t* T;
  int M;
  int leaf_triangles = 0;   // 0: default
  std::vector<Box> tbox;
  std::vector<float> cen;   // 3 per triangle
  std::vector<int> order;
  std::vector<BNode> nodes;

  void run() {
    tbox.resize(M);
    cen.resize(3 * (size_t)M);
    order.resize(M);
    for (int i = 0; i < M; ++i) {
      Box b;
      b.reset();
      for (int k = 0; k < 3; ++k) b.grow(V + 3 * (size_t)T[3 * (size_t)i + k]);
      tbox[i] = b;
      for (int k = 0; k < 3; ++k) cen[3 * (size_t)i + k] = 0.5f * (b.lo[k] + b.hi[k]);
      order[i] = i;
    }
    nodes.reserve(M);
    nodes.push_back(BNode{});
    struct Job { int node, first, count; };
    std::vector<Job> todo;
    todo.push_back({0, 0, M});
    constexpr int kBins = 16;
    // triangles per leaf (<= 4: the leaf code holds count - 1 in two bits); ELG_BVH_LEAF overrides it for measurements.  Measured on
    // the 1.6 M-triangle terrain (profiles/r3r_bvh_leaf.txt; depth camera far clip 2 m / 10 m / incoherent rays, Mrays/s):
    // 1: 5348 / 3606 / 2969, 2: 6599 / 4734 / 3595, 3: 6716 / 4830 / 3626, 4: 6149 / 4264 / 3408 -- an fp64 triangle test costs about as much
    // as a node step, so smaller leaves win until the tree gets a level deeper for nothing
    // (closest-point queries prefer 4: SDF 300 vs 266 Mpoints/s -- elg_mesh_create_ex lets the caller say which queries the mesh serves)
    int leaf_max = leaf_triangles > 0 ? leaf_triangles : 3;
    if (const char* ev = getenv("ELG_BVH_LEAF")) { const int v = atoi(ev); if (v >= 1 && v <= 4) leaf_max = v; }
    while (!todo.empty()) {
      const Job j = todo.back();
      todo.pop_back();
      Box nb, cb;
      nb.reset();
      cb.reset();
      for (int i = j.first; i < j.first + j.count; ++i) {
        nb.grow(tbox[order[i]]);
        cb.grow(&cen[3 * (size_t)order[i]]);
      }
      BNode& N = nodes[j.node];
      N.box = nb;
      N.left = N.right = -1;
      N.first = j.first;
      N.count = j.count;
      if (j.count <= leaf_max) continue;
      // best binned split over the three axes
      int best_axis = -1, best_bin = -1;
      double best_cost = DBL_MAX;
      for (int ax = 0; ax < 3; ++ax) {
        const float lo = cb.lo[ax], ext = cb.hi[ax] - cb.lo[ax];
        if (!(ext > 0.0f)) continue;
        Box bb[kBins];
        int bc[kBins];
        for (int b = 0; b < kBins; ++b) { bb[b].reset(); bc[b] = 0; }
        const float scale = kBins / ext;
        for (int i = j.first; i < j.first + j.count; ++i) {
          const int t = order[i];
          int b = (int)((cen[3 * (size_t)t + ax] - lo) * scale);
          b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
          bb[b].grow(tbox[t]);
          bc[b]++;
        }
        double la[kBins], ra[kBins];
        int lc[kBins], rc[kBins];
        Box acc;
        acc.reset();
        int c = 0;
        for (int b = 0; b < kBins; ++b) { acc.grow(bb[b]); c += bc[b]; la[b] = acc.area(); lc[b] = c; }
        acc.reset();
        c = 0;
        for (int b = kBins - 1; b >= 0; --b) { acc.grow(bb[b]); c += bc[b]; ra[b] = acc.area(); rc[b] = c; }
        for (int b = 0; b + 1 < kBins; ++b) {
          if (lc[b] == 0 || rc[b + 1] == 0) continue;
          const double cost = la[b] * lc[b] + ra[b + 1] * rc[b + 1];
          if (cost < best_cost) { best_cost = cost; best_axis = ax; best_bin = b; }
        }
      }
      int mid;
      if (best_axis < 0) {   // all centroids coincide: split the range in the middle
        mid = j.first + j.count / 2;
      } else {
        const float lo = cb.lo[best_axis], scale = kBins / (cb.hi[best_axis] - cb.lo[best_axis]);
        auto it = std::partition(order.begin() + j.first, order.begin() + j.first + j.count, [&](int t) {
          int b = (int)((cen[3 * (size_t)t + best_axis] - lo) * scale);
          b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
          return b <= best_bin;
        });
        mid = (int)(it - order.begin());
        if (mid == j.first || mid == j.first + j.count) mid = j.first + j.count / 2;
      }
      const int l = (int)nodes.size();
      nodes.push_back(BNode{});
      nodes.push_back(BNode{});
      nodes[j.node].left = l;
      nodes[j.node].right = l + 1;
      todo.push_back({l, j.first, mid - j.first});
      todo.push_back({l + 1, mid, j.first + j.count - mid});
    }
  }
};

}  // namespace elg

namespace {
int mfail(int code, const char* msg) { return elg::set_error(code, msg); }
}  // namespace

extern "C" {

int elg_mesh_create_ex(const float* vertices, int32_t num_vertices, const int32_t* triangles, int32_t num_triangles, int32_t leaf_triangles,
                       ElgMesh** out) {
  if (!out) return mfail(ELG_ERR_NULL_POINTER, "out is NULL");
  if (leaf_triangles < 0 || leaf_triangles > 4) return mfail(ELG_ERR_INVALID_ARGUMENT, "leaf_triangles must be 0 (default) or 1 ... 4");
  *out = nullptr;
  if (!vertices || !triangles) return mfail(ELG_ERR_NULL_POINTER, "vertices/triangles is NULL (host pointers expected)");
  if (num_vertices < 3 || num_triangles < 1) return mfail(ELG_ERR_INVALID_ARGUMENT, "a mesh needs >= 3 vertices and >= 1 triangle");
  if (num_triangles > (1 << 28)) return mfail(ELG_ERR_UNSUPPORTED, "more than 2^28 triangles");
  for (long long i = 0; i < 3LL * num_triangles; ++i)
    if (triangles[i] < 0 || triangles[i] >= num_vertices) return mfail(ELG_ERR_INVALID_ARGUMENT, "triangle index out of range");
  for (long long i = 0; i < 3LL * num_vertices; ++i)
    if (!std::isfinite(vertices[i])) return mfail(ELG_ERR_INVALID_ARGUMENT, "non-finite vertex coordinate");

  elg::Builder B;
  B.V = vertices;
  B.T = triangles;
  B.M = num_triangles;
  B.leaf_triangles = leaf_triangles;
  B.run();

  // collapse: every 4-wide node adopts grandchildren, widest-area child first
  struct WNode { int child[4]; elg::Box box[4]; };
  std::vector<WNode> wide;
  std::vector<int> bin_of;   // binary node of each wide node
  wide.reserve(B.nodes.size() / 2 + 1);
  wide.push_back(WNode{});
  bin_of.push_back(0);
  const auto leaf_code = [&](const elg::BNode& n) { return ~((n.first << 2) | (n.count - 1)); };
  for (size_t w = 0; w < wide.size(); ++w) {
    const int bn = bin_of[w];
    int slots[4], ns = 0;
    if (B.nodes[bn].left < 0) {
      slots[ns++] = bn;   // the root itself is a leaf (tiny mesh)
    } else {
      slots[ns++] = B.nodes[bn].left;
      slots[ns++] = B.nodes[bn].right;
      while (ns < 4) {
        int pick = -1;
        double best = -1.0;
        for (int s = 0; s < ns; ++s)
          if (B.nodes[slots[s]].left >= 0 && B.nodes[slots[s]].box.area() > best) { best = B.nodes[slots[s]].box.area(); pick = s; }
        if (pick < 0) break;
        const int p = slots[pick];
        slots[pick] = B.nodes[p].left;
        slots[ns++] = B.nodes[p].right;
      }
    }
    WNode wn;
    for (int s = 0; s < 4; ++s) {
      wn.child[s] = elg::kEmpty;
      wn.box[s].reset();
    }
    for (int s = 0; s < ns; ++s) {
      const elg::BNode& c = B.nodes[slots[s]];
      wn.box[s] = c.box;
      if (c.left < 0) {
        wn.child[s] = leaf_code(c);
      } else {
        wn.child[s] = (int)wide.size();
        wide.push_back(WNode{});
        bin_of.push_back(slots[s]);
      }
    }
    wide[w] = wn;
  }

  // pack
  elg::Box scene = B.nodes[0].box;
  float ext = 0.0f;
  for (int k = 0; k < 3; ++k) ext = std::max(ext, std::max(fabsf(scene.lo[k]), fabsf(scene.hi[k])));
  const float pad = 4e-7f * ext + 1e-30f;   // > 3 ulp of the largest coordinate: the fp32 slab test stays conservative
  std::vector<float> hn(32 * wide.size());
  for (size_t w = 0; w < wide.size(); ++w) {
    float* p = hn.data() + 32 * w;
    for (int s = 0; s < 4; ++s) {
      const bool e = wide[w].child[s] == elg::kEmpty;
      for (int k = 0; k < 3; ++k) {
        p[4 * k + s] = e ? FLT_MAX : wide[w].box[s].lo[k] - pad;
        p[12 + 4 * k + s] = e ? -FLT_MAX : wide[w].box[s].hi[k] + pad;
      }
      memcpy(p + 24 + s, &wide[w].child[s], 4);
    }
    for (int s = 28; s < 32; ++s) p[s] = 0.0f;
  }
  std::vector<float> ht(12 * (size_t)num_triangles);
  double edge_sum = 0.0;
  for (int i = 0; i < num_triangles; ++i) {
    const int t = B.order[i];
    float* p = ht.data() + 12 * (size_t)i;
    const float* v[3];
    for (int k = 0; k < 3; ++k) v[k] = vertices + 3 * (size_t)triangles[3 * (size_t)t + k];
    for (int k = 0; k < 3; ++k) { p[k] = v[0][k]; p[3 + k] = v[1][k]; p[6 + k] = v[2][k]; }
    memcpy(p + 9, &t, 4);
    p[10] = p[11] = 0.0f;
    for (int e = 0; e < 3; ++e) {
      const float* a = v[e];
      const float* b = v[(e + 1) % 3];
      edge_sum += sqrt(((double)a[0] - b[0]) * ((double)a[0] - b[0]) + ((double)a[1] - b[1]) * ((double)a[1] - b[1]) +
                       ((double)a[2] - b[2]) * ((double)a[2] - b[2]));
    }
  }

  // ---- regular-grid detection (height-field-derived meshes: GridView above).  V = layers x nx x ny vertices with vertex (l, i, j) at
  // (xs[i], ys[j]) exactly, xs / ys increasing and uniform to within a small fraction of a cell, every triangle inside one cell of one
  // layer, exactly two triangles per cell.  Anything else (OBJ scenes, slope-corrected meshes whose vertices were moved sideways)
  // keeps the BVH walk.
  std::vector<float> h_cellz;
  std::vector<int> h_celltri;
  int g_layers = 0, g_nx = 0, g_ny = 0;
  float g_x0 = 0, g_y0 = 0, g_dx = 0, g_dy = 0, g_pad = 0;
  {
    std::vector<int> pos_of(num_triangles);
    for (int i = 0; i < num_triangles; ++i) pos_of[B.order[i]] = i;
    for (int L = 1; L <= 4 && g_layers == 0; ++L) {
      if (num_vertices % L) continue;
      const int P = num_vertices / L;
      int ny = 1;
      while (ny < P && vertices[3 * (size_t)ny] == vertices[0]) ++ny;
      if (ny < 2 || P % ny) continue;
      const int nx = P / ny;
      if (nx < 2 || (long long)L * (nx - 1) * (ny - 1) * 2 != num_triangles) continue;
      bool ok = true;
      for (int l = 0; l < L && ok; ++l)
        for (int i = 0; i < nx && ok; ++i)
          for (int j = 0; j < ny; ++j) {
            const float* v = vertices + 3 * ((size_t)l * P + (size_t)i * ny + j);
            if (v[0] != vertices[3 * (size_t)i * ny] || v[1] != vertices[3 * (size_t)j + 1]) { ok = false; break; }
          }
      if (!ok) continue;
      const double x0 = vertices[0], y0 = vertices[1];
      const double dx = ((double)vertices[3 * (size_t)(nx - 1) * ny] - x0) / (nx - 1), dy = ((double)vertices[3 * (size_t)(ny - 1) + 1] - y0) / (ny - 1);
      if (!(dx > 0.0) || !(dy > 0.0)) continue;
      double dev = 0.0;
      for (int i = 0; i < nx; ++i) dev = std::max(dev, fabs((double)vertices[3 * (size_t)i * ny] - (x0 + i * dx)));
      for (int j = 0; j < ny; ++j) dev = std::max(dev, fabs((double)vertices[3 * (size_t)j + 1] - (y0 + j * dy)));
      if (dev > 0.01 * std::min(dx, dy)) continue;
      const size_t ncell = (size_t)L * (nx - 1) * (ny - 1);
      std::vector<int> tri(2 * ncell, -1);
      for (int t = 0; t < num_triangles && ok; ++t) {
        int lmin = INT32_MAX, lmax = -1, imin = INT32_MAX, imax = -1, jmin = INT32_MAX, jmax = -1;
        for (int k = 0; k < 3; ++k) {
          const int vi = triangles[3 * (size_t)t + k];
          const int l = vi / P, r = vi % P, i = r / ny, j = r % ny;
          lmin = std::min(lmin, l); lmax = std::max(lmax, l);
          imin = std::min(imin, i); imax = std::max(imax, i);
          jmin = std::min(jmin, j); jmax = std::max(jmax, j);
        }
        if (lmin != lmax || imax - imin > 1 || jmax - jmin > 1 || imin > nx - 2 || jmin > ny - 2) { ok = false; break; }
        // (a degenerate triangle on one grid line still belongs to the cell at its lower corner)
        const size_t c = ((size_t)lmin * (nx - 1) + imin) * (ny - 1) + jmin;
        if (tri[2 * c] < 0) tri[2 * c] = pos_of[t];
        else if (tri[2 * c + 1] < 0) tri[2 * c + 1] = pos_of[t];
        else ok = false;
      }
      for (size_t c = 0; c < 2 * ncell && ok; ++c) ok = tri[c] >= 0;
      if (!ok) continue;
      h_cellz.resize(2 * ncell);
      for (int l = 0; l < L; ++l)
        for (int i = 0; i < nx - 1; ++i)
          for (int j = 0; j < ny - 1; ++j) {
            float lo = FLT_MAX, hi = -FLT_MAX;
            for (int a = 0; a < 2; ++a)
              for (int b = 0; b < 2; ++b) {
                const float z = vertices[3 * ((size_t)l * P + (size_t)(i + a) * ny + (j + b)) + 2];
                lo = std::min(lo, z); hi = std::max(hi, z);
              }
            const size_t c = ((size_t)l * (nx - 1) + i) * (ny - 1) + j;
            h_cellz[2 * c] = lo; h_cellz[2 * c + 1] = hi;
          }
      h_celltri.swap(tri);
      g_layers = L; g_nx = nx; g_ny = ny;
      g_x0 = (float)x0; g_y0 = (float)y0; g_dx = (float)dx; g_dy = (float)dy;
      // footprint padding: the grid-line deviation + a thousandth of a cell + 16 ulp of the largest coordinate (fp32 slab arithmetic)
      g_pad = (float)(dev + 1e-3 * std::min(dx, dy)) + 2e-6f * ext + 1e-30f;
    }
  }

  ElgMesh* m = new ElgMesh();
  m->num_vertices = num_vertices;
  m->num_triangles = num_triangles;
  m->num_nodes = (int)wide.size();
  m->avg_edge = edge_sum / (3.0 * num_triangles);
  for (int k = 0; k < 3; ++k) { m->lo[k] = scene.lo[k]; m->hi[k] = scene.hi[k]; }
  m->nodes = nullptr;
  m->tris = nullptr;
  m->grid_layers = 0;
  m->grid_cellz = nullptr;
  m->grid_celltri = nullptr;
  cudaGetDevice(&m->device);
  if (cudaMalloc(&m->nodes, hn.size() * 4) != cudaSuccess || cudaMalloc(&m->tris, ht.size() * 4) != cudaSuccess ||
      cudaMemcpy(m->nodes, hn.data(), hn.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(m->tris, ht.data(), ht.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaGetLastError();
    if (m->nodes) cudaFree(m->nodes);
    if (m->tris) cudaFree(m->tris);
    delete m;
    return mfail(ELG_ERR_CUDA, "elg_mesh_create: cannot allocate / upload the BVH (is a CUDA device present?)");
  }
  if (g_layers > 0) {      // the accelerator is optional: a failed upload just leaves the BVH walk
    if (cudaMalloc(&m->grid_cellz, h_cellz.size() * 4) == cudaSuccess && cudaMalloc(&m->grid_celltri, h_celltri.size() * 4) == cudaSuccess &&
        cudaMemcpy(m->grid_cellz, h_cellz.data(), h_cellz.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
        cudaMemcpy(m->grid_celltri, h_celltri.data(), h_celltri.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess) {
      m->grid_layers = g_layers; m->grid_nx = g_nx; m->grid_ny = g_ny;
      m->grid_x0 = g_x0; m->grid_y0 = g_y0; m->grid_dx = g_dx; m->grid_dy = g_dy; m->grid_pad = g_pad;
    } else {
      cudaGetLastError();
      if (m->grid_cellz) cudaFree(m->grid_cellz);
      if (m->grid_celltri) cudaFree(m->grid_celltri);
      m->grid_cellz = nullptr; m->grid_celltri = nullptr;
    }
  }
  *out = m;
  return ELG_OK;
}

// Measured on B200 (profiles/README.md r2i, 1.6 M-triangle terrain at 0.1 m cells): the grid walk returns the BVH walk's hits bit for bit
// but is SLOWER -- depth camera 4.37 vs 5.45 Grays/s (far clip 2 m), 1.74 vs 3.82 (10 m), incoherent rays 0.93 vs 3.26: a ray crosses
// 10 cells per metre one by one where the 4-wide tree skips the empty air above the terrain in a few boxes.  The tree stays the
// default; elg_set_mesh_tuning(1) selects the grid walk (A/B runs, the bit-identity test).
static int g_mesh_use_grid = 0;
static elg::GridView grid_view(const ElgMesh* m) {
  elg::GridView g{};
  if (m->grid_layers > 0 && g_mesh_use_grid) {
    g.layers = m->grid_layers; g.nx = m->grid_nx; g.ny = m->grid_ny;
    g.x0 = m->grid_x0; g.y0 = m->grid_y0; g.dx = m->grid_dx; g.dy = m->grid_dy; g.pad = m->grid_pad;
    g.cellz = m->grid_cellz; g.celltri = m->grid_celltri;
  }
  return g;
}

int elg_mesh_create(const float* vertices, int32_t num_vertices, const int32_t* triangles, int32_t num_triangles, ElgMesh** out) {
  return elg_mesh_create_ex(vertices, num_vertices, triangles, num_triangles, 0, out);
}

int elg_set_mesh_tuning(int use_grid) {
  g_mesh_use_grid = use_grid != 0;
  return ELG_OK;
}

int elg_mesh_grid_info(const ElgMesh* mesh, int32_t* layers, int32_t* nx, int32_t* ny) {
  if (!mesh) return mfail(ELG_ERR_NULL_POINTER, "mesh is NULL");
  if (layers) *layers = mesh->grid_layers;
  if (nx) *nx = mesh->grid_nx;
  if (ny) *ny = mesh->grid_ny;
  return ELG_OK;
}

int elg_mesh_free(ElgMesh* mesh) {
  if (!mesh) return ELG_OK;
  cudaFree(mesh->nodes);
  cudaFree(mesh->tris);
  if (mesh->grid_cellz) cudaFree(mesh->grid_cellz);
  if (mesh->grid_celltri) cudaFree(mesh->grid_celltri);
  delete mesh;
  return ELG_OK;
}

int elg_mesh_info(const ElgMesh* mesh, int32_t* num_triangles, int32_t* num_nodes, float* bounds6) {
  if (!mesh) return mfail(ELG_ERR_NULL_POINTER, "mesh is NULL");
  if (num_triangles) *num_triangles = mesh->num_triangles;
  if (num_nodes) *num_nodes = mesh->num_nodes;
  if (bounds6)
    for (int k = 0; k < 3; ++k) { bounds6[k] = mesh->lo[k]; bounds6[3 + k] = mesh->hi[k]; }
  return ELG_OK;
}

int elg_raycast(const ElgMesh* mesh, const float* ray_origins, const float* ray_directions, int64_t num_rays, float max_dist,
                float* ray_hits, uint8_t* hits_found, float* hit_distance, int32_t* hit_triangle, void* stream) {
  if (!mesh) return mfail(ELG_ERR_NULL_POINTER, "Mesh cannot be None");
  if (num_rays < 0) return mfail(ELG_ERR_INVALID_ARGUMENT, "num_rays < 0");
  if (num_rays == 0) return ELG_OK;
  if (!ray_origins || !ray_directions || !ray_hits || !hits_found) return mfail(ELG_ERR_NULL_POINTER, "a ray buffer is NULL");
  if (!(max_dist >= 0.0f)) return mfail(ELG_ERR_INVALID_ARGUMENT, "max_dist must be >= 0");
  const int threads = 128;
  const long long blocks = (num_rays + threads - 1) / threads;
  if (blocks > 0x7fffffffLL) return mfail(ELG_ERR_UNSUPPORTED, "too many rays for one launch");
  const elg::GridView gv = grid_view(mesh);
  (gv.layers > 0 ? elg::elg_raycast_kernel<true> : elg::elg_raycast_kernel<false>)<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
      gv, mesh->nodes, mesh->tris, ray_origins, ray_directions, num_rays, max_dist, ray_hits, hits_found, hit_distance, hit_triangle);
  return elg::check_launch("elg_raycast");
}

int elg_raycast_sensor(const ElgMesh* mesh, const float* pattern_origins, const float* pattern_directions, int32_t num_rays,
                       const float* sensor_pos, const float* sensor_quat, const int64_t* env_ids, int64_t num_sensors, int yaw_only,
                       float max_dist, float* ray_hits, uint8_t* hits_found, void* stream) {
  if (!mesh) return mfail(ELG_ERR_NULL_POINTER, "Mesh cannot be None");
  if (num_rays < 0 || num_sensors < 0) return mfail(ELG_ERR_INVALID_ARGUMENT, "negative ray / sensor count");
  if (num_rays == 0 || num_sensors == 0) return ELG_OK;
  if (!pattern_origins || !pattern_directions || !sensor_pos || !sensor_quat || !ray_hits || !hits_found)
    return mfail(ELG_ERR_NULL_POINTER, "a sensor / ray buffer is NULL");
  if (!(max_dist >= 0.0f)) return mfail(ELG_ERR_INVALID_ARGUMENT, "max_dist must be >= 0");
  const int threads = 128;
  const long long blocks = (num_sensors * num_rays + threads - 1) / threads;
  if (blocks > 0x7fffffffLL) return mfail(ELG_ERR_UNSUPPORTED, "too many rays for one launch");
  const elg::GridView gv = grid_view(mesh);
  (gv.layers > 0 ? elg::elg_raycast_sensor_kernel<true> : elg::elg_raycast_sensor_kernel<false>)<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
      gv, mesh->nodes, mesh->tris, pattern_origins, pattern_directions, num_rays, sensor_pos, sensor_quat, env_ids, num_sensors, yaw_only,
      max_dist, ray_hits, hits_found, nullptr, 0, 0, nullptr, 0);
  return elg::check_launch("elg_raycast_sensor");
}

int elg_raycast_sensor_obs(const ElgMesh* mesh, const float* pattern_origins, const float* pattern_directions, int32_t num_rays,
                           const float* sensor_pos, const float* sensor_quat, const int64_t* env_ids, int64_t num_sensors, int yaw_only,
                           float max_dist, float* ray_hits, uint8_t* hits_found, const float* dist_origins, int32_t dist_origin_stride,
                           int32_t normalize, float* distances, int64_t distances_row_stride, void* stream) {
  if (!mesh) return mfail(ELG_ERR_NULL_POINTER, "Mesh cannot be None");
  if (distances_row_stride < num_rays) return mfail(ELG_ERR_INVALID_ARGUMENT, "distances_row_stride < num_rays");
  if (num_rays < 0 || num_sensors < 0) return mfail(ELG_ERR_INVALID_ARGUMENT, "negative ray / sensor count");
  if (num_rays == 0 || num_sensors == 0) return ELG_OK;
  if (!pattern_origins || !pattern_directions || !sensor_pos || !sensor_quat || !ray_hits || !hits_found || !dist_origins || !distances)
    return mfail(ELG_ERR_NULL_POINTER, "a sensor / ray / distance buffer is NULL");
  if (!(max_dist > 0.0f) || dist_origin_stride < 3) return mfail(ELG_ERR_INVALID_ARGUMENT, "max_dist must be > 0 and the origin stride >= 3");
  const int threads = 128;
  const long long blocks = (num_sensors * num_rays + threads - 1) / threads;
  if (blocks > 0x7fffffffLL) return mfail(ELG_ERR_UNSUPPORTED, "too many rays for one launch");
  const elg::GridView gv = grid_view(mesh);
  (gv.layers > 0 ? elg::elg_raycast_sensor_kernel<true> : elg::elg_raycast_sensor_kernel<false>)<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
      gv, mesh->nodes, mesh->tris, pattern_origins, pattern_directions, num_rays, sensor_pos, sensor_quat, env_ids, num_sensors, yaw_only,
      max_dist, ray_hits, hits_found, dist_origins, dist_origin_stride, normalize, distances, (long long)distances_row_stride);
  return elg::check_launch("elg_raycast_sensor_obs");
}

int elg_sizeof_cam_params(void) { return (int)sizeof(ElgCamParams); }

int elg_camera_pose(const float* sensor_pos, const float* sensor_quat, const int64_t* env_ids, int64_t num, const float* offset_pos3,
                    const float* offset_quat4, float* camera_pos, float* camera_rot, void* stream) {
  if (num < 0) return mfail(ELG_ERR_INVALID_ARGUMENT, "num < 0");
  if (num == 0) return ELG_OK;
  if (!sensor_pos || !sensor_quat || !offset_pos3 || !offset_quat4 || !camera_pos || !camera_rot) return mfail(ELG_ERR_NULL_POINTER, "a camera pose buffer is NULL");
  elg::elg_camera_pose_kernel<<<(unsigned)((num + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      sensor_pos, sensor_quat, env_ids, num, offset_pos3[0], offset_pos3[1], offset_pos3[2], offset_quat4[0], offset_quat4[1], offset_quat4[2],
      offset_quat4[3], camera_pos, camera_rot);
  return elg::check_launch("elg_camera_pose");
}

int elg_depth_camera(const ElgMesh* mesh, const ElgCamParams* cam, const float* ray_directions, const float* camera_pos, const float* camera_rot,
                     const int64_t* episode_length_buf, const float* noise_u, const int32_t* resize_x_start, const float* resize_x_weights,
                     const int32_t* resize_y_start, const float* resize_y_weights, int64_t num_envs, float* depth_buffer, float* raw_depth,
                     void* stream) {
  if (!mesh) return mfail(ELG_ERR_NULL_POINTER, "No meshes available for ray casting");
  if (!cam) return mfail(ELG_ERR_NULL_POINTER, "camera params is NULL");
  if (cam->width < 1 || cam->height < 1 || cam->out_width < 1 || cam->out_height < 1 || cam->buffer_len < 1)
    return mfail(ELG_ERR_INVALID_ARGUMENT, "image sizes and buffer_len must be >= 1");
  if (!cam->resize && (cam->width != cam->out_width || cam->height != cam->out_height))
    return mfail(ELG_ERR_INVALID_ARGUMENT, "original and resized image sizes differ but no resize tables were given");
  if (cam->resize && (!resize_x_start || !resize_x_weights || !resize_y_start || !resize_y_weights || cam->max_taps < 1))
    return mfail(ELG_ERR_NULL_POINTER, "resize needs the four tap tables");
  if (!(cam->far_clip > cam->near_clip)) return mfail(ELG_ERR_INVALID_ARGUMENT, "far_clip must exceed near_clip");
  if (num_envs < 0) return mfail(ELG_ERR_INVALID_ARGUMENT, "num_envs < 0");
  if (num_envs == 0) return ELG_OK;
  if (!ray_directions || !camera_pos || !camera_rot || !episode_length_buf || !depth_buffer) return mfail(ELG_ERR_NULL_POINTER, "a camera buffer is NULL");
  const size_t in_px = (size_t)cam->width * cam->height, out_px = (size_t)cam->out_width * cam->out_height;
  const size_t smem = 4 * ((in_px > out_px ? in_px : out_px) + (cam->resize ? (size_t)cam->height * cam->out_width : 0));
  if (smem > 200 * 1024) return mfail(ELG_ERR_UNSUPPORTED, "image too large for the fused depth kernel (> 200 KB of shared memory)");
  static elg::SmemCache smem_cache = {};
  size_t& smem_set = elg::smem_slot(smem_cache);
  if (smem > 48 * 1024 && smem > smem_set) {
    if (cudaFuncSetAttribute(elg::elg_depth_camera_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(elg::elg_depth_camera_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return mfail(ELG_ERR_CUDA, "cannot reserve shared memory for elg_depth_camera_kernel");
    smem_set = smem;
  }
  const elg::GridView gv = grid_view(mesh);
  (gv.layers > 0 ? elg::elg_depth_camera_kernel<true> : elg::elg_depth_camera_kernel<false>)<<<(unsigned)num_envs, 256, smem, (cudaStream_t)stream>>>(
      gv, mesh->nodes, mesh->tris, *cam, ray_directions, camera_pos, camera_rot, episode_length_buf, noise_u, resize_x_start, resize_x_weights,
      resize_y_start, resize_y_weights, depth_buffer, raw_depth);
  return elg::check_launch("elg_depth_camera");
}

int elg_sdf_query(const ElgMesh* mesh, const float* points, int64_t num_points, float max_distance, float epsilon, float* sdf, float* grad,
                  float* closest_points, int32_t* closest_face, void* stream) {
  if (!mesh) return mfail(ELG_ERR_NULL_POINTER, "No meshes available for SDF queries");
  if (num_points < 0) return mfail(ELG_ERR_INVALID_ARGUMENT, "num_points < 0");
  if (num_points == 0) return ELG_OK;
  if (!points || !sdf || !grad) return mfail(ELG_ERR_NULL_POINTER, "points/sdf/grad is NULL");
  if (!(max_distance >= 0.0f) || !(epsilon >= 0.0f)) return mfail(ELG_ERR_INVALID_ARGUMENT, "max_distance and epsilon must be >= 0");
  const int threads = 128;
  const long long blocks = (num_points + threads - 1) / threads;
  if (blocks > 0x7fffffffLL) return mfail(ELG_ERR_UNSUPPORTED, "too many points for one launch");
  elg::elg_sdf_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(mesh->nodes, mesh->tris, points, num_points, max_distance,
                                                                              (double)epsilon * mesh->avg_edge, sdf, grad, closest_points,
                                                                              closest_face);
  return elg::check_launch("elg_sdf_query");
}

int elg_sdf_query_bodies(const ElgMesh* mesh, const float* rigid_body_state, int32_t num_bodies, const int32_t* body_indices,
                         const float* sphere_offsets, int32_t num_query_bodies, const int64_t* env_ids, int64_t num_rows, float max_distance,
                         float epsilon, float* sdf, int64_t sdf_row_stride, float* grad, float* nearest_points, float* query_points, void* stream) {
  if (!mesh) return mfail(ELG_ERR_NULL_POINTER, "No meshes available for SDF queries");
  if (num_rows < 0 || num_bodies < 1 || num_query_bodies < 0 || num_query_bodies > elg::kMaxSdfBodies)
    return mfail(ELG_ERR_INVALID_ARGUMENT, "bad body counts (at most 16 query bodies)");
  if (num_rows == 0 || num_query_bodies == 0) return ELG_OK;
  if (!rigid_body_state || !body_indices || !sdf) return mfail(ELG_ERR_NULL_POINTER, "an SDF body buffer is NULL");
  if (sdf_row_stride < num_query_bodies) return mfail(ELG_ERR_INVALID_ARGUMENT, "sdf_row_stride < num_query_bodies");
  if (!(max_distance >= 0.0f) || !(epsilon >= 0.0f)) return mfail(ELG_ERR_INVALID_ARGUMENT, "max_distance and epsilon must be >= 0");
  elg::SdfBodies sb{};
  sb.num_bodies = num_bodies;
  sb.count = num_query_bodies;
  for (int k = 0; k < num_query_bodies; ++k) {
    if (body_indices[k] < 0 || body_indices[k] >= num_bodies) return mfail(ELG_ERR_INVALID_ARGUMENT, "query body index out of range");
    sb.body_idx[k] = body_indices[k];
    sb.has_offset[k] = 0;
    if (sphere_offsets && sphere_offsets[3 * k] == sphere_offsets[3 * k]) {   // NaN in the first component: no offset for this body
      sb.has_offset[k] = 1;
      for (int c = 0; c < 3; ++c) sb.offset[k][c] = sphere_offsets[3 * k + c];
    }
  }
  const long long total = num_rows * num_query_bodies;
  const int threads = 128;
  elg::elg_sdf_bodies_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      mesh->nodes, mesh->tris, rigid_body_state, sb, env_ids, num_rows, max_distance, (double)epsilon * mesh->avg_edge, sdf,
      (long long)sdf_row_stride, grad, nearest_points, query_points);
  return elg::check_launch("elg_sdf_query_bodies");
}

double elg_mesh_mean_edge(const ElgMesh* mesh) { return mesh ? mesh->avg_edge : 0.0; }

}  // extern "C"
