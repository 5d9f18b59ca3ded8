// Brute-force signed distance from a batch of points to a triangle mesh.
//
// What it replaces: trimesh.proximity.signed_distance / closest_point as used by the reference's obstacle-free tube
// (drone3d/obstacles/mesh_obstacle.py:38-42 signed distance, :110-145 largest-sphere search: 800 collocation points x
// (1 + 5 radii x 8 angles) samples against the 8884 faces of the arena mesh).  trimesh walks an r-tree on the CPU; here
// every point visits every triangle: 33 k points x 8.9 k faces = 0.3 G point-triangle tests, a few milliseconds.
//
// One thread per point; triangles stream through shared memory in tiles (coalesced loads, every triangle read once
// per CTA).  Distance: closest point on the triangle by the Voronoi-region walk (Ericson, Real-Time Collision Detection
// 5.1.5).  Sign: parity of the crossings of a ray from the point (fixed, generic direction) with the mesh
// (Moeller-Trumbore) -- inside = odd, like trimesh's `contains`.  Output sign convention of the reference
// (mesh_obstacle.py:38-42): positive OUTSIDE the mesh.
#pragma once
#include "common.cuh"

#define RB_SDF_THREADS 128
#define RB_SDF_TILE 128   // triangles per shared-memory tile

struct RbVec3 {
  double x, y, z;
};
__device__ __forceinline__ RbVec3 v3(double x, double y, double z) { return RbVec3{x, y, z}; }
__device__ __forceinline__ RbVec3 operator-(RbVec3 a, RbVec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ RbVec3 operator+(RbVec3 a, RbVec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ RbVec3 operator*(RbVec3 a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ double dot3(RbVec3 a, RbVec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ RbVec3 cross3(RbVec3 a, RbVec3 b) {
  return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// closest point to p on triangle (a, b, c)
__device__ __forceinline__ RbVec3 closest_on_triangle(RbVec3 p, RbVec3 a, RbVec3 b, RbVec3 c) {
  const RbVec3 ab = b - a, ac = c - a, ap = p - a;
  const double d1 = dot3(ab, ap), d2 = dot3(ac, ap);
  if (d1 <= 0.0 && d2 <= 0.0) return a;
  const RbVec3 bp = p - b;
  const double d3 = dot3(ab, bp), d4 = dot3(ac, bp);
  if (d3 >= 0.0 && d4 <= d3) return b;
  const double vc = d1 * d4 - d3 * d2;
  if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) return a + ab * (d1 / (d1 - d3));
  const RbVec3 cp = p - c;
  const double d5 = dot3(ab, cp), d6 = dot3(ac, cp);
  if (d6 >= 0.0 && d5 <= d6) return c;
  const double vb = d5 * d2 - d1 * d6;
  if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) return a + ac * (d2 / (d2 - d6));
  const double va = d3 * d6 - d5 * d4;
  if (va <= 0.0 && (d4 - d3) >= 0.0 && (d5 - d6) >= 0.0) return b + (c - b) * ((d4 - d3) / ((d4 - d3) + (d5 - d6)));
  const double denom = 1.0 / (va + vb + vc);
  return a + ab * (vb * denom) + ac * (vc * denom);
}

// does the ray p + t dir (t > 0) cross triangle (a, b, c)?
__device__ __forceinline__ bool ray_hits(RbVec3 p, RbVec3 dir, RbVec3 a, RbVec3 b, RbVec3 c) {
  const RbVec3 e1 = b - a, e2 = c - a;
  const RbVec3 h = cross3(dir, e2);
  const double det = dot3(e1, h);
  if (fabs(det) < 1e-300) return false;
  const double inv = 1.0 / det;
  const RbVec3 s = p - a;
  const double u = dot3(s, h) * inv;
  if (u < 0.0 || u > 1.0) return false;
  const RbVec3 q = cross3(s, e1);
  const double v = dot3(dir, q) * inv;
  if (v < 0.0 || u + v > 1.0) return false;
  return dot3(e2, q) * inv > 0.0;
}

// tri [nt][9] (three vertices), pts [np][3]; dist [np] signed (positive outside), closest [np][3] (optional)
__global__ void __launch_bounds__(RB_SDF_THREADS)
mesh_sdf_kernel(const double* __restrict__ tri, int nt, const double* __restrict__ pts, int np,
                double* __restrict__ dist, double* __restrict__ closest) {
  __shared__ double tile[RB_SDF_TILE * 9];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < np;
  const RbVec3 p = live ? v3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]) : v3(0, 0, 0);
  const RbVec3 dir = v3(0.5773502691896258, 0.7071067811865476 * 0.8164965809277260, 0.40824829046386296);   // generic
  double best = 1e300;
  RbVec3 bestc = p;
  int crossings = 0;
  for (int t0 = 0; t0 < nt; t0 += RB_SDF_TILE) {
    const int nn = nt - t0 < RB_SDF_TILE ? nt - t0 : RB_SDF_TILE;
    __syncthreads();
    for (int k = threadIdx.x; k < nn * 9; k += blockDim.x) tile[k] = tri[(size_t)t0 * 9 + k];
    __syncthreads();
    if (live) {
      for (int k = 0; k < nn; ++k) {
        const double* q = tile + 9 * k;
        const RbVec3 a = v3(q[0], q[1], q[2]), b = v3(q[3], q[4], q[5]), c = v3(q[6], q[7], q[8]);
        const RbVec3 cp = closest_on_triangle(p, a, b, c);
        const RbVec3 d = p - cp;
        const double d2 = dot3(d, d);
        if (d2 < best) {
          best = d2;
          bestc = cp;
        }
        crossings += ray_hits(p, dir, a, b, c) ? 1 : 0;
      }
    }
  }
  if (live) {
    const double r = sqrt(best);
    dist[i] = (crossings & 1) ? -r : r;
    if (closest) {
      closest[3 * i] = bestc.x;
      closest[3 * i + 1] = bestc.y;
      closest[3 * i + 2] = bestc.z;
    }
  }
}
