// mesh_kernels.cuh -- image-space surface normals from a triangle mesh.
//
// The reference takes the normal that scores a voxel from an off-screen OpenGL render of the Poisson mesh
// (enable_opengl = true, the planner's default): ViewpointOffscreenRenderer::drawPoissonMeshNormals
// (viewpoint_offscreen_renderer.cpp:361-397) draws every triangle with the projection of :276-293, depth test on,
// no culling, clear colour (1,1,1,1); the fragment shader writes 0.5 * n + 0.5 of the interpolated vertex normal
// (shaders/triangles_normals.{v,f}.glsl) into an 8-bit RGBA target; computePoissonMeshNormalVector (:499-527) reads
// the pixel back through QImage and decodes it.  ViewpointPlanner::computeNormalVector
// (viewpoint_planner_scoring.cpp:26-38) looks the voxel's KEPT pixel up in that image.
//
// Here the image is defined by casting, for pixel (x, y), the ray through the pixel CENTRE as that projection maps it
//     c = ( ((x + 0.5) * 2 / W - 1) / (fx / cx),  ((y + 0.5) * 2 / H - 1) / (fy / cy),  1 ),   d = R c   (not normalised)
// against every triangle (Moeller-Trumbore in binary32, fixed operation order, no FMA): the ray parameter t is then
// the camera-space depth, accepted inside [near, far] (the clip planes); the smallest t wins, equal t goes to the
// LOWER triangle index (GL_LESS keeps the fragment drawn first); the normal is the barycentric (= perspective
// correct) blend of the three vertex normals, encoded like the shader + an 8-bit UNORM target (round to nearest) into
// QImage's 0xAARRGGBB; no hit = the clear colour.  This is a ray-cast restatement of a rasteriser: it agrees with
// OpenGL except at pixels whose centre lies within rounding distance of a triangle edge or of a depth tie, where GL
// implementations differ among themselves (fill rule on snapped sub-pixel coordinates, depth-buffer precision).
//
// Kernels:
//   mesh_render_kernel        one thread per pixel of an n x H x W image batch (8x4 pixel footprints per warp)
//   mesh_kept_normals_kernel  (score_kernels.cuh) one thread per entry of the chunk's hit lists: the code of the
//                             entry's kept pixel only (~7 % of the pixels of a viewpoint), written next to the list
//                             for score_blocks_kernel
// Both walk a binary BVH over the triangles (64-byte nodes holding both children's boxes, built on the host by
// mesh_build.h).  The result is independent of the tree: boxes are padded and a subtree is skipped only when it lies
// beyond the best t by more than kMeshSlack, far above the rounding error of t.
#pragma once

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

namespace q3dr {

constexpr int kMeshStack = 40;            // median-split tree over <= 2^31 triangles, <= 4 per leaf
constexpr float kMeshSlackRel = 1e-3f;    // subtree skipped iff entry t > best t * (1 + kMeshSlackRel) + pad
constexpr uint32_t kMeshLeafMaxTris = 4;

struct MeshDev {
  const float4* __restrict__ nodes;    // [n_nodes][4]: layout in mesh_build.h
  const float4* __restrict__ tris;     // [nf][3] in BVH order: {v0 xyz, bits(original index)}, {e1 xyz, 0}, {e2 xyz, 0}
  const float4* __restrict__ normals;  // [nf][3] by ORIGINAL index: n0, n1, n2
  uint32_t n_tris;
  int32_t root_ref;                    // >= 0 node index, < 0 leaf (a mesh of <= 4 triangles)
  float pad;                           // absolute padding already applied to the boxes
};

struct RenderOptsDev {
  float near_plane, far_plane;
  uint32_t clear_argb;
};

struct MeshRay {
  float ox, oy, oz, dx, dy, dz;
};

// pixel-centre ray of the reference's projection matrix (viewpoint_offscreen_renderer.cpp:276-293) -- see header
__host__ __device__ __forceinline__ void mesh_camera_ray(const float* rt, float fx, float fy, float cx, float cy,
                                                         uint32_t width, uint32_t height, uint32_t x, uint32_t y,
                                                         MeshRay& r) {
  const float ndc_x = ((float)x + 0.5f) * 2.0f / (float)width - 1.0f;
  const float ndc_y = ((float)y + 0.5f) * 2.0f / (float)height - 1.0f;
  const float c0 = ndc_x / (fx / cx);
  const float c1 = ndc_y / (fy / cy);
  r.dx = rt[0] * c0 + (rt[1] * c1 + rt[2]);
  r.dy = rt[4] * c0 + (rt[5] * c1 + rt[6]);
  r.dz = rt[8] * c0 + (rt[9] * c1 + rt[10]);
  r.ox = rt[3];
  r.oy = rt[7];
  r.oz = rt[11];
}

// Moeller-Trumbore; (u, v, t) valid when it returns true.  Three-term sums associate as a0 + (a1 + a2).
__host__ __device__ __forceinline__ bool mesh_tri_test(const MeshRay& r, float v0x, float v0y, float v0z, float e1x,
                                                       float e1y, float e1z, float e2x, float e2y, float e2z,
                                                       float near_t, float far_t, float& u, float& v, float& t) {
  const float px = r.dy * e2z - r.dz * e2y;
  const float py = r.dz * e2x - r.dx * e2z;
  const float pz = r.dx * e2y - r.dy * e2x;
  const float det = e1x * px + (e1y * py + e1z * pz);
  if (!(det != 0.0f)) return false;  // parallel (or NaN)
  const float inv = 1.0f / det;
  const float sx = r.ox - v0x, sy = r.oy - v0y, sz = r.oz - v0z;
  u = (sx * px + (sy * py + sz * pz)) * inv;
  if (!(u >= 0.0f && u <= 1.0f)) return false;
  const float qx = sy * e1z - sz * e1y;
  const float qy = sz * e1x - sx * e1z;
  const float qz = sx * e1y - sy * e1x;
  v = (r.dx * qx + (r.dy * qy + r.dz * qz)) * inv;
  if (!(v >= 0.0f && u + v <= 1.0f)) return false;
  t = (e2x * qx + (e2y * qy + e2z * qz)) * inv;
  return t >= near_t && t <= far_t;
}

__host__ __device__ __forceinline__ uint32_t mesh_unorm8(float n) {
  float c = 0.5f * n + 0.5f;  // shaders/triangles_normals.f.glsl
  if (!(c == c)) c = 0.0f;
  c = c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c);
#ifdef __CUDA_ARCH__
  return (uint32_t)__float2int_rn(c * 255.0f);
#else
  return (uint32_t)nearbyintf(c * 255.0f);
#endif
}

// interpolated normal -> 0xAARRGGBB
__host__ __device__ __forceinline__ uint32_t mesh_encode_normal(float u, float v, const float* n0, const float* n1,
                                                                const float* n2) {
  const float w0 = (1.0f - u) - v;
  const float nx = (w0 * n0[0] + u * n1[0]) + v * n2[0];
  const float ny = (w0 * n0[1] + u * n1[1]) + v * n2[1];
  const float nz = (w0 * n0[2] + u * n1[2]) + v * n2[2];
  return 0xFF000000u | (mesh_unorm8(nx) << 16) | (mesh_unorm8(ny) << 8) | mesh_unorm8(nz);
}

// computePoissonMeshNormalVector's decode (viewpoint_offscreen_renderer.cpp:519-526): QColor channels / 255,
// 2 (v - 0.5), zero when the squared norm is below 0.5, Eigen normalize() (divides only when the norm is > 0).
__host__ __device__ __forceinline__ void decode_normal_code(uint32_t argb, float& nx, float& ny, float& nz) {
  const float r = (float)((argb >> 16) & 255u) / 255.0f;
  const float g = (float)((argb >> 8) & 255u) / 255.0f;
  const float b = (float)(argb & 255u) / 255.0f;
  nx = 2.0f * (r - 0.5f);
  ny = 2.0f * (g - 0.5f);
  nz = 2.0f * (b - 0.5f);
  float sq = nx * nx + (ny * ny + nz * nz);
  if (sq < 0.5f) {
    nx = ny = nz = 0.0f;
    sq = 0.0f;
  }
  if (sq > 0.0f) {
#ifdef __CUDA_ARCH__
    const float n = __fsqrt_rn(sq);
#else
    const float n = sqrtf(sq);
#endif
    nx = nx / n;
    ny = ny / n;
    nz = nz / n;
  }
}

#ifdef __CUDA_ARCH__
#define Q3DR_MESH_LDG(p) __ldg(p)
#else
#define Q3DR_MESH_LDG(p) (*(p))  // tests/cpp/mesh_host_harness.cu runs this traversal on the host against the oracle
#endif

__host__ __device__ __forceinline__ int32_t mesh_bits_i32(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_int(f);
#else
  int32_t i;
  memcpy(&i, &f, 4);
  return i;
#endif
}

// entry parameter of the (padded) box along the ray, or +inf when the ray misses it or leaves it before t_lo
__host__ __device__ __forceinline__ float mesh_box_entry(const MeshRay& r, float ix, float iy, float iz, float mnx, float mny,
                                                float mnz, float mxx, float mxy, float mxz, float t_lo, float t_hi) {
  const float ax = (mnx - r.ox) * ix, bx = (mxx - r.ox) * ix;
  const float ay = (mny - r.oy) * iy, by = (mxy - r.oy) * iy;
  const float az = (mnz - r.oz) * iz, bz = (mxz - r.oz) * iz;
  // fminf / fmaxf drop a NaN operand (0 * inf: origin exactly on a slab plane of an axis the ray is parallel to),
  // which keeps the test conservative
  const float tmin = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), t_lo));
  const float tmax = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), t_hi));
  return (tmin <= tmax) ? tmin : FLT_MAX * 2.0f;
}

// ---- traversal, as steps over an explicit state so that a kernel can interleave rays (mesh_kept_normals_kernel
// hands a finished lane its next ray while the other lanes of the warp go on) ----
constexpr int32_t kMeshDone = 0x7FFFFFFF;

struct MeshTrav {
  MeshRay r;
  float ix, iy, iz;
  float best_t, best_u, best_v;
  uint32_t best_tri;
  float t_lo, t_hi;
  int32_t ref;  // >= 0 (and != kMeshDone): inner node to visit; < 0: leaf to test; kMeshDone: finished
  int sp;
};

__host__ __device__ __forceinline__ void mesh_trav_init(MeshTrav& s, const MeshDev& m, const RenderOptsDev& o,
                                                        const MeshRay& r) {
  s.r = r;
  s.ix = 1.0f / r.dx;
  s.iy = 1.0f / r.dy;
  s.iz = 1.0f / r.dz;
  s.best_t = FLT_MAX * 2.0f;
  s.best_u = s.best_v = 0.0f;
  s.best_tri = 0xFFFFFFFFu;
  s.t_lo = o.near_plane - (o.near_plane * kMeshSlackRel + m.pad);
  s.t_hi = o.far_plane + (o.far_plane * kMeshSlackRel + m.pad);
  s.ref = m.root_ref;
  s.sp = 0;
}

// next subtree from the stack; a stacked subtree carries its entry parameter and is dropped without a fetch when a
// closer hit has been found since it was pushed
__host__ __device__ __forceinline__ void mesh_trav_pop(MeshTrav& s, const int32_t* stack, const float* stack_t) {
  s.ref = kMeshDone;
  while (s.sp > 0) {
    --s.sp;
    if (stack_t[s.sp] <= s.t_hi) {
      s.ref = stack[s.sp];
      break;
    }
  }
}

// visits inner node s.ref: nearer child next, the other one stacked
__host__ __device__ __forceinline__ void mesh_trav_inner(MeshTrav& s, const MeshDev& m, int32_t* stack, float* stack_t) {
  const float4* nd = m.nodes + 4 * (size_t)s.ref;
  const float4 a = Q3DR_MESH_LDG(nd), b = Q3DR_MESH_LDG(nd + 1), c = Q3DR_MESH_LDG(nd + 2), d = Q3DR_MESH_LDG(nd + 3);
  const float tl = mesh_box_entry(s.r, s.ix, s.iy, s.iz, a.x, a.y, a.z, a.w, b.x, b.y, s.t_lo, s.t_hi);
  const float tr = mesh_box_entry(s.r, s.ix, s.iy, s.iz, b.z, b.w, c.x, c.y, c.z, c.w, s.t_lo, s.t_hi);
  const int32_t lref = mesh_bits_i32(d.x), rref = mesh_bits_i32(d.y);
  const bool hl = tl <= FLT_MAX, hr = tr <= FLT_MAX;
  if (hl && hr) {
    const bool left_first = tl <= tr;
    stack[s.sp] = left_first ? rref : lref;
    stack_t[s.sp] = left_first ? tr : tl;
    ++s.sp;
    s.ref = left_first ? lref : rref;
  } else if (hl || hr) {
    s.ref = hl ? lref : rref;
  } else {
    mesh_trav_pop(s, stack, stack_t);
  }
}

// tests the triangles of leaf s.ref, then pops
__host__ __device__ __forceinline__ void mesh_trav_leaf(MeshTrav& s, const MeshDev& m, const RenderOptsDev& o,
                                                        const int32_t* stack, const float* stack_t) {
  const uint32_t code = (uint32_t)(~s.ref);
  const uint32_t first = code >> 2, cnt = (code & 3u) + 1u;
  for (uint32_t k = 0; k < cnt; ++k) {
    const float4* tp = m.tris + 3 * (size_t)(first + k);
    const float4 v0 = Q3DR_MESH_LDG(tp), e1 = Q3DR_MESH_LDG(tp + 1), e2 = Q3DR_MESH_LDG(tp + 2);
    float u, v, t;
    if (mesh_tri_test(s.r, v0.x, v0.y, v0.z, e1.x, e1.y, e1.z, e2.x, e2.y, e2.z, o.near_plane, o.far_plane, u, v, t)) {
      const uint32_t idx = (uint32_t)mesh_bits_i32(v0.w);
      if (t < s.best_t || (t == s.best_t && idx < s.best_tri)) {
        s.best_t = t;
        s.best_u = u;
        s.best_v = v;
        s.best_tri = idx;
        s.t_hi = fminf(s.t_hi, t + (t * kMeshSlackRel + m.pad));
      }
    }
  }
  mesh_trav_pop(s, stack, stack_t);
}

__host__ __device__ __forceinline__ uint32_t mesh_trav_result(const MeshTrav& s, const MeshDev& m, const RenderOptsDev& o) {
  if (s.best_tri == 0xFFFFFFFFu) return o.clear_argb;
  const float4* np = m.normals + 3 * (size_t)s.best_tri;
  const float4 n0 = Q3DR_MESH_LDG(np), n1 = Q3DR_MESH_LDG(np + 1), n2 = Q3DR_MESH_LDG(np + 2);
  const float a0[3] = {n0.x, n0.y, n0.z}, a1[3] = {n1.x, n1.y, n1.z}, a2[3] = {n2.x, n2.y, n2.z};
  return mesh_encode_normal(s.best_u, s.best_v, a0, a1, a2);
}

// closest triangle of the definition in the header; returns the ARGB code.
// "while-while" traversal: a lane first descends through inner nodes until it holds a leaf (or is done), then all
// lanes of the warp test their leaf's triangles together.
__host__ __device__ __forceinline__ uint32_t mesh_cast(const MeshDev& m, const RenderOptsDev& o, const MeshRay& r) {
  int32_t stack[kMeshStack];
  float stack_t[kMeshStack];
  MeshTrav s;
  mesh_trav_init(s, m, o, r);
  while (true) {
    while (s.ref >= 0 && s.ref != kMeshDone) mesh_trav_inner(s, m, stack, stack_t);
    if (s.ref == kMeshDone) break;
    mesh_trav_leaf(s, m, o, stack, stack_t);
  }
  return mesh_trav_result(s, m, o);
}

#ifdef __CUDACC__

struct MeshRenderParams {
  MeshDev mesh;
  RenderOptsDev opts;
  const float* __restrict__ rt;  // [n][12]
  float fx, fy, cx, cy;
  uint32_t width, height;
  uint32_t tiles_x, tiles_y;     // 8x4-pixel warp footprints per image
  uint32_t* out;                 // [n][height][width]
};

// grid = (ceil(tiles / 8), n): a CTA of 8 warps renders 8 consecutive footprints of one image
static __global__ void __launch_bounds__(256) mesh_render_kernel(const MeshRenderParams p) {
  const uint32_t vp = blockIdx.y;
  const uint32_t tile = blockIdx.x * 8u + (threadIdx.x >> 5);
  if (tile >= p.tiles_x * p.tiles_y) return;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t x = (tile % p.tiles_x) * 8u + (lane & 7u);
  const uint32_t y = (tile / p.tiles_x) * 4u + (lane >> 3);
  if (x >= p.width || y >= p.height) return;
  const float* rt = p.rt + 12 * (size_t)vp;
  float m[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) m[i] = __ldg(rt + i);
  MeshRay r;
  mesh_camera_ray(m, p.fx, p.fy, p.cx, p.cy, p.width, p.height, x, y, r);
  p.out[((size_t)vp * p.height + y) * p.width + x] = mesh_cast(p.mesh, p.opts, r);
}

#endif  // __CUDACC__

}  // namespace q3dr
