// mesh_host_harness.cu -- TEST INFRASTRUCTURE (not shipped, not linked into libq3dr_raycast.so).
//
// Compiles the product's mesh BVH builder (quad3dr_b200/csrc/mesh_build.h) and its traversal (mesh_cast in
// mesh_kernels.cuh, a __host__ __device__ function) for the HOST, so that the tree-independence of the traversal --
// padded boxes, slack on the pruning bound, tie rule -- can be checked against the oracle's brute force over all
// triangles on a machine without a GPU (tests/test_mesh_normals.py).  The GPU tests check the same on the device.
#include <cstdint>
#include <vector>

#include "../../quad3dr_b200/csrc/mesh_build.h"
#include "../../quad3dr_b200/csrc/mesh_kernels.cuh"

extern "C" int mesh_host_cast_pixels(const float* tri_vertices, const float* tri_normals, size_t nf, const float K[4],
                                     uint32_t width, uint32_t height, const float* Rt, float near_plane,
                                     float far_plane, uint32_t clear_argb, const uint32_t* xs, const uint32_t* ys,
                                     size_t n, uint32_t* out_argb, uint32_t* tree_depth, uint64_t* n_nodes) {
  q3dr::MeshHost hm;
  q3dr::build_mesh_host(tri_vertices, tri_normals, nf, &hm);
  if (hm.depth + 2 > (uint32_t)q3dr::kMeshStack) return 1;
  q3dr::MeshDev md;
  md.nodes = hm.nodes.data();
  md.tris = hm.tris.data();
  md.normals = hm.normals.data();
  md.n_tris = (uint32_t)nf;
  md.root_ref = hm.root_ref;
  md.pad = hm.pad;
  q3dr::RenderOptsDev ro;
  ro.near_plane = near_plane;
  ro.far_plane = far_plane;
  ro.clear_argb = clear_argb;
#pragma omp parallel for schedule(dynamic, 256)
  for (long i = 0; i < (long)n; ++i) {
    q3dr::MeshRay r;
    q3dr::mesh_camera_ray(Rt, K[0], K[1], K[2], K[3], width, height, xs[i], ys[i], r);
    out_argb[i] = q3dr::mesh_cast(md, ro, r);
  }
  if (tree_depth) *tree_depth = hm.depth;
  if (n_nodes) *n_nodes = hm.nodes.size() / 4;
  return 0;
}

// traversal statistics of the same rays: inner nodes visited, leaves and triangles tested (sums over the rays)
extern "C" int mesh_host_stats(const float* tri_vertices, size_t nf, const float K[4], uint32_t width, uint32_t height,
                               const float* Rt, float near_plane, float far_plane, const uint32_t* xs,
                               const uint32_t* ys, size_t n, uint64_t out[3]) {
  q3dr::MeshHost hm;
  q3dr::build_mesh_host(tri_vertices, nullptr, nf, &hm);
  q3dr::MeshDev md;
  md.nodes = hm.nodes.data();
  md.tris = hm.tris.data();
  md.normals = hm.normals.data();
  md.n_tris = (uint32_t)nf;
  md.root_ref = hm.root_ref;
  md.pad = hm.pad;
  q3dr::RenderOptsDev ro;
  ro.near_plane = near_plane;
  ro.far_plane = far_plane;
  ro.clear_argb = 0xFFFFFFFFu;
  uint64_t inner = 0, leaves = 0, tris = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : inner, leaves, tris)
  for (long i = 0; i < (long)n; ++i) {
    q3dr::MeshRay r;
    q3dr::mesh_camera_ray(Rt, K[0], K[1], K[2], K[3], width, height, xs[i], ys[i], r);
    int32_t stack[q3dr::kMeshStack];
    float stack_t[q3dr::kMeshStack];
    q3dr::MeshTrav s;
    q3dr::mesh_trav_init(s, md, ro, r);
    while (true) {
      while (s.ref >= 0 && s.ref != q3dr::kMeshDone) {
        q3dr::mesh_trav_inner(s, md, stack, stack_t);
        ++inner;
      }
      if (s.ref == q3dr::kMeshDone) break;
      ++leaves;
      tris += ((uint32_t)(~s.ref) & 3u) + 1u;
      q3dr::mesh_trav_leaf(s, md, ro, stack, stack_t);
    }
  }
  out[0] = inner;
  out[1] = leaves;
  out[2] = tris;
  return 0;
}
