// mesh_build.h -- host-side BVH (binned SAH, object median as the fallback) over the triangles of a q3dr_mesh (see mesh_kernels.cuh for the
// traversal and q3dr_mesh.cu for the upload).
#pragma once

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "mesh_kernels.cuh"

namespace q3dr {

struct Box3 {
  float mn[3], mx[3];
  void reset() {
    for (int i = 0; i < 3; ++i) {
      mn[i] = FLT_MAX;
      mx[i] = -FLT_MAX;
    }
  }
  void grow(const float* p) {
    for (int i = 0; i < 3; ++i) {
      mn[i] = std::min(mn[i], p[i]);
      mx[i] = std::max(mx[i], p[i]);
    }
  }
  void grow(const Box3& b) {
    for (int i = 0; i < 3; ++i) {
      mn[i] = std::min(mn[i], b.mn[i]);
      mx[i] = std::max(mx[i], b.mx[i]);
    }
  }
};

// 64-byte inner node: both children's (padded) boxes + their references.
//   a = (l.min xyz, l.max x)  b = (l.max yz, r.min xy)  c = (r.min z, r.max xyz)  d = (bits lref, bits rref, 0, 0)
// reference >= 0: inner node index; < 0: leaf, ~ref = first triangle (BVH order) << 2 | (count - 1), count <= 4
struct MeshBuilder {
  const float* v;  // nf x 9
  std::vector<Box3> tri_box;
  std::vector<float> centroid;  // nf x 3
  std::vector<uint32_t> order;
  std::vector<float4> nodes;
  uint32_t depth = 0;
  float pad = 0.f;
  static constexpr int kExactSweep = 32;
  size_t leaf_max = q3dr::kMeshLeafMaxTris;  // Q3DR_MESH_LEAF=1..4 (for measurements)
  bool use_sah = true;  // Q3DR_MESH_BUILD=median switches back to the object-median split (for measurements)

  static float half_area(const Box3& b) {
    const float dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
    return dx * dy + dy * dz + dz * dx;
  }

  Box3 padded(Box3 b) const {
    for (int i = 0; i < 3; ++i) {
      b.mn[i] = b.mn[i] - (pad + 1e-6f * std::fabs(b.mn[i]));
      b.mx[i] = b.mx[i] + (pad + 1e-6f * std::fabs(b.mx[i]));
    }
    return b;
  }

  int32_t build(size_t begin, size_t end, uint32_t level, Box3* box_out) {
    depth = std::max(depth, level);
    Box3 box;
    box.reset();
    for (size_t i = begin; i < end; ++i) box.grow(tri_box[order[i]]);
    *box_out = box;
    const size_t cnt = end - begin;
    if (cnt <= leaf_max) return ~(int32_t)((begin << 2) | (cnt - 1));
    float cmn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, cmx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (size_t i = begin; i < end; ++i) {
      const float* c = &centroid[3 * (size_t)order[i]];
      for (int a = 0; a < 3; ++a) {
        cmn[a] = std::min(cmn[a], c[a]);
        cmx[a] = std::max(cmx[a], c[a]);
      }
    }
    int axis = 0;
    if (cmx[1] - cmn[1] > cmx[axis] - cmn[axis]) axis = 1;
    if (cmx[2] - cmn[2] > cmx[axis] - cmn[axis]) axis = 2;
    size_t mid = begin + cnt / 2;
    bool split_done = false;
    if (use_sah && cnt <= (size_t)kExactSweep) {
      // small ranges: every cut position along every axis (sorted by centroid)
      float best_cost = FLT_MAX;
      int best_axis = -1;
      size_t best_k = 0;
      uint32_t tmp[kExactSweep];
      float ra[kExactSweep];
      for (int a = 0; a < 3; ++a) {
        for (size_t i = 0; i < cnt; ++i) tmp[i] = order[begin + i];
        std::sort(tmp, tmp + cnt, [&](uint32_t x, uint32_t y) {
          const float cx = centroid[3 * (size_t)x + a], cy = centroid[3 * (size_t)y + a];
          return cx < cy || (cx == cy && x < y);
        });
        Box3 acc;
        acc.reset();
        for (size_t i = cnt; i-- > 1;) {
          acc.grow(tri_box[tmp[i]]);
          ra[i] = half_area(acc);
        }
        acc.reset();
        for (size_t k = 1; k < cnt; ++k) {  // left = first k
          acc.grow(tri_box[tmp[k - 1]]);
          const float cost = half_area(acc) * (float)k + ra[k] * (float)(cnt - k);
          if (cost < best_cost) {
            best_cost = cost;
            best_axis = a;
            best_k = k;
          }
        }
      }
      if (best_axis >= 0) {
        std::sort(order.begin() + begin, order.begin() + end, [&](uint32_t x, uint32_t y) {
          const float cx = centroid[3 * (size_t)x + best_axis], cy = centroid[3 * (size_t)y + best_axis];
          return cx < cy || (cx == cy && x < y);
        });
        mid = begin + best_k;
        split_done = true;
      }
    } else if (use_sah) {
      // binned surface-area heuristic over the three axes (32 bins of the centroid range): the cut with the least
      // area(left) * count(left) + area(right) * count(right); the traversal's answer does not depend on the tree
      // (mesh_kernels.cuh), so this only changes how many nodes a ray visits
      constexpr int kBins = 32;
      float best_cost = FLT_MAX;
      int best_axis = -1, best_bin = -1;
      for (int a = 0; a < 3; ++a) {
        const float ext = cmx[a] - cmn[a];
        if (!(ext > 0.0f)) continue;
        const float scale = (float)kBins / ext;
        Box3 bb[kBins];
        uint32_t bc[kBins];
        for (int k = 0; k < kBins; ++k) {
          bb[k].reset();
          bc[k] = 0;
        }
        for (size_t i = begin; i < end; ++i) {
          const uint32_t t = order[i];
          int k = (int)((centroid[3 * (size_t)t + a] - cmn[a]) * scale);
          k = std::min(std::max(k, 0), kBins - 1);
          bb[k].grow(tri_box[t]);
          ++bc[k];
        }
        float right_area[kBins];
        uint32_t right_cnt[kBins];
        Box3 acc;
        acc.reset();
        uint32_t n_acc = 0;
        for (int k = kBins - 1; k > 0; --k) {
          if (bc[k]) acc.grow(bb[k]);
          n_acc += bc[k];
          right_area[k] = n_acc ? half_area(acc) : 0.0f;
          right_cnt[k] = n_acc;
        }
        acc.reset();
        n_acc = 0;
        for (int k = 0; k < kBins - 1; ++k) {
          if (bc[k]) acc.grow(bb[k]);
          n_acc += bc[k];
          if (n_acc == 0 || right_cnt[k + 1] == 0) continue;
          const float cost = half_area(acc) * (float)n_acc + right_area[k + 1] * (float)right_cnt[k + 1];
          if (cost < best_cost) {
            best_cost = cost;
            best_axis = a;
            best_bin = k;
          }
        }
      }
      if (best_axis >= 0) {
        const float scale = (float)kBins / (cmx[best_axis] - cmn[best_axis]);
        const float lo = cmn[best_axis];
        auto it = std::stable_partition(order.begin() + begin, order.begin() + end, [&](uint32_t t) {
          int k = (int)((centroid[3 * (size_t)t + best_axis] - lo) * scale);
          k = std::min(std::max(k, 0), kBins - 1);
          return k <= best_bin;
        });
        mid = (size_t)(it - order.begin());
        split_done = mid > begin && mid < end;
        if (!split_done) mid = begin + cnt / 2;
      }
    }
    if (!split_done) {
      std::nth_element(order.begin() + begin, order.begin() + mid, order.begin() + end, [&](uint32_t a, uint32_t b) {
        const float ca = centroid[3 * (size_t)a + axis], cb = centroid[3 * (size_t)b + axis];
        return ca < cb || (ca == cb && a < b);
      });
    }
    const size_t me = nodes.size() / 4;
    nodes.resize(nodes.size() + 4);
    Box3 lb, rb;
    const int32_t lref = build(begin, mid, level + 1, &lb);
    const int32_t rref = build(mid, end, level + 1, &rb);
    lb = padded(lb);
    rb = padded(rb);
    float li, ri;
    std::memcpy(&li, &lref, 4);
    std::memcpy(&ri, &rref, 4);
    nodes[4 * me + 0] = make_float4(lb.mn[0], lb.mn[1], lb.mn[2], lb.mx[0]);
    nodes[4 * me + 1] = make_float4(lb.mx[1], lb.mx[2], rb.mn[0], rb.mn[1]);
    nodes[4 * me + 2] = make_float4(rb.mn[2], rb.mx[0], rb.mx[1], rb.mx[2]);
    nodes[4 * me + 3] = make_float4(li, ri, 0.f, 0.f);
    return (int32_t)me;
  }
};


// everything q3dr_mesh_create uploads
struct MeshHost {
  std::vector<float4> nodes;    // [n_nodes][4]
  std::vector<float4> tris;     // [nf][3] BVH order
  std::vector<float4> normals;  // [nf][3] original order
  int32_t root_ref = 0;
  uint32_t depth = 0;
  float pad = 0.f;
};

// tri_vertices nf x 9 (finite), tri_normals nf x 9 or null (=> the reference's face normals)
inline void build_mesh_host(const float* tri_vertices, const float* tri_normals, size_t nf, MeshHost* out) {
  MeshBuilder b;
  b.v = tri_vertices;
  if (const char* e = std::getenv("Q3DR_MESH_BUILD")) b.use_sah = std::strcmp(e, "median") != 0;
  if (const char* e = std::getenv("Q3DR_MESH_LEAF")) b.leaf_max = (size_t)std::min<long>(q3dr::kMeshLeafMaxTris, std::max<long>(1, std::strtol(e, nullptr, 10)));
  b.tri_box.resize(nf);
  b.centroid.resize(3 * nf);
  b.order.resize(nf);
  for (size_t i = 0; i < nf; ++i) b.order[i] = (uint32_t)i;
  Box3 scene;
  scene.reset();
  for (size_t i = 0; i < nf; ++i) {
    Box3& tb = b.tri_box[i];
    tb.reset();
    for (int k = 0; k < 3; ++k) tb.grow(tri_vertices + 9 * i + 3 * k);
    for (int a = 0; a < 3; ++a) b.centroid[3 * i + a] = 0.5f * tb.mn[a] + 0.5f * tb.mx[a];
    scene.grow(tb);
  }
  float extent = 0.f;
  for (int a = 0; a < 3; ++a) extent = std::max(extent, scene.mx[a] - scene.mn[a]);
  b.pad = 1e-4f * extent + 1e-30f;
  b.nodes.reserve(4 * (nf / 2 + 1));
  Box3 root_box;
  out->root_ref = b.build(0, nf, 0, &root_box);
  uint32_t max_depth = (uint32_t)kMeshStack - 2;  // Q3DR_MESH_MAX_DEPTH lowers it (tests of the fallback below)
  if (const char* e = std::getenv("Q3DR_MESH_MAX_DEPTH")) max_depth = std::min<uint32_t>(max_depth, (uint32_t)std::max<long>(1, std::strtol(e, nullptr, 10)));
  if (b.use_sah && b.depth > max_depth) {
    // an adversarial triangle distribution made the SAH tree deeper than the traversal stack: the balanced tree instead
    b.use_sah = false;
    b.depth = 0;
    b.nodes.clear();
    for (size_t i = 0; i < nf; ++i) b.order[i] = (uint32_t)i;
    out->root_ref = b.build(0, nf, 0, &root_box);
  }
  out->depth = b.depth;
  out->pad = b.pad;
  out->nodes.swap(b.nodes);
  out->tris.resize(3 * nf);
  out->normals.resize(3 * nf);
  for (size_t k = 0; k < nf; ++k) {
    const uint32_t i = b.order[k];
    const float* p = tri_vertices + 9 * (size_t)i;
    float w;
    std::memcpy(&w, &i, 4);
    out->tris[3 * k + 0] = make_float4(p[0], p[1], p[2], w);
    out->tris[3 * k + 1] = make_float4(p[3] - p[0], p[4] - p[1], p[5] - p[2], 0.f);
    out->tris[3 * k + 2] = make_float4(p[6] - p[0], p[7] - p[1], p[8] - p[2], 0.f);
  }
  for (size_t i = 0; i < nf; ++i) {
    if (tri_normals) {
      for (int k = 0; k < 3; ++k) {
        const float* n = tri_normals + 9 * i + 3 * k;
        out->normals[3 * i + k] = make_float4(n[0], n[1], n[2], 0.f);
      }
    } else {
      // viewpoint_offscreen_renderer.cpp:162-166: normalize(cross(v2 - v1, v3 - v2)) on all three corners
      const float* p = tri_vertices + 9 * i;
      const float ax = p[3] - p[0], ay = p[4] - p[1], az = p[5] - p[2];
      const float bx = p[6] - p[3], by = p[7] - p[4], bz = p[8] - p[5];
      const float cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
      const float len = std::sqrt((cx * cx + cy * cy) + cz * cz);
      const float4 n = make_float4(cx / len, cy / len, cz / len, 0.f);
      out->normals[3 * i + 0] = out->normals[3 * i + 1] = out->normals[3 * i + 2] = n;
    }
  }
}

}  // namespace q3dr
