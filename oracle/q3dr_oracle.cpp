// q3dr_oracle.cpp -- CPU ORACLE (test infrastructure, not the product; see q3dr_oracle.h).
//
// Dependency-free restatement of the reference CPU raycast + score path in strict IEEE
// binary32.  Build: g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math -fopenmp (oracle/Makefile).
// Every function cites the reference file:line it follows (paths relative to the reference root).
#include "q3dr_oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <stack>
#include <unordered_map>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#if defined(__FAST_MATH__)
#error "the oracle must not be built with fast-math"
#endif

namespace {

// std::min / std::max semantics (ternaries), as used by geometry.h:336-343.
inline float smin(float a, float b) { return (b < a) ? b : a; }
inline float smax(float a, float b) { return (a < b) ? b : a; }

struct Box {
  float mn[3];
  float mx[3];
};

// bvh.h:46-119: node = bounding box + left/right + object.  Stored in one array; `leaf` is the
// input index of the object for leaves, -1 for inner nodes.
struct Node {
  Box box;
  int32_t left;
  int32_t right;
  int32_t leaf;
};

struct Ray {  // geometry.h:57-110 (RayData)
  float o[3];
  float d[3];
  float inv[3];
};

struct Hit {  // bvh.h:162-172 (IntersectionResult)
  float p[3];
  int32_t node;
  uint32_t depth;
  float dist_sq;
};

}  // namespace

struct orc_tree {
  std::vector<Node> nodes;
  std::vector<Box> leaf_boxes;       // input order
  std::vector<uint32_t> dfs_order;   // k -> input index
  std::vector<uint32_t> dfs_rank;    // input index -> k
  std::vector<uint32_t> leaf_depth;  // input index -> depth
  size_t depth = 0;
  int32_t root = -1;
};

namespace {

// geometry.h:227-229  getCenter(i) = (min(i) + max(i)) / 2
inline float center(const Box& b, int axis) { return (b.mn[axis] + b.mx[axis]) / 2.0f; }

// bvh.h:764-788 splitMedian.  `idx` plays the role of the objects vector (it is permuted in place,
// exactly like std::stable_sort permutes the ObjectWithBoundingBox range).
int32_t splitMedian(orc_tree* t, std::vector<uint32_t>& idx, size_t begin, size_t end, int sort_axis,
                    size_t cur_depth) {
  const int32_t me = (int32_t)t->nodes.size();
  t->nodes.emplace_back();
  t->nodes[me].left = t->nodes[me].right = t->nodes[me].leaf = -1;
  if (end - begin > 1) {
    const std::vector<Box>& boxes = t->leaf_boxes;
    std::stable_sort(idx.begin() + begin, idx.begin() + end, [&](uint32_t a, uint32_t b) -> bool {
      return center(boxes[a], sort_axis) < center(boxes[b], sort_axis);
    });
    const int next_sort_axis = (sort_axis + 1) % 3;
    const size_t mid = begin + (end - begin) / 2;
    const int32_t l = splitMedian(t, idx, begin, mid, next_sort_axis, cur_depth + 1);
    const int32_t r = splitMedian(t, idx, mid, end, next_sort_axis, cur_depth + 1);
    Node& n = t->nodes[me];
    n.left = l;
    n.right = r;
    // bvh.h:96-106 computeBoundingBox -> geometry.h:353-357 getUnion (cwiseMin / cwiseMax)
    const Box& a = t->nodes[l].box;
    const Box& b = t->nodes[r].box;
    for (int i = 0; i < 3; ++i) {
      n.box.mn[i] = smin(a.mn[i], b.mn[i]);
      n.box.mx[i] = smax(a.mx[i], b.mx[i]);
    }
  } else {
    Node& n = t->nodes[me];
    n.box = t->leaf_boxes[idx[begin]];
    n.leaf = (int32_t)idx[begin];
    t->leaf_depth[idx[begin]] = (uint32_t)cur_depth;
    if (cur_depth > t->depth) t->depth = cur_depth;
  }
  return me;
}

// geometry.h:235-238
inline bool isOutside(const Box& b, const float* p) {
  return (p[0] < b.mn[0]) || (p[1] < b.mn[1]) || (p[2] < b.mn[2]) || (p[0] > b.mx[0]) ||
         (p[1] > b.mx[1]) || (p[2] > b.mx[2]);
}

// geometry.h:319-351 BoundingBox3D::intersects(RayData, Vector3*)
inline bool boxIntersects(const Box& b, const Ray& r, float* intersection) {
  float t_min = -std::numeric_limits<float>::max();
  float t_max = std::numeric_limits<float>::max();
  for (int i = 0; i < 3; ++i) {
    float t0 = (b.mn[i] - r.o[i]) * r.inv[i];
    float t1 = (b.mx[i] - r.o[i]) * r.inv[i];
    t_min = smax(t_min, smin(t0, t1));
    t_max = smin(t_max, smax(t0, t1));
  }
  bool intersect = t_max > smax(t_min, 0.0f);
  if (intersect) {
    t_min = smax(t_min, 0.0f);
    for (int i = 0; i < 3; ++i) {
      float s = r.d[i] * t_min;
      intersection[i] = r.o[i] + s;
    }
  }
  return intersect;
}

// Association of (a - b).squaredNorm() in the query.  0 (default, THE contract): Eigen's fixed-size-3
// reduction a0 + (a1 + a2), the reference's CPU path.  1: the sequential loop ((0 + a0) + a1) + a2 of
// the reference's device code (include/bh/cuda_math.h:135-141) -- used only by the test that pins this
// oracle bit for bit against the reference's own bvh.cu compiled under oracle/ref_cuda.
int g_sqnorm_sequential = 0;

// (a - b).squaredNorm()
inline float distSq(const float* a, const float* b) {
  float e0 = a[0] - b[0], e1 = a[1] - b[1], e2 = a[2] - b[2];
  float s0 = e0 * e0, s1 = e1 * e1, s2 = e2 * e2;
  if (g_sqnorm_sequential) {
    float acc = 0.0f + s0;
    acc = acc + s1;
    return acc + s2;
  }
  float s12 = s1 + s2;
  return s0 + s12;
}

// bvh.h:842-909 intersectsRecursive
bool intersectsRecursive(const orc_tree* t, const Ray& ray, int32_t cur, uint32_t cur_depth, Hit* result,
                         uint64_t* visits) {
  ++*visits;
  const Node& n = t->nodes[cur];
  const bool outside = isOutside(n.box, ray.o);
  float isect[3] = {0, 0, 0};
  float isect_dist_sq = 0;
  if (outside) {
    const bool hit = boxIntersects(n.box, ray, isect);
    if (hit) {
      isect_dist_sq = distSq(ray.o, isect);
      if (isect_dist_sq > result->dist_sq) {
        return false;
      }
    } else {
      return false;
    }
  }
  if (n.left < 0 && n.right < 0) {
    if (!outside) {
      isect[0] = ray.o[0];
      isect[1] = ray.o[1];
      isect[2] = ray.o[2];
      isect_dist_sq = 0;
    }
    result->p[0] = isect[0];
    result->p[1] = isect[1];
    result->p[2] = isect[2];
    result->node = cur;
    result->depth = cur_depth;
    result->dist_sq = isect_dist_sq;
    return true;
  }
  bool l = false, r = false;
  if (n.left >= 0) l = intersectsRecursive(t, ray, n.left, cur_depth + 1, result, visits);
  if (n.right >= 0) r = intersectsRecursive(t, ray, n.right, cur_depth + 1, result, visits);
  return l || r;
}

// bvh.h:376-391 Tree::intersects(ray, min_range, max_range).  min_range is stored, never read.
inline bool treeIntersects(const orc_tree* t, const Ray& ray, float /*min_range*/, float max_range, Hit* result,
                           uint64_t* visits) {
  result->p[0] = result->p[1] = result->p[2] = 0;
  result->node = -1;
  result->depth = 0;
  if (max_range > 0) {
    result->dist_sq = max_range * max_range;
  } else {
    result->dist_sq = std::numeric_limits<float>::max();
  }
  return intersectsRecursive(t, ray, t->root, 0, result, visits);
}

// geometry.h:65-66 Ray(origin, direction): direction.normalized(); RayData: cwiseInverse (bvh.h:380)
inline void makeRay(const float* origin, const float* direction, Ray* r) {
  float s0 = direction[0] * direction[0], s1 = direction[1] * direction[1], s2 = direction[2] * direction[2];
  float n2 = s0 + (s1 + s2);
  for (int i = 0; i < 3; ++i) r->o[i] = origin[i];
  if (n2 > 0) {
    float n = std::sqrt(n2);
    for (int i = 0; i < 3; ++i) r->d[i] = direction[i] / n;
  } else {
    for (int i = 0; i < 3; ++i) r->d[i] = direction[i];
  }
  for (int i = 0; i < 3; ++i) r->inv[i] = 1.0f / r->d[i];
}

// viewpoint.cpp:88-96 + sparse_reconstruction.cpp:148-154: camera ray for integer pixel (x, y).
inline void cameraRay(const float K[4], const float Rt[12], uint32_t x, uint32_t y, Ray* r) {
  const float fx = K[0], fy = K[1], cx = K[2], cy = K[3];
  float c[3];
  c[0] = ((float)x - cx) / fx;
  c[1] = ((float)y - cy) / fy;
  c[2] = 1.0f;
  float origin[3] = {Rt[3], Rt[7], Rt[11]};
  float d[3];
  for (int i = 0; i < 3; ++i) {
    // Eigen 3x3 * vec3 lazy product: row(i).cwiseProduct(v).sum() = p0 + (p1 + p2)
    float p0 = Rt[4 * i + 0] * c[0];
    float p1 = Rt[4 * i + 1] * c[1];
    float p2 = Rt[4 * i + 2] * c[2];
    float p12 = p1 + p2;
    d[i] = p0 + p12;
  }
  makeRay(origin, d, r);
}

struct Derived {  // viewpoint_score.cpp:20-28, viewpoint_planner.cpp:107-109
  float thr, kf, a_thr, ka;
};

inline Derived derive(const orc_score_opts* o) {
  Derived d;
  d.thr = o->voxel_sensor_size_ratio_threshold;
  d.kf = 1 / o->voxel_sensor_size_ratio_inv_falloff_factor;
  d.a_thr = o->incidence_angle_threshold_degrees * float(M_PI) / float(180);
  d.ka = 1 / (o->incidence_angle_inv_falloff_factor_degrees * float(M_PI) / float(180));
  return d;
}

// viewpoint_planner_scoring.cpp:11-24,67-73,116-122 ; viewpoint_score.cpp:45-68
float voxelInformation(const Box& b, float weight, const float* nrm, const float* cam, float fx,
                       uint32_t image_width, const orc_score_opts* opts, const Derived& dv) {
  float g[3];
  for (int i = 0; i < 3; ++i) {
    float m = (b.mn[i] + b.mx[i]) / 2.0f;  // getCenter, geometry.h:223-225
    g[i] = cam[i] - m;
  }
  float s0 = g[0] * g[0], s1 = g[1] * g[1], s2 = g[2] * g[2];
  float n2 = s0 + (s1 + s2);
  float distance = std::sqrt(n2);                  // .norm()
  float size_x = b.mx[0] - b.mn[0];                // getExtent(0), geometry.h:203-205
  float rel = fx * size_x / distance / (float)image_width;  // sparse_reconstruction.cpp:132-135
  float res;
  if (rel >= dv.thr) {
    res = 1;
  } else {
    res = std::exp(-dv.kf * (dv.thr - rel));
  }
  float inc;
  if (nrm[0] == 0 && nrm[1] == 0 && nrm[2] == 0) {
    inc = 1;
  } else {
    float v[3];
    if (n2 > 0) {
      for (int i = 0; i < 3; ++i) v[i] = g[i] / distance;  // normalized()
    } else {
      for (int i = 0; i < 3; ++i) v[i] = g[i];
    }
    float p0 = v[0] * nrm[0], p1 = v[1] * nrm[1], p2 = v[2] * nrm[2];
    float dot = p0 + (p1 + p2);
    dot = smax(smin(dot, 1.0f), -1.0f);  // ait::clamp, include/ait/utilities.h:65-68
    bool zero = false;
    if (dot <= 0) {
      if (opts->incidence_ignore_dot_product_sign) {
        dot = std::fabs(dot);
      } else {
        zero = true;
      }
    }
    if (zero) {
      inc = 0;
    } else {
      float angle = std::acos(dot);
      if (angle <= dv.a_thr) {
        inc = 1;
      } else {
        inc = std::exp(-dv.ka * (angle - dv.a_thr));
      }
    }
  }
  float factor = res * inc;
  return factor * weight;
}

struct Kept {
  int32_t leaf;
  uint32_t pixel;
  float dsq;
};

// viewpoint_raycast.cpp:166-219 pixel loop + :321-335 first-writer dedup
void castAndDedup(const orc_tree* t, const float K[4], const float Rt[12], uint32_t x0, uint32_t x1, uint32_t y0,
                  uint32_t y1, float min_range, float max_range, std::vector<Kept>* kept) {
  const uint32_t w = x1 - x0;
  std::unordered_map<int32_t, Kept> raycast_map;
  uint64_t visits = 0;
  for (uint32_t y = y0; y < y1; ++y) {
    for (uint32_t x = x0; x < x1; ++x) {
      Ray ray;
      cameraRay(K, Rt, x, y, &ray);
      Hit h;
      if (treeIntersects(t, ray, min_range, max_range, &h, &visits)) {
        Kept k;
        k.leaf = t->nodes[h.node].leaf;
        k.pixel = (y - y0) * w + (x - x0);
        k.dsq = h.dist_sq;
        raycast_map.emplace(k.leaf, k);  // emplace keeps the first pixel
      }
    }
  }
  kept->clear();
  kept->reserve(raycast_map.size());
  for (const auto& kv : raycast_map) kept->push_back(kv.second);
  // The reference's output order is hash order (unspecified); pinned here: ascending leaf index.
  std::sort(kept->begin(), kept->end(), [](const Kept& a, const Kept& b) { return a.leaf < b.leaf; });
}

}  // namespace

extern "C" {

orc_tree* orc_tree_build(const float* boxes, size_t n) {
  if (n == 0) return nullptr;
  orc_tree* t = new orc_tree;
  t->leaf_boxes.resize(n);
  std::memcpy(t->leaf_boxes.data(), boxes, n * sizeof(Box));
  t->leaf_depth.assign(n, 0);
  t->nodes.reserve(2 * n);
  std::vector<uint32_t> idx(n);
  for (size_t i = 0; i < n; ++i) idx[i] = (uint32_t)i;
  t->root = splitMedian(t, idx, 0, n, 0, 0);
  t->dfs_order = idx;  // after all the in-place sorts the range is in left-to-right leaf order
  t->dfs_rank.resize(n);
  for (size_t k = 0; k < n; ++k) t->dfs_rank[idx[k]] = (uint32_t)k;
  return t;
}

void orc_tree_free(orc_tree* t) { delete t; }
size_t orc_tree_num_leaves(const orc_tree* t) { return t->leaf_boxes.size(); }
size_t orc_tree_num_nodes(const orc_tree* t) { return t->nodes.size(); }
size_t orc_tree_depth(const orc_tree* t) { return t->depth; }

void orc_tree_dfs_order(const orc_tree* t, uint32_t* order) {
  std::memcpy(order, t->dfs_order.data(), t->dfs_order.size() * sizeof(uint32_t));
}

void orc_tree_leaf_depths(const orc_tree* t, uint32_t* depth_out) {
  std::memcpy(depth_out, t->leaf_depth.data(), t->leaf_depth.size() * sizeof(uint32_t));
}

// bvh.h:203-246 all-node iterator: stack, push left then right => root, right subtree, left subtree.
void orc_tree_voxel_index(const orc_tree* t, uint32_t* idx_out) {
  std::stack<int32_t> st;
  st.push(t->root);
  uint32_t voxel_index = 0;
  while (!st.empty()) {
    int32_t n = st.top();
    st.pop();
    const Node& node = t->nodes[n];
    if (node.leaf >= 0) idx_out[node.leaf] = voxel_index;
    ++voxel_index;
    if (node.left >= 0) st.push(node.left);
    if (node.right >= 0) st.push(node.right);
  }
}

void orc_tree_root_box(const orc_tree* t, float* box_out) {
  std::memcpy(box_out, &t->nodes[t->root].box, sizeof(Box));
}

// Eigen QuaternionBase::toRotationMatrix (Eigen 3.3.1, Geometry/Quaternion.h), used by pose.h:122-127.
void orc_pose_to_rt(const float q[4], const float trans[3], float rt[12]) {
  const float w = q[0], x = q[1], y = q[2], z = q[3];
  const float tx = 2.0f * x, ty = 2.0f * y, tz = 2.0f * z;
  const float twx = tx * w, twy = ty * w, twz = tz * w;
  const float txx = tx * x, txy = ty * x, txz = tz * x;
  const float tyy = ty * y, tyz = tz * y, tzz = tz * z;
  rt[0] = 1.0f - (tyy + tzz);
  rt[1] = txy - twz;
  rt[2] = txz + twy;
  rt[4] = txy + twz;
  rt[5] = 1.0f - (txx + tzz);
  rt[6] = tyz - twx;
  rt[8] = txz - twy;
  rt[9] = tyz + twx;
  rt[10] = 1.0f - (txx + tyy);
  rt[3] = trans[0];
  rt[7] = trans[1];
  rt[11] = trans[2];
}

uint64_t orc_raycast_window(const orc_tree* t, const float K[4], const float Rt[12], uint32_t x0, uint32_t x1,
                            uint32_t y0, uint32_t y1, float min_range, float max_range, int32_t* leaf, float* dsq,
                            float* xyz, uint32_t* depth) {
  const uint32_t w = x1 - x0;
  uint64_t visits = 0;
  for (uint32_t y = y0; y < y1; ++y) {
    for (uint32_t x = x0; x < x1; ++x) {
      Ray ray;
      cameraRay(K, Rt, x, y, &ray);
      Hit h;
      const bool ok = treeIntersects(t, ray, min_range, max_range, &h, &visits);
      const size_t i = (size_t)(y - y0) * w + (x - x0);
      if (leaf) leaf[i] = ok ? t->nodes[h.node].leaf : -1;
      if (dsq) dsq[i] = ok ? h.dist_sq : 0.0f;
      if (xyz) {
        xyz[3 * i + 0] = ok ? h.p[0] : 0.0f;
        xyz[3 * i + 1] = ok ? h.p[1] : 0.0f;
        xyz[3 * i + 2] = ok ? h.p[2] : 0.0f;
      }
      if (depth) depth[i] = ok ? h.depth : 0u;
    }
  }
  return visits;
}

void orc_raycast_window_brute(const orc_tree* t, const float K[4], const float Rt[12], uint32_t x0, uint32_t x1,
                              uint32_t y0, uint32_t y1, float /*min_range*/, float max_range, int32_t* leaf,
                              float* dsq) {
  const uint32_t w = x1 - x0;
  const size_t n = t->leaf_boxes.size();
  for (uint32_t y = y0; y < y1; ++y) {
    for (uint32_t x = x0; x < x1; ++x) {
      Ray ray;
      cameraRay(K, Rt, x, y, &ray);
      float best = (max_range > 0) ? max_range * max_range : std::numeric_limits<float>::max();
      int32_t best_leaf = -1;
      // leaves in DFS order; "<=" so that a later leaf at equal distance overwrites (bvh.h:867,881-895)
      for (size_t k = 0; k < n; ++k) {
        const uint32_t li = t->dfs_order[k];
        const Box& b = t->leaf_boxes[li];
        float d2;
        if (isOutside(b, ray.o)) {
          float p[3];
          if (!boxIntersects(b, ray, p)) continue;
          d2 = distSq(ray.o, p);
          if (d2 > best) continue;
        } else {
          d2 = 0;
        }
        best = d2;
        best_leaf = (int32_t)li;
      }
      const size_t i = (size_t)(y - y0) * w + (x - x0);
      leaf[i] = best_leaf;
      if (dsq) dsq[i] = best_leaf >= 0 ? best : 0.0f;
    }
  }
}

void orc_intersect_rays(const orc_tree* t, const float* origins, const float* directions, size_t n,
                        float min_range, float max_range, int32_t* leaf, float* dsq, float* xyz,
                        uint32_t* depth) {
  uint64_t visits = 0;
  for (size_t i = 0; i < n; ++i) {
    Ray ray;
    makeRay(origins + 3 * i, directions + 3 * i, &ray);
    Hit h;
    const bool ok = treeIntersects(t, ray, min_range, max_range, &h, &visits);
    if (leaf) leaf[i] = ok ? t->nodes[h.node].leaf : -1;
    if (dsq) dsq[i] = ok ? h.dist_sq : 0.0f;
    if (xyz) {
      xyz[3 * i + 0] = ok ? h.p[0] : 0.0f;
      xyz[3 * i + 1] = ok ? h.p[1] : 0.0f;
      xyz[3 * i + 2] = ok ? h.p[2] : 0.0f;
    }
    if (depth) depth[i] = ok ? h.depth : 0u;
  }
}

void orc_set_sqnorm_sequential(int on) { g_sqnorm_sequential = on ? 1 : 0; }

void orc_tree_export_nodes(const orc_tree* t, float* boxes, int32_t* left, int32_t* right, int32_t* leaf,
                           int32_t* root) {
  for (size_t i = 0; i < t->nodes.size(); ++i) {
    const Node& n = t->nodes[i];
    for (int a = 0; a < 3; ++a) {
      boxes[6 * i + a] = n.box.mn[a];
      boxes[6 * i + 3 + a] = n.box.mx[a];
    }
    left[i] = n.left;
    right[i] = n.right;
    leaf[i] = n.leaf;
  }
  if (root) *root = t->root;
}

void orc_camera_rays(const float K[4], const float Rt[12], uint32_t x0, uint32_t x1, uint32_t y0, uint32_t y1,
                     float* origins, float* directions) {
  const uint32_t w = x1 - x0;
  for (uint32_t y = y0; y < y1; ++y) {
    for (uint32_t x = x0; x < x1; ++x) {
      Ray r;
      cameraRay(K, Rt, x, y, &r);
      const size_t i = (size_t)(y - y0) * w + (x - x0);
      for (int a = 0; a < 3; ++a) {
        origins[3 * i + a] = r.o[a];
        directions[3 * i + a] = r.d[a];
      }
    }
  }
}

void orc_intersect_rays_raw(const orc_tree* t, const float* origins, const float* directions, size_t n,
                            float min_range, float max_range, int32_t* leaf, float* dsq, float* xyz,
                            uint32_t* depth) {
#pragma omp parallel for schedule(dynamic, 256)
  for (size_t i = 0; i < n; ++i) {
    uint64_t visits = 0;
    Ray ray;
    for (int a = 0; a < 3; ++a) {
      ray.o[a] = origins[3 * i + a];
      ray.d[a] = directions[3 * i + a];
      ray.inv[a] = 1.0f / ray.d[a];  // cwiseInverse (bvh.h:380; bvh.cu:458)
    }
    Hit h;
    const bool ok = treeIntersects(t, ray, min_range, max_range, &h, &visits);
    if (leaf) leaf[i] = ok ? t->nodes[h.node].leaf : -1;
    if (dsq) dsq[i] = ok ? h.dist_sq : 0.0f;
    if (xyz) {
      xyz[3 * i + 0] = ok ? h.p[0] : 0.0f;
      xyz[3 * i + 1] = ok ? h.p[1] : 0.0f;
      xyz[3 * i + 2] = ok ? h.p[2] : 0.0f;
    }
    if (depth) depth[i] = ok ? h.depth : 0u;
  }
}

float orc_voxel_information(const float box[6], float weight, const float normal[3], const float cam[3],
                            float fx, uint32_t image_width, const orc_score_opts* opts) {
  Box b;
  std::memcpy(&b, box, sizeof(Box));
  return voxelInformation(b, weight, normal, cam, fx, image_width, opts, derive(opts));
}

size_t orc_score_viewpoint(const orc_tree* t, const float* weights, const float* normals, const float K[4],
                           uint32_t image_width, const float Rt[12], uint32_t x0, uint32_t x1, uint32_t y0,
                           uint32_t y1, const orc_score_opts* opts, size_t cap, int32_t* out_leaf,
                           float* out_info, uint32_t* out_pixel, float* out_dsq, uint32_t* hit_count,
                           float* total, float* total_f32) {
  std::vector<Kept> kept;
  castAndDedup(t, K, Rt, x0, x1, y0, y1, opts->raycast_min_range, opts->raycast_max_range, &kept);
  const Derived dv = derive(opts);
  const float cam[3] = {Rt[3], Rt[7], Rt[11]};
  // viewpoint_planner_raycast.cpp:128-137
  size_t n_out = 0;
  double sum64 = 0.0;
  float sum32 = 0.0f;
  bool overflow = false;
  const bool counting_only = !out_leaf && !out_info && !out_pixel && !out_dsq;
  for (const Kept& k : kept) {
    const float info = voxelInformation(t->leaf_boxes[k.leaf], weights[k.leaf], normals + 3 * (size_t)k.leaf, cam,
                                        K[0], image_width, opts, dv);
    if (!opts->ignore_voxels_with_zero_information || info > 0) {
      if (n_out < cap) {
        if (out_leaf) out_leaf[n_out] = k.leaf;
        if (out_info) out_info[n_out] = info;
        if (out_pixel) out_pixel[n_out] = k.pixel;
        if (out_dsq) out_dsq[n_out] = k.dsq;
      } else if (counting_only == false) {
        overflow = true;
      }
      ++n_out;
      sum64 += (double)info;
      sum32 += info;
    }
  }
  if (hit_count) *hit_count = (uint32_t)kept.size();
  if (total) *total = (float)sum64;
  if (total_f32) *total_f32 = sum32;
  return overflow ? (size_t)-1 : n_out;
}

// ---- image-space normals ----
//
// The reference renders the mesh normals with OpenGL (viewpoint_offscreen_renderer.cpp:361-397, projection :276-293,
// shaders/triangles_normals.{v,f}.glsl) and decodes one pixel (:519-526).  A rasteriser cannot be run or pinned here;
// the image is restated as a brute-force ray cast over ALL triangles through the pixel centres of that projection
// (PARITY UNPINNED against OpenGL; the decode and the scoring with the decoded normal follow the reference line by
// line).  Arithmetic: binary32, no FMA, three-term sums a0 + (a1 + a2).
namespace {

struct MRay {
  float o[3], d[3];
};

// pixel centre (x + 0.5, y + 0.5) through proj(0,0) = fx / cx, proj(1,1) = fy / cy (:285-286); camera z = 1, so the
// ray parameter is the camera-space depth the clip planes and the depth test act on
inline void meshRay(const float K[4], const float Rt[12], uint32_t W, uint32_t H, uint32_t x, uint32_t y, MRay* r) {
  const float fx = K[0], fy = K[1], cx = K[2], cy = K[3];
  float ndc_x = ((float)x + 0.5f) * 2.0f / (float)W - 1.0f;
  float ndc_y = ((float)y + 0.5f) * 2.0f / (float)H - 1.0f;
  float c0 = ndc_x / (fx / cx);
  float c1 = ndc_y / (fy / cy);
  for (int i = 0; i < 3; ++i) {
    float p0 = Rt[4 * i + 0] * c0;
    float p1 = Rt[4 * i + 1] * c1;
    float p2 = Rt[4 * i + 2];
    float p12 = p1 + p2;
    r->d[i] = p0 + p12;
    r->o[i] = Rt[4 * i + 3];
  }
}

inline float dot3(const float* a, const float* b) {
  float p0 = a[0] * b[0], p1 = a[1] * b[1], p2 = a[2] * b[2];
  float p12 = p1 + p2;
  return p0 + p12;
}

inline void cross3(const float* a, const float* b, float* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// Moeller-Trumbore against the triangle with corners v (9 floats)
inline bool meshTriangle(const MRay& r, const float* v, float near_t, float far_t, float* u, float* vv, float* t) {
  float e1[3], e2[3], p[3], s[3], q[3];
  for (int i = 0; i < 3; ++i) {
    e1[i] = v[3 + i] - v[i];
    e2[i] = v[6 + i] - v[i];
  }
  cross3(r.d, e2, p);
  const float det = dot3(e1, p);
  if (!(det != 0.0f)) return false;
  const float inv = 1.0f / det;
  for (int i = 0; i < 3; ++i) s[i] = r.o[i] - v[i];
  *u = dot3(s, p) * inv;
  if (!(*u >= 0.0f && *u <= 1.0f)) return false;
  cross3(s, e1, q);
  *vv = dot3(r.d, q) * inv;
  if (!(*vv >= 0.0f && *u + *vv <= 1.0f)) return false;
  *t = dot3(e2, q) * inv;
  return *t >= near_t && *t <= far_t;
}

inline uint32_t unorm8(float n) {
  float c = 0.5f * n + 0.5f;  // triangles_normals.f.glsl
  if (c != c) c = 0.0f;
  if (c < 0.0f) c = 0.0f;
  if (c > 1.0f) c = 1.0f;
  return (uint32_t)std::nearbyint(c * 255.0f);  // 8-bit UNORM colour attachment, round to nearest
}

uint32_t meshPixel(const float* tv, const float* tn, size_t nf, const MRay& r, float near_t, float far_t,
                   uint32_t clear_argb) {
  bool found = false;
  float bt = 0, bu = 0, bv = 0;
  size_t bi = 0;
  for (size_t i = 0; i < nf; ++i) {
    float u, v, t;
    if (meshTriangle(r, tv + 9 * i, near_t, far_t, &u, &v, &t)) {
      if (!found || t < bt) {  // GL_LESS: an equal depth keeps the triangle drawn first
        found = true;
        bt = t;
        bu = u;
        bv = v;
        bi = i;
      }
    }
  }
  if (!found) return clear_argb;
  const float* n = tn + 9 * bi;
  const float w0 = (1.0f - bu) - bv;
  uint32_t c[3];
  for (int k = 0; k < 3; ++k) {
    float a = w0 * n[k] + bu * n[3 + k];
    c[k] = unorm8(a + bv * n[6 + k]);
  }
  return 0xFF000000u | (c[0] << 16) | (c[1] << 8) | c[2];
}

// viewpoint_offscreen_renderer.cpp:519-526
inline void decodeNormal(uint32_t argb, float* n) {
  const int red = (argb >> 16) & 255, green = (argb >> 8) & 255, blue = argb & 255;  // QColor(QRgb)
  float v[3] = {(float)red / 255.f, (float)green / 255.f, (float)blue / 255.f};
  for (int i = 0; i < 3; ++i) n[i] = 2 * (v[i] - 0.5f);
  float s0 = n[0] * n[0], s1 = n[1] * n[1], s2 = n[2] * n[2];
  float sq = s0 + (s1 + s2);
  if (sq < 0.5f) {
    n[0] = n[1] = n[2] = 0;
    sq = 0;
  }
  if (sq > 0) {  // Eigen 3.3 normalize(): divides only when squaredNorm() > 0
    float len = std::sqrt(sq);
    for (int i = 0; i < 3; ++i) n[i] = n[i] / len;
  }
}

}  // namespace

void orc_decode_normal(uint32_t argb, float normal_out[3]) { decodeNormal(argb, normal_out); }

void orc_mesh_face_normals(const float* tri_vertices, size_t nf, float* tri_normals) {
  for (size_t i = 0; i < nf; ++i) {  // viewpoint_offscreen_renderer.cpp:162-166
    const float* p = tri_vertices + 9 * i;
    float a[3], b[3], c[3];
    for (int k = 0; k < 3; ++k) {
      a[k] = p[3 + k] - p[k];
      b[k] = p[6 + k] - p[3 + k];
    }
    cross3(a, b, c);
    const float len = std::sqrt((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]);
    for (int corner = 0; corner < 3; ++corner) {
      for (int k = 0; k < 3; ++k) tri_normals[9 * i + 3 * corner + k] = c[k] / len;
    }
  }
}

void orc_mesh_cast_pixels(const float* tri_vertices, const float* tri_normals, size_t nf, const float K[4],
                          uint32_t width, uint32_t height, const float Rt[12], float near_plane, float far_plane,
                          uint32_t clear_argb, const uint32_t* xs, const uint32_t* ys, size_t n, uint32_t* out_argb) {
#pragma omp parallel for schedule(dynamic, 64)
  for (long i = 0; i < (long)n; ++i) {
    MRay r;
    meshRay(K, Rt, width, height, xs[i], ys[i], &r);
    out_argb[i] = meshPixel(tri_vertices, tri_normals, nf, r, near_plane, far_plane, clear_argb);
  }
}

void orc_mesh_render_normals(const float* tri_vertices, const float* tri_normals, size_t nf, const float K[4],
                             uint32_t width, uint32_t height, const float Rt[12], float near_plane, float far_plane,
                             uint32_t clear_argb, uint32_t* out_argb) {
#pragma omp parallel for schedule(dynamic, 64)
  for (long i = 0; i < (long)width * height; ++i) {
    MRay r;
    meshRay(K, Rt, width, height, (uint32_t)(i % width), (uint32_t)(i / width), &r);
    out_argb[i] = meshPixel(tri_vertices, tri_normals, nf, r, near_plane, far_plane, clear_argb);
  }
}

// getRaycastHitVoxelsWithInformationScore with enable_opengl = true (viewpoint_planner_scoring.cpp:26-38):
// every voxel is scored with the decoded pixel of `normal_image` (image_height x image_width ARGB32) at its kept
// pixel (window origin added back: screen_coordinates are full-image pixel coordinates, viewpoint_raycast.cpp:203).
size_t orc_score_viewpoint_image_normals(const orc_tree* t, const float* weights, const uint32_t* normal_image,
                                         uint32_t image_height, const float K[4], uint32_t image_width,
                                         const float Rt[12], uint32_t x0, uint32_t x1, uint32_t y0, uint32_t y1,
                                         const orc_score_opts* opts, size_t cap, int32_t* out_leaf, float* out_info,
                                         uint32_t* out_pixel, float* out_dsq, uint32_t* hit_count, float* total) {
  (void)image_height;
  std::vector<Kept> kept;
  castAndDedup(t, K, Rt, x0, x1, y0, y1, opts->raycast_min_range, opts->raycast_max_range, &kept);
  const Derived dv = derive(opts);
  const float cam[3] = {Rt[3], Rt[7], Rt[11]};
  const uint32_t w = x1 - x0;
  size_t n_out = 0;
  double sum64 = 0.0;
  for (const Kept& k : kept) {
    const size_t x = x0 + k.pixel % w, y = y0 + k.pixel / w;  // static_cast<size_t>(image_coordinates)
    float nrm[3];
    decodeNormal(normal_image[y * image_width + x], nrm);
    const float info = voxelInformation(t->leaf_boxes[k.leaf], weights[k.leaf], nrm, cam, K[0], image_width, opts, dv);
    if (!opts->ignore_voxels_with_zero_information || info > 0) {
      if (n_out >= cap) return (size_t)-1;
      if (out_leaf) out_leaf[n_out] = k.leaf;
      if (out_info) out_info[n_out] = info;
      if (out_pixel) out_pixel[n_out] = k.pixel;
      if (out_dsq) out_dsq[n_out] = k.dsq;
      ++n_out;
      sum64 += (double)info;
    }
  }
  if (hit_count) *hit_count = (uint32_t)kept.size();
  if (total) *total = (float)sum64;
  return n_out;
}

// ViewpointPlannerData::updateWeightsWithRealViewpoints for ONE real image (viewpoint_planner_data.cpp:498-571): every
// pixel of the depth camera is cast, remove_duplicates = false, misses dropped (viewpoint_raycast.cpp:166-219; the
// reference's removeInvalidRaycastHitVoxels is buggy, :287-303 -- the intended compaction is restated); then, in scan
// order, every hit pixel whose depth-map value is <= 0 or infinite lowers the weight of the voxel it hit:
//   information = computeViewpointObservationScore(viewpoint, voxel, pixel)   = (res * inc) * CURRENT weight
//   weight      = max(0, weight - invalid_pixel_observation_factor * information)
// weights (n_leaves) are updated in place; min_range is accepted and ignored (Q1); max_range >= FLT_MAX-ish means
// unlimited like the reference's default.  normal_image (nullable): enable_opengl = true, the normal of the pixel itself
// (static_cast<size_t>(image_coordinates), :535-538); otherwise the voxel normals.  Returns the number of updates.
uint64_t orc_downweight_real_viewpoint(const orc_tree* t, float* weights, const float* normals,
                                       const uint32_t* normal_image, const float K[4], uint32_t width, uint32_t height,
                                       const float Rt[12], const float* depth_map, const orc_score_opts* opts,
                                       float invalid_pixel_observation_factor) {
  const Derived dv = derive(opts);
  const float cam[3] = {Rt[3], Rt[7], Rt[11]};
  uint64_t visits = 0, updates = 0;
  // pass 1: the raycast results vector (all pixels, scan order, hits only)
  struct PixelHit {
    int32_t leaf;
    uint32_t x, y;
  };
  std::vector<PixelHit> results;
  for (uint32_t y = 0; y < height; ++y) {
    for (uint32_t x = 0; x < width; ++x) {
      Ray ray;
      cameraRay(K, Rt, x, y, &ray);
      Hit h;
      if (treeIntersects(t, ray, opts->raycast_min_range, opts->raycast_max_range, &h, &visits)) {
        results.push_back(PixelHit{t->nodes[h.node].leaf, x, y});
      }
    }
  }
  // pass 2: viewpoint_planner_data.cpp:551-568
  for (const PixelHit& ir : results) {
    const float depth = depth_map[(size_t)ir.y * width + ir.x];
    if (depth <= 0 || std::isinf(depth)) {
      float nrm_img[3];
      const float* nrm = normals + 3 * (size_t)ir.leaf;
      if (normal_image) {
        decodeNormal(normal_image[(size_t)ir.y * width + ir.x], nrm_img);
        nrm = nrm_img;
      }
      const float information = voxelInformation(t->leaf_boxes[ir.leaf], weights[ir.leaf], nrm, cam, K[0], width, opts, dv);
      const float new_weight = std::max<float>(0, weights[ir.leaf] - invalid_pixel_observation_factor * information);
      weights[ir.leaf] = new_weight;
      ++updates;
    }
  }
  return updates;
}

// order-independent 64-bit checksum of a voxel set: sum over the entries of splitmix64(leaf << 32 | pixel)
static inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

uint64_t orc_score_batch_checksum(const orc_tree* t, const float* weights, const float* normals, const float K[4],
                                  uint32_t width, uint32_t height, const float* Rt, size_t n_vp,
                                  const orc_score_opts* opts, int threads, uint32_t* kept_count, uint32_t* hit_count,
                                  float* total, uint64_t* set_checksum) {
  if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (long v = 0; v < (long)n_vp; ++v) {
    uint32_t hc = 0;
    float tot = 0, tot32 = 0;
    size_t kc;
    if (set_checksum) {
      const size_t cap = (size_t)width * height;
      std::vector<int32_t> leaf(cap);
      std::vector<uint32_t> pix(cap);
      kc = orc_score_viewpoint(t, weights, normals, K, width, Rt + 12 * v, 0, width, 0, height, opts, cap, leaf.data(),
                               nullptr, pix.data(), nullptr, &hc, &tot, &tot32);
      uint64_t h = 0;
      for (size_t i = 0; i < kc; ++i) h += mix64(((uint64_t)(uint32_t)leaf[i] << 32) | pix[i]);
      set_checksum[v] = h;
    } else {
      // all output arrays NULL => counting mode (no capacity check)
      kc = orc_score_viewpoint(t, weights, normals, K, width, Rt + 12 * v, 0, width, 0, height, opts, 0, nullptr,
                               nullptr, nullptr, nullptr, &hc, &tot, &tot32);
    }
    if (hit_count) hit_count[v] = hc;
    if (total) total[v] = tot;
    if (kept_count) kept_count[v] = (uint32_t)kc;
  }
  return (uint64_t)n_vp * width * height;
}

uint64_t orc_score_batch(const orc_tree* t, const float* weights, const float* normals, const float K[4],
                         uint32_t width, uint32_t height, const float* Rt, size_t n_vp,
                         const orc_score_opts* opts, int threads, uint32_t* kept_count, uint32_t* hit_count,
                         float* total) {
  return orc_score_batch_checksum(t, weights, normals, K, width, height, Rt, n_vp, opts, threads, kept_count, hit_count,
                                  total, nullptr);
}

// ---- occupancy octree -> BVH leaves, grid weights ----

// viewpoint_planner_data.cpp:820-863 for one octree leaf.  Returns 1 and the box if it becomes a BVH leaf.
int orc_octree_leaf_to_box(const float center[3], float size, float occupancy, uint32_t observation_count,
                           const float bvh_bbox[6], float obstacle_free_height, float occupancy_threshold,
                           uint32_t observation_threshold, float box_out[6]) {
  // occupancy_map.h:75-106
  const bool is_occupied = occupancy > occupancy_threshold;
  const bool is_free = !is_occupied;
  const bool is_known = observation_count >= observation_threshold;
  if (is_free && is_known) return 0;
  // geometry.h:159-160 BoundingBox3D(center, size)
  float mn[3], mx[3];
  for (int i = 0; i < 3; ++i) {
    mn[i] = center[i] - size / 2;
    mx[i] = center[i] + size / 2;
  }
  // geometry.h:373-376 constrainTo
  for (int i = 0; i < 3; ++i) {
    mn[i] = std::max(mn[i], bvh_bbox[i]);
    mx[i] = std::min(mx[i], bvh_bbox[3 + i]);
  }
  // geometry.h:175-177 isEmpty: (min >= max).any()
  auto is_empty = [&]() { return mn[0] >= mx[0] || mn[1] >= mx[1] || mn[2] >= mx[2]; };
  if (is_empty()) return 0;
  if (mx[2] >= obstacle_free_height) {
    mn[2] = std::min(obstacle_free_height, mn[2]);
    mx[2] = obstacle_free_height;
  }
  if (is_empty()) return 0;
  for (int i = 0; i < 3; ++i) {
    box_out[i] = mn[i];
    box_out[3 + i] = mx[i];
  }
  return 1;
}

// viewpoint_planner_data.cpp:438-482,573-577 updateWeights, BVH part: literally every cell in ix, iy, iz order,
// every overlapping leaf (brute force over the leaves; test sizes only).
void orc_update_weights_from_grid(const float* boxes, size_t n, const float grid_origin[3], float grid_increment,
                                  const int32_t grid_dim[3], const float* grid_weight, float lambda,
                                  const uint32_t* observation_count, float* weight /* in/out */) {
  for (int ix = 0; ix < grid_dim[0]; ++ix) {
    for (int iy = 0; iy < grid_dim[1]; ++iy) {
      for (int iz = 0; iz < grid_dim[2]; ++iz) {
        // getGridPosition: grid_origin_ + grid_increment_ * indices.cast<float>()
        const float xyz[3] = {grid_origin[0] + grid_increment * (float)ix, grid_origin[1] + grid_increment * (float)iy,
                              grid_origin[2] + grid_increment * (float)iz};
        Box cell;
        for (int a = 0; a < 3; ++a) {
          cell.mn[a] = xyz[a] - grid_increment / 2;
          cell.mx[a] = xyz[a] + grid_increment / 2;
        }
        const float w = grid_weight[((size_t)ix * grid_dim[1] + iy) * grid_dim[2] + iz];
        for (size_t k = 0; k < n; ++k) {
          Box leaf;
          for (int a = 0; a < 3; ++a) {
            leaf.mn[a] = boxes[6 * k + a];
            leaf.mx[a] = boxes[6 * k + 3 + a];
          }
          bool hit = true;  // geometry.h:277-285, node box vs cell
          for (int a = 0; a < 3; ++a) {
            if (leaf.mx[a] < cell.mn[a]) hit = false;
          }
          for (int a = 0; a < 3; ++a) {
            if (cell.mx[a] < leaf.mn[a]) hit = false;
          }
          if (hit) {
            const float observation_count_factor = std::exp(-lambda * observation_count[k]);
            weight[k] = w * observation_count_factor;
          }
        }
      }
    }
  }
}

// ---- box-vs-map overlap queries ----

namespace {
// geometry.h:277-285 BoundingBox3D::intersects(other)
inline bool boxIntersectsBox(const Box& a, const Box& other) {
  for (int i = 0; i < 3; ++i) {
    if (a.mx[i] < other.mn[i]) return false;
  }
  for (int i = 0; i < 3; ++i) {
    if (other.mx[i] < a.mn[i]) return false;
  }
  return true;
}

// bvh.h:911-939 intersectsRecursive(bbox, node, depth, results)
void bboxRecursive(const orc_tree* t, const Box& bbox, int32_t cur, std::vector<int32_t>* results) {
  const Node& n = t->nodes[cur];
  if (!boxIntersectsBox(n.box, bbox)) return;
  if (n.left < 0 && n.right < 0) {
    results->push_back(n.leaf);
    return;
  }
  if (n.left >= 0) bboxRecursive(t, bbox, n.left, results);
  if (n.right >= 0) bboxRecursive(t, bbox, n.right, results);
}
}  // namespace

// bvh.h:544-554 Tree::intersects(bbox): input leaf indices in the reference's visiting order.  Returns the count;
// writes at most cap of them.
size_t orc_overlap_box(const orc_tree* t, const float box[6], size_t cap, int32_t* out_leaf) {
  Box b;
  for (int i = 0; i < 3; ++i) {
    b.mn[i] = box[i];
    b.mx[i] = box[3 + i];
  }
  std::vector<int32_t> res;
  bboxRecursive(t, b, t->root, &res);
  for (size_t i = 0; i < res.size() && i < cap; ++i) out_leaf[i] = res[i];
  return res.size();
}

// viewpoint_planner_data.cpp:209-235 isValidObjectPosition without the no-fly zones
int orc_valid_object_position(const orc_tree* t, const float position[3], const float object_bbox[6],
                              const float bvh_bbox[6], float obstacle_free_height) {
  const float* p = position;
  if (p[2] >= obstacle_free_height) {
    // geometry.h:240-243 isInside
    return p[0] >= bvh_bbox[0] && p[1] >= bvh_bbox[1] && p[2] >= bvh_bbox[2] && p[0] <= bvh_bbox[3] &&
           p[1] <= bvh_bbox[4] && p[2] <= bvh_bbox[5];
  }
  const Box& root = t->nodes[t->root].box;
  const bool inside_root = p[0] >= root.mn[0] && p[1] >= root.mn[1] && p[2] >= root.mn[2] && p[0] <= root.mx[0] &&
                           p[1] <= root.mx[1] && p[2] <= root.mx[2];
  if (!inside_root) return 0;
  Box c;  // object_bbox + position (geometry.h:179-181)
  for (int i = 0; i < 3; ++i) {
    c.mn[i] = object_bbox[i] + p[i];
    c.mx[i] = object_bbox[3 + i] + p[i];
  }
  if (c.mx[2] >= obstacle_free_height) {
    // Vector3 cropped_maximum = object_bbox.getMaximum(); cropped_maximum(2) = obstacle_free_height  (sic: un-centred)
    c.mx[0] = object_bbox[3];
    c.mx[1] = object_bbox[4];
    c.mx[2] = obstacle_free_height;
  }
  std::vector<int32_t> res;
  bboxRecursive(t, c, t->root, &res);
  return res.empty() ? 1 : 0;
}

// ---- consumers of the voxel sets (path search) ----

// ViewpointPlanner::computeNewInformation (viewpoint_planner_scoring.cpp:124-147).  observed: dense over the
// leaves, NaN = not in observed_voxel_map.  The reference accumulates in hash order; here ascending entry order.
float orc_new_information(const int32_t* leaf, const float* info, size_t n, const float* observed, const float* weight,
                          float factor, double* sum_f64) {
  float acc = 0.0f;
  double acc64 = 0.0;
  for (size_t i = 0; i < n; ++i) {
    const float observation_information = factor * info[i];
    float novel_information = observation_information;
    const float seen = observed[leaf[i]];
    if (!std::isnan(seen)) {
      const float voxel_weight = weight[leaf[i]];
      novel_information = smin(observation_information, voxel_weight - seen);
    }
    acc = acc + novel_information;
    acc64 += (double)novel_information;
  }
  if (sum_f64) *sum_f64 = acc64;
  return acc;
}

// addObservedVoxelsToViewpointPath (viewpoint_planner_path.cpp:203-233); observed is updated in place.
float orc_observed_add(const int32_t* leaf, const float* info, size_t n, float* observed, const float* weight,
                       float factor, double* sum_f64) {
  float new_information = 0.0f;
  double acc64 = 0.0;
  for (size_t i = 0; i < n; ++i) {
    const float observation_information = factor * info[i];
    float novel_observation_information;
    float& slot = observed[leaf[i]];
    if (std::isnan(slot)) {
      slot = observation_information;
      novel_observation_information = observation_information;
    } else {
      const float voxel_weight = weight[leaf[i]];
      novel_observation_information = smin(observation_information, voxel_weight - slot);
      if (slot <= voxel_weight) {
        slot += observation_information;
        if (slot > voxel_weight) slot = voxel_weight;
      }
    }
    new_information += novel_observation_information;
    acc64 += (double)novel_observation_information;
  }
  if (sum_f64) *sum_f64 = acc64;
  return new_information;
}

// compute_overlap_information_lambda (viewpoint_planner_path.cpp:838-848): for every element of set 1, look it up in
// set 2 with VoxelWithInformation's equality (voxel AND information, occupied_tree.h:73-79); both lists ascending.
float orc_overlap_information(const int32_t* leaf1, const float* info1, size_t n1, const int32_t* leaf2,
                              const float* info2, size_t n2, double* sum_f64) {
  float overlap_information = 0.0f;
  double acc64 = 0.0;
  size_t j = 0;
  for (size_t i = 0; i < n1; ++i) {
    while (j < n2 && leaf2[j] < leaf1[i]) ++j;
    if (j < n2 && leaf2[j] == leaf1[i] && info2[j] == info1[i]) {
      const float v = smin(info2[j], info1[i]);
      overlap_information += v;
      acc64 += (double)v;
    }
  }
  if (sum_f64) *sum_f64 = acc64;
  return overlap_information;
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
