// test_shim.cpp -- exercises include/q3dr/occupied_tree_cuda.hpp the way the reference's callers use bvh::Tree,
// ViewpointRaycast and ViewpointPlanner::getRaycastHitVoxelsWithInformationScore, and checks every result against
// the CPU oracle (test infrastructure, linked only here).
//
// A mock host tree with the reference's accessor names (getLeftChild / getRightChild / isLeaf / getBoundingBox /
// getObject) is built by median splits with std::stable_sort, like bvh::Tree::build would; the shim walks it.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/q3dr/occupied_tree_cuda.hpp"
#include "../../oracle/q3dr_oracle.h"

namespace mock {

struct Vec3 {
  float v[3];
  float operator()(int i) const { return v[i]; }
};
struct BBox {
  float mn[3], mx[3];
  float getMinimum(int i) const { return mn[i]; }
  float getMaximum(int i) const { return mx[i]; }
  float getCenter(int i) const { return (mn[i] + mx[i]) / 2; }
};
struct NodeObject {  // planner/occupied_tree.h:17-35
  float occupancy;
  uint16_t observation_count;
  float weight;
  Vec3 normal;
};
struct Node {
  BBox box;
  Node* left = nullptr;
  Node* right = nullptr;
  NodeObject* object = nullptr;
  const BBox& getBoundingBox() const { return box; }
  Node* getLeftChild() { return left; }
  Node* getRightChild() { return right; }
  bool isLeaf() const { return left == nullptr && right == nullptr; }
  NodeObject* getObject() { return object; }
};
struct Obj {
  BBox box;
  NodeObject* object;
};

struct Tree {
  std::vector<std::unique_ptr<Node>> pool;
  std::vector<std::unique_ptr<NodeObject>> objects;
  Node* root = nullptr;
  Node* alloc() {
    pool.emplace_back(new Node);
    return pool.back().get();
  }
  void split(Node* node, std::vector<Obj>::iterator b, std::vector<Obj>::iterator e, int axis) {
    if (e - b > 1) {
      std::stable_sort(b, e, [axis](const Obj& x, const Obj& y) { return x.box.getCenter(axis) < y.box.getCenter(axis); });
      node->left = alloc();
      node->right = alloc();
      split(node->left, b, b + (e - b) / 2, (axis + 1) % 3);
      split(node->right, b + (e - b) / 2, e, (axis + 1) % 3);
      for (int i = 0; i < 3; ++i) {
        node->box.mn[i] = std::min(node->left->box.mn[i], node->right->box.mn[i]);
        node->box.mx[i] = std::max(node->left->box.mx[i], node->right->box.mx[i]);
      }
    } else {
      node->box = b->box;
      node->object = b->object;
    }
  }
  void build(std::vector<Obj> objs) {
    root = alloc();
    split(root, objs.begin(), objs.end(), 0);
  }
};

struct Mat44 {
  float m[4][4] = {};
  float operator()(int r, int c) const { return m[r][c]; }
};
struct Mat34 {
  float m[3][4] = {};
  float operator()(int r, int c) const { return m[r][c]; }
};
struct Camera {
  Mat44 K;
  std::size_t w, h;
  const Mat44& intrinsics() const { return K; }
  std::size_t width() const { return w; }
  std::size_t height() const { return h; }
};
struct Pose {
  Mat34 T;
  Mat34 getTransformationImageToWorld() const { return T; }
};
struct Viewpoint {
  const Camera* cam;
  Pose p;
  const Camera& camera() const { return *cam; }
  const Pose& pose() const { return p; }
};

}  // namespace mock

#define CHECK(cond)                                                              \
  do {                                                                           \
    if (!(cond)) {                                                               \
      std::fprintf(stderr, "CHECK failed at line %d: %s\n", __LINE__, #cond);    \
      return 1;                                                                  \
    }                                                                            \
  } while (0)

int main(int argc, char** argv) {
  const bool compile_only = argc > 1 && std::string(argv[1]) == "--no-run";
  if (compile_only) {
    std::puts("shim compiled and linked");
    return 0;
  }
  // --- scene: ground plate + two walls of 0.25 m voxels, one big "unknown" cube -------------------------------
  mock::Tree host;
  std::vector<mock::Obj> objs;
  auto add = [&](float x, float y, float z, float s, float nx, float ny, float nz, int count) {
    host.objects.emplace_back(new mock::NodeObject);
    mock::NodeObject* o = host.objects.back().get();
    o->occupancy = count ? 0.97f : 0.5f;
    o->observation_count = (uint16_t)count;
    o->weight = std::exp(-0.1f * count);
    o->normal = mock::Vec3{{nx, ny, nz}};
    mock::Obj ob;
    ob.box = mock::BBox{{x, y, z}, {x + s, y + s, z + s}};
    ob.object = o;
    objs.push_back(ob);
  };
  for (int i = -24; i < 24; ++i)
    for (int j = -24; j < 24; ++j) add(0.25f * i, 0.25f * j, -0.25f, 0.25f, 0, 0, 1, 1 + (i * 7 + j * 3 + 100) % 9);
  for (int i = -10; i < 10; ++i)
    for (int k = 0; k < 12; ++k) {
      add(0.25f * i, 2.0f, 0.25f * k, 0.25f, 0, -1, 0, 2 + (i + k + 40) % 5);
      add(-3.0f, 0.25f * i, 0.25f * k, 0.25f, 1, 0, 0, 3);
    }
  add(0.0f, 2.25f, 0.0f, 2.0f, 0, 0, 0, 0);  // unknown space behind the wall
  const std::size_t n = objs.size();
  host.build(objs);

  // --- the shim: walk the host tree, upload --------------------------------------------------------------------
  using Tree = q3dr::OccupiedTreeCuda<mock::Node>;
  Tree tree;
  try {
    tree.ensureCudaTreeIsInitialized(host.root, 0);
  } catch (const Tree::Error& e) {
    std::fprintf(stderr, "upload failed: %s\n", e.what());
    return 2;
  }
  CHECK(tree.getNumOfLeafNodes() == n);

  // --- the oracle over the same leaves, in the same (left-to-right) order --------------------------------------
  std::vector<q3dr::LeafRecord<mock::Node>> leaves;
  q3dr::collectLeaves<mock::Node>(host.root, 0, &leaves);
  std::vector<float> boxes(6 * n), weights(n), normals(3 * n);
  for (std::size_t k = 0; k < n; ++k) {
    for (int i = 0; i < 3; ++i) {
      boxes[6 * k + i] = leaves[k].bbox_min[i];
      boxes[6 * k + 3 + i] = leaves[k].bbox_max[i];
      normals[3 * k + i] = leaves[k].normal[i];
    }
    weights[k] = leaves[k].weight;
  }
  orc_tree* oracle = orc_tree_build(boxes.data(), n);
  std::vector<uint32_t> order(n), odepth(n);
  orc_tree_dfs_order(oracle, order.data());
  orc_tree_leaf_depths(oracle, odepth.data());
  for (std::size_t k = 0; k < n; ++k) {
    CHECK(order[k] == k);                      // the host tree's leaf order is the oracle's
    CHECK(odepth[k] == leaves[k].depth);
  }

  // --- cameras --------------------------------------------------------------------------------------------------
  mock::Camera cam;
  cam.w = 96;
  cam.h = 72;
  cam.K.m[0][0] = cam.K.m[1][1] = 0.85f * 96;
  cam.K.m[0][2] = 48;
  cam.K.m[1][2] = 36;
  cam.K.m[2][2] = cam.K.m[3][3] = 1;
  const float Kf[4] = {cam.K.m[0][0], cam.K.m[1][1], cam.K.m[0][2], cam.K.m[1][2]};
  std::vector<mock::Viewpoint> vps;
  const float quats[3][4] = {{0.6532815f, -0.6532815f, 0.2705981f, -0.2705981f}, {0.5f, -0.5f, 0.5f, -0.5f}, {0.70710678f, -0.70710678f, 0.0f, 0.0f}};
  const float pos[3][3] = {{1.3f, -4.1f, 2.2f}, {3.7f, 0.4f, 1.1f}, {-0.6f, -3.3f, 0.9f}};
  for (int v = 0; v < 3; ++v) {
    float rt[12];
    CHECK(q3dr_pose_to_rt(quats[v], pos[v], rt) == Q3DR_OK);
    mock::Viewpoint vp;
    vp.cam = &cam;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) vp.p.T.m[r][c] = rt[4 * r + c];
    vps.push_back(vp);
  }

  q3dr::ViewpointRaycastCuda<mock::Node> raycaster(&tree, 0.0f, 60.0f);
  orc_score_opts oo = {0.002f, 0.001f, 30.0f, 30.0f, 0, 1, 0.0f, 60.0f};
  std::size_t total_hits = 0;
  for (const mock::Viewpoint& vp : vps) {
    float Rt[12];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) Rt[4 * r + c] = vp.p.T.m[r][c];
    const std::size_t np = cam.w * cam.h;
    std::vector<int32_t> leaf(np);
    std::vector<float> dsq(np), xyz(3 * np);
    std::vector<uint32_t> depth(np);
    orc_raycast_window(oracle, Kf, Rt, 0, cam.w, 0, cam.h, 0.0f, 60.0f, leaf.data(), dsq.data(), xyz.data(), depth.data());

    // bvh::Tree::raycastWithScreenCoordinatesCuda
    auto px = tree.raycastWithScreenCoordinatesCuda(cam.intrinsics(), vp.pose().getTransformationImageToWorld(), 0, cam.w, 0,
                                                    cam.h, 0.0f, 60.0f);
    CHECK(px.size() == np);
    for (std::size_t i = 0; i < np; ++i) {
      const auto& r = px[i].intersection_result;
      CHECK(r.node == (leaf[i] < 0 ? nullptr : leaves[leaf[i]].node));
      CHECK(r.dist_sq == dsq[i]);
      CHECK(r.depth == depth[i]);
      CHECK(r.intersection[0] == xyz[3 * i] && r.intersection[1] == xyz[3 * i + 1] && r.intersection[2] == xyz[3 * i + 2]);
      CHECK(px[i].screen_coordinates[0] == (float)(i % cam.w) && px[i].screen_coordinates[1] == (float)(i / cam.w));
      total_hits += r.node != nullptr;
    }
    auto px2 = tree.raycastCuda(cam.intrinsics(), vp.pose().getTransformationImageToWorld(), 8, 40, 4, 20, 0.0f, 60.0f);
    CHECK(px2.size() == 32 * 16);
    for (std::size_t y = 0; y < 16; ++y)
      for (std::size_t x = 0; x < 32; ++x) CHECK(px2[y * 32 + x].node == px[(y + 4) * cam.w + x + 8].intersection_result.node);

    // ViewpointRaycast::getRaycastHitVoxelsWithScreenCoordinates(remove_duplicates = true)
    auto uniq = raycaster.getRaycastHitVoxelsWithScreenCoordinates(vp, 0, cam.w, 0, cam.h, true);
    std::vector<int32_t> o_leaf(np);
    std::vector<float> o_info(np), o_dsq(np);
    std::vector<uint32_t> o_pix(np);
    uint32_t hc = 0;
    float tot = 0, tot32 = 0;
    orc_score_opts all = oo;
    all.ignore_voxels_with_zero_information = 0;
    size_t k = orc_score_viewpoint(oracle, weights.data(), normals.data(), Kf, (uint32_t)cam.w, Rt, 0, cam.w, 0, cam.h, &all, np,
                                   o_leaf.data(), o_info.data(), o_pix.data(), o_dsq.data(), &hc, &tot, &tot32);
    CHECK(uniq.size() == k && k == hc);
    for (size_t i = 0; i < k; ++i) {
      CHECK(uniq[i].intersection_result.node == leaves[o_leaf[i]].node);
      CHECK(uniq[i].intersection_result.dist_sq == o_dsq[i]);
      CHECK(uniq[i].screen_coordinates[0] == (float)(o_pix[i] % cam.w));
      CHECK(uniq[i].screen_coordinates[1] == (float)(o_pix[i] / cam.w));
    }
    auto nodup = raycaster.getRaycastHitVoxelsWithScreenCoordinates(vp, 0, cam.w, 0, cam.h, false);
    std::size_t nh = 0;
    for (std::size_t i = 0; i < np; ++i) nh += leaf[i] >= 0;
    CHECK(nodup.size() == nh);
    {
      // remove_duplicates = false (the reference's updateWeightsWithRealViewpoints mode, viewpoint_planner_data.cpp:522):
      // every hit pixel in scan order, misses dropped -- the properly compacted list the reference's buggy
      // removeInvalidRaycastHitVoxels (viewpoint_raycast.cpp:287-303) was meant to produce
      std::size_t j = 0;
      for (std::size_t i = 0; i < np; ++i) {
        if (leaf[i] < 0) continue;
        const auto& r = nodup[j].intersection_result;
        CHECK(r.node == leaves[leaf[i]].node && r.dist_sq == dsq[i] && r.depth == depth[i]);
        CHECK(r.intersection[0] == xyz[3 * i] && r.intersection[1] == xyz[3 * i + 1] && r.intersection[2] == xyz[3 * i + 2]);
        CHECK(nodup[j].screen_coordinates[0] == (float)(i % cam.w) && nodup[j].screen_coordinates[1] == (float)(i / cam.w));
        ++j;
      }
      CHECK(j == nh);
      // ... also on a sub-window
      auto sub = raycaster.getRaycastHitVoxelsWithScreenCoordinates(vp, 8, 40, 4, 20, false);
      j = 0;
      for (std::size_t y = 4; y < 20; ++y)
        for (std::size_t x = 8; x < 40; ++x) {
          const std::size_t i = y * cam.w + x;
          if (leaf[i] < 0) continue;
          CHECK(j < sub.size());
          CHECK(sub[j].intersection_result.node == leaves[leaf[i]].node && sub[j].intersection_result.dist_sq == dsq[i]);
          CHECK(sub[j].screen_coordinates[0] == (float)x && sub[j].screen_coordinates[1] == (float)y);
          ++j;
        }
      CHECK(j == sub.size());
    }

    // ViewpointPlanner::getRaycastHitVoxelsWithInformationScore(viewpoint, ignore_zero = true)
    auto scored = raycaster.getRaycastHitVoxelsWithInformationScore(vp, true);
    k = orc_score_viewpoint(oracle, weights.data(), normals.data(), Kf, (uint32_t)cam.w, Rt, 0, cam.w, 0, cam.h, &oo, np,
                            o_leaf.data(), o_info.data(), o_pix.data(), o_dsq.data(), &hc, &tot, &tot32);
    CHECK(scored.first.size() == k);
    {
      // the set is hashed by voxel pointer but compares (voxel, information), like planner/occupied_tree.h:62-92
      std::unordered_map<const mock::Node*, float> by_voxel;
      for (const auto& vi : scored.first) by_voxel[vi.voxel] = vi.information;
      CHECK(by_voxel.size() == k);
      for (size_t i = 0; i < k; ++i) {
        auto it = by_voxel.find(leaves[o_leaf[i]].node);
        CHECK(it != by_voxel.end());
        CHECK(std::fabs(it->second - o_info[i]) <= 1e-5f * std::fabs(o_info[i]));
      }
    }
    CHECK(std::fabs(scored.second - tot) <= 1e-5f * std::fabs(tot));
  }
  CHECK(total_hits > 5000);

  // --- options_.enable_opengl = true: normals from the Poisson mesh at the kept pixel (row N2) ------------------
  {
    // ground top (z = 0), the wall at y = 2 and the wall at x = -2.75 as six triangles with tilted corner normals
    std::vector<float> tv, tn;
    auto quad = [&](const float a[3], const float b[3], const float c[3], const float d[3], float nx, float ny, float nz) {
      const float* tri[2][3] = {{a, b, c}, {a, c, d}};
      for (int t = 0; t < 2; ++t)
        for (int k = 0; k < 3; ++k) {
          for (int i = 0; i < 3; ++i) tv.push_back(tri[t][k][i]);
          const float tilt = 0.1f * (float)(k + t);
          const float n[3] = {nx + tilt * ny, ny + tilt * nz, nz + tilt * nx};
          const float len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
          for (int i = 0; i < 3; ++i) tn.push_back(n[i] / len);
        }
    };
    const float g0[3] = {-6, -6, 0}, g1[3] = {6, -6, 0}, g2[3] = {6, 6, 0}, g3[3] = {-6, 6, 0};
    const float w0[3] = {-2.5f, 2, 0}, w1[3] = {2.5f, 2, 0}, w2[3] = {2.5f, 2, 3}, w3[3] = {-2.5f, 2, 3};
    const float x0[3] = {-2.75f, -2.5f, 0}, x1[3] = {-2.75f, 2.5f, 0}, x2[3] = {-2.75f, 2.5f, 3}, x3[3] = {-2.75f, -2.5f, 3};
    quad(g0, g1, g2, g3, 0, 0, 1);
    quad(w0, w1, w2, w3, 0, -1, 0);
    quad(x0, x1, x2, x3, 1, 0, 0);
    q3dr::PoissonMeshNormalsCuda mesh_normals(tv, tn, 0);
    raycaster.setPoissonMeshNormals(&mesh_normals);
    const std::size_t np = cam.w * cam.h;
    for (const mock::Viewpoint& vp : vps) {
      float Rt[12];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) Rt[4 * r + c] = vp.p.T.m[r][c];
      // drawPoissonMeshNormals == the oracle's brute-force render, pixel for pixel
      std::vector<uint32_t> image = mesh_normals.drawPoissonMeshNormals(vp);
      std::vector<uint32_t> ref_image(np);
      orc_mesh_render_normals(tv.data(), tn.data(), tv.size() / 9, Kf, (uint32_t)cam.w, (uint32_t)cam.h, Rt, 0.5f, 1e5f,
                              0xFFFFFFFFu, ref_image.data());
      CHECK(image == ref_image);
      float nrm[3], ref_nrm[3];
      mesh_normals.computePoissonMeshNormalVector(vp, 17, 40, nrm);
      orc_decode_normal(ref_image[40 * cam.w + 17], ref_nrm);
      CHECK(nrm[0] == ref_nrm[0] && nrm[1] == ref_nrm[1] && nrm[2] == ref_nrm[2]);
      // scoring with those normals
      auto scored = raycaster.getRaycastHitVoxelsWithInformationScore(vp, true);
      std::vector<int32_t> o_leaf(np);
      std::vector<float> o_info(np);
      uint32_t hc = 0;
      float tot = 0;
      const size_t k = orc_score_viewpoint_image_normals(oracle, weights.data(), ref_image.data(), (uint32_t)cam.h, Kf,
                                                         (uint32_t)cam.w, Rt, 0, cam.w, 0, cam.h, &oo, np, o_leaf.data(),
                                                         o_info.data(), nullptr, nullptr, &hc, &tot);
      CHECK(scored.first.size() == k);
      std::unordered_map<const mock::Node*, float> by_voxel;
      for (const auto& vi : scored.first) by_voxel[vi.voxel] = vi.information;
      for (size_t i = 0; i < k; ++i) {
        auto it = by_voxel.find(leaves[o_leaf[i]].node);
        CHECK(it != by_voxel.end());
        CHECK(std::fabs(it->second - o_info[i]) <= 1e-5f * std::fabs(o_info[i]));
      }
      CHECK(std::fabs(scored.second - tot) <= 1e-5f * std::fabs(tot));
    }
    raycaster.setPoissonMeshNormals(nullptr);
  }

  // bvh::Tree::intersectsCuda
  std::vector<Tree::Ray> rays;
  for (int i = 0; i < 257; ++i) {
    Tree::Ray r;
    r.origin[0] = 0.11f * (i % 17) - 1.0f;
    r.origin[1] = -4.0f + 0.05f * (i % 5);
    r.origin[2] = 0.3f + 0.01f * i;
    r.direction[0] = 0.02f * (i % 11) - 0.1f;
    r.direction[1] = 1.0f;
    r.direction[2] = -0.002f * i;
    rays.push_back(r);
  }
  auto hits = tree.intersectsCuda(rays, 0.0f, -1.0f);
  {
    std::vector<float> o(3 * rays.size()), d(3 * rays.size()), dsq(rays.size()), xyz(3 * rays.size());
    std::vector<int32_t> leaf(rays.size());
    std::vector<uint32_t> depth(rays.size());
    for (size_t i = 0; i < rays.size(); ++i)
      for (int k = 0; k < 3; ++k) {
        o[3 * i + k] = rays[i].origin[k];
        d[3 * i + k] = rays[i].direction[k];
      }
    orc_intersect_rays(oracle, o.data(), d.data(), rays.size(), 0.0f, -1.0f, leaf.data(), dsq.data(), xyz.data(), depth.data());
    for (size_t i = 0; i < rays.size(); ++i) {
      CHECK(hits[i].first == (leaf[i] >= 0));
      CHECK(hits[i].second.node == (leaf[i] < 0 ? nullptr : leaves[leaf[i]].node));
      CHECK(hits[i].second.dist_sq == dsq[i]);
    }
  }

  // bbox queries (bvh.h:544-554) and isValidObjectPosition batches (viewpoint_planner_data.cpp:209-235)
  {
    const float qmin[3] = {-0.6f, 1.6f, 0.1f}, qmax[3] = {0.6f, 2.6f, 0.9f};
    auto bb = tree.intersects(qmin, qmax);
    const float q6[6] = {qmin[0], qmin[1], qmin[2], qmax[0], qmax[1], qmax[2]};
    std::vector<int32_t> ref(n);
    const size_t cnt = orc_overlap_box(oracle, q6, n, ref.data());
    CHECK(cnt > 4 && bb.size() == cnt);
    for (size_t i = 0; i < cnt; ++i) {
      CHECK(bb[i].node == leaves[ref[i]].node);
      CHECK(bb[i].depth == odepth[ref[i]]);
    }
    const float obj[6] = {-0.3f, -0.3f, -0.15f, 0.3f, 0.3f, 0.15f};
    float root[6];
    orc_tree_root_box(oracle, root);
    std::vector<float> pos;
    for (int i = 0; i < 400; ++i) {
      pos.push_back(-5.5f + 0.027f * i);
      pos.push_back(-4.0f + 0.019f * i);
      pos.push_back(-0.2f + 0.011f * i);
    }
    auto valid = tree.isValidObjectPositionBatch(pos.data(), pos.size() / 3, obj, root, 3.5f);
    size_t n_valid = 0;
    for (size_t i = 0; i < valid.size(); ++i) {
      CHECK((int)valid[i] == orc_valid_object_position(oracle, &pos[3 * i], obj, root, 3.5f));
      n_valid += valid[i];
    }
    CHECK(n_valid > 20 && n_valid < valid.size() - 20);
  }

  // the path search's information sums on the device-resident batch result
  {
    std::vector<const mock::Viewpoint*> batch;
    for (const mock::Viewpoint& vp : vps) batch.push_back(&vp);
    auto scored = raycaster.getRaycastHitVoxelsWithInformationScoreBatch(batch, 0, cam.w, 0, cam.h, true);
    std::unordered_map<const mock::Node*, int32_t> leaf_of;
    for (std::size_t k = 0; k < n; ++k) leaf_of[leaves[k].node] = (int32_t)k;
    q3dr::ObservedVoxelMapCuda<mock::Node> observed(&tree);
    std::vector<float> obs_ref(n, std::nanf(""));
    std::vector<uint32_t> all;
    for (uint32_t v = 0; v < batch.size(); ++v) all.push_back(v);
    for (uint32_t round = 0; round < 3; ++round) {
      auto ni = observed.computeNewInformation(all);
      uint32_t best = 0;
      for (uint32_t v = 0; v < batch.size(); ++v) {
        // the same voxel set as (leaf, info) arrays in ascending leaf order
        std::vector<std::pair<int32_t, float> > set;
        for (const auto& vi : scored[v].first) set.push_back(std::make_pair((int32_t)leaf_of[vi.voxel], vi.information));
        std::sort(set.begin(), set.end());
        std::vector<int32_t> sl;
        std::vector<float> si;
        for (const auto& e : set) { sl.push_back(e.first); si.push_back(e.second); }
        double s64 = 0;
        orc_new_information(sl.data(), si.data(), sl.size(), obs_ref.data(), weights.data(), 1.0f / 3.0f, &s64);
        CHECK(std::fabs(ni[v] - s64) <= 1e-5 * std::fabs(s64) + 1e-30);
        if (ni[v] > ni[best]) best = v;
      }
      std::vector<std::pair<int32_t, float> > set;
      for (const auto& vi : scored[best].first) set.push_back(std::make_pair((int32_t)leaf_of[vi.voxel], vi.information));
      std::sort(set.begin(), set.end());
      std::vector<int32_t> sl;
      std::vector<float> si;
      for (const auto& e : set) { sl.push_back(e.first); si.push_back(e.second); }
      double s64 = 0;
      orc_observed_add(sl.data(), si.data(), sl.size(), obs_ref.data(), weights.data(), 1.0f / 3.0f, &s64);
      const float added = observed.addObservedVoxels(best);
      CHECK(std::fabs(added - s64) <= 1e-5 * std::fabs(s64) + 1e-30);
    }
    // the same through the reference's own set type (a stored ViewpointEntry::voxel_set)
    {
      q3dr::ObservedVoxelMapCuda<mock::Node> by_set(&tree);
      q3dr::ObservedVoxelMapCuda<mock::Node> by_index(&tree);
      for (uint32_t v = 0; v < batch.size(); ++v) {
        const float g1 = by_set.computeNewInformation(scored[v].first);
        const float g2 = by_index.computeNewInformation(std::vector<uint32_t>(1, v))[0];
        CHECK(std::fabs(g1 - g2) <= 1e-6f * std::fabs(g2));
        const float a1 = by_set.addObservedVoxels(scored[v].first);
        const float a2 = by_index.addObservedVoxels(v);
        CHECK(std::fabs(a1 - a2) <= 1e-6f * std::fabs(a2));
      }
    }
  }

  // the CSR-backed voxel sets: same container contract as the reference's unordered_set, no per-voxel host work
  {
    std::vector<const mock::Viewpoint*> batch;
    for (const mock::Viewpoint& vp : vps) batch.push_back(&vp);
    for (int ignore_zero = 0; ignore_zero < 2; ++ignore_zero) {
      auto sets = raycaster.getRaycastHitVoxelsWithInformationScoreBatch(batch, 0, cam.w, 0, cam.h, ignore_zero != 0);
      auto views = raycaster.getRaycastHitVoxelsWithInformationScoreBatchView(batch, 0, cam.w, 0, cam.h, ignore_zero != 0);
      CHECK(sets.size() == batch.size() && views.size() == batch.size());
      for (std::size_t v = 0; v < batch.size(); ++v) {
        const auto& set = sets[v].first;
        const auto& view = views[v].first;
        if (view.size() != set.size()) std::fprintf(stderr, "viewpoint %zu: view %zu vs set %zu voxels\n", v, view.size(), set.size());
        CHECK(view.size() == set.size());
        CHECK(view.empty() == set.empty());
        CHECK(views[v].second == sets[v].second);
        std::size_t cnt = 0;
        int32_t prev = -1;
        float sum = 0;
        for (const q3dr::VoxelWithInformation<mock::Node>& vi : view) {  // the reference's loop, verbatim
          CHECK(set.find(vi) != set.end());
          sum += vi.information;
          ++cnt;
        }
        CHECK(cnt == set.size());
        for (auto it = view.begin(); it != view.end(); ++it) {
          CHECK(it.leaf() > prev);
          prev = it.leaf();
          CHECK(it->voxel == leaves[(std::size_t)it.leaf()].node && it->information == (*it).information);
        }
        for (const auto& vi : set) {
          const auto it = view.find(vi);  // viewpoint_planner_path.cpp:842-843
          CHECK(it != view.end());
          CHECK(it->voxel == vi.voxel && it->information == vi.information);
          CHECK(view.count(vi) == 1);
        }
        if (view.empty()) continue;
        // same voxel, other information: not "in" the set (VoxelWithInformation::operator==, occupied_tree.h:73-79)
        const q3dr::VoxelWithInformation<mock::Node> first = *view.begin();
        CHECK(view.find(q3dr::VoxelWithInformation<mock::Node>(first.voxel, first.information + 1.0f)) == view.end());
        mock::Node stranger;
        CHECK(view.find(q3dr::VoxelWithInformation<mock::Node>(&stranger, first.information)) == view.end());
        CHECK(view.toUnorderedSet() == set);
        CHECK((std::size_t)(view.end() - view.begin()) == view.size());
        // ViewpointEntry keeps the set by value: moving the view keeps it valid after the batch vector is gone
        q3dr::VoxelWithInformationSetView<mock::Node> kept = std::move(views[v].first);
        CHECK(kept.size() == set.size() && kept.find(first) != kept.end());
      }
      // views of an earlier call stay intact when the map's host buffers are reused by the next call
      auto again = raycaster.getRaycastHitVoxelsWithInformationScoreBatchView(batch, 0, cam.w, 0, cam.h, ignore_zero != 0);
      auto other = raycaster.getRaycastHitVoxelsWithInformationScoreBatchView(batch, 8, 40, 4, 20, ignore_zero != 0);
      for (std::size_t v = 0; v < batch.size(); ++v) {
        CHECK(again[v].first.toUnorderedSet() == sets[v].first);
        CHECK(other[v].first.size() <= again[v].first.size());
      }
    }
    // a batch larger than one piece of the pipelined conversion (helper thread drives the engine, this thread's team
    // fills the sets): every set equals the view of the same viewpoint, copies and moves of arena-backed sets behave
    {
      std::vector<const mock::Viewpoint*> big;
      for (int rep = 0; rep < 210; ++rep)
        for (const mock::Viewpoint& vp : vps) big.push_back(&vp);
      CHECK(big.size() > 600);
      auto sets_big = raycaster.getRaycastHitVoxelsWithInformationScoreBatch(big, 0, cam.w, 0, cam.h, true);
      auto views_big = raycaster.getRaycastHitVoxelsWithInformationScoreBatchView(big, 0, cam.w, 0, cam.h, true);
      CHECK(sets_big.size() == big.size() && views_big.size() == big.size());
      std::size_t total_voxels = 0;
      for (std::size_t v = 0; v < big.size(); ++v) {
        CHECK(sets_big[v].second == views_big[v].second);
        CHECK(sets_big[v].first.size() == views_big[v].first.size());
        if (v % 37 == 0) CHECK(views_big[v].first.toUnorderedSet() == sets_big[v].first);
        CHECK(sets_big[v].first == sets_big[v % vps.size()].first);  // the same viewpoint again
        total_voxels += sets_big[v].first.size();
      }
      CHECK(total_voxels > 10000);
      q3dr::VoxelWithInformationSet<mock::Node> copy = sets_big[3].first;  // a copy owns an arena of its own
      q3dr::VoxelWithInformationSet<mock::Node> moved = std::move(sets_big[3].first);
      CHECK(copy == moved && copy.size() == views_big[3].first.size());
      sets_big.clear();  // the moved-from vector dies; copy and moved live on
      CHECK(copy == moved);
      copy.erase(copy.begin());
      CHECK(copy.size() + 1 == moved.size());
      copy.insert(*moved.begin());
      copy.insert(*views_big[3].first.begin());
      CHECK(copy.size() >= moved.size());
    }
    // the set consumers take a view as well
    auto views = raycaster.getRaycastHitVoxelsWithInformationScoreBatchView(batch, 0, cam.w, 0, cam.h, true);
    auto sets = raycaster.getRaycastHitVoxelsWithInformationScoreBatch(batch, 0, cam.w, 0, cam.h, true);
    q3dr::ObservedVoxelMapCuda<mock::Node> a(&tree), b(&tree);
    for (std::size_t v = 0; v < batch.size(); ++v) {
      CHECK(a.computeNewInformation(views[v].first) == b.computeNewInformation(sets[v].first));
      CHECK(a.addObservedVoxels(views[v].first) == b.addObservedVoxels(sets[v].first));
    }
  }

  // updateWeightsWithRealViewpoints through the shim against the oracle's literal loop (must stay the LAST user of the
  // tree's weights: it lowers them)
  {
    const std::size_t np = cam.w * cam.h;
    std::vector<float> Rt, depth(2 * np, 3.0f);
    for (int v = 0; v < 2; ++v)
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) Rt.push_back(vps[v].p.T.m[r][c]);
    for (std::size_t y = 0; y < cam.h; ++y)
      for (std::size_t x = 0; x < cam.w; ++x) {
        if ((x / 8 + y / 8) % 2 == 0) depth[y * cam.w + x] = 0.0f;               // image 0: a checkerboard of holes
        depth[np + y * cam.w + x] = (x % 3 == 0) ? -1.0f : 2.0f;                 // image 1: every third column
      }
    q3dr_score_opts so;
    q3dr_score_opts_default(&so);
    so.raycast_min_range = 5.0f;
    so.raycast_max_range = 3.0e38f;  // the reference's default for this pass: unlimited
    std::uint64_t n_updates = 0;
    std::vector<float> lowered = tree.updateWeightsWithRealViewpoints(Kf, cam.w, cam.h, Rt, depth, so, 0.5f, &n_updates);
    std::vector<float> ref_w = weights;
    orc_score_opts ro = oo;
    ro.raycast_min_range = 5.0f;
    ro.raycast_max_range = -1.0f;
    std::uint64_t ref_updates = 0;
    for (int v = 0; v < 2; ++v) {
      ref_updates += orc_downweight_real_viewpoint(oracle, ref_w.data(), normals.data(), nullptr, Kf, (uint32_t)cam.w, (uint32_t)cam.h,
                                                   Rt.data() + 12 * v, depth.data() + v * np, &ro, 0.5f);
    }
    if (n_updates != ref_updates) std::fprintf(stderr, "down-weighting: %llu updates vs %llu in the oracle\n", (unsigned long long)n_updates, (unsigned long long)ref_updates);
    CHECK(n_updates == ref_updates);
    CHECK(n_updates > 100);
    std::size_t touched = 0;
    for (std::size_t k = 0; k < n; ++k) {
      CHECK(std::fabs(lowered[k] - ref_w[k]) <= 1e-5f * std::fabs(ref_w[k]) + 1e-30f);
      if (ref_w[k] != weights[k]) ++touched;
      if (ref_w[k] == weights[k]) CHECK(lowered[k] == weights[k]);
    }
    CHECK(touched > 50);
  }

  // error convention: callers catch OccupiedTreeType::Error
  bool threw = false;
  try {
    tree.raycastCuda(cam.intrinsics(), vps[0].pose().getTransformationImageToWorld(), 10, 10, 0, 5);
  } catch (const Tree::Error& e) {
    threw = true;
  }
  CHECK(threw);
  tree.clear();
  CHECK(!tree.isInitialized());
  threw = false;
  try {
    tree.intersectsCuda(rays);
  } catch (const Tree::Error&) {
    threw = true;
  }
  CHECK(threw);

  orc_tree_free(oracle);
  std::printf("shim parity OK: %zu leaves, %zu pixel hits checked\n", n, total_hits);
  return 0;
}
