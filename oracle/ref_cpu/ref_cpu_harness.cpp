// ref_cpu_harness.cpp -- C entry points over the REFERENCE'S OWN CPU code: bvh::Tree (viewpoint_planner/src/bvh/bvh.h)
// and bh::BoundingBox3D / bh::Ray (include/bh/math/geometry.h), compiled from the reference checkout where it lies
// (nothing is copied into this repository) against the stand-in Eigen / Boost headers in stubs/.  Test infrastructure
// only: tests/test_refcpu_pin.py compares these outputs with the oracle's restatement (oracle/q3dr_oracle.cpp).
//
// What this pins: Tree::build -> splitMedian (leaf order = tie-break order, bvh.h:361-373,764-788), the union boxes,
// Tree::intersects(ray) -> intersectsRecursive -> BoundingBox3D::isOutside / intersects (bvh.h:376-391,842-909,
// geometry.h:235-238,319-351), the all-node iterator order (bvh.h:203-246 = the serialised voxel index), and
// Tree::intersects(bbox) (bvh.h:544-554,911-939).  What it cannot pin: Eigen's own arithmetic (stubs/Eigen/Dense says
// which convention it uses) and the reference's fast-math build flag -- the recipe compiles strict IEEE, which is the
// contract.
#include <cstdint>
#include <cstdio>
#include <sstream>
#include <unordered_map>
#include <stack>
#include <vector>

#include <bvh.h>

namespace {

struct Obj {
  uint32_t index;
};
typedef bvh::Tree<Obj, float> Tree;

struct Handle {
  Tree tree;
  std::vector<Obj> objects;
  std::unordered_map<const Tree::NodeType*, uint32_t> dfs_rank;  // leaf node -> position in the left-to-right leaf order
};

void collect(const Tree::NodeType* node, uint32_t depth, std::vector<const Tree::NodeType*>* leaves, std::vector<uint32_t>* depths) {
  if (node->isLeaf()) {
    leaves->push_back(node);
    depths->push_back(depth);
    return;
  }
  if (node->hasLeftChild()) collect(node->getLeftChild(), depth + 1, leaves, depths);
  if (node->hasRightChild()) collect(node->getRightChild(), depth + 1, leaves, depths);
}

}  // namespace

extern "C" {

// boxes: n x 6 (min xyz, max xyz) in the order handed to Tree::build
void* refcpu_tree_build(const float* boxes, size_t n) {
  Handle* h = new Handle;
  h->objects.resize(n);
  std::vector<Tree::ObjectWithBoundingBox> objs(n);
  for (size_t i = 0; i < n; ++i) {
    h->objects[i].index = (uint32_t)i;
    objs[i].bounding_box = Tree::BoundingBoxType(Eigen::Vector3f(boxes[6 * i], boxes[6 * i + 1], boxes[6 * i + 2]),
                                                 Eigen::Vector3f(boxes[6 * i + 3], boxes[6 * i + 4], boxes[6 * i + 5]));
    objs[i].object = &h->objects[i];
  }
  std::streambuf* old = std::cout.rdbuf();
  std::ostringstream sink;
  std::cout.rdbuf(sink.rdbuf());  // build() prints tree statistics
  h->tree.build(objs, /*take_ownership=*/false);
  std::cout.rdbuf(old);
  std::vector<const Tree::NodeType*> leaves;
  std::vector<uint32_t> depths;
  collect(h->tree.getRoot(), 0, &leaves, &depths);
  for (size_t k = 0; k < leaves.size(); ++k) h->dfs_rank[leaves[k]] = (uint32_t)k;
  return h;
}

void refcpu_tree_free(void* hv) { delete static_cast<Handle*>(hv); }

size_t refcpu_tree_depth(void* hv) { return static_cast<Handle*>(hv)->tree.getDepth(); }
size_t refcpu_tree_num_nodes(void* hv) { return static_cast<Handle*>(hv)->tree.getNumOfNodes(); }
size_t refcpu_tree_num_leaves(void* hv) { return static_cast<Handle*>(hv)->tree.getNumOfLeafNodes(); }

// order[k] = input index of the k-th leaf, left to right; depth[k] = its depth; boxes_out[k] = its box (nullable)
void refcpu_leaf_order(void* hv, uint32_t* order, uint32_t* depth, float* boxes_out) {
  Handle* h = static_cast<Handle*>(hv);
  std::vector<const Tree::NodeType*> leaves;
  std::vector<uint32_t> depths;
  collect(h->tree.getRoot(), 0, &leaves, &depths);
  for (size_t k = 0; k < leaves.size(); ++k) {
    order[k] = leaves[k]->getObject()->index;
    if (depth) depth[k] = depths[k];
    if (boxes_out) {
      for (int i = 0; i < 3; ++i) {
        boxes_out[6 * k + i] = leaves[k]->getBoundingBox().getMinimum(i);
        boxes_out[6 * k + 3 + i] = leaves[k]->getBoundingBox().getMaximum(i);
      }
    }
  }
}

void refcpu_root_box(void* hv, float* box) {
  Handle* h = static_cast<Handle*>(hv);
  for (int i = 0; i < 3; ++i) {
    box[i] = h->tree.getRoot()->getBoundingBox().getMinimum(i);
    box[3 + i] = h->tree.getRoot()->getBoundingBox().getMaximum(i);
  }
}

// voxel_index[k] = position of the k-th leaf (left to right) in the reference's all-node iterator (begin() .. end())
void refcpu_voxel_index(void* hv, uint32_t* voxel_index) {
  Handle* h = static_cast<Handle*>(hv);
  uint32_t pos = 0;
  for (Tree::ConstIterator it = const_cast<const Tree&>(h->tree).begin(); it != const_cast<const Tree&>(h->tree).end(); ++it, ++pos) {
    auto f = h->dfs_rank.find(&*it);
    if (f != h->dfs_rank.end()) voxel_index[f->second] = pos;
  }
}

// Tree::intersects(bh::Ray(origin, direction), min_range, max_range) per ray.  leaf = left-to-right leaf rank or -1.
void refcpu_intersect_rays(void* hv, const float* origins, const float* directions, size_t n, float min_range, float max_range,
                           int32_t* leaf, float* dist_sq, float* xyz, uint32_t* depth) {
  Handle* h = static_cast<Handle*>(hv);
  for (size_t i = 0; i < n; ++i) {
    const bh::Ray<float> ray(Eigen::Vector3f(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]),
                             Eigen::Vector3f(directions[3 * i], directions[3 * i + 1], directions[3 * i + 2]));
    const std::pair<bool, Tree::IntersectionResult> r = h->tree.intersects(ray, min_range, max_range);
    if (r.first) {
      leaf[i] = (int32_t)h->dfs_rank.at(r.second.node);
      dist_sq[i] = r.second.dist_sq;
      depth[i] = (uint32_t)r.second.depth;
      for (int k = 0; k < 3; ++k) xyz[3 * i + k] = r.second.intersection(k);
    } else {
      leaf[i] = -1;
      dist_sq[i] = 0;
      depth[i] = 0;
      for (int k = 0; k < 3; ++k) xyz[3 * i + k] = 0;
    }
  }
}

// Tree::intersects(bbox): left-to-right leaf ranks in the reference's visiting order; returns the count (<= cap written)
size_t refcpu_overlap_box(void* hv, const float box[6], size_t cap, int32_t* out_leaf) {
  Handle* h = static_cast<Handle*>(hv);
  const Tree::BoundingBoxType bbox(Eigen::Vector3f(box[0], box[1], box[2]), Eigen::Vector3f(box[3], box[4], box[5]));
  const std::vector<Tree::BBoxIntersectionResult> res = h->tree.intersects(bbox);
  size_t k = 0;
  for (const Tree::BBoxIntersectionResult& r : res) {
    if (k < cap) out_leaf[k] = (int32_t)h->dfs_rank.at(r.node);
    ++k;
  }
  return k;
}

// BoundingBox3D(center, size) / constrainTo / isEmpty as generateBVHTree uses them (viewpoint_planner_data.cpp:833-860)
int refcpu_leaf_box(const float center[3], float size, const float constrain[6], float* box_out) {
  bh::BoundingBox3D<float> b(Eigen::Vector3f(center[0], center[1], center[2]), size);
  if (constrain) {
    b.constrainTo(bh::BoundingBox3D<float>(Eigen::Vector3f(constrain[0], constrain[1], constrain[2]),
                                           Eigen::Vector3f(constrain[3], constrain[4], constrain[5])));
  }
  for (int i = 0; i < 3; ++i) {
    box_out[i] = b.getMinimum(i);
    box_out[3 + i] = b.getMaximum(i);
  }
  return b.isEmpty() ? 1 : 0;
}

}  // extern "C"
