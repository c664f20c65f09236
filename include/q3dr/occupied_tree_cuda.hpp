// occupied_tree_cuda.hpp -- header-only C++ host side above the C ABI (q3dr_raycast.h).
//
// Mirrors, name for name, the CUDA bridge of the reference's bvh::Tree and the raycast / scoring calls
// layered on it, so that the reference's callers keep compiling unchanged (INTEGRATION.md shows the
// three-line patch).  Paths below are relative to the reference checkout.
//
//   q3dr::OccupiedTreeCuda<NodeT>            <- bvh::Tree<...>::{raycastCuda, raycastWithScreenCoordinatesCuda,
//                                               intersectsCuda, ensureCudaTreeIsInitialized, clear}
//                                               viewpoint_planner/src/bvh/bvh.h:396-542,949-955,296-314
//   ...::IntersectionResult[WithScreenCoordinates]   <- bvh.h:162-182
//   ...::Error / CudaError                   <- bvh.h:137-155 (callers catch OccupiedTreeType::Error)
//   q3dr::ViewpointRaycastCuda<NodeT>        <- viewpoint_planner::ViewpointRaycast (enable_cuda path)
//                                               viewpoint_planner/src/planner/viewpoint_raycast.h:19-143,
//                                               viewpoint_raycast.cpp:94-113,221-283,305-335
//   ...::getRaycastHitVoxelsWithInformationScore[Batch]
//                                            <- ViewpointPlanner::getRaycastHitVoxelsWithInformationScore
//                                               viewpoint_planner/src/planner/viewpoint_planner_raycast.cpp:89-148
//   q3dr::VoxelWithInformation / VoxelWithInformationSet  <- planner/occupied_tree.h:62-92
//   q3dr::VoxelWithInformationSetView        <- the same container contract (iterate / size / find / move) as a
//                                               window into the engine's CSR result
//   ...::intersects(bbox) / isValidObjectPositionBatch    <- bvh.h:544-554 ; viewpoint_planner_data.cpp:209-235
//   q3dr::ObservedVoxelMapCuda<NodeT>        <- ViewpointPath::observed_voxel_map + computeNewInformation /
//                                               addObservedVoxelsToViewpointPath / the stereo overlap sum
//                                               viewpoint_planner_scoring.cpp:124-147, viewpoint_planner_path.cpp:203-233,838-848
//
// Matrices are taken as templates that only need operator()(row, col): Eigen::Matrix4f / Matrix<float,3,4> in the
// reference tree, any small POD matrix elsewhere.  No Eigen, Boost or CUDA headers are needed to include this file.
#ifndef Q3DR_OCCUPIED_TREE_CUDA_HPP_
#define Q3DR_OCCUPIED_TREE_CUDA_HPP_

#include <algorithm>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <deque>
#include <exception>
#include <iterator>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

#include "../q3dr_raycast.h"

namespace q3dr {

/// Voxel node and its information (planner/occupied_tree.h:62-92): hashed and compared by voxel pointer.
template <typename NodeT>
struct VoxelWithInformation {
  VoxelWithInformation(const NodeT* voxel, float information) : voxel(voxel), information(information) {}
  const NodeT* voxel;
  float information;
  bool operator==(const VoxelWithInformation& other) const {
    return voxel == other.voxel && information == other.information;
  }
  struct VoxelHash {
    std::size_t operator()(const VoxelWithInformation& v) const { return std::hash<const NodeT*>()(v.voxel); }
  };
};

/// Monotonic arena behind one voxel set: hash nodes and bucket arrays are carved out of a few large chunks and freeing
/// is a no-op; the memory goes when the last container sharing the arena dies.  A scoring batch fills 10^7 hash nodes:
/// with the default allocator their allocation (and later their release, one free() per voxel) costs several times
/// what the GPU does.
class SetArena {
 public:
  SetArena() : cur_(nullptr), left_(0), next_chunk_(4096) {}
  SetArena(const SetArena&) = delete;
  SetArena& operator=(const SetArena&) = delete;
  /// makes room for `bytes` more without another chunk (called with the known size of a set before it is filled)
  void reserve(std::size_t bytes) {
    if (bytes > left_) add_chunk(bytes);
  }
  void* take(std::size_t bytes, std::size_t align) {
    std::size_t pad = padding(align);
    if (pad + bytes > left_) {
      add_chunk(bytes + align);
      pad = padding(align);
    }
    cur_ += pad;
    void* p = cur_;
    cur_ += bytes;
    left_ -= pad + bytes;
    return p;
  }

 private:
  std::size_t padding(std::size_t align) const {
    return (align - (reinterpret_cast<std::uintptr_t>(cur_) & (align - 1))) & (align - 1);
  }
  void add_chunk(std::size_t at_least) {
    const std::size_t n = std::max(at_least, next_chunk_);
    chunks_.emplace_back(new char[n]);
    cur_ = chunks_.back().get();
    left_ = n;
    next_chunk_ = std::min<std::size_t>(2 * next_chunk_, std::size_t(1) << 20);
  }
  std::vector<std::unique_ptr<char[]> > chunks_;
  char* cur_;
  std::size_t left_, next_chunk_;
};

/// Standard allocator over a SetArena.  A moved or swapped container takes its arena along; a COPY gets an arena of its
/// own (two copies may be filled by two threads).  Not thread-safe per arena, like the container it serves.
template <typename T>
class SetArenaAllocator {
 public:
  typedef T value_type;
  typedef std::true_type propagate_on_container_move_assignment;
  typedef std::true_type propagate_on_container_swap;
  typedef std::false_type propagate_on_container_copy_assignment;
  SetArenaAllocator() : arena_(std::make_shared<SetArena>()) {}
  template <typename U>
  SetArenaAllocator(const SetArenaAllocator<U>& o) : arena_(o.arena_ptr()) {}
  SetArenaAllocator select_on_container_copy_construction() const { return SetArenaAllocator(); }
  T* allocate(std::size_t n) { return static_cast<T*>(arena_->take(n * sizeof(T), alignof(T))); }
  void deallocate(T*, std::size_t) {}
  SetArena& arena() const { return *arena_; }
  const std::shared_ptr<SetArena>& arena_ptr() const { return arena_; }
  template <typename U>
  bool operator==(const SetArenaAllocator<U>& o) const { return arena_ == o.arena_ptr(); }
  template <typename U>
  bool operator!=(const SetArenaAllocator<U>& o) const { return arena_ != o.arena_ptr(); }

 private:
  std::shared_ptr<SetArena> arena_;
};

/// The reference's container (planner/occupied_tree.h:91: std::unordered_set<VoxelWithInformation, VoxelHash>) with the
/// arena allocator as its fourth template argument; everything a consumer does with the set is unchanged.
template <typename NodeT>
using VoxelWithInformationSet =
    std::unordered_set<VoxelWithInformation<NodeT>, typename VoxelWithInformation<NodeT>::VoxelHash,
                       std::equal_to<VoxelWithInformation<NodeT> >, SetArenaAllocator<VoxelWithInformation<NodeT> > >;

/// The per-viewpoint voxel sets of one batched scoring call, CSR form, shared by the views below.  The two arrays are
/// the engine's own pinned host arrays, handed over by q3dr_result_detach_host (no copy on the host); they go back to
/// the library's pool when the last view is gone.
template <typename NodeT>
struct VoxelSetBlock {
  VoxelSetBlock() : leaf(nullptr), information(nullptr), n_entries(0), nodes(nullptr), leaf_of(nullptr), tree(nullptr) {
    std::memset(&host, 0, sizeof(host));
  }
  ~VoxelSetBlock() { q3dr_host_block_release(&host); }
  VoxelSetBlock(const VoxelSetBlock&) = delete;
  VoxelSetBlock& operator=(const VoxelSetBlock&) = delete;
  q3dr_host_block host;
  const std::int32_t* leaf;     // ascending leaf index inside every viewpoint's range
  const float* information;
  std::size_t n_entries;
  std::vector<NodeT*> const* nodes;   // leaf index -> host node (owned by the OccupiedTreeCuda, which must outlive the views)
  std::unordered_map<const NodeT*, std::int32_t> const* (*leaf_of)(const void*);  // host node -> leaf index, built on first find()
  const void* tree;
};

/// A VoxelWithInformationSet (planner/occupied_tree.h:62-92) that is a window into the engine's CSR result instead of
/// a hash table of heap nodes: what the reference's consumers do with a voxel set -- range-for
/// (viewpoint_planner_graph.cpp:46-52, viewpoint_planner_path.cpp:208,456,841), size() / empty()
/// (viewpoint_planner_graph.cpp:29,53, viewpoint_planner_path.cpp:791), find() against end()
/// (viewpoint_planner_path.cpp:842-843), begin() / end() (:985-986), move into a ViewpointEntry -- works unchanged,
/// building it costs nothing per voxel.  Immutable; toUnorderedSet() gives the reference's own container.
/// Iteration order is ascending leaf index (the reference's is hash order, i.e. unspecified).
template <typename NodeT>
class VoxelWithInformationSetView {
 public:
  typedef VoxelWithInformation<NodeT> value_type;
  typedef std::size_t size_type;

  class const_iterator {
   public:
    typedef std::random_access_iterator_tag iterator_category;
    typedef VoxelWithInformation<NodeT> value_type;
    typedef std::ptrdiff_t difference_type;
    struct pointer {  // arrow proxy: it->voxel, it->information
      value_type v;
      const value_type* operator->() const { return &v; }
    };
    typedef value_type reference;  // by value: elements are assembled from (leaf, information) on the fly

    const_iterator() : block_(nullptr), i_(0) {}
    const_iterator(const VoxelSetBlock<NodeT>* block, std::size_t i) : block_(block), i_(i) {}
    value_type operator*() const {
      return value_type((*block_->nodes)[static_cast<std::size_t>(block_->leaf[i_])], block_->information[i_]);
    }
    pointer operator->() const { return pointer{**this}; }
    value_type operator[](difference_type k) const { return *(*this + k); }
    const_iterator& operator++() { ++i_; return *this; }
    const_iterator operator++(int) { const_iterator t(*this); ++i_; return t; }
    const_iterator& operator--() { --i_; return *this; }
    const_iterator operator--(int) { const_iterator t(*this); --i_; return t; }
    const_iterator& operator+=(difference_type k) { i_ += k; return *this; }
    const_iterator& operator-=(difference_type k) { i_ -= k; return *this; }
    const_iterator operator+(difference_type k) const { return const_iterator(block_, i_ + k); }
    const_iterator operator-(difference_type k) const { return const_iterator(block_, i_ - k); }
    difference_type operator-(const const_iterator& o) const {
      return static_cast<difference_type>(i_) - static_cast<difference_type>(o.i_);
    }
    bool operator==(const const_iterator& o) const { return i_ == o.i_; }
    bool operator!=(const const_iterator& o) const { return i_ != o.i_; }
    bool operator<(const const_iterator& o) const { return i_ < o.i_; }
    std::int32_t leaf() const { return block_->leaf[i_]; }

   private:
    const VoxelSetBlock<NodeT>* block_;
    std::size_t i_;
  };
  typedef const_iterator iterator;

  VoxelWithInformationSetView() : begin_(0), end_(0) {}
  VoxelWithInformationSetView(std::shared_ptr<const VoxelSetBlock<NodeT> > block, std::size_t begin, std::size_t end)
      : block_(std::move(block)), begin_(begin), end_(end) {}

  size_type size() const { return end_ - begin_; }
  bool empty() const { return end_ == begin_; }
  const_iterator begin() const { return const_iterator(block_.get(), begin_); }
  const_iterator end() const { return const_iterator(block_.get(), end_); }
  const_iterator cbegin() const { return begin(); }
  const_iterator cend() const { return end(); }

  /// Like the reference's unordered_set::find with VoxelWithInformation::operator== : the voxel must be in the set
  /// AND carry the same information (occupied_tree.h:73-79).
  const_iterator find(const value_type& vi) const {
    if (empty()) return end();
    const std::unordered_map<const NodeT*, std::int32_t>* table = block_->leaf_of(block_->tree);
    typename std::unordered_map<const NodeT*, std::int32_t>::const_iterator it = table->find(vi.voxel);
    if (it == table->end()) return end();
    const std::int32_t* first = block_->leaf + begin_;
    const std::int32_t* last = block_->leaf + end_;
    const std::int32_t* p = std::lower_bound(first, last, it->second);
    if (p == last || *p != it->second) return end();
    const std::size_t i = static_cast<std::size_t>(p - block_->leaf);
    return block_->information[i] == vi.information ? const_iterator(block_.get(), i) : end();
  }
  size_type count(const value_type& vi) const { return find(vi) == end() ? 0 : 1; }

  /// the engine's view of the set: ascending leaf indices and their informations
  const std::int32_t* leafData() const { return block_ ? block_->leaf + begin_ : nullptr; }
  const float* informationData() const { return block_ ? block_->information + begin_ : nullptr; }

  /// the reference's own container (one hash-node allocation per voxel)
  VoxelWithInformationSet<NodeT> toUnorderedSet() const {
    VoxelWithInformationSet<NodeT> set;
    set.reserve(size());
    for (const_iterator it = begin(); it != end(); ++it) set.insert(*it);
    return set;
  }

 private:
  std::shared_ptr<const VoxelSetBlock<NodeT> > block_;
  std::size_t begin_, end_;
};

/// One BVH leaf as generateBVHTree emits it (viewpoint_planner_data.cpp:820-897) plus its host node.
template <typename NodeT>
struct LeafRecord {
  float bbox_min[3];
  float bbox_max[3];
  float weight;      // NodeObject::weight
  float normal[3];   // NodeObject::normal (zero vector: no normal)
  NodeT* node;       // host bvh::Node, echoed back in results (the reference's CudaNode::ptr_)
  std::uint32_t depth;  // depth of the leaf in the host tree (IntersectionResult::depth)
};

/// Walks a reference-style host tree left to right (the order of the recursion in bvh.h:899-905) and emits its
/// leaves.  NodeT needs getLeftChild(), getRightChild(), isLeaf(), getBoundingBox().getMinimum(i)/getMaximum(i),
/// getObject()->weight / ->normal(i).
template <typename NodeT>
void collectLeaves(NodeT* node, std::uint32_t depth, std::vector<LeafRecord<NodeT>>* out) {
  if (node == nullptr) return;
  if (node->isLeaf()) {
    LeafRecord<NodeT> r;
    for (int i = 0; i < 3; ++i) {
      r.bbox_min[i] = node->getBoundingBox().getMinimum(i);
      r.bbox_max[i] = node->getBoundingBox().getMaximum(i);
      r.normal[i] = node->getObject()->normal(i);
    }
    r.weight = node->getObject()->weight;
    r.node = node;
    r.depth = depth;
    out->push_back(r);
    return;
  }
  collectLeaves(node->getLeftChild(), depth + 1, out);
  collectLeaves(node->getRightChild(), depth + 1, out);
}

template <typename NodeT>
class OccupiedTreeCuda {
 public:
  using NodeType = NodeT;

  /// bvh::Tree::Error / CudaError (bvh.h:137-155); the planner catches Error and drops the candidate.
  class Error : public std::runtime_error {
   public:
    Error(int status, const std::string& msg) : std::runtime_error(msg), status_(status) {}
    int status() const { return status_; }
   private:
    int status_;
  };
  class CudaError : public Error {
   public:
    using Error::Error;
  };

  struct IntersectionResult {  // bvh.h:162-172
    IntersectionResult() : intersection{0, 0, 0}, node(nullptr), depth(0), dist_sq(0) {}
    float intersection[3];
    NodeT* node;
    std::size_t depth;
    float dist_sq;
  };
  struct IntersectionResultWithScreenCoordinates {  // bvh.h:174-182
    IntersectionResultWithScreenCoordinates() : screen_coordinates{0, 0} {}
    IntersectionResult intersection_result;
    float screen_coordinates[2];
  };
  struct Ray {  // bh::Ray: origin + direction (normalised by the engine like bh::Ray's constructor)
    float origin[3];
    float direction[3];
  };

  OccupiedTreeCuda() = default;
  OccupiedTreeCuda(const OccupiedTreeCuda&) = delete;
  OccupiedTreeCuda& operator=(const OccupiedTreeCuda&) = delete;
  ~OccupiedTreeCuda() { clear(); }

  /// ensureCudaTreeIsInitialized (bvh.h:949-955): flatten + upload.  `leaves` in left-to-right order of the host tree.
  void ensureCudaTreeIsInitialized(const std::vector<LeafRecord<NodeT>>& leaves, int cuda_gpu_id = 0) {
    if (map_ != nullptr) return;
    const std::size_t n = leaves.size();
    std::vector<float> boxes(6 * n), weights(n), normals(3 * n);
    std::vector<std::uint32_t> depth(n);
    nodes_.resize(n);
    for (std::size_t k = 0; k < n; ++k) {
      for (int i = 0; i < 3; ++i) {
        boxes[6 * k + i] = leaves[k].bbox_min[i];
        boxes[6 * k + 3 + i] = leaves[k].bbox_max[i];
        normals[3 * k + i] = leaves[k].normal[i];
      }
      weights[k] = leaves[k].weight;
      depth[k] = leaves[k].depth;
      nodes_[k] = leaves[k].node;
    }
    depth_ = depth;
    check(q3dr_map_create(boxes.data(), weights.data(), normals.data(), depth.data(), n, cuda_gpu_id, &map_));
  }

  template <typename RootT>
  void ensureCudaTreeIsInitialized(RootT* host_root, int cuda_gpu_id = 0) {
    if (map_ != nullptr) return;
    std::vector<LeafRecord<NodeT>> leaves;
    collectLeaves<NodeT>(host_root, 0, &leaves);
    ensureCudaTreeIsInitialized(leaves, cuda_gpu_id);
  }

  bool isInitialized() const { return map_ != nullptr; }

  /// Tree::clear() frees the device tree (bvh.h:296-314).
  void clear() {
    if (map_ != nullptr) q3dr_map_destroy(map_);
    map_ = nullptr;
    nodes_.clear();
    depth_.clear();
    leaf_of_.clear();
  }

  /// ViewpointPlannerData::updateWeights*: NodeObject::weight changed on the host, refresh the device copy.
  void updateWeights(const std::vector<float>& weights_left_to_right) {
    check(q3dr_map_update_weights(map_, weights_left_to_right.data()));
  }

  /// ViewpointPlannerData::updateWeightsWithRealViewpoints (viewpoint_planner_data.cpp:498-571) for real images that
  /// share one depth camera: poses (Rt: n x 12, image-to-world [R|t] rows), depth maps (n x height x width, <= 0 or inf =
  /// nothing reconstructed at that pixel).  Lowers the weights on the device and returns them (left-to-right leaf
  /// order) so that the caller can write them back into NodeObject::weight and the octree, as the reference does.
  std::vector<float> updateWeightsWithRealViewpoints(const float K[4], std::size_t width, std::size_t height,
                                                     const std::vector<float>& Rt, const std::vector<float>& depth_maps,
                                                     const q3dr_score_opts& opts, float invalid_pixel_observation_factor,
                                                     std::uint64_t* n_updates = nullptr) {
    requireMap();
    std::vector<float> weights(nodes_.size());
    check(q3dr_map_downweight_real_viewpoints(map_, K, static_cast<std::uint32_t>(width), static_cast<std::uint32_t>(height),
                                              Rt.data(), depth_maps.data(), Rt.size() / 12, &opts,
                                              invalid_pixel_observation_factor, nullptr, nullptr, nullptr, weights.data(),
                                              n_updates));
    return weights;
  }

  std::size_t getNumOfLeafNodes() const { return nodes_.size(); }
  /// leaf index of a host node (the inverse of leafNode); -1 if the node is not a leaf of this tree
  std::int32_t leafIndex(const NodeT* node) const {
    if (leaf_of_.empty()) {
      leaf_of_.reserve(nodes_.size());
      for (std::size_t k = 0; k < nodes_.size(); ++k) leaf_of_[nodes_[k]] = static_cast<std::int32_t>(k);
    }
    typename std::unordered_map<const NodeT*, std::int32_t>::const_iterator it = leaf_of_.find(node);
    return it == leaf_of_.end() ? -1 : it->second;
  }
  NodeT* leafNode(std::int32_t leaf) const { return leaf < 0 ? nullptr : nodes_[static_cast<std::size_t>(leaf)]; }
  const std::vector<NodeT*>& leafNodes() const { return nodes_; }
  /// host node -> leaf index for all leaves (built once, on first use)
  const std::unordered_map<const NodeT*, std::int32_t>& leafIndexTable() const {
    leafIndex(nullptr);
    return leaf_of_;
  }
  static const std::unordered_map<const NodeT*, std::int32_t>* leafIndexTableOf(const void* tree) {
    return &static_cast<const OccupiedTreeCuda*>(tree)->leafIndexTable();
  }
  q3dr_map* handle() const { return map_; }

  /// bvh.h:396-451.  intrinsics: 4x4 (fx, fy, cx, cy are read, bvh.cu:329-330); extrinsics: 3x4 [R|t] image-to-world.
  template <typename Matrix4x4, typename Matrix3x4>
  std::vector<IntersectionResult> raycastCuda(const Matrix4x4& intrinsics, const Matrix3x4& extrinsics,
                                              std::size_t x_start, std::size_t x_end, std::size_t y_start,
                                              std::size_t y_end, float min_range = 0, float max_range = -1,
                                              bool fail_on_error = false) {
    std::vector<q3dr_pixel_hit> px = castWindow(intrinsics, extrinsics, x_start, x_end, y_start, y_end, min_range,
                                                max_range, fail_on_error);
    std::vector<IntersectionResult> out(px.size());
    for (std::size_t i = 0; i < px.size(); ++i) out[i] = convert(px[i]);
    return out;
  }

  /// bvh.h:453-509
  template <typename Matrix4x4, typename Matrix3x4>
  std::vector<IntersectionResultWithScreenCoordinates> raycastWithScreenCoordinatesCuda(
      const Matrix4x4& intrinsics, const Matrix3x4& extrinsics, std::size_t x_start, std::size_t x_end,
      std::size_t y_start, std::size_t y_end, float min_range = 0, float max_range = -1, bool fail_on_error = false) {
    std::vector<q3dr_pixel_hit> px = castWindow(intrinsics, extrinsics, x_start, x_end, y_start, y_end, min_range,
                                                max_range, fail_on_error);
    std::vector<IntersectionResultWithScreenCoordinates> out(px.size());
    for (std::size_t i = 0; i < px.size(); ++i) {
      out[i].intersection_result = convert(px[i]);
      out[i].screen_coordinates[0] = px[i].screen_x;
      out[i].screen_coordinates[1] = px[i].screen_y;
    }
    return out;
  }

  /// bvh.h:512-541: pair.first = hit?
  std::vector<std::pair<bool, IntersectionResult>> intersectsCuda(const std::vector<Ray>& rays, float min_range = 0,
                                                                  float max_range = -1) {
    requireMap();
    std::vector<float> o(3 * rays.size()), d(3 * rays.size());
    for (std::size_t i = 0; i < rays.size(); ++i) {
      for (int k = 0; k < 3; ++k) {
        o[3 * i + k] = rays[i].origin[k];
        d[3 * i + k] = rays[i].direction[k];
      }
    }
    std::vector<q3dr_pixel_hit> px(rays.size());
    check(q3dr_intersect_rays(map_, o.data(), d.data(), rays.size(), min_range, max_range, px.data()));
    std::vector<std::pair<bool, IntersectionResult>> out(rays.size());
    for (std::size_t i = 0; i < px.size(); ++i) out[i] = std::make_pair(px[i].leaf >= 0, convert(px[i]));
    return out;
  }

  /// bvh::Tree::intersects(bbox) (bvh.h:544-554): the leaf nodes whose box overlaps [bbox_min, bbox_max], in the
  /// reference's visiting order (left to right).
  struct BBoxIntersectionResult {  // bvh.h:184-195
    NodeT* node;
    std::size_t depth;
  };
  std::vector<BBoxIntersectionResult> intersects(const float bbox_min[3], const float bbox_max[3]) {
    float b[6] = {bbox_min[0], bbox_min[1], bbox_min[2], bbox_max[0], bbox_max[1], bbox_max[2]};
    return intersectsBatch(b, 1)[0];
  }
  /// Many boxes at once (boxes: n x 6, min xyz then max xyz).
  std::vector<std::vector<BBoxIntersectionResult> > intersectsBatch(const float* boxes, std::size_t n) {
    requireMap();
    q3dr_overlap_result r;
    check(q3dr_overlap_boxes(map_, boxes, n, &r));
    std::vector<std::vector<BBoxIntersectionResult> > out(n);
    for (std::size_t i = 0; i < n; ++i) {
      for (std::uint64_t e = r.offsets[i]; e < r.offsets[i + 1]; ++e) {
        BBoxIntersectionResult br;
        br.node = nodes_[static_cast<std::size_t>(r.leaf[e])];
        br.depth = depth_.empty() ? 0 : depth_[static_cast<std::size_t>(r.leaf[e])];
        out[i].push_back(br);
      }
    }
    return out;
  }

  /// ViewpointPlannerData::isValidObjectPosition (viewpoint_planner_data.cpp:209-235) for a batch of positions
  /// (n x 3); no-fly zones are the caller's business.  object_bbox / bvh_bbox: min xyz, max xyz.
  std::vector<bool> isValidObjectPositionBatch(const float* positions, std::size_t n, const float object_bbox[6],
                                               const float bvh_bbox[6], float obstacle_free_height) {
    requireMap();
    std::vector<std::uint8_t> v(n);
    check(q3dr_valid_object_positions(map_, positions, n, object_bbox, bvh_bbox, obstacle_free_height, v.data()));
    return std::vector<bool>(v.begin(), v.end());
  }

  IntersectionResult convert(const q3dr_pixel_hit& h) const {
    IntersectionResult r;
    r.intersection[0] = h.intersection[0];
    r.intersection[1] = h.intersection[1];
    r.intersection[2] = h.intersection[2];
    r.node = leafNode(h.leaf);
    r.depth = h.depth;
    r.dist_sq = h.dist_sq;
    return r;
  }

  void check(int status) const {
    if (status == Q3DR_OK) return;
    const std::string msg = std::string(q3dr_status_string(status)) + ": " + q3dr_last_error();
    if (status == Q3DR_ERR_CUDA || status == Q3DR_ERR_NO_DEVICE || status == Q3DR_ERR_OUT_OF_MEMORY) {
      // the reference frees and re-uploads the device tree before rethrowing (bvh.h:443-449)
      if (map_ != nullptr && status == Q3DR_ERR_CUDA) q3dr_map_reset(map_);
      throw CudaError(status, msg);
    }
    throw Error(status, msg);
  }

  void requireMap() const {
    if (map_ == nullptr) throw Error(Q3DR_ERR_INVALID_ARG, "device tree not initialised");
  }

 private:
  template <typename Matrix4x4, typename Matrix3x4>
  std::vector<q3dr_pixel_hit> castWindow(const Matrix4x4& intrinsics, const Matrix3x4& extrinsics, std::size_t x_start,
                                         std::size_t x_end, std::size_t y_start, std::size_t y_end, float min_range,
                                         float max_range, bool /*fail_on_error*/) {
    requireMap();
    float K[4], Rt[12];
    packCamera(intrinsics, extrinsics, K, Rt);
    std::vector<q3dr_pixel_hit> px((x_end - x_start) * (y_end - y_start));
    check(q3dr_raycast_window(map_, K, Rt, static_cast<std::uint32_t>(x_start), static_cast<std::uint32_t>(x_end),
                              static_cast<std::uint32_t>(y_start), static_cast<std::uint32_t>(y_end), min_range,
                              max_range, px.data()));
    return px;
  }

 public:
  template <typename Matrix4x4, typename Matrix3x4>
  static void packCamera(const Matrix4x4& intrinsics, const Matrix3x4& extrinsics, float K[4], float Rt[12]) {
    K[0] = intrinsics(0, 0);
    K[1] = intrinsics(1, 1);
    K[2] = intrinsics(0, 2);
    K[3] = intrinsics(1, 2);
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 4; ++c) Rt[4 * r + c] = extrinsics(r, c);
    }
  }

 private:
  q3dr_map* map_ = nullptr;
  std::vector<NodeT*> nodes_;  // leaf index -> host node
  std::vector<std::uint32_t> depth_;  // leaf index -> depth in the host tree
  mutable std::unordered_map<const NodeT*, std::int32_t> leaf_of_;  // host node -> leaf index (built on first use)
};

/// The normals part of viewpoint_planner::ViewpointOffscreenRenderer (enable_opengl = true): the Poisson mesh uploaded
/// once (uploadPoissonMesh, viewpoint_offscreen_renderer.cpp:138-200), drawPoissonMeshNormals (:361-397) and
/// computePoissonMeshNormalVector (:499-527) with its one-pose image cache.  Images are defined by the engine's ray cast
/// (q3dr_raycast.h, "image-space normals"), not by an OpenGL context.
class PoissonMeshNormalsCuda {
 public:
  /// tri_vertices / tri_normals: 9 floats per triangle (what uploadPoissonMesh puts into OGLTriangleWithNormalData);
  /// empty tri_normals => the reference's per-face normals
  PoissonMeshNormalsCuda(const std::vector<float>& tri_vertices, const std::vector<float>& tri_normals,
                         int cuda_gpu_id = 0)
      : mesh_(nullptr), cached_valid_(false), cached_w_(0), cached_h_(0) {
    q3dr_render_opts_default(&ropts_);
    if (!tri_normals.empty() && tri_normals.size() != tri_vertices.size()) {
      throw std::runtime_error("tri_normals must be empty or match tri_vertices");
    }
    const int st = q3dr_mesh_create(tri_vertices.data(), tri_normals.empty() ? nullptr : tri_normals.data(),
                                    tri_vertices.size() / 9, cuda_gpu_id, &mesh_);
    if (st != Q3DR_OK) throw std::runtime_error(std::string("q3dr_mesh_create: ") + q3dr_last_error());
  }
  PoissonMeshNormalsCuda(const PoissonMeshNormalsCuda&) = delete;
  PoissonMeshNormalsCuda& operator=(const PoissonMeshNormalsCuda&) = delete;
  ~PoissonMeshNormalsCuda() { q3dr_mesh_destroy(mesh_); }

  /// viewpoint_offscreen_renderer.cpp:76-79
  void setNearFarPlane(float near_plane, float far_plane) {
    ropts_.near_plane = near_plane;
    ropts_.far_plane = far_plane;
    cached_valid_ = false;
  }
  /// viewpoint_offscreen_renderer.cpp:43-45 (r, g, b, a in [0, 1])
  void setClearColor(float r, float g, float b, float a) {
    ropts_.clear_argb = (unorm8(a) << 24) | (unorm8(r) << 16) | (unorm8(g) << 8) | unorm8(b);
    cached_valid_ = false;
  }
  const q3dr_mesh* handle() const { return mesh_; }
  const q3dr_render_opts& renderOptions() const { return ropts_; }

  /// drawPoissonMeshNormals: width * height QImage::Format_ARGB32 pixels, row 0 on top
  template <typename ViewpointT>
  std::vector<std::uint32_t> drawPoissonMeshNormals(const ViewpointT& viewpoint) const {
    float K[4], Rt[12];
    packCamera(viewpoint, K, Rt);
    const std::uint32_t w = static_cast<std::uint32_t>(viewpoint.camera().width());
    const std::uint32_t h = static_cast<std::uint32_t>(viewpoint.camera().height());
    std::vector<std::uint32_t> image(static_cast<std::size_t>(w) * h);
    const int st = q3dr_mesh_render_normals(mesh_, K, w, h, Rt, 1, &ropts_, image.data());
    if (st != Q3DR_OK) throw std::runtime_error(std::string("q3dr_mesh_render_normals: ") + q3dr_last_error());
    return image;
  }

  /// computePoissonMeshNormalVector(viewpoint, x, y): re-renders only when the pose changed (:505-513)
  template <typename ViewpointT>
  void computePoissonMeshNormalVector(const ViewpointT& viewpoint, std::size_t x, std::size_t y, float normal_out[3]) const {
    float K[4], Rt[12];
    packCamera(viewpoint, K, Rt);
    const std::uint32_t w = static_cast<std::uint32_t>(viewpoint.camera().width());
    const std::uint32_t h = static_cast<std::uint32_t>(viewpoint.camera().height());
    bool same = cached_valid_ && w == cached_w_ && h == cached_h_;
    for (int i = 0; same && i < 12; ++i) same = Rt[i] == cached_rt_[i];
    if (!same) {
      cached_image_ = drawPoissonMeshNormals(viewpoint);
      for (int i = 0; i < 12; ++i) cached_rt_[i] = Rt[i];
      cached_w_ = w;
      cached_h_ = h;
      cached_valid_ = true;
    }
    q3dr_decode_normal(cached_image_[y * w + x], normal_out);
  }

 private:
  static std::uint32_t unorm8(float c) {
    c = c < 0.f ? 0.f : (c > 1.f ? 1.f : c);
    return static_cast<std::uint32_t>(c * 255.0f + 0.5f);
  }
  template <typename ViewpointT>
  static void packCamera(const ViewpointT& viewpoint, float K[4], float Rt[12]) {
    const auto& intrinsics = viewpoint.camera().intrinsics();
    const auto extrinsics = viewpoint.pose().getTransformationImageToWorld();
    K[0] = intrinsics(0, 0);
    K[1] = intrinsics(1, 1);
    K[2] = intrinsics(0, 2);
    K[3] = intrinsics(1, 2);
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 4; ++c) Rt[4 * r + c] = extrinsics(r, c);
    }
  }
  q3dr_mesh* mesh_;
  q3dr_render_opts ropts_;
  mutable bool cached_valid_;
  mutable std::uint32_t cached_w_, cached_h_;
  mutable float cached_rt_[12];
  mutable std::vector<std::uint32_t> cached_image_;
};

/// viewpoint_planner::ViewpointRaycast with enable_cuda = true, plus the planner's scoring facade.
/// ViewpointT needs camera().intrinsics() (4x4), camera().width(), camera().height() and
/// pose().getTransformationImageToWorld() (3x4).
template <typename NodeT>
class ViewpointRaycastCuda {
 public:
  using TreeType = OccupiedTreeCuda<NodeT>;
  using IntersectionResult = typename TreeType::IntersectionResult;
  using IntersectionResultWithScreenCoordinates = typename TreeType::IntersectionResultWithScreenCoordinates;
  using VoxelSet = VoxelWithInformationSet<NodeT>;

  /// viewpoint_raycast.h:28-31 (min_range is accepted and, like in the reference, never used by the query)
  ViewpointRaycastCuda(TreeType* bvh_tree, float min_range = 0, float max_range = 60)
      : bvh_tree_(bvh_tree), min_range_(min_range), max_range_(max_range), mesh_normals_(nullptr) {
    q3dr_score_opts_default(&opts_);
    opts_.raycast_min_range = min_range;
    opts_.raycast_max_range = max_range;
  }

  q3dr_score_opts& scoreOptions() { return opts_; }

  /// options_.enable_opengl = true (ViewpointPlanner::computeNormalVector, viewpoint_planner_scoring.cpp:26-38): score
  /// every voxel with the mesh normal at its kept pixel instead of NodeObject::normal.  nullptr switches back.
  void setPoissonMeshNormals(const PoissonMeshNormalsCuda* mesh_normals) { mesh_normals_ = mesh_normals; }

  /// viewpoint_raycast.cpp:94-113,254-283,321-335
  template <typename ViewpointT>
  std::vector<IntersectionResultWithScreenCoordinates> getRaycastHitVoxelsWithScreenCoordinates(
      const ViewpointT& viewpoint, std::size_t x_start, std::size_t x_end, std::size_t y_start, std::size_t y_end,
      bool remove_duplicates = true, bool fail_on_error = false) const {
    if (!remove_duplicates) {
      // every hit pixel, misses dropped (the reference's removeInvalidRaycastHitVoxels erases a single element
      // only -- viewpoint_raycast.cpp:287-303; this returns the properly compacted list)
      auto all = bvh_tree_->raycastWithScreenCoordinatesCuda(viewpoint.camera().intrinsics(),
                                                             viewpoint.pose().getTransformationImageToWorld(), x_start,
                                                             x_end, y_start, y_end, min_range_, max_range_, fail_on_error);
      std::vector<IntersectionResultWithScreenCoordinates> out;
      for (const auto& r : all) {
        if (r.intersection_result.node != nullptr) out.push_back(r);
      }
      return out;
    }
    bvh_tree_->requireMap();
    float K[4], Rt[12];
    TreeType::packCamera(viewpoint.camera().intrinsics(), viewpoint.pose().getTransformationImageToWorld(), K, Rt);
    std::vector<q3dr_pixel_hit> px((x_end - x_start) * (y_end - y_start));
    std::size_t count = 0;
    bvh_tree_->check(q3dr_raycast_unique(bvh_tree_->handle(), K, Rt, static_cast<std::uint32_t>(x_start),
                                         static_cast<std::uint32_t>(x_end), static_cast<std::uint32_t>(y_start),
                                         static_cast<std::uint32_t>(y_end), min_range_, max_range_, px.data(), px.size(),
                                         &count));
    std::vector<IntersectionResultWithScreenCoordinates> out(count);
    for (std::size_t i = 0; i < count; ++i) {
      out[i].intersection_result = bvh_tree_->convert(px[i]);
      out[i].screen_coordinates[0] = px[i].screen_x;
      out[i].screen_coordinates[1] = px[i].screen_y;
    }
    return out;
  }

  /// ViewpointPlanner::getRaycastHitVoxelsWithInformationScore (viewpoint_planner_raycast.cpp:116-148)
  template <typename ViewpointT>
  std::pair<VoxelSet, float> getRaycastHitVoxelsWithInformationScore(
      const ViewpointT& viewpoint, std::size_t x_start, std::size_t x_end, std::size_t y_start, std::size_t y_end,
      bool ignore_voxels_with_zero_information = false) const {
    std::vector<const ViewpointT*> one(1, &viewpoint);
    auto all = getRaycastHitVoxelsWithInformationScoreBatch(one, x_start, x_end, y_start, y_end,
                                                            ignore_voxels_with_zero_information);
    return std::move(all[0]);
  }

  /// Same, whole image (viewpoint_planner_raycast.cpp:89-101)
  template <typename ViewpointT>
  std::pair<VoxelSet, float> getRaycastHitVoxelsWithInformationScore(
      const ViewpointT& viewpoint, bool ignore_voxels_with_zero_information = false) const {
    return getRaycastHitVoxelsWithInformationScore(viewpoint, 0, viewpoint.camera().width(), 0,
                                                   viewpoint.camera().height(), ignore_voxels_with_zero_information);
  }

  typedef VoxelWithInformationSetView<NodeT> VoxelSetView;

  /// Batched form: all candidates of one graph-building round in one call (one camera for all viewpoints), the
  /// reference's own container per viewpoint.  The sets are filled by all host threads (OpenMP over viewpoints, like the
  /// reference's result conversion, bvh.h:430,487) out of per-set arenas, and large batches are cut into pieces of
  /// kSetBatchPiece viewpoints: a helper thread drives the engine through the pieces while this thread's team turns the
  /// finished ones into sets -- the 10^7 hash-table insertions of a 10^3-viewpoint batch overlap with the GPU work.
  /// Consumers that only read the sets take ...BatchView below (no host work per voxel at all).
  template <typename ViewpointT>
  std::vector<std::pair<VoxelSet, float>> getRaycastHitVoxelsWithInformationScoreBatch(
      const std::vector<const ViewpointT*>& viewpoints, std::size_t x_start, std::size_t x_end, std::size_t y_start,
      std::size_t y_end, bool ignore_voxels_with_zero_information = false) const {
    std::vector<std::pair<VoxelSet, float>> out(viewpoints.size());
    if (viewpoints.empty()) return out;
    const std::size_t n = viewpoints.size();
    if (n <= kSetBatchPiece) {
      const q3dr_result r = scoreBatch(viewpoints, x_start, x_end, y_start, y_end, ignore_voxels_with_zero_information);
      fillSets(r.offsets, r.total, r.leaf, r.information, 0, n, &out);
      return out;
    }
    // producer: the engine calls, piece after piece; every piece's host arrays are detached from the map at once
    struct Piece {
      std::size_t first, count;
      std::vector<std::uint64_t> offsets;
      std::vector<float> total;
      q3dr_host_block host;
    };
    std::mutex mutex;
    std::condition_variable ready;
    std::deque<std::shared_ptr<Piece> > queue;
    bool done = false;
    std::exception_ptr producer_error;
    std::thread producer([&]() {
      try {
        for (std::size_t first = 0; first < n; first += kSetBatchPiece) {
          const std::size_t count = std::min(kSetBatchPiece, n - first);
          const std::vector<const ViewpointT*> sub(viewpoints.begin() + static_cast<std::ptrdiff_t>(first),
                                                   viewpoints.begin() + static_cast<std::ptrdiff_t>(first + count));
          const q3dr_result r = scoreBatch(sub, x_start, x_end, y_start, y_end, ignore_voxels_with_zero_information);
          std::shared_ptr<Piece> p = std::make_shared<Piece>();
          p->first = first;
          p->count = count;
          p->offsets.assign(r.offsets, r.offsets + count + 1);
          p->total.assign(r.total, r.total + count);
          std::memset(&p->host, 0, sizeof(p->host));
          bvh_tree_->check(q3dr_result_detach_host(bvh_tree_->handle(), &p->host));
          {
            std::lock_guard<std::mutex> lock(mutex);
            queue.push_back(p);
          }
          ready.notify_one();
        }
      } catch (...) {
        producer_error = std::current_exception();
      }
      {
        std::lock_guard<std::mutex> lock(mutex);
        done = true;
      }
      ready.notify_one();
    });
    // consumer: this thread and its OpenMP team
    std::exception_ptr consumer_error;
    for (;;) {
      std::shared_ptr<Piece> p;
      {
        std::unique_lock<std::mutex> lock(mutex);
        ready.wait(lock, [&]() { return !queue.empty() || done; });
        if (queue.empty()) break;
        p = queue.front();
        queue.pop_front();
      }
      if (!consumer_error) {
        try {
          fillSets(p->offsets.data(), p->total.data(), p->host.leaf, p->host.information, p->first, p->count, &out);
        } catch (...) {
          consumer_error = std::current_exception();
        }
      }
      q3dr_host_block_release(&p->host);
    }
    producer.join();
    if (producer_error) std::rethrow_exception(producer_error);
    if (consumer_error) std::rethrow_exception(consumer_error);
    return out;
  }

  /// Batched form returning views into the engine's own host arrays: no per-voxel work on the host at all.
  template <typename ViewpointT>
  std::vector<std::pair<VoxelSetView, float>> getRaycastHitVoxelsWithInformationScoreBatchView(
      const std::vector<const ViewpointT*>& viewpoints, std::size_t x_start, std::size_t x_end, std::size_t y_start,
      std::size_t y_end, bool ignore_voxels_with_zero_information = false) const {
    std::vector<std::pair<VoxelSetView, float>> out(viewpoints.size());
    if (viewpoints.empty()) return out;
    const q3dr_result r = scoreBatch(viewpoints, x_start, x_end, y_start, y_end, ignore_voxels_with_zero_information);
    std::shared_ptr<VoxelSetBlock<NodeT> > block = std::make_shared<VoxelSetBlock<NodeT> >();
    // the map hands its two big host arrays over (it takes others for its next call): nothing is copied on the host
    bvh_tree_->check(q3dr_result_detach_host(bvh_tree_->handle(), &block->host));
    block->leaf = block->host.leaf;
    block->information = block->host.information;
    block->n_entries = static_cast<std::size_t>(r.n_entries);
    block->nodes = &bvh_tree_->leafNodes();
    block->leaf_of = &TreeType::leafIndexTableOf;
    block->tree = bvh_tree_;
    for (std::size_t v = 0; v < viewpoints.size(); ++v) {
      out[v].first = VoxelSetView(block, static_cast<std::size_t>(r.offsets[v]), static_cast<std::size_t>(r.offsets[v + 1]));
      out[v].second = r.total[v];
    }
    return out;
  }

 private:
  static const std::size_t kSetBatchPiece = 256;

  /// sets of viewpoints [first, first + count) of `out` from a CSR piece (offsets / total indexed from 0)
  void fillSets(const std::uint64_t* offsets, const float* total, const std::int32_t* leaf, const float* information,
                std::size_t first, std::size_t count, std::vector<std::pair<VoxelSet, float>>* out) const {
    const std::vector<NodeT*>& nodes = bvh_tree_->leafNodes();
    const std::ptrdiff_t nv = static_cast<std::ptrdiff_t>(count);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4)
#endif
    for (std::ptrdiff_t v = 0; v < nv; ++v) {
      VoxelSet& set = (*out)[first + static_cast<std::size_t>(v)].first;
      const std::size_t k = static_cast<std::size_t>(offsets[v + 1] - offsets[v]);
      // one arena chunk for the whole set: 32-byte hash nodes + the bucket array
      set.get_allocator().arena().reserve(k * (sizeof(VoxelWithInformation<NodeT>) + 2 * sizeof(void*)) + 2 * k * sizeof(void*) + 256);
      set.reserve(k);
      for (std::uint64_t i = offsets[v]; i < offsets[v + 1]; ++i) {
        set.insert(VoxelWithInformation<NodeT>(nodes[static_cast<std::size_t>(leaf[i])], information[i]));
      }
      (*out)[first + static_cast<std::size_t>(v)].second = total[v];
    }
  }

  /// the engine call behind both batch forms; only (leaf, information) travel back (8 bytes per voxel)
  template <typename ViewpointT>
  q3dr_result scoreBatch(const std::vector<const ViewpointT*>& viewpoints, std::size_t x_start, std::size_t x_end,
                         std::size_t y_start, std::size_t y_end, bool ignore_voxels_with_zero_information) const {
    bvh_tree_->requireMap();
    float K[4];
    std::vector<float> Rt(12 * viewpoints.size());
    for (std::size_t v = 0; v < viewpoints.size(); ++v) {
      TreeType::packCamera(viewpoints[v]->camera().intrinsics(), viewpoints[v]->pose().getTransformationImageToWorld(),
                           K, Rt.data() + 12 * v);
    }
    q3dr_score_opts o = opts_;
    o.ignore_voxels_with_zero_information = ignore_voxels_with_zero_information ? 1 : 0;
    q3dr_result r;
    bvh_tree_->check(q3dr_map_set_result_fields(bvh_tree_->handle(), 0u));
    int st;
    if (mesh_normals_ != nullptr) {
      st = q3dr_score_viewpoints_mesh_normals(
          bvh_tree_->handle(), mesh_normals_->handle(), &mesh_normals_->renderOptions(), K,
          static_cast<std::uint32_t>(viewpoints[0]->camera().width()),
          static_cast<std::uint32_t>(viewpoints[0]->camera().height()), static_cast<std::uint32_t>(x_start),
          static_cast<std::uint32_t>(x_end), static_cast<std::uint32_t>(y_start), static_cast<std::uint32_t>(y_end),
          Rt.data(), viewpoints.size(), &o, &r);
    } else {
      st = q3dr_score_viewpoints(bvh_tree_->handle(), K, static_cast<std::uint32_t>(viewpoints[0]->camera().width()),
                                 static_cast<std::uint32_t>(x_start), static_cast<std::uint32_t>(x_end),
                                 static_cast<std::uint32_t>(y_start), static_cast<std::uint32_t>(y_end), Rt.data(),
                                 viewpoints.size(), &o, &r);
    }
    q3dr_map_set_result_fields(bvh_tree_->handle(), Q3DR_RESULT_FIRST_PIXEL | Q3DR_RESULT_DIST_SQ);
    bvh_tree_->check(st);
    return r;
  }

  TreeType* bvh_tree_;
  float min_range_;
  float max_range_;
  q3dr_score_opts opts_;
  const PoissonMeshNormalsCuda* mesh_normals_;
};

/// One ViewpointPath::observed_voxel_map on the device plus the information sums the path search evaluates against
/// it (viewpoint_planner_scoring.cpp:124-147, viewpoint_planner_path.cpp:203-233,838-848).  Viewpoint indices refer to
/// the last (batched) scoring call on the tree.
template <typename NodeT>
class ObservedVoxelMapCuda {
 public:
  typedef OccupiedTreeCuda<NodeT> TreeType;
  explicit ObservedVoxelMapCuda(TreeType* tree, float viewpoint_information_factor = 1.0f / 3.0f)
      : tree_(tree), factor_(viewpoint_information_factor), obs_(nullptr) {
    check(q3dr_observed_create(tree_->handle(), &obs_));
  }
  ObservedVoxelMapCuda(const ObservedVoxelMapCuda&) = delete;
  ObservedVoxelMapCuda& operator=(const ObservedVoxelMapCuda&) = delete;
  ~ObservedVoxelMapCuda() { q3dr_observed_destroy(obs_); }

  void clear() { check(q3dr_observed_clear(obs_)); }
  /// computeNewInformation for the given viewpoints of the last result
  std::vector<float> computeNewInformation(const std::vector<std::uint32_t>& viewpoints) const {
    std::vector<float> out(viewpoints.size());
    check(q3dr_new_information(tree_->handle(), obs_, factor_, viewpoints.data(), viewpoints.size(), out.data()));
    return out;
  }
  /// addObservedVoxelsToViewpointPath; returns the novel information
  float addObservedVoxels(std::uint32_t viewpoint) {
    float v = 0;
    check(q3dr_observed_add_viewpoint(tree_->handle(), obs_, factor_, viewpoint, &v));
    return v;
  }
  /// computeNewInformation / addObservedVoxelsToViewpointPath for a stored ViewpointEntry::voxel_set
  float computeNewInformation(const VoxelWithInformationSet<NodeT>& voxel_set) const { return voxelSet(voxel_set, 0); }
  float addObservedVoxels(const VoxelWithInformationSet<NodeT>& voxel_set) { return voxelSet(voxel_set, 1); }
  /// ... and for a view (its leaf / information arrays go to the device as they are)
  float computeNewInformation(const VoxelWithInformationSetView<NodeT>& voxel_set) const { return voxelView(voxel_set, 0); }
  float addObservedVoxels(const VoxelWithInformationSetView<NodeT>& voxel_set) { return voxelView(voxel_set, 1); }

  /// compute_overlap_information_lambda(set1 = others[i], set2 = ref)
  std::vector<float> computeOverlapInformation(std::uint32_t ref, const std::vector<std::uint32_t>& others) const {
    std::vector<float> out(others.size());
    check(q3dr_overlap_information(tree_->handle(), ref, others.data(), others.size(), out.data()));
    return out;
  }

 private:
  float voxelSet(const VoxelWithInformationSet<NodeT>& voxel_set, int add) const {
    std::vector<std::int32_t> leaf;
    std::vector<float> info;
    leaf.reserve(voxel_set.size());
    info.reserve(voxel_set.size());
    for (typename VoxelWithInformationSet<NodeT>::const_iterator it = voxel_set.begin(); it != voxel_set.end(); ++it) {
      leaf.push_back(tree_->leafIndex(it->voxel));
      info.push_back(it->information);
    }
    float v = 0;
    check(q3dr_observed_voxel_set(tree_->handle(), obs_, factor_, leaf.data(), info.data(), leaf.size(), add, &v));
    return v;
  }
  float voxelView(const VoxelWithInformationSetView<NodeT>& voxel_set, int add) const {
    float v = 0;
    check(q3dr_observed_voxel_set(tree_->handle(), obs_, factor_, voxel_set.leafData(), voxel_set.informationData(),
                                  voxel_set.size(), add, &v));
    return v;
  }
  void check(int status) const {
    if (status != Q3DR_OK) throw typename TreeType::CudaError(status, q3dr_last_error());
  }
  TreeType* tree_;
  float factor_;
  q3dr_observed* obs_;
};

}  // namespace q3dr

#endif  // Q3DR_OCCUPIED_TREE_CUDA_HPP_
