// bvh_build.h -- host-side flattening of the occupancy leaves into the packed 4-wide node array the
// sm_100a traversal kernel walks, and the reference's splitMedian leaf order (the tie-break order).
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

namespace q3dr {

// 128-byte, 128-byte-aligned node: child boxes in SoA so that every field is one 16-byte fetch.
// child[c]:  internal -> node index;  leaf -> kLeafFlag | leaf index;  unused slot -> kEmptyChild
// (unused slots carry a degenerate box that the slab test can never accept).
struct alignas(128) Node4 {
  float mn[3][4];     // mn[axis][child]
  float mx[3][4];     // mx[axis][child]
  uint32_t child[4];
  uint32_t meta[4];   // meta[0] = number of used slots, meta[1] = leaves below, meta[2] = depth, meta[3] = order axis
};
static_assert(sizeof(Node4) == 128, "Node4 must be one 128-byte line");

// child word: bit 31 = leaf flag; bits 0..26 = node index (inner) or leaf index (leaf).  child[0] additionally
// carries, in bits 27..30, the LEAF MASK of the whole node (bit 27 + c set: slot c holds a leaf or is unused), so the
// traversal reads the node's shape with one shift.  Unused slots are kEmptyChild (a leaf flag with an all-ones
// index) and carry a degenerate box that no ray can hit.
constexpr uint32_t kLeafFlag = 0x80000000u;
constexpr uint32_t kPayloadBits = 27;
constexpr uint32_t kPayloadMask = (1u << kPayloadBits) - 1u;
constexpr uint32_t kLeafMaskShift = kPayloadBits;
constexpr uint32_t kOrderAxisShift = kPayloadBits;  // child[1], bits 27..28: the axis the four children are sorted along (ascending box centres)
constexpr uint32_t kEmptyChild = kLeafFlag | kPayloadMask;
constexpr uint32_t kNodeIndexMask = kPayloadMask;

struct FlatTree {
  std::vector<Node4> nodes;  // nodes[0] is the root; the first `top_nodes` are the top levels in BFS order
  uint32_t top_nodes = 0;
  uint32_t depth = 0;        // depth of the 4-wide tree (root = 0, deepest internal node)
  uint32_t max_stack = 0;    // upper bound of traversal-stack entries a packet can hold
};

// Builds the 4-wide BVH over n leaf boxes (n x 6 floats).  Leaf k is referenced as kLeafFlag | k.
// top_levels: how many levels from the root are laid out breadth-first at the front of the array.
void build_flat_tree(const float* boxes, size_t n, int top_levels, FlatTree* out);

// Reference Tree::build leaf order (bvh.h:764-788): order[k] = input index of the k-th leaf.
// depth (nullable): depth[k] = depth of that leaf in the reference tree.
void splitmedian_order(const float* boxes, size_t n, uint32_t* order, uint32_t* depth);

// Position of every leaf in the reference's all-node iterator (bvh.h:226-235: stack DFS pushing left then right,
// i.e. root, right subtree, left subtree) -- the voxel index its serialisation stores
// (viewpoint_planner_serialization.h:88-105).  Depends on n only: splitMedian halves ranges exactly.
void splitmedian_voxel_indices(size_t n, uint32_t* out);

// Depth of the k-th leaf (left to right) of a splitMedian tree over n leaves: ranges halve exactly.
uint32_t splitmedian_leaf_depth(size_t n, size_t k);

}  // namespace q3dr
