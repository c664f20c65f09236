// octree_io.cpp -- dependency-free reader of OctoMap's ".ot" octree stream and the depth-first leaf walk that feeds
// the leaf rule of generateBVHTree ("lets real Octomap files feed the engine").
//
// The reference reads its occupancy map with octomap::AbstractOcTree::read (viewpoint_planner/src/octree/
// occupancy_map.hxx:1123-1127) and walks it with octree->begin_tree() (viewpoint_planner/src/planner/
// viewpoint_planner_data.cpp:820-863).  OctoMap itself is an un-vendored dependency of the reference (found through
// CMake, viewpoint_planner/CMakeLists.txt:9, no version pinned; the node classes follow the 1.7 / 1.8 API:
// AbstractOcTreeNode with a children array, file recursion inside OcTreeBaseImpl).  What is restated here is OctoMap's
// published file format and iterator arithmetic:
//   header   "# Octomap OcTree file", comment lines, "id <type>", "size <nodes>", "res <metres>", "data\n"
//   data     pre-order: node payload, one byte with a bit per existing child, then the existing children 0 .. 7
//            (OcTreeBaseImpl::readNodesRecurs / writeNodesRecurs)
//   payload  Ait_OccupancyMap with OccupancyNode:           float occupancy, uint32 observation_count       (8 bytes)
//            Ait_OccupancyMap with AugmentedOccupancyNode:  ... + float weight + uint64 observation_count_sum (20 bytes)
//            (viewpoint_planner/src/octree/occupancy_node.cpp:139-211; both trees carry the same id, occupancy_map.h:264)
//            OcTree: float log-odds (4 bytes); occupancy = 1 / (1 + exp(-l)), every stored node counts as observed once
//   order    begin_tree() is a pre-order walk that pushes the children 7 .. 0, i.e. visits them 0 .. 7: file order
//   geometry tree depth 16, root key 32768 per axis; child i adds (bit set) or subtracts the centre offset
//            32768 >> depth on axis x (bit 0), y (bit 1), z (bit 2) (computeChildKey); centre = keyToCoord(key, depth)
//            in double, then float (point3d); size = resolution * 2^(16 - depth)
// Parity of this file against OctoMap's own reader is UNPINNED (the library is not available here); the format is
// checked against an independent Python restatement and a synthetic writer (oracle/octree_ot.py, tests/test_octree_io.py).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <new>
#include <sstream>
#include <string>
#include <vector>

#include "q3dr_internal.cuh"

struct q3dr_octree {
  q3dr_octree_info info{};
  std::vector<float> center;  // n x 3
  std::vector<float> size, occupancy, weight;
  std::vector<uint32_t> count, depth;
};

namespace {

constexpr uint32_t kTreeDepth = 16;
constexpr uint32_t kTreeMaxVal = 32768;

struct Parser {
  const unsigned char* p;
  const unsigned char* end;
  uint32_t payload;  // Q3DR_OCTREE_*
  double res;
  q3dr_octree* out;
  uint64_t nodes = 0;
  bool truncated = false;

  static size_t payload_bytes(uint32_t kind) {
    return kind == Q3DR_OCTREE_OCCUPANCY ? 8 : (kind == Q3DR_OCTREE_AUGMENTED ? 20 : 4);
  }

  // keyToCoord(key, depth) of OcTreeBaseImpl, double arithmetic as OctoMap does it
  double coord(uint32_t key, uint32_t depth) const {
    if (depth == 0) return 0.0;
    if (depth == kTreeDepth) return ((double)((int)key - (int)kTreeMaxVal) + 0.5) * res;
    const double node_size = res * (double)(1u << (kTreeDepth - depth));
    return (std::floor(((double)key - (double)kTreeMaxVal) / (double)(1u << (kTreeDepth - depth))) + 0.5) * node_size;
  }

  // iterative pre-order walk with an explicit stack (depth <= 16, at most 8 pending siblings per level)
  bool walk(bool store) {
    struct Item {
      uint16_t key[3];
      uint8_t depth;
    };
    std::vector<Item> stack;
    stack.push_back(Item{{(uint16_t)kTreeMaxVal, (uint16_t)kTreeMaxVal, (uint16_t)kTreeMaxVal}, 0});
    const size_t pb = payload_bytes(payload);
    while (!stack.empty()) {
      const Item it = stack.back();
      stack.pop_back();
      if ((size_t)(end - p) < pb + 1) {
        truncated = true;
        return false;
      }
      float occ = 0.f, w = 0.f;
      uint32_t cnt = 0;
      if (payload == Q3DR_OCTREE_LOGODDS) {
        float l;
        std::memcpy(&l, p, 4);
        occ = (float)(1.0 - 1.0 / (1.0 + std::exp((double)l)));  // octomap::probability
        cnt = 1;
      } else {
        std::memcpy(&occ, p, 4);
        std::memcpy(&cnt, p + 4, 4);
        if (payload == Q3DR_OCTREE_AUGMENTED) std::memcpy(&w, p + 8, 4);
      }
      p += pb;
      const unsigned children = *p++;
      ++nodes;
      if (children == 0) {
        if (store) {
          for (int a = 0; a < 3; ++a) out->center.push_back((float)coord(it.key[a], it.depth));
          out->size.push_back((float)(res * (double)(1u << (kTreeDepth - it.depth))));
          out->occupancy.push_back(occ);
          out->count.push_back(cnt);
          out->weight.push_back(w);
          out->depth.push_back(it.depth);
        }
        continue;
      }
      if (it.depth >= kTreeDepth) return false;  // children below the finest level: not an octree stream
      const uint32_t d = it.depth + 1u;
      const uint32_t off = kTreeMaxVal >> d;
      for (int i = 7; i >= 0; --i) {  // pushed 7 .. 0: child 0 is read next, like the stream stores them
        if (!(children & (1u << i))) continue;
        Item c;
        c.depth = (uint8_t)d;
        for (int a = 0; a < 3; ++a) {
          const uint32_t k = it.key[a];
          c.key[a] = (uint16_t)((i & (1 << a)) ? k + off : k - off - (off ? 0u : 1u));
        }
        stack.push_back(c);
      }
    }
    return true;
  }
};

}  // namespace

extern "C" {

int q3dr_octree_read(const char* path, uint32_t payload, q3dr_octree** out) {
  if (!out) return fail(Q3DR_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  if (!path) return fail(Q3DR_ERR_INVALID_ARG, "path is NULL");
  if (payload > Q3DR_OCTREE_LOGODDS) return fail(Q3DR_ERR_INVALID_ARG, "unknown payload kind");
  std::ifstream f(path, std::ios::binary);
  if (!f) return fail(Q3DR_ERR_INVALID_ARG, std::string("cannot open ") + path);
  std::string line;
  if (!std::getline(f, line) || line.compare(0, 21, "# Octomap OcTree file") != 0) {
    return fail(Q3DR_ERR_INVALID_ARG, "not an OctoMap .ot stream (first line must be '# Octomap OcTree file')");
  }
  std::string id;
  uint64_t size = 0;
  double res = 0.0;
  bool have_data = false;
  while (std::getline(f, line)) {
    if (line.empty() || line[0] == '#') continue;
    std::istringstream ls(line);
    std::string token;
    ls >> token;
    if (token == "data") {
      have_data = true;
      break;
    } else if (token == "id") {
      ls >> id;
    } else if (token == "size") {
      ls >> size;
    } else if (token == "res") {
      ls >> res;
    }  // unknown header tokens are skipped, like AbstractOcTree::readHeader does
  }
  if (!have_data || id.empty() || !(res > 0.0)) return fail(Q3DR_ERR_INVALID_ARG, "incomplete .ot header (id / res / data)");
  std::vector<unsigned char> blob((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  q3dr_octree* t = new (std::nothrow) q3dr_octree;
  if (!t) return fail(Q3DR_ERR_OUT_OF_MEMORY, "host allocation failed");
  // payload kinds to try: the caller's, or by id (the reference's two node types share one id: the one whose node
  // count and stream length agree with the header wins)
  std::vector<uint32_t> kinds;
  if (payload != Q3DR_OCTREE_AUTO) {
    kinds.push_back(payload);
  } else if (id == "OcTree") {
    kinds.push_back(Q3DR_OCTREE_LOGODDS);
  } else {
    kinds.push_back(Q3DR_OCTREE_AUGMENTED);
    kinds.push_back(Q3DR_OCTREE_OCCUPANCY);
  }
  bool ok = false;
  try {
    for (uint32_t kind : kinds) {
      Parser ps{blob.data(), blob.data() + blob.size(), kind, res, t};
      if (size == 0) {  // empty tree: no root in the stream
        ok = blob.empty();
        t->info.payload = kind;
        break;
      }
      if (ps.walk(false) && ps.nodes == size && ps.p == ps.end) {
        Parser fill{blob.data(), blob.data() + blob.size(), kind, res, t};
        fill.walk(true);
        t->info.payload = kind;
        ok = true;
        break;
      }
    }
  } catch (const std::bad_alloc&) {
    delete t;
    return fail(Q3DR_ERR_OUT_OF_MEMORY, "host allocation failed");
  }
  if (!ok) {
    delete t;
    return fail(Q3DR_ERR_INVALID_ARG, "octree stream does not match its header (node count / length / payload size)");
  }
  std::memset(t->info.id, 0, sizeof(t->info.id));
  std::strncpy(t->info.id, id.c_str(), sizeof(t->info.id) - 1);
  t->info.resolution = res;
  t->info.n_nodes = size;
  t->info.n_leaves = t->size.size();
  t->info.tree_depth = kTreeDepth;
  *out = t;
  return Q3DR_OK;
}

int q3dr_octree_get_info(const q3dr_octree* t, q3dr_octree_info* info) {
  if (!t || !info) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  *info = t->info;
  return Q3DR_OK;
}

int q3dr_octree_leaves(const q3dr_octree* t, float* center, float* size, float* occupancy, uint32_t* observation_count,
                       float* weight, uint32_t* depth) {
  if (!t) return fail(Q3DR_ERR_INVALID_ARG, "octree is NULL");
  const size_t n = t->size.size();
  if (center) std::memcpy(center, t->center.data(), 3 * n * sizeof(float));
  if (size) std::memcpy(size, t->size.data(), n * sizeof(float));
  if (occupancy) std::memcpy(occupancy, t->occupancy.data(), n * sizeof(float));
  if (observation_count) std::memcpy(observation_count, t->count.data(), n * sizeof(uint32_t));
  if (weight) std::memcpy(weight, t->weight.data(), n * sizeof(float));
  if (depth) std::memcpy(depth, t->depth.data(), n * sizeof(uint32_t));
  return Q3DR_OK;
}

int q3dr_octree_destroy(q3dr_octree* t) {
  delete t;
  return Q3DR_OK;
}

}  // extern "C"
