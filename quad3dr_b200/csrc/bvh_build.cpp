// bvh_build.cpp -- host-side map flattening (no CUDA here).
//
//  * splitmedian_order(): the leaf order of the reference's bvh::Tree::build
//    (viewpoint_planner/src/bvh/bvh.h:361-373,764-788 of the reference): recursive median split,
//    stable sort by box centre on axes x -> y -> z -> x ...  The engine needs it only as the
//    tie-break order between leaves at equal distance (bvh.h:867,881-895); it is computed with
//    radix sorts and OpenMP tasks instead of N log^2 N comparator sorts.
//  * build_flat_tree(): the engine's own acceleration structure.  The closest-hit result depends
//    only on the set of leaf boxes, so the tree is free to differ from the
//    reference's: binned-SAH binary build, collapsed into 128-byte 4-wide nodes, top levels laid
//    out breadth-first (they are staged in shared memory), the rest depth-first.
#include "bvh_build.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace q3dr {

namespace {

// ------------------------------------------------------------------------------------------
// splitMedian order
// ------------------------------------------------------------------------------------------

inline uint32_t float_sort_key(float f) {
  f = f + 0.0f;  // -0 -> +0: the reference's comparator (a < b) treats them as equal
  uint32_t u;
  std::memcpy(&u, &f, 4);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct OrderCtx {
  const float* centre[3];  // centre[axis][leaf]
  uint32_t* idx;           // permuted in place
  uint32_t* depth;         // nullable, by position
};

// Stable sort of idx[b,e) by centre[axis].  LSD radix (3 x 11 bits) for large ranges.
void stable_sort_range(const OrderCtx& c, size_t b, size_t e, int axis, std::vector<uint32_t>& tmp_key,
                       std::vector<uint32_t>& tmp_idx) {
  const size_t n = e - b;
  const float* key = c.centre[axis];
  if (n < 4096) {
    std::stable_sort(c.idx + b, c.idx + e, [key](uint32_t x, uint32_t y) { return key[x] < key[y]; });
    return;
  }
  tmp_key.resize(2 * n);
  tmp_idx.resize(2 * n);
  uint32_t* k0 = tmp_key.data();
  uint32_t* k1 = k0 + n;
  uint32_t* i0 = tmp_idx.data();
  uint32_t* i1 = i0 + n;
  for (size_t i = 0; i < n; ++i) {
    i0[i] = c.idx[b + i];
    k0[i] = float_sort_key(key[i0[i]]);
  }
  constexpr int kBits = 11;
  constexpr uint32_t kBuckets = 1u << kBits;
  for (int pass = 0; pass < 3; ++pass) {
    const int shift = pass * kBits;
    size_t hist[kBuckets + 1];
    std::memset(hist, 0, sizeof(hist));
    for (size_t i = 0; i < n; ++i) ++hist[((k0[i] >> shift) & (kBuckets - 1)) + 1];
    if (hist[((k0[0] >> shift) & (kBuckets - 1)) + 1] == n) continue;  // all in one bucket
    for (uint32_t j = 0; j < kBuckets; ++j) hist[j + 1] += hist[j];
    for (size_t i = 0; i < n; ++i) {
      const size_t d = hist[(k0[i] >> shift) & (kBuckets - 1)]++;
      k1[d] = k0[i];
      i1[d] = i0[i];
    }
    std::swap(k0, k1);
    std::swap(i0, i1);
  }
  std::memcpy(c.idx + b, i0, n * sizeof(uint32_t));
}

void order_recursive(const OrderCtx& c, size_t b, size_t e, int axis, uint32_t cur_depth) {
  if (e - b <= 1) {
    if (c.depth && e > b) c.depth[b] = cur_depth;
    return;
  }
  {
    std::vector<uint32_t> tk, ti;
    stable_sort_range(c, b, e, axis, tk, ti);
  }
  const size_t m = b + (e - b) / 2;
  const int next = (axis + 1) % 3;
  if (e - b > 32768) {
#pragma omp task default(shared)
    order_recursive(c, b, m, next, cur_depth + 1);
#pragma omp task default(shared)
    order_recursive(c, m, e, next, cur_depth + 1);
#pragma omp taskwait
  } else {
    order_recursive(c, b, m, next, cur_depth + 1);
    order_recursive(c, m, e, next, cur_depth + 1);
  }
}

// ------------------------------------------------------------------------------------------
// binned-SAH binary build
// ------------------------------------------------------------------------------------------

struct Box3 {
  float mn[3];
  float mx[3];
  void reset() {
    mn[0] = mn[1] = mn[2] = FLT_MAX;
    mx[0] = mx[1] = mx[2] = -FLT_MAX;
  }
  void grow(const float* b6) {
    for (int a = 0; a < 3; ++a) {
      mn[a] = std::min(mn[a], b6[a]);
      mx[a] = std::max(mx[a], b6[3 + a]);
    }
  }
  void grow(const Box3& o) {
    for (int a = 0; a < 3; ++a) {
      mn[a] = std::min(mn[a], o.mn[a]);
      mx[a] = std::max(mx[a], o.mx[a]);
    }
  }
  double half_area() const {
    const double dx = (double)mx[0] - mn[0], dy = (double)mx[1] - mn[1], dz = (double)mx[2] - mn[2];
    if (dx < 0 || dy < 0 || dz < 0) return 0.0;
    return dx * dy + dy * dz + dz * dx;
  }
};

// Binary build node.  Terminal nodes (left < 0) are GROUPS of 1..4 leaves idx[first, first + count): the
// bottom of the tree is built in fours so that the 4-wide nodes over the leaves are full.
struct BinNode {
  Box3 box;
  int64_t left = -1, right = -1;
  uint32_t first = 0;
  uint32_t count = 1;  // leaves below
};

struct BuildCtx {
  const float* boxes;     // n x 6
  std::vector<float> cx;  // centroids (they only steer the split; the boxes themselves stay exact)
  std::vector<float> cy;
  std::vector<float> cz;
  std::vector<uint32_t> idx;
  std::vector<BinNode> nodes;  // 2n - 1 slots
  const float* cen(int a) const { return a == 0 ? cx.data() : (a == 1 ? cy.data() : cz.data()); }
};

constexpr int kMaxBins = 64;
constexpr size_t kSmallRange = 64;

// build-quality knobs (experiments only; defaults are what ships)
inline int env_int(const char* name, int dflt) {
  const char* e = std::getenv(name);
  return e ? atoi(e) : dflt;
}
const int kBins = std::max(4, std::min(kMaxBins, env_int("Q3DR_BINS", 16)));
const int kCollapse = env_int("Q3DR_COLLAPSE", 0);  // 0: largest area first, 1: most leaves first, 2: area x leaves

// Exact SAH sweep for a small range with the cut restricted to multiples of 4 leaves.
size_t split_small(BuildCtx& c, size_t b, size_t e) {
  const size_t n = e - b;
  uint32_t sorted[3][kSmallRange];
  double best_cost = DBL_MAX;
  int best_axis = 0;
  size_t best_k = 4;
  for (int a = 0; a < 3; ++a) {
    const float* cen = c.cen(a);
    uint32_t* s = sorted[a];
    for (size_t i = 0; i < n; ++i) s[i] = c.idx[b + i];
    std::stable_sort(s, s + n, [cen](uint32_t x, uint32_t y) { return cen[x] < cen[y]; });
    double right_area[kSmallRange];
    Box3 acc;
    acc.reset();
    for (size_t i = n; i-- > 1;) {
      acc.grow(c.boxes + 6 * (size_t)s[i]);
      right_area[i] = acc.half_area();
    }
    acc.reset();
    for (size_t k = 1; k < n; ++k) {
      acc.grow(c.boxes + 6 * (size_t)s[k - 1]);
      if (k % 4 != 0) continue;
      const double cost = acc.half_area() * (double)k + right_area[k] * (double)(n - k);
      if (cost < best_cost) {
        best_cost = cost;
        best_axis = a;
        best_k = k;
      }
    }
  }
  for (size_t i = 0; i < n; ++i) c.idx[b + i] = sorted[best_axis][i];
  return b + best_k;
}

// Builds the subtree over idx[b,e) into slots [base, base + 2(e-b) - 1); returns the root slot.
int64_t build_binary(BuildCtx& c, size_t b, size_t e, int64_t base) {
  const size_t n = e - b;
  const int64_t me = base + 2 * (int64_t)n - 2;
  BinNode& node = c.nodes[me];
  if (n <= 4) {
    node.box.reset();
    for (size_t i = b; i < e; ++i) node.box.grow(c.boxes + 6 * (size_t)c.idx[i]);
    node.left = node.right = -1;
    node.first = (uint32_t)b;
    node.count = (uint32_t)n;
    return me;
  }
  size_t mid;
  if (n <= kSmallRange) {
    mid = split_small(c, b, e);
  } else {
    // centroid bounds
    float cmn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, cmx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (size_t i = b; i < e; ++i) {
      const uint32_t li = c.idx[i];
      const float p[3] = {c.cx[li], c.cy[li], c.cz[li]};
      for (int a = 0; a < 3; ++a) {
        cmn[a] = std::min(cmn[a], p[a]);
        cmx[a] = std::max(cmx[a], p[a]);
      }
    }
    mid = b + n / 2;
    int best_axis = -1;
    int best_split = -1;
    double best_cost = DBL_MAX;
    float best_scale = 0;
    for (int a = 0; a < 3; ++a) {
      const float ext = cmx[a] - cmn[a];
      if (!(ext > 0)) continue;
      const float scale = (float)kBins / ext;
      Box3 bin_box[kMaxBins];
      uint32_t bin_cnt[kMaxBins];
      for (int k = 0; k < kBins; ++k) {
        bin_box[k].reset();
        bin_cnt[k] = 0;
      }
      const float* cen = c.cen(a);
      for (size_t i = b; i < e; ++i) {
        const uint32_t li = c.idx[i];
        int k = (int)((cen[li] - cmn[a]) * scale);
        k = k < 0 ? 0 : (k >= kBins ? kBins - 1 : k);
        bin_box[k].grow(c.boxes + 6 * (size_t)li);
        ++bin_cnt[k];
      }
      double right_area[kMaxBins];
      uint32_t right_cnt[kMaxBins];
      Box3 acc;
      acc.reset();
      uint32_t cnt = 0;
      for (int k = kBins - 1; k > 0; --k) {
        acc.grow(bin_box[k]);
        cnt += bin_cnt[k];
        right_area[k] = acc.half_area();
        right_cnt[k] = cnt;
      }
      acc.reset();
      cnt = 0;
      for (int k = 0; k < kBins - 1; ++k) {
        acc.grow(bin_box[k]);
        cnt += bin_cnt[k];
        if (cnt == 0 || right_cnt[k + 1] == 0) continue;
        const double cost = acc.half_area() * cnt + right_area[k + 1] * right_cnt[k + 1];
        if (cost < best_cost) {
          best_cost = cost;
          best_axis = a;
          best_split = k;
          best_scale = scale;
        }
      }
    }
    if (best_axis >= 0) {
      const float* cen = c.cen(best_axis);
      const float lo = cmn[best_axis];
      uint32_t* first = c.idx.data() + b;
      uint32_t* last = c.idx.data() + e;
      uint32_t* p = std::partition(first, last, [&](uint32_t li) {
        int k = (int)((cen[li] - lo) * best_scale);
        k = k < 0 ? 0 : (k >= kBins ? kBins - 1 : k);
        return k <= best_split;
      });
      mid = b + (size_t)(p - first);
      if (mid == b || mid == e) mid = b + n / 2;  // cannot happen; keep the build total anyway
    } else {
      // all centroids identical: split by position, on a multiple of four
      mid = b + ((n / 2 + 3) / 4) * 4;
      if (mid >= e) mid = b + n / 2;
    }
  }
  const int64_t lbase = base;
  const int64_t rbase = base + 2 * (int64_t)(mid - b) - 1;
  int64_t l, r;
  if (n > 65536) {
#pragma omp task default(shared)
    l = build_binary(c, b, mid, lbase);
#pragma omp task default(shared)
    r = build_binary(c, mid, e, rbase);
#pragma omp taskwait
  } else {
    l = build_binary(c, b, mid, lbase);
    r = build_binary(c, mid, e, rbase);
  }
  BinNode& nd = c.nodes[me];
  nd.left = l;
  nd.right = r;
  nd.box = c.nodes[l].box;
  nd.box.grow(c.nodes[r].box);
  nd.count = c.nodes[l].count + c.nodes[r].count;
  return me;
}

// ------------------------------------------------------------------------------------------
// collapse to 4-wide nodes + layout
// ------------------------------------------------------------------------------------------

struct Child {
  int64_t bin;   // binary node that becomes a 4-wide node of its own, or -1
  int64_t leaf;  // leaf index referenced directly, or -1
};

// Up to four children of the 4-wide node made from binary node `node`.
int gather_children(const BuildCtx& c, int64_t node, Child out[4]) {
  const std::vector<BinNode>& bn = c.nodes;
  int cnt = 0;
  if (bn[node].left < 0) {  // a group: its leaves
    for (uint32_t i = 0; i < bn[node].count; ++i) out[cnt++] = Child{-1, (int64_t)c.idx[bn[node].first + i]};
    return cnt;
  }
  out[cnt++] = Child{bn[node].left, -1};
  out[cnt++] = Child{bn[node].right, -1};
  // pull grandchildren up, largest box first
  while (cnt < 4) {
    int pick = -1;
    double best = -1.0;
    for (int i = 0; i < cnt; ++i) {
      if (out[i].bin < 0 || bn[out[i].bin].left < 0) continue;
      double a = bn[out[i].bin].box.half_area();
      if (kCollapse == 1) a = (double)bn[out[i].bin].count;
      if (kCollapse == 2) a *= (double)bn[out[i].bin].count;
      if (a > best) {
        best = a;
        pick = i;
      }
    }
    if (pick < 0) break;
    const int64_t p = out[pick].bin;
    out[pick] = Child{bn[p].left, -1};
    out[cnt++] = Child{bn[p].right, -1};
  }
  // inline groups that fit into the free slots (smallest first): saves a node visit
  while (true) {
    int pick = -1;
    uint32_t best = 5;
    for (int i = 0; i < cnt; ++i) {
      if (out[i].bin < 0 || bn[out[i].bin].left >= 0) continue;
      const uint32_t k = bn[out[i].bin].count;
      if (k <= (uint32_t)(4 - cnt) + 1u && k < best) {
        best = k;
        pick = i;
      }
    }
    if (pick < 0) break;
    const BinNode& g = bn[out[pick].bin];
    out[pick] = Child{-1, (int64_t)c.idx[g.first]};
    for (uint32_t i = 1; i < g.count; ++i) out[cnt++] = Child{-1, (int64_t)c.idx[g.first + i]};
  }
  return cnt;
}

void child_box(const BuildCtx& c, const Child& ch, Box3* b) {
  if (ch.bin >= 0) {
    *b = c.nodes[ch.bin].box;
  } else {
    b->reset();
    b->grow(c.boxes + 6 * (size_t)ch.leaf);
  }
}

// Sorts the gathered children by box centre along the axis on which the centres spread most; returns it.
int order_children(const BuildCtx& c, Child ch[4], int cnt) {
  float ctr[4][3];
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int s = 0; s < cnt; ++s) {
    Box3 b;
    child_box(c, ch[s], &b);
    for (int a = 0; a < 3; ++a) {
      ctr[s][a] = 0.5f * b.mn[a] + 0.5f * b.mx[a];
      lo[a] = std::min(lo[a], ctr[s][a]);
      hi[a] = std::max(hi[a], ctr[s][a]);
    }
  }
  int axis = 0;
  for (int a = 1; a < 3; ++a) {
    if (hi[a] - lo[a] > hi[axis] - lo[axis]) axis = a;
  }
  for (int i = 1; i < cnt; ++i) {  // insertion sort, stable
    const Child n = ch[i];
    const float keys[3] = {ctr[i][0], ctr[i][1], ctr[i][2]};
    int j = i - 1;
    while (j >= 0 && ctr[j][axis] > keys[axis]) {
      ch[j + 1] = ch[j];
      for (int a = 0; a < 3; ++a) ctr[j + 1][a] = ctr[j][a];
      --j;
    }
    ch[j + 1] = n;
    for (int a = 0; a < 3; ++a) ctr[j + 1][a] = keys[a];
  }
  return axis;
}

void fill_node(const BuildCtx& c, const Child ch[4], int cnt, uint32_t depth, uint32_t leaves, uint32_t axis,
               Node4* n4) {
  for (int s = 0; s < 4; ++s) {
    if (s < cnt) {
      Box3 b;
      child_box(c, ch[s], &b);
      for (int a = 0; a < 3; ++a) {
        n4->mn[a][s] = b.mn[a];
        n4->mx[a][s] = b.mx[a];
      }
      n4->child[s] = (ch[s].bin < 0) ? (kLeafFlag | (uint32_t)ch[s].leaf) : 0u;  // inner refs are patched later
    } else {
      // a point far away: tmax > max(tmin, 0) can never hold for a degenerate box seen from outside
      for (int a = 0; a < 3; ++a) {
        n4->mn[a][s] = FLT_MAX;
        n4->mx[a][s] = FLT_MAX;
      }
      n4->child[s] = kEmptyChild;
    }
  }
  uint32_t leaf_mask = 0;
  for (int s = 0; s < 4; ++s) {
    if (n4->child[s] & kLeafFlag) leaf_mask |= 1u << s;
  }
  n4->child[0] |= leaf_mask << kLeafMaskShift;
  n4->child[1] |= (axis & 3u) << kOrderAxisShift;
  n4->meta[0] = (uint32_t)cnt;
  n4->meta[1] = leaves;
  n4->meta[2] = depth;
  n4->meta[3] = axis;
}

struct Pending {
  int64_t bin;       // binary node to expand
  uint32_t parent;   // flat index of the parent (or ~0u for the root)
  uint32_t slot;
  uint32_t depth;
};

}  // namespace

uint32_t splitmedian_leaf_depth(size_t n, size_t k) {
  size_t b = 0, e = n;
  uint32_t d = 0;
  while (e - b > 1) {
    const size_t m = b + (e - b) / 2;
    if (k < m) {
      e = m;
    } else {
      b = m;
    }
    ++d;
  }
  return d;
}

void splitmedian_voxel_indices(size_t n, uint32_t* out) {
  struct Range {
    size_t b, e;
  };
  std::vector<Range> stack;
  stack.push_back(Range{0, n});
  uint32_t voxel_index = 0;
  while (!stack.empty()) {
    const Range r = stack.back();
    stack.pop_back();
    if (r.e - r.b == 1) out[r.b] = voxel_index;
    ++voxel_index;
    if (r.e - r.b > 1) {
      const size_t m = r.b + (r.e - r.b) / 2;
      stack.push_back(Range{r.b, m});  // left pushed first: the right subtree is visited first
      stack.push_back(Range{m, r.e});
    }
  }
}

void splitmedian_order(const float* boxes, size_t n, uint32_t* order, uint32_t* depth) {
  std::vector<float> cen[3];
  for (int a = 0; a < 3; ++a) cen[a].resize(n);
  for (size_t i = 0; i < n; ++i) {
    for (int a = 0; a < 3; ++a) {
      // BoundingBox3D::getCenter(i) = (min(i) + max(i)) / 2   (include/bh/math/geometry.h:227-229)
      const float s = boxes[6 * i + a] + boxes[6 * i + 3 + a];
      cen[a][i] = s / 2.0f;
    }
  }
  for (size_t i = 0; i < n; ++i) order[i] = (uint32_t)i;
  OrderCtx c;
  c.centre[0] = cen[0].data();
  c.centre[1] = cen[1].data();
  c.centre[2] = cen[2].data();
  c.idx = order;
  c.depth = depth;
#pragma omp parallel
#pragma omp single nowait
  order_recursive(c, 0, n, 0, 0);
}

void build_flat_tree(const float* boxes, size_t n, int top_levels, FlatTree* out) {
  out->nodes.clear();
  out->top_nodes = 0;
  out->depth = 0;
  out->max_stack = 4;
  if (n == 0) return;

  BuildCtx c;
  c.boxes = boxes;
  c.cx.resize(n);
  c.cy.resize(n);
  c.cz.resize(n);
  c.idx.resize(n);
  for (size_t i = 0; i < n; ++i) {
    c.cx[i] = 0.5f * boxes[6 * i + 0] + 0.5f * boxes[6 * i + 3];
    c.cy[i] = 0.5f * boxes[6 * i + 1] + 0.5f * boxes[6 * i + 4];
    c.cz[i] = 0.5f * boxes[6 * i + 2] + 0.5f * boxes[6 * i + 5];
    c.idx[i] = (uint32_t)i;
  }
  c.nodes.resize(2 * n - 1);
  int64_t root = -1;
#pragma omp parallel
#pragma omp single nowait
  root = build_binary(c, 0, n, 0);

  const std::vector<BinNode>& bn = c.nodes;
  std::vector<Node4>& flat = out->nodes;
  flat.reserve(n / 3 + 16);

  // phase 1: breadth-first over the top levels
  std::vector<Pending> queue;
  std::vector<Pending> dfs_roots;
  queue.push_back(Pending{root, ~0u, 0u, 0u});
  size_t head = 0;
  uint32_t max_depth = 0;
  auto emit = [&](const Pending& p, std::vector<Pending>* same_phase, std::vector<Pending>* next_phase, bool reversed) {
    const uint32_t me = (uint32_t)flat.size();
    flat.emplace_back();
    Child ch[4];
    const int cnt = gather_children(c, p.bin, ch);
    const uint32_t axis = (uint32_t)order_children(c, ch, cnt);
    if (p.parent != ~0u) {
      flat[p.parent].child[p.slot] |= me;  // slot 0 already holds the parent's leaf mask in its top bits
    }
    fill_node(c, ch, cnt, p.depth, bn[p.bin].count, axis, &flat[me]);
    max_depth = std::max(max_depth, p.depth);
    for (int k = 0; k < cnt; ++k) {
      const int s = reversed ? cnt - 1 - k : k;
      if (ch[s].bin < 0) continue;
      Pending q{ch[s].bin, me, (uint32_t)s, p.depth + 1};
      if (same_phase != nullptr && (int)(p.depth + 1) < top_levels) {
        same_phase->push_back(q);
      } else {
        next_phase->push_back(q);
      }
    }
  };
  while (head < queue.size()) {
    const Pending p = queue[head++];
    emit(p, &queue, &dfs_roots, false);
  }
  out->top_nodes = (uint32_t)flat.size();

  // phase 2: depth-first (pre-order) below the top levels, explicit stack
  std::vector<Pending> stack;
  for (size_t r = 0; r < dfs_roots.size(); ++r) {
    stack.push_back(dfs_roots[r]);
    while (!stack.empty()) {
      const Pending p = stack.back();
      stack.pop_back();
      emit(p, nullptr, &stack, true);  // reversed so that slot 0's subtree follows its parent
    }
  }
  out->depth = max_depth;
  out->max_stack = 3 * (max_depth + 1) + 1;
}

}  // namespace q3dr
