// overlap_kernels.cuh -- box-vs-map overlap queries on the flattened tree:
//   bvh::Tree::intersects(bbox) / intersectsRecursive(bbox, ...)      viewpoint_planner/src/bvh/bvh.h:544-554,911-939
//   BoundingBox3D::intersects(other)                                  include/bh/math/geometry.h:277-285
//   ViewpointPlannerData::isValidObjectPosition                       viewpoint_planner_data.cpp:209-235
// The queries only compare (a leaf overlaps unless max < other.min or other.max < min on some axis; touching boxes
// overlap), so any tree over the same leaves gives the reference's answer; inner boxes are exact unions of leaf boxes.
// One thread per query: candidate positions are spatially incoherent and a drone-sized box meets a handful of nodes.
#pragma once

#include "raycast_kernels.cuh"

namespace q3dr {

constexpr int kOverlapThreads = 128;

struct QueryBox {
  float mnx, mny, mnz, mxx, mxy, mxz;
};

// Visits every leaf whose box overlaps q; f(leaf) returns false to stop.  Returns false if stopped.
template <typename F>
__device__ __forceinline__ bool for_each_overlapping_leaf(const float4* __restrict__ nodes, const QueryBox& q, F f) {
  uint32_t stack[kStackEntries];
  int sp = 0;
  uint32_t node = 0;
  while (true) {
    const float4* np = nodes + (size_t)node * 8u;
    const float4 mnx = __ldg(np + 0), mny = __ldg(np + 1), mnz = __ldg(np + 2);
    const float4 mxx = __ldg(np + 3), mxy = __ldg(np + 4), mxz = __ldg(np + 5);
    const uint4 ch = __ldg(reinterpret_cast<const uint4*>(np + 6));
#define Q3DR_OVL(C, CH)                                                                                    \
    {                                                                                                      \
      /* node.getBoundingBox().intersects(bbox): !(max < other.min).any() && !(other.max < min).any() */  \
      const bool hit = !(mxx.C < q.mnx) && !(mxy.C < q.mny) && !(mxz.C < q.mnz) && !(q.mxx < mnx.C) &&     \
                       !(q.mxy < mny.C) && !(q.mxz < mnz.C);                                               \
      if (hit) {                                                                                           \
        const uint32_t c = (CH) & (kLeafFlag | kPayloadMask);                                              \
        if (c & kLeafFlag) {                                                                               \
          if (c != kEmptyChild) {                                                                          \
            if (!f((int32_t)(c & kPayloadMask))) return false;                                             \
          }                                                                                                \
        } else {                                                                                           \
          stack[sp++] = c;                                                                                 \
        }                                                                                                  \
      }                                                                                                    \
    }
    Q3DR_OVL(x, ch.x)
    Q3DR_OVL(y, ch.y)
    Q3DR_OVL(z, ch.z)
    Q3DR_OVL(w, ch.w)
#undef Q3DR_OVL
    if (sp == 0) return true;
    node = stack[--sp];
  }
}

// counts[i] = number of leaves overlapping boxes[i]; with out_leaf != nullptr also writes them (unordered) at
// out_leaf[offsets[i] ...].
static __global__ void __launch_bounds__(kOverlapThreads) overlap_boxes_kernel(const float4* __restrict__ nodes,
                                                                        const float* __restrict__ boxes, uint32_t n,
                                                                        uint32_t* counts, const uint64_t* __restrict__ offsets,
                                                                        int32_t* out_leaf) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  QueryBox q;
  q.mnx = boxes[6 * i + 0]; q.mny = boxes[6 * i + 1]; q.mnz = boxes[6 * i + 2];
  q.mxx = boxes[6 * i + 3]; q.mxy = boxes[6 * i + 4]; q.mxz = boxes[6 * i + 5];
  uint32_t cnt = 0;
  if (out_leaf == nullptr) {
    for_each_overlapping_leaf(nodes, q, [&](int32_t) { ++cnt; return true; });
    counts[i] = cnt;
  } else {
    int32_t* dst = out_leaf + offsets[i];
    for_each_overlapping_leaf(nodes, q, [&](int32_t leaf) { dst[cnt++] = leaf; return true; });
  }
}

struct ValidParams {
  float obj[6];        // object_bbox (drone) relative to its position
  float bvh_bbox[6];   // ViewpointPlannerData::bvh_bbox_
  float root_box[6];   // occupied_bvh_.getRoot()->getBoundingBox()
  float obstacle_free_height;
};

// ViewpointPlannerData::isValidObjectPosition without the no-fly-zone polygons (those stay on the host).
static __global__ void __launch_bounds__(kOverlapThreads) valid_positions_kernel(const float4* __restrict__ nodes,
                                                                          const float* __restrict__ pos, uint32_t n,
                                                                          const ValidParams p, uint8_t* valid) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
  bool ok;
  if (z >= p.obstacle_free_height) {
    // bvh_bbox_.isInside(position): inclusive on both sides (geometry.h:240-243)
    ok = x >= p.bvh_bbox[0] && y >= p.bvh_bbox[1] && z >= p.bvh_bbox[2] && x <= p.bvh_bbox[3] && y <= p.bvh_bbox[4] &&
         z <= p.bvh_bbox[5];
  } else if (x >= p.root_box[0] && y >= p.root_box[1] && z >= p.root_box[2] && x <= p.root_box[3] && y <= p.root_box[4] &&
             z <= p.root_box[5]) {
    QueryBox q;  // object_bbox + position (geometry.h:179-181)
    q.mnx = __fadd_rn(p.obj[0], x); q.mny = __fadd_rn(p.obj[1], y); q.mnz = __fadd_rn(p.obj[2], z);
    q.mxx = __fadd_rn(p.obj[3], x); q.mxy = __fadd_rn(p.obj[4], y); q.mxz = __fadd_rn(p.obj[5], z);
    if (q.mxz >= p.obstacle_free_height) {
      // the reference crops with the UN-centred object maximum (viewpoint_planner_data.cpp:222-226); kept as is
      q.mxx = p.obj[3];
      q.mxy = p.obj[4];
      q.mxz = p.obstacle_free_height;
    }
    ok = for_each_overlapping_leaf(nodes, q, [](int32_t) { return false; });
  } else {
    ok = false;
  }
  valid[i] = ok ? 1 : 0;
}

}  // namespace q3dr
