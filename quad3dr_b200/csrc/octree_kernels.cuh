// octree_kernels.cuh -- occupancy-octree leaves -> BVH leaf list, and the grid weight update.
//   leaf rule + box construction   ViewpointPlannerData::generateBVHTree     viewpoint_planner_data.cpp:820-863
//                                  BoundingBox3D(center, size), constrainTo, isEmpty   include/bh/math/geometry.h:159-160,175-177,373-376
//                                  isNodeFree / isNodeKnown                  viewpoint_planner/src/octree/occupancy_map.h:75-106
//   grid weights                   ViewpointPlannerData::updateWeights       viewpoint_planner_data.cpp:438-482,573-577
// Input is what the reference's loop sees per octree leaf (it.getCoordinate(), it.getSize(), the node payload), in
// begin_tree() order; walking the octomap structure itself stays with octomap on the host.
#pragma once

#include "raycast_kernels.cuh"

namespace q3dr {

struct LeafRuleParams {
  const float* __restrict__ center;           // [n][3]
  const float* __restrict__ size;             // [n]
  const float* __restrict__ occupancy;        // [n]
  const uint32_t* __restrict__ observation_count;  // [n]
  uint32_t n;
  float bvh_bbox[6];
  float obstacle_free_height;
  float occupancy_threshold;       // occ_prob_thres_ (0.51)
  uint32_t observation_threshold;  // observation_thres_ (1)
};

// keep[i] = 1 and box[i] = the BVH leaf box if octree leaf i becomes a BVH leaf
static __global__ void leaf_rule_kernel(const LeafRuleParams p, uint32_t* keep, float* box /* [n][6] */) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  bool k = true;
  // if (octree->isNodeFree(node) && octree->isNodeKnown(node)) continue;
  const bool is_free = !(p.occupancy[i] > p.occupancy_threshold);
  const bool is_known = p.observation_count[i] >= p.observation_threshold;
  if (is_free && is_known) k = false;
  // BoundingBox3D(center, size): min = center - size / 2, max = center + size / 2
  const float half = __fdiv_rn(p.size[i], 2.0f);
  float mn[3], mx[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float c = p.center[3 * (size_t)i + a];
    mn[a] = __fsub_rn(c, half);
    mx[a] = __fadd_rn(c, half);
    // constrainTo(bvh_bbox_): cwiseMax / cwiseMin
    mn[a] = fmaxf(mn[a], p.bvh_bbox[a]);
    mx[a] = fminf(mx[a], p.bvh_bbox[3 + a]);
  }
  // isEmpty(): (min >= max).any()
  if (mn[0] >= mx[0] || mn[1] >= mx[1] || mn[2] >= mx[2]) k = false;
  if (mx[2] >= p.obstacle_free_height) {
    mn[2] = smin(p.obstacle_free_height, mn[2]);
    mx[2] = p.obstacle_free_height;
  }
  if (mn[0] >= mx[0] || mn[1] >= mx[1] || mn[2] >= mx[2]) k = false;
  keep[i] = k ? 1u : 0u;
  float* b = box + 6 * (size_t)i;
  b[0] = mn[0]; b[1] = mn[1]; b[2] = mn[2];
  b[3] = mx[0]; b[4] = mx[1]; b[5] = mx[2];
}

// ordered compaction given the exclusive scan of keep[]
static __global__ void leaf_compact_kernel(const uint32_t* __restrict__ keep, const uint32_t* __restrict__ pos, const float* __restrict__ box,
                                    uint32_t n, float* out_box, uint32_t* out_src) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !keep[i]) return;
  const uint32_t d = pos[i];
#pragma unroll
  for (int a = 0; a < 6; ++a) out_box[6 * (size_t)d + a] = box[6 * (size_t)i + a];
  out_src[d] = i;
}

struct GridParams {
  float origin[3];
  float increment;
  int32_t dim[3];
  float lambda;  // options_.voxel_information_lambda
};

// Largest grid index along one axis whose cell [p - inc/2, p + inc/2], p = origin + inc * float(i), overlaps
// [lmin, lmax] (touching counts), or -1.  Cell bounds are non-decreasing in i, so the largest i with cell_min <= lmax
// is the only candidate.
__device__ __forceinline__ int last_cell(float origin, float inc, float half, int dim, float lmin, float lmax) {
  int lo = 0, hi = dim - 1, best = -1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const float p = __fadd_rn(origin, __fmul_rn(inc, (float)mid));
    const float cmin = __fsub_rn(p, half);
    if (!(lmax < cmin)) {
      best = mid;
      lo = mid + 1;
    } else {
      hi = mid - 1;
    }
  }
  if (best < 0) return -1;
  const float p = __fadd_rn(origin, __fmul_rn(inc, (float)best));
  const float cmax = __fadd_rn(p, half);
  return (cmax < lmin) ? -1 : best;
}

// updateWeights: every grid cell writes weight(cell) * exp(-lambda * observation_count) into the leaves its box
// overlaps, cells visited in ix, iy, iz order, so a leaf ends up with the value of the LAST overlapping cell;
// leaves no cell overlaps keep their weight.  weight slot = leaf_rec[stride * leaf].w.
static __global__ void grid_weights_kernel(float4* leaf_rec, uint32_t rec_stride, uint32_t n, const GridParams g,
                                    const float* __restrict__ grid_weight, const uint32_t* __restrict__ observation_count,
                                    float* weight_out /* [n], mirror for the host copy */) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 a = leaf_rec[(size_t)rec_stride * i], b = leaf_rec[(size_t)rec_stride * i + 1];
  const float half = __fdiv_rn(g.increment, 2.0f);
  const int ix = last_cell(g.origin[0], g.increment, half, g.dim[0], a.x, b.x);
  const int iy = last_cell(g.origin[1], g.increment, half, g.dim[1], a.y, b.y);
  const int iz = last_cell(g.origin[2], g.increment, half, g.dim[2], a.z, b.z);
  float w = a.w;
  if (ix >= 0 && iy >= 0 && iz >= 0) {
    const float gw = grid_weight[((size_t)ix * g.dim[1] + iy) * g.dim[2] + iz];
    const float f = expf(__fmul_rn(-g.lambda, (float)observation_count[i]));
    w = __fmul_rn(gw, f);
    leaf_rec[(size_t)rec_stride * i].w = w;
  }
  weight_out[i] = w;
}

}  // namespace q3dr
