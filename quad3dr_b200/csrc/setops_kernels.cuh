// setops_kernels.cuh -- consumers of the per-viewpoint voxel sets: the information
// sums the planner's path search evaluates over the gathered hit lists.  They run on the CSR result of the last
// scoring call while it is still resident in HBM; a viewpoint path's observed_voxel_map
// (std::unordered_map<voxel, information>, viewpoint_planner.h) is a dense float array over the leaves, NaN = absent.
//
//   new_information_kernel   ViewpointPlanner::computeNewInformation        viewpoint_planner_scoring.cpp:124-147
//   observed_add_kernel      addObservedVoxelsToViewpointPath               viewpoint_planner_path.cpp:203-233
//   overlap kernels          compute_overlap_information_lambda             viewpoint_planner_path.cpp:838-848
//
// Per-voxel arithmetic is the reference's, operation for operation (fp32, std::min ternary); the sums over a voxel
// set have no defined order in the reference (hash iteration), here they are fp64 fixed-tree sums rounded once.
#pragma once

#include "raycast_kernels.cuh"

namespace q3dr {

constexpr int kSetThreads = 256;

struct SetParams {
  const uint64_t* __restrict__ offsets;  // CSR of the last scoring result
  const int32_t* __restrict__ leaf;
  const float* __restrict__ info;
  const float4* __restrict__ leaf_rec;   // weight = leaf_rec[kRecStride * leaf].w
  uint32_t rec_stride;
  float factor;                          // options_.viewpoint_information_factor
};

__device__ __forceinline__ double block_sum_f64(double v, double* s_part) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(Q3DR_FULL_MASK, v, o);
  __syncthreads();
  if (lane == 0) s_part[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kSetThreads / 32; ++i) t += s_part[i];
  }
  return t;  // valid in thread 0
}

// novel information of one voxel against the observed map (computeNewInformation's lambda)
__device__ __forceinline__ float novel_information(float obs_info, float observed, float weight) {
  if (observed != observed) return obs_info;  // not in the map
  return smin(obs_info, __fsub_rn(weight, observed));
}

// one CTA per requested viewpoint
static __global__ void __launch_bounds__(kSetThreads) new_information_kernel(const SetParams p, const float* __restrict__ observed,
                                                                      const uint32_t* __restrict__ vp_index /* nullable */,
                                                                      float* out) {
  __shared__ double s_part[kSetThreads / 32];
  const uint32_t vp = vp_index ? vp_index[blockIdx.x] : blockIdx.x;
  const uint64_t b = p.offsets[vp], e = p.offsets[vp + 1];
  double acc = 0.0;
  for (uint64_t i = b + threadIdx.x; i < e; i += kSetThreads) {
    const int32_t l = p.leaf[i];
    const float obs_info = __fmul_rn(p.factor, p.info[i]);
    const float w = __ldg(&p.leaf_rec[(size_t)p.rec_stride * l].w);
    acc += (double)novel_information(obs_info, observed[l], w);
  }
  const double t = block_sum_f64(acc, s_part);
  if (threadIdx.x == 0) out[blockIdx.x] = (float)t;
}

// folds viewpoint vp's voxel set into the observed map; each voxel occurs once in the set, so the update is free of
// conflicts.  Grid-stride over the entries, per-CTA partial sums in part[blockIdx.x].
static __global__ void __launch_bounds__(kSetThreads) observed_add_kernel(const SetParams p, float* observed, uint32_t vp,
                                                                   double* part) {
  __shared__ double s_part[kSetThreads / 32];
  const uint64_t b = p.offsets[vp], e = p.offsets[vp + 1];
  double acc = 0.0;
  for (uint64_t i = b + (uint64_t)blockIdx.x * kSetThreads + threadIdx.x; i < e; i += (uint64_t)gridDim.x * kSetThreads) {
    const int32_t l = p.leaf[i];
    const float obs_info = __fmul_rn(p.factor, p.info[i]);
    const float cur = observed[l];
    float novel;
    if (cur != cur) {
      observed[l] = obs_info;
      novel = obs_info;
    } else {
      const float w = __ldg(&p.leaf_rec[(size_t)p.rec_stride * l].w);
      novel = smin(obs_info, __fsub_rn(w, cur));
      if (cur <= w) {
        float nv = __fadd_rn(cur, obs_info);
        if (nv > w) nv = w;
        observed[l] = nv;
      }
    }
    acc += (double)novel;
  }
  const double t = block_sum_f64(acc, s_part);
  if (threadIdx.x == 0) part[blockIdx.x] = t;
}

// The same two sums for a voxel set handed over explicitly (device arrays leaf / info of n entries) instead of a
// viewpoint of the resident result.
static __global__ void __launch_bounds__(kSetThreads) new_information_list_kernel(const SetParams p, const float* __restrict__ observed,
                                                                                  const int32_t* __restrict__ leaf,
                                                                                  const float* __restrict__ info, uint32_t n,
                                                                                  double* part) {
  __shared__ double s_part[kSetThreads / 32];
  double acc = 0.0;
  for (uint32_t i = blockIdx.x * kSetThreads + threadIdx.x; i < n; i += gridDim.x * kSetThreads) {
    const int32_t l = leaf[i];
    const float obs_info = __fmul_rn(p.factor, info[i]);
    const float w = __ldg(&p.leaf_rec[(size_t)p.rec_stride * l].w);
    acc += (double)novel_information(obs_info, observed[l], w);
  }
  const double t = block_sum_f64(acc, s_part);
  if (threadIdx.x == 0) part[blockIdx.x] = t;
}

static __global__ void __launch_bounds__(kSetThreads) observed_add_list_kernel(const SetParams p, float* observed,
                                                                               const int32_t* __restrict__ leaf,
                                                                               const float* __restrict__ info, uint32_t n,
                                                                               double* part) {
  __shared__ double s_part[kSetThreads / 32];
  double acc = 0.0;
  for (uint32_t i = blockIdx.x * kSetThreads + threadIdx.x; i < n; i += gridDim.x * kSetThreads) {
    const int32_t l = leaf[i];
    const float obs_info = __fmul_rn(p.factor, info[i]);
    const float cur = observed[l];
    float novel;
    if (cur != cur) {
      observed[l] = obs_info;
      novel = obs_info;
    } else {
      const float w = __ldg(&p.leaf_rec[(size_t)p.rec_stride * l].w);
      novel = smin(obs_info, __fsub_rn(w, cur));
      if (cur <= w) {
        float nv = __fadd_rn(cur, obs_info);
        if (nv > w) nv = w;
        observed[l] = nv;
      }
    }
    acc += (double)novel;
  }
  const double t = block_sum_f64(acc, s_part);
  if (threadIdx.x == 0) part[blockIdx.x] = t;
}

static __global__ void sum_parts_kernel(const double* __restrict__ part, uint32_t n, float* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (uint32_t i = 0; i < n; ++i) t += part[i];
    out[0] = (float)t;
  }
}

// dense[leaf] = info for every entry of viewpoint vp (value = NaN: remove them again)
static __global__ void scatter_set_kernel(const SetParams p, uint32_t vp, float* dense, int clear) {
  const uint64_t b = p.offsets[vp], e = p.offsets[vp + 1];
  const float nan = __int_as_float(0x7FC00000);
  for (uint64_t i = b + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += (uint64_t)gridDim.x * blockDim.x) {
    dense[p.leaf[i]] = clear ? nan : p.info[i];
  }
}

// compute_overlap_information_lambda(set1 = viewpoint b, set2 = the scattered viewpoint a): the reference looks b's
// entry up in a with VoxelWithInformation::operator== (occupied_tree.h:73-79), which compares voxel AND information,
// so only voxels carrying bit-identical information in both sets count; each adds min(info_a, info_b).
static __global__ void __launch_bounds__(kSetThreads) overlap_information_kernel(const SetParams p, const float* __restrict__ dense_a,
                                                                          const uint32_t* __restrict__ vp_index, float* out) {
  __shared__ double s_part[kSetThreads / 32];
  const uint32_t vp = vp_index[blockIdx.x];
  const uint64_t b = p.offsets[vp], e = p.offsets[vp + 1];
  double acc = 0.0;
  for (uint64_t i = b + threadIdx.x; i < e; i += kSetThreads) {
    const float ib = p.info[i];
    const float ia = dense_a[p.leaf[i]];
    if (ia == ib) acc += (double)smin(ia, ib);
  }
  const double t = block_sum_f64(acc, s_part);
  if (threadIdx.x == 0) out[blockIdx.x] = (float)t;
}

static __global__ void fill_f32_kernel(float* p, size_t n, float v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace q3dr
