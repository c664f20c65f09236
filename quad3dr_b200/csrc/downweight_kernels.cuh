// downweight_kernels.cuh -- ViewpointPlannerData::updateWeightsWithRealViewpoints on the device (SURVEY.md section 8f,
// row N4; viewpoint_planner_data.cpp:498-571).
//
// Per real image the reference casts every pixel of the depth camera (remove_duplicates = false), and then, in scan
// order, every hit pixel whose depth-map value is invalid (<= 0 or infinite: nothing was reconstructed there) lowers the
// weight of the voxel it hit:  information = (res * inc) * CURRENT weight;  weight = max(0, weight - f * information).
// The updates of one voxel form a sequential chain in pixel order, chains of different voxels are independent, and
// images follow each other.  Here: the per-pixel hits come from the camera-window kernel (raycast2_kernel<kModePixels>),
// dw_keys_kernel turns the invalid-depth hits into 64-bit keys (leaf << 32 | pixel), a radix sort groups them by voxel
// in ascending pixel order, and dw_apply_kernel walks each voxel's run with one thread in exactly the reference's order
// and arithmetic.  Once per map, off the hot path.
#pragma once

#include "score_kernels.cuh"

namespace q3dr {

constexpr uint64_t kNoKey = ~0ull;

static __global__ void dw_keys_kernel(const q3dr_pixel_hit* __restrict__ pixels, const float* __restrict__ depth, uint32_t n_pix,
                                      uint64_t* __restrict__ keys) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pix) return;
  const int32_t leaf = pixels[i].leaf;
  const float d = depth[i];
  // depth <= 0 || std::isinf(depth)   (viewpoint_planner_data.cpp:554; a NaN depth is neither)
  const bool invalid = (d <= 0.0f) || (fabsf(d) == __int_as_float(0x7F800000));
  keys[i] = (leaf >= 0 && invalid) ? (((uint64_t)(uint32_t)leaf << 32) | (uint64_t)i) : kNoKey;
}

struct DownweightParams {
  float4* leaf_rec;             // weights are updated in place (record word 3)
  const float* rt;              // 12 floats of this image's pose
  const uint32_t* normal_image; // kNormalImage: ARGB32 [h][w] of this image
  uint32_t n_keys;
  float factor;                 // invalid_pixel_observation_factor
};

// One thread per run of equal leaf: the reference's sequential updates of that voxel, pixel after pixel.
template <int NMODE>
static __global__ void dw_apply_kernel(const ScoreParams2 p, const DownweightParams q, const uint64_t* __restrict__ keys,
                                       const MeshDev mesh, const RenderOptsDev ropts, unsigned long long* n_updates) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= q.n_keys) return;
  const uint64_t k0 = keys[i];
  if (k0 == kNoKey) return;
  const uint32_t leaf = (uint32_t)(k0 >> 32);
  if (i > 0 && (uint32_t)(keys[i - 1] >> 32) == leaf) return;  // not the start of this voxel's run
  float4* rec = q.leaf_rec + kLeafRecStride * (size_t)leaf;
  float4 ra = rec[0];
  const float4 rb = rec[1];
  const float4 rc = rec[2];
  const float camx = q.rt[3], camy = q.rt[7], camz = q.rt[11];
  float m[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) m[k] = q.rt[k];
  uint32_t count = 0;
  for (uint32_t j = i; j < q.n_keys; ++j) {
    const uint64_t key = keys[j];
    if ((uint32_t)(key >> 32) != leaf) break;  // kNoKey ends the last run as well
    float nx = rc.x, ny = rc.y, nz = rc.z;
    if (NMODE != kNormalVoxel) {
      const uint32_t pix = (uint32_t)key;
      uint32_t code;
      if (NMODE == kNormalImage) {
        code = q.normal_image[pix];
      } else {
        const uint32_t py = pix / p.img_w, px = pix - py * p.img_w;
        MeshRay r;
        mesh_camera_ray(m, p.fx, p.fy, p.cx, p.cy, p.img_w, p.img_h, px, py, r);
        code = mesh_cast(mesh, ropts, r);
      }
      decode_normal_code(code, nx, ny, nz);
    }
    // computeViewpointObservationScore with the voxel's CURRENT weight, then viewpoint_planner_data.cpp:558-560
    const float information = voxel_information_rec(p, ra, rb, nx, ny, nz, camx, camy, camz);
    ra.w = smax(0.0f, __fsub_rn(ra.w, __fmul_rn(q.factor, information)));
    ++count;
  }
  rec[0].w = ra.w;
  atomicAdd(n_updates, (unsigned long long)count);
}

static __global__ void gather_weights_kernel(const float4* __restrict__ leaf_rec, uint32_t n, float* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = leaf_rec[kLeafRecStride * (size_t)i].w;
}

}  // namespace q3dr
