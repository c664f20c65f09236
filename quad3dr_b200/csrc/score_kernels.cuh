// score_kernels.cuh -- dedup tables -> ordered, scored, compacted per-viewpoint hit lists.
//
// Input (left by raycast2_kernel<kModeScore> for a chunk of B viewpoints):
//   bitmap[B][W]     one bit per leaf, set when some pixel of the viewpoint hit it
//   first_pix[B][n]  row-major index of the first pixel that hit the leaf (0xFFFFFFFF = not hit)
//   seg_hits[B][S]   number of set bits per SEGMENT of kSegWords bitmap words (counted by the raycast kernel when it
//                    sets a bit for the first time)
// Output: per viewpoint the voxels in ascending leaf order with their information
// (computeViewpointObservationScore, viewpoint_planner_scoring.cpp:11-38,67-73,116-122; viewpoint_score.cpp:45-68),
// first pixel and that pixel's dist_sq; entries with information <= 0 dropped when ignore_zero; total = sum of the
// kept informations.
//
// Visible voxels cluster in a few contiguous leaf ranges per viewpoint, so neither one CTA per viewpoint nor one
// per bitmap segment balances; the scoring and packing passes therefore run over a flat numbering of 256-entry
// blocks of all hit lists of the chunk.
//   seg_scan_kernel             exclusive scan of seg_hits per viewpoint, hit_count           (B CTAs)
//   emit_list_kernel            bits -> ordered leaf list of the viewpoint                    (S x B CTAs, empty ones exit)
//   block_offsets_kernel        numbering of the 256-entry scoring blocks of the chunk        (1 CTA)
//   score_blocks_kernel         information / first pixel / dist_sq per entry, in place       (persistent grid)
//   finalize_viewpoints_kernel  kept counts + offsets per block, fp64 total per viewpoint     (B CTAs)
//   csr_offsets_kernel          CSR offsets of the chunk                                      (1 CTA)
//   pack_blocks_kernel          surviving entries -> dense CSR pool                           (persistent grid)
// Everything is deterministic: fixed assignment of entries to threads, fixed reduction trees.
#pragma once

#include "mesh_kernels.cuh"
#include "raycast_kernels.cuh"

namespace q3dr {

// where the normal that scores a voxel comes from (ViewpointPlanner::computeNormalVector,
// viewpoint_planner_scoring.cpp:26-38)
enum NormalMode : int {
  kNormalVoxel = 0,  // enable_opengl = false: NodeObject::normal of the voxel
  kNormalImage = 1,  // enable_opengl = true: the caller's normal images (QImage ARGB32), looked up at the kept pixel
  kNormalCode = 2    // enable_opengl = true: the code mesh_kept_normals_kernel left next to the hit list
};

constexpr int kSegThreads = 256;
constexpr int kLeafRecStride = 4;  // float4s per leaf record: 64 bytes, so a record never straddles a 128-byte line
constexpr uint32_t kSegWords = 4u * kSegThreads;   // 1024 bitmap words
constexpr uint32_t kSegLeaves = 32u * kSegWords;   // 32768 leaves per segment
constexpr uint32_t kSegShift = 15;                 // leaf >> kSegShift = segment

struct ScoreParams2 {
  const float4* __restrict__ leaf_rec;  // [n][4] (one 64-byte line half): {min xyz, weight}, {max xyz, -}, {normal xyz, -}, pad
  const float* __restrict__ rt;         // [B][12]
  float fx, fy, cx, cy;
  float image_width_f;
  uint32_t x0, y0, w;
  float thr, kf, a_thr, ka;  // derived constants (viewpoint_score.cpp:20-28)
  int32_t ignore_sign;
  int32_t ignore_zero;
  uint32_t n_segs;
  uint32_t* first_pix;
  uint32_t* bitmap;
  uint64_t leaf_stride, word_stride;
  uint64_t cap;  // entries per viewpoint in the tmp arrays
  int32_t* tmp_leaf;
  float* tmp_info;
  uint32_t* tmp_pix;
  float* tmp_dsq;
  uint32_t* seg_hits;   // [B][S]
  uint32_t* seg_off;    // [B][S] exclusive scan of seg_hits
  uint32_t* blk_off;    // [B + 1] first scoring block of each viewpoint (see block_offsets_kernel)
  uint32_t* blk_kept;   // [total blocks] entries of the block that survive the filter
  uint32_t* blk_koff;   // [total blocks] exclusive scan of blk_kept inside the viewpoint
  double* blk_total;    // [total blocks]
  uint32_t* kept_count;  // [B]
  uint32_t* hit_count;   // [n_vp_total] (already offset to the chunk)
  float* total;          // [n_vp_total] (already offset to the chunk)
  // image-space normals (row N2)
  uint32_t pix_mask;   // first_pix entry & pix_mask = pixel index (tagged entries, see RaycastParams::pix_tag)
  int32_t reset_fp;    // untagged mode: write kEmptyPixel back after reading an entry
  int32_t want_dsq;    // compute dist_sq of the kept pixel (Q3DR_RESULT_DIST_SQ)
  const uint32_t* normal_images;  // kNormalImage: [B][img_h][img_w] ARGB32 of the chunk's viewpoints
  uint32_t* tmp_code;             // kNormalCode: [B][cap] next to tmp_leaf
  uint32_t img_w, img_h;
};

// Block-wide exclusive scan for kSegThreads threads; returns the exclusive prefix, `total` = block sum.
// s_warp: kSegThreads / 32 + 1 words.
__device__ __forceinline__ uint32_t seg_block_scan(uint32_t v, uint32_t* s_warp, uint32_t& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = kSegThreads / 32;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(Q3DR_FULL_MASK, inc, o);
    if (lane >= o) inc += y;
  }
  __syncthreads();  // s_warp may still be read by the previous call
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint32_t before = 0, all = 0;
#pragma unroll
  for (int i = 0; i < kWarps; ++i) {
    const uint32_t c = s_warp[i];
    if (i < warp) before += c;
    all += c;
  }
  total = all;
  return before + inc - v;
}

static __global__ void __launch_bounds__(kSegThreads) seg_scan_kernel(const uint32_t* __restrict__ counts, uint32_t* offsets,
                                                               uint32_t n_segs, uint32_t* totals /* [B] or null */) {
  __shared__ uint32_t s_warp[kSegThreads / 32 + 1];
  const uint32_t vp = blockIdx.x;
  const uint32_t* c = counts + (size_t)vp * n_segs;
  uint32_t* o = offsets + (size_t)vp * n_segs;
  uint32_t run = 0;
  for (uint32_t s0 = 0; s0 < n_segs; s0 += kSegThreads) {
    const uint32_t s = s0 + threadIdx.x;
    const uint32_t v = (s < n_segs) ? c[s] : 0u;
    uint32_t tot;
    const uint32_t ex = seg_block_scan(v, s_warp, tot);
    if (s < n_segs) o[s] = run + ex;
    run += tot;
  }
  if (totals != nullptr && threadIdx.x == 0) totals[vp] = run;
}

// bits -> ordered leaf list of the viewpoint (thread t owns 4 consecutive words of the segment); the bitmap is
// cleared on the way and the segment counter reset for the next chunk.
static __global__ void __launch_bounds__(kSegThreads) emit_list_kernel(const ScoreParams2 p) {
  __shared__ uint32_t s_warp[kSegThreads / 32 + 1];
  const uint32_t seg = blockIdx.x, vp = blockIdx.y, tid = threadIdx.x;
  const size_t sidx = (size_t)vp * p.n_segs + seg;
  if (p.seg_hits[sidx] == 0u) return;
  int32_t* list = p.tmp_leaf + (size_t)vp * p.cap + p.seg_off[sidx];
  uint4* words = reinterpret_cast<uint4*>(p.bitmap + (size_t)vp * p.word_stride + (size_t)seg * kSegWords) + tid;
  const uint4 w4 = *words;
  const uint32_t cnt = __popc(w4.x) + __popc(w4.y) + __popc(w4.z) + __popc(w4.w);
  uint32_t tot;
  uint32_t pos = seg_block_scan(cnt, s_warp, tot);
  if (cnt) {
    *words = make_uint4(0u, 0u, 0u, 0u);
    const uint32_t leaf0 = seg * kSegLeaves + tid * 128u;
    uint32_t bits[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t b = bits[j];
      while (b) {
        const int k = __ffs((int)b) - 1;
        b &= b - 1u;
        list[pos++] = (int32_t)(leaf0 + 32u * j + k);
      }
    }
  }
  if (tid == 0) p.seg_hits[sidx] = 0u;
}

// Work decomposition of the scoring / packing passes: viewpoint v owns ceil(hit_count[v] / 256) BLOCKS of 256 list
// entries; blocks of all viewpoints of the chunk are numbered consecutively (blk_off = exclusive scan) and handed
// out round-robin to a persistent grid, so a viewpoint with many visible voxels is spread over many CTAs.
static __global__ void __launch_bounds__(kSegThreads) block_offsets_kernel(const uint32_t* __restrict__ hit_count, uint32_t n_vp,
                                                                    uint32_t* blk_off /* [n_vp + 1] */) {
  __shared__ uint32_t s_warp[kSegThreads / 32 + 1];
  uint32_t run = 0;
  for (uint32_t i0 = 0; i0 < n_vp; i0 += kSegThreads) {
    const uint32_t i = i0 + threadIdx.x;
    const uint32_t v = (i < n_vp) ? (hit_count[i] + kSegThreads - 1u) / kSegThreads : 0u;
    uint32_t tot;
    const uint32_t ex = seg_block_scan(v, s_warp, tot);
    if (i < n_vp) blk_off[i] = run + ex;
    run += tot;
  }
  if (threadIdx.x == 0) blk_off[n_vp] = run;
}

constexpr uint32_t kMaxChunkViewpoints = 4096;  // choose_chunk() never exceeds it: blk_off fits in shared memory

// blk_off (n_vp + 1 entries) staged in shared memory once per CTA; every thread then finds the viewpoint of a block
// with a dozen shared-memory probes instead of a chain of dependent global loads.
__device__ __forceinline__ void stage_block_offsets(const uint32_t* __restrict__ blk_off, uint32_t n_vp, uint32_t* s_off) {
  for (uint32_t i = threadIdx.x; i <= n_vp; i += blockDim.x) s_off[i] = __ldg(blk_off + i);
  __syncthreads();
}

__device__ __forceinline__ uint32_t find_viewpoint(const uint32_t* s_off, uint32_t n_vp, uint32_t blk) {
  uint32_t lo = 0, hi = n_vp;  // largest v with s_off[v] <= blk
  while (hi - lo > 1u) {
    const uint32_t mid = (lo + hi) >> 1;
    if (s_off[mid] <= blk) {
      lo = mid;
    } else {
      hi = mid;
    }
  }
  return lo;
}

// information of one voxel from its record (computeViewpointObservationScore)
__device__ __forceinline__ float voxel_information_rec(const ScoreParams2& p, const float4& ra, const float4& rb,
                                                       float nx, float ny, float nz, float camx, float camy,
                                                       float camz) {
  const float gx = __fsub_rn(camx, __fdiv_rn(__fadd_rn(ra.x, rb.x), 2.0f));
  const float gy = __fsub_rn(camy, __fdiv_rn(__fadd_rn(ra.y, rb.y), 2.0f));
  const float gz = __fsub_rn(camz, __fdiv_rn(__fadd_rn(ra.z, rb.z), 2.0f));
  const float n2 = __fadd_rn(__fmul_rn(gx, gx), __fadd_rn(__fmul_rn(gy, gy), __fmul_rn(gz, gz)));
  const float dist = __fsqrt_rn(n2);
  const float size_x = __fsub_rn(rb.x, ra.x);
  const float rel = __fdiv_rn(__fdiv_rn(__fmul_rn(p.fx, size_x), dist), p.image_width_f);
  float res;
  if (rel >= p.thr) {
    res = 1.0f;
  } else {
    res = expf(__fmul_rn(-p.kf, __fsub_rn(p.thr, rel)));
  }
  float inc;
  if (nx == 0.0f && ny == 0.0f && nz == 0.0f) {
    inc = 1.0f;
  } else {
    float vx = gx, vy = gy, vz = gz;
    if (n2 > 0.0f) {
      vx = __fdiv_rn(gx, dist);
      vy = __fdiv_rn(gy, dist);
      vz = __fdiv_rn(gz, dist);
    }
    float dot = __fadd_rn(__fmul_rn(vx, nx), __fadd_rn(__fmul_rn(vy, ny), __fmul_rn(vz, nz)));
    dot = smax(smin(dot, 1.0f), -1.0f);
    bool zero = false;
    if (dot <= 0.0f) {
      if (p.ignore_sign) {
        dot = fabsf(dot);
      } else {
        zero = true;
      }
    }
    if (zero) {
      inc = 0.0f;
    } else {
      const float angle = acosf(dot);
      if (angle <= p.a_thr) {
        inc = 1.0f;
      } else {
        inc = expf(__fmul_rn(-p.ka, __fsub_rn(angle, p.a_thr)));
      }
    }
  }
  return __fmul_rn(__fmul_rn(res, inc), ra.w);
}

// Scores every entry of every hit list of the chunk in place (position i of the list keeps position i of the
// tmp arrays); per block: number of entries that survive the information > 0 filter and their fp64 sum.
// The pass is bound by the latency of two dependent random fetches per voxel (list -> first pixel + record), so a
// CTA works on kScoreUnroll blocks at once and issues all their fetches before any arithmetic.
constexpr int kScoreUnroll = 4;

template <int NMODE>
static __global__ void __launch_bounds__(kSegThreads) score_blocks_kernel(const ScoreParams2 p, uint32_t n_vp) {
  __shared__ uint32_t s_cnt[kScoreUnroll][kSegThreads / 32];
  __shared__ double s_dsum[kScoreUnroll][kSegThreads / 32];
  __shared__ uint32_t s_off[kMaxChunkViewpoints + 1];
  const uint32_t tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  stage_block_offsets(p.blk_off, n_vp, s_off);
  const uint32_t total_blocks = s_off[n_vp];
  for (uint32_t blk0 = blockIdx.x * kScoreUnroll; blk0 < total_blocks; blk0 += gridDim.x * kScoreUnroll) {
    uint32_t vp[kScoreUnroll];
    size_t at[kScoreUnroll];
    bool valid[kScoreUnroll];
    int32_t leaf[kScoreUnroll];
#pragma unroll
    for (int u = 0; u < kScoreUnroll; ++u) {
      const uint32_t blk = blk0 + u;
      valid[u] = false;
      vp[u] = 0;
      at[u] = 0;
      leaf[u] = 0;
      if (blk < total_blocks) {
        vp[u] = find_viewpoint(s_off, n_vp, blk);
        const uint32_t i = (blk - s_off[vp[u]]) * kSegThreads + tid;
        valid[u] = i < p.hit_count[vp[u]];
        at[u] = (size_t)vp[u] * p.cap + i;
      }
    }
#pragma unroll
    for (int u = 0; u < kScoreUnroll; ++u) {
      if (valid[u]) leaf[u] = p.tmp_leaf[at[u]];
    }
    uint32_t pix[kScoreUnroll];
    float4 ra[kScoreUnroll], rb[kScoreUnroll], rc[kScoreUnroll];
#pragma unroll
    for (int u = 0; u < kScoreUnroll; ++u) {
      pix[u] = 0;
      ra[u] = rb[u] = rc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid[u]) {
        if (NMODE == kNormalCode && !p.reset_fp) {
          pix[u] = p.tmp_pix[at[u]];  // mesh_kept_normals_kernel already fetched the kept pixel: no second random read
        } else {
          uint32_t* fp = p.first_pix + (size_t)vp[u] * p.leaf_stride + (uint32_t)leaf[u];
          pix[u] = *fp & p.pix_mask;
          if (p.reset_fp) *fp = kEmptyPixel;
        }
        const float4* rec = p.leaf_rec + kLeafRecStride * (size_t)leaf[u];
        ra[u] = __ldg(rec);
        rb[u] = __ldg(rec + 1);
        if (NMODE == kNormalVoxel) {
          rc[u] = __ldg(rec + 2);
        } else {
          uint32_t code;
          if (NMODE == kNormalImage) {
            const uint32_t py = pix[u] / p.w, px = pix[u] - py * p.w;
            code = __ldg(p.normal_images + ((size_t)vp[u] * p.img_h + (p.y0 + py)) * p.img_w + (p.x0 + px));
          } else {
            code = p.tmp_code[at[u]];
          }
          decode_normal_code(code, rc[u].x, rc[u].y, rc[u].z);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kScoreUnroll; ++u) {
      bool keep = false;
      float info = 0.0f;
      if (valid[u]) {
        const float* rt = p.rt + 12 * (size_t)vp[u];
        info = voxel_information_rec(p, ra[u], rb[u], rc[u].x, rc[u].y, rc[u].z, __ldg(rt + 3), __ldg(rt + 7),
                                     __ldg(rt + 11));
        keep = !p.ignore_zero || info > 0.0f;
        float dsq = 0.0f;
        if (keep && p.want_dsq) {
          // dist_sq of the kept pixel: the reference's own test of that pixel's ray against the voxel
          const uint32_t py = pix[u] / p.w, px = pix[u] - py * p.w;
          Ray ray;
          camera_ray(rt, p.fx, p.fy, p.cx, p.cy, p.x0 + px, p.y0 + py, ray);
          float t;
          box_test(ray, ra[u].x, ra[u].y, ra[u].z, rb[u].x, rb[u].y, rb[u].z, dsq, t);
        }
        p.tmp_info[at[u]] = info;
        p.tmp_pix[at[u]] = pix[u];
        p.tmp_dsq[at[u]] = dsq;
      }
      // fixed-tree reductions: warp level here, block level below
      const uint32_t wc = __popc(__ballot_sync(Q3DR_FULL_MASK, keep));
      double part = keep ? (double)info : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(Q3DR_FULL_MASK, part, o);
      if (lane == 0) {
        s_cnt[u][warp] = wc;
        s_dsum[u][warp] = part;
      }
    }
    __syncthreads();
    if (tid < kScoreUnroll && blk0 + tid < total_blocks) {
      uint32_t c = 0;
      double t = 0.0;
      for (int w = 0; w < kSegThreads / 32; ++w) {
        c += s_cnt[tid][w];
        t += s_dsum[tid][w];
      }
      p.blk_kept[blk0 + tid] = c;
      p.blk_total[blk0 + tid] = t;
    }
    __syncthreads();
  }
}

// enable_opengl = true with a mesh: the normal code of every list entry's kept pixel (mesh_kernels.cuh), one thread
// per entry over the same flat block numbering; runs before score_blocks_kernel (which resets first_pix).
static __global__ void __launch_bounds__(kSegThreads) mesh_kept_normals_kernel(const ScoreParams2 p, uint32_t n_vp,
                                                                              const MeshDev mesh,
                                                                              const RenderOptsDev ropts) {
  __shared__ uint32_t s_off[kMaxChunkViewpoints + 1];
  stage_block_offsets(p.blk_off, n_vp, s_off);
  const uint32_t total_blocks = s_off[n_vp];
  for (uint32_t blk = blockIdx.x; blk < total_blocks; blk += gridDim.x) {
    const uint32_t vp = find_viewpoint(s_off, n_vp, blk);
    const uint32_t i = (blk - s_off[vp]) * kSegThreads + threadIdx.x;
    if (i >= p.hit_count[vp]) continue;
    const size_t at = (size_t)vp * p.cap + i;
    const int32_t leaf = p.tmp_leaf[at];
    const uint32_t pix = p.first_pix[(size_t)vp * p.leaf_stride + (uint32_t)leaf] & p.pix_mask;
    const uint32_t py = pix / p.w, px = pix - py * p.w;
    const float* rt = p.rt + 12 * (size_t)vp;
    float m[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) m[k] = __ldg(rt + k);
    MeshRay r;
    mesh_camera_ray(m, p.fx, p.fy, p.cx, p.cy, p.img_w, p.img_h, p.x0 + px, p.y0 + py, r);
    p.tmp_code[at] = mesh_cast(mesh, ropts, r);
    p.tmp_pix[at] = pix;  // score_blocks_kernel<kNormalCode> takes the kept pixel from here
  }
}

// per viewpoint: exclusive scan of its blocks' kept counts, fp64 total over its blocks in ascending order
static __global__ void __launch_bounds__(kSegThreads) finalize_viewpoints_kernel(const ScoreParams2 p) {
  __shared__ uint32_t s_warp[kSegThreads / 32 + 1];
  __shared__ double s_dsum[kSegThreads];
  const uint32_t vp = blockIdx.x;
  const uint32_t b0 = p.blk_off[vp], nb = p.blk_off[vp + 1] - b0;
  uint32_t run = 0;
  double tsum = 0.0;  // thread t sums blocks t, t + 256, ... ; threads are then added in index order
  for (uint32_t s0 = 0; s0 < nb; s0 += kSegThreads) {
    const uint32_t s = s0 + threadIdx.x;
    const uint32_t v = (s < nb) ? p.blk_kept[b0 + s] : 0u;
    if (s < nb) tsum += p.blk_total[b0 + s];
    uint32_t tot;
    const uint32_t ex = seg_block_scan(v, s_warp, tot);
    if (s < nb) p.blk_koff[b0 + s] = run + ex;
    run += tot;
  }
  s_dsum[threadIdx.x] = tsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < kSegThreads; ++i) t += s_dsum[i];
    p.kept_count[vp] = run;
    p.total[vp] = (float)t;
  }
}

struct PackParams2 {
  const uint64_t* __restrict__ offsets;  // CSR offsets at the chunk's first viewpoint
  const uint32_t* __restrict__ blk_off;
  const uint32_t* __restrict__ blk_koff;
  const uint32_t* __restrict__ hit_count;
  uint32_t n_vp;
  int32_t ignore_zero;
  uint64_t cap;
  const int32_t* __restrict__ tmp_leaf;
  const float* __restrict__ tmp_info;
  const uint32_t* __restrict__ tmp_pix;
  const float* __restrict__ tmp_dsq;
  int32_t* out_leaf;
  float* out_info;
  uint32_t* out_pix;  // nullptr: field not requested (q3dr_map_set_result_fields)
  float* out_dsq;
  uint64_t pool_cap;  // entries the out_* pools hold
};

// kept entries -> dense CSR pool, ascending leaf order preserved (block-local scan of the keep flags)
static __global__ void __launch_bounds__(kSegThreads) pack_blocks_kernel(const PackParams2 p) {
  __shared__ uint32_t s_warp[kSegThreads / 32 + 1];
  __shared__ uint32_t s_off[kMaxChunkViewpoints + 1];
  const uint32_t tid = threadIdx.x;
  stage_block_offsets(p.blk_off, p.n_vp, s_off);
  const uint32_t total_blocks = s_off[p.n_vp];
  for (uint32_t blk = blockIdx.x; blk < total_blocks; blk += gridDim.x) {
    const uint32_t vp = find_viewpoint(s_off, p.n_vp, blk);
    const uint32_t j = blk - s_off[vp];
    const uint32_t i = j * kSegThreads + tid;
    const size_t at = (size_t)vp * p.cap + i;
    float info = 0.0f;
    bool keep = false;
    if (i < p.hit_count[vp]) {
      info = p.tmp_info[at];
      keep = !p.ignore_zero || info > 0.0f;
    }
    uint32_t tot;
    const uint32_t rank = seg_block_scan(keep ? 1u : 0u, s_warp, tot);
    const uint64_t dst = keep ? p.offsets[vp] + p.blk_koff[blk] + rank : 0;
    if (keep && dst < p.pool_cap) {  // a pool that turns out too small is grown by the host and the chunk packed again
      p.out_leaf[dst] = p.tmp_leaf[at];
      p.out_info[dst] = info;
      if (p.out_pix != nullptr) p.out_pix[dst] = p.tmp_pix[at];
      if (p.out_dsq != nullptr) p.out_dsq[dst] = p.tmp_dsq[at];
    }
    __syncthreads();
  }
}

// offsets[i + 1] = offsets[0] + inclusive_scan(kept)[i]; one block, n <= 4096 viewpoints per chunk.
static __global__ void __launch_bounds__(kSegThreads) csr_offsets_kernel(const uint32_t* __restrict__ kept, uint32_t n,
                                                                  uint64_t* offsets /* at the chunk's first viewpoint */) {
  __shared__ uint32_t s_warp[kSegThreads / 32 + 1];
  uint64_t run = offsets[0];
  for (uint32_t i0 = 0; i0 < n; i0 += kSegThreads) {
    const uint32_t i = i0 + threadIdx.x;
    const uint32_t v = (i < n) ? kept[i] : 0u;
    uint32_t tot;
    const uint32_t ex = seg_block_scan(v, s_warp, tot);
    if (i < n) offsets[i + 1] = run + ex + v;
    run += tot;
  }
}

}  // namespace q3dr
