// q3dr_api.cu -- the C ABI (include/q3dr_raycast.h) over the sm_100a kernels: map lifecycle, raycast and scoring
// entry points.  (q3dr_queries.cu holds the set consumers, overlap queries, octree helpers and q3dr_multi.)
//
// No CPU compute path exists in this library: without a usable CUDA device every compute entry point
// returns Q3DR_ERR_NO_DEVICE.  Host code here only flattens the map (bvh_build.cpp), moves buffers and
// launches kernels.
#include "q3dr_internal.cuh"

#include "raycast2_kernels.cuh"
#include "score_kernels.cuh"

namespace q3dr_detail {
std::string& last_error() {
  thread_local std::string msg;
  return msg;
}
}  // namespace q3dr_detail

namespace {

// Process-wide pool of released host blocks (pairs of pinned arrays, q3dr_result_detach_host): a caller that keeps a few
// results alive and drops them again allocates pinned memory a few times in total, not once per call.
struct HostPair {
  int32_t* leaf = nullptr;
  float* info = nullptr;
  size_t cap = 0;
};
std::mutex& host_pool_mutex() {
  static std::mutex mu;
  return mu;
}
std::vector<HostPair>& host_pool() {
  static std::vector<HostPair> pool;
  return pool;
}
constexpr size_t kHostPoolBlocks = 3;

// the smallest pooled pair with cap >= n; with n == 0 the largest one
bool host_pool_take(size_t n, HostPair* out) {
  std::lock_guard<std::mutex> lock(host_pool_mutex());
  std::vector<HostPair>& pool = host_pool();
  int pick = -1;
  for (size_t i = 0; i < pool.size(); ++i) {
    if (pool[i].cap < n) continue;
    if (pick < 0 || (n == 0 ? pool[i].cap > pool[(size_t)pick].cap : pool[i].cap < pool[(size_t)pick].cap)) pick = (int)i;
  }
  if (pick < 0) return false;
  *out = pool[(size_t)pick];
  pool.erase(pool.begin() + pick);
  return true;
}

void host_pool_give(const HostPair& hp) {
  HostPair drop;
  {
    std::lock_guard<std::mutex> lock(host_pool_mutex());
    std::vector<HostPair>& pool = host_pool();
    pool.push_back(hp);
    if (pool.size() <= kHostPoolBlocks) return;
    size_t smallest = 0;
    for (size_t i = 1; i < pool.size(); ++i) {
      if (pool[i].cap < pool[smallest].cap) smallest = i;
    }
    drop = pool[smallest];
    pool.erase(pool.begin() + (std::ptrdiff_t)smallest);
  }
  if (drop.leaf) cudaFreeHost(drop.leaf);
  if (drop.info) cudaFreeHost(drop.info);
  cudaGetLastError();
}

// a map without host arrays for (leaf, information) -- they were detached -- takes a pooled pair that is large enough
void adopt_pooled_host_pair(q3dr_map* m, size_t n) {
  if (m->h_leaf.p != nullptr || m->h_info.p != nullptr) return;
  HostPair hp;
  if (!host_pool_take(n, &hp)) return;
  m->h_leaf.p = hp.leaf;
  m->h_leaf.cap = hp.cap;
  m->h_info.p = hp.info;
  m->h_info.cap = hp.cap;
}

void release_device(q3dr_map* m) {
  m->d_nodes.release();
  m->d_leaf_rec.release();
  m->d_q_boxes.release();
  m->d_q_counts.release();
  m->d_q_offsets.release();
  m->d_q_leaf.release();
  m->d_q_sorted.release();
  m->d_q_tmp.release();
  m->d_q_valid.release();
  m->d_dense.release();
  m->d_set_out.release();
  m->d_set_idx.release();
  m->d_set_part.release();
  m->dense_ready = false;
  m->d_seg_u32.release();
  m->d_blk_u32.release();
  m->d_seg_f64.release();
  m->max_blocks = 0;
  m->d_depth.release();
  m->d_counter.release();
  m->d_planes.release();
  m->d_vp_flags.release();
  m->d_vp_list.release();
  m->d_first_pix.release();
  m->d_bitmap.release();
  m->d_kept.release();
  m->slots = 0;
  m->t_leaf.release();
  m->t_info.release();
  m->t_dsq.release();
  m->t_pix.release();
  m->t_code.release();
  m->d_nimg.release();
  m->tmp_cap = 0;
  m->d_offsets.release();
  m->d_total.release();
  m->d_hit_count.release();
  m->r_leaf.release();
  m->r_info.release();
  m->r_dsq.release();
  m->r_pix.release();
  m->d_rt.release();
  m->d_ray_o.release();
  m->d_ray_d.release();
  m->d_pixels.release();
  m->last_n_vp = m->last_n_entries = 0;
  m->free_retired();
  m->fp_epoch = 0;
  m->fp_tagged_garbage = false;
  m->scratch_dirty = false;
}

int configure_rays_kernel(q3dr_map* m) {
  if (m->smem_bytes > 48 * 1024) {
    Q3DR_CUDA(cudaFuncSetAttribute(q3dr::raycast_rays_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)m->smem_bytes));
  }
  int per_sm = 0;
  Q3DR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, q3dr::raycast_rays_kernel, q3dr::kThreadsPerCta,
                                                          m->smem_bytes));
  if (per_sm < 1) return fail(Q3DR_ERR_INTERNAL, "raycast kernel does not fit on an SM");
  m->grid_rays_ctas = per_sm * m->sm_count;
  return Q3DR_OK;
}

// The camera-window kernel exists in several packet shapes (raycast2_kernels.cuh): rays per lane, tile width, packed
// binary32 arithmetic.  All compute the same results bit for bit; which one a map uses is fixed at creation
// (Q3DR_RAY_VARIANT overrides the default, for measurements).
typedef void (*RaycastKernel)(const q3dr::RaycastParams);
struct RayVariant {
  const char* name;
  int nr, tw, warps;         // rays per lane, tile width, warps per CTA
  bool smem_top;             // top tree levels staged in shared memory (dynamic smem = top nodes x 128 bytes)
  RaycastKernel kernel[2];   // [MODE], with the on-face check (any viewpoint)
  RaycastKernel score_fast;  // kModeScore without the check: viewpoints whose centre lies on no box plane
};
#define Q3DR_VARIANT_SO(NAME, NR, TW, PK, MC, WP, ST, SO)                                                      \
  {NAME, NR, TW, WP, ST,                                                                                          \
   {q3dr::raycast2_kernel<q3dr::kModeScore, NR, TW, PK, true, MC, WP, ST, SO>,                                    \
    q3dr::raycast2_kernel<q3dr::kModePixels, NR, TW, PK, true, MC, WP, ST, SO>},                                  \
   q3dr::raycast2_kernel<q3dr::kModeScore, NR, TW, PK, false, MC, WP, ST, SO>}
#define Q3DR_VARIANT(NAME, NR, TW, PK, MC, WP, ST) Q3DR_VARIANT_SO(NAME, NR, TW, PK, MC, WP, ST, false)
const RayVariant kRayVariants[] = {
    Q3DR_VARIANT("2x8x8", 2, 8, false, 4, 8, false),       // 0: two rays per lane, 8 x 8 pixel packets, 4 CTAs x 8 warps x 64 registers
    Q3DR_VARIANT("2x8x8p", 2, 8, true, 4, 8, false),       // 1: ... with FADD2 / FMUL2
    Q3DR_VARIANT("4x8x16", 4, 8, false, 2, 8, false),      // 2: four rays per lane, 8 x 16 pixel packets, 2 CTAs x 128 registers
    Q3DR_VARIANT("4x8x16p", 4, 8, true, 2, 8, false),      // 3
    Q3DR_VARIANT("2x8x8p/3", 2, 8, true, 3, 8, false),     // 4: like 1 with 3 CTAs x 80 registers (no spills, 24 warps per SM)
    Q3DR_VARIANT("2x8x8/3", 2, 8, false, 3, 8, false),     // 5: like 0 with 3 CTAs x 80 registers
    Q3DR_VARIANT("2x8x8p/7x4", 2, 8, true, 7, 4, false),   // 6: like 1 with 7 CTAs x 4 warps x 72 registers (28 warps per SM)
    Q3DR_VARIANT("2x8x8p/7x4/smem", 2, 8, true, 7, 4, true),  // 7: like 6 with the top tree levels in shared memory
    Q3DR_VARIANT("2x8x8p/8x4", 2, 8, true, 8, 4, false),   // 8: like 1 in CTAs of 4 warps (8 CTAs x 4 warps x 64 registers)
    Q3DR_VARIANT("2x8x8p/9x4", 2, 8, true, 9, 4, false),   // 9: 9 CTAs x 4 warps x 56 registers (36 warps per SM)
    Q3DR_VARIANT_SO("2x8x8p/8x4/op", 2, 8, true, 8, 4, false, true),  // 10: like 8, far siblings stacked nearest-on-top
};
#undef Q3DR_VARIANT
#undef Q3DR_VARIANT_SO
constexpr int kNumRayVariants = (int)(sizeof(kRayVariants) / sizeof(kRayVariants[0]));
constexpr uint32_t kDefaultTileRun = 1;
constexpr int kDefaultRayVariant = 10;  // measured on C2 / C3 (profiles/r02_variants_*.jsonl): packed beats scalar, 32 warps per SM at 64 registers beat 28 at 72 and 36 at 56 by 5 %, ordered push gains 1 %, four rays per lane lose a third

const RayVariant& variant_of(const q3dr_map* m) { return kRayVariants[m->ray_variant]; }

int configure_kernel2(q3dr_map* m) {
  m->ray_variant = kDefaultRayVariant;
  if (const char* e = std::getenv("Q3DR_RAY_VARIANT")) {
    const long v = std::strtol(e, nullptr, 10);
    if (v < 0 || v >= kNumRayVariants) return fail(Q3DR_ERR_INVALID_ARG, "Q3DR_RAY_VARIANT out of range");
    m->ray_variant = (int)v;
  }
  m->tile_run = kDefaultTileRun;
  if (const char* e = std::getenv("Q3DR_TILE_RUN")) m->tile_run = (uint32_t)std::min<long>(64, std::max<long>(1, std::strtol(e, nullptr, 10)));
  const RayVariant& rv = variant_of(m);
  for (int mode = 0; mode < 2; ++mode) {
    int per_sm = 0;
    Q3DR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rv.kernel[mode], rv.warps * 32, rv.smem_top ? m->smem_bytes : 0));
    if (per_sm < 1) return fail(Q3DR_ERR_INTERNAL, "raycast2 kernel does not fit on an SM");
    m->grid2_ctas[mode] = per_sm * m->sm_count;
  }
  {
    int per_sm = 0;
    Q3DR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rv.score_fast, rv.warps * 32, rv.smem_top ? m->smem_bytes : 0));
    if (per_sm < 1) return fail(Q3DR_ERR_INTERNAL, "raycast2 kernel does not fit on an SM");
    m->grid2_ctas[q3dr::kModeScore] = std::min(m->grid2_ctas[q3dr::kModeScore], per_sm * m->sm_count);
  }
  return Q3DR_OK;
}

int upload_map(q3dr_map* m) {
  Q3DR_CUDA(cudaSetDevice(m->device));
  const size_t n = m->n;
  const size_t nn = m->tree->nodes.size();
  Q3DR_TRY(m->d_nodes.ensure(nn * 8));
  Q3DR_CUDA(cudaMemcpy(m->d_nodes.p, m->tree->nodes.data(), nn * sizeof(q3dr::Node4), cudaMemcpyHostToDevice));
  {
    constexpr size_t R = q3dr::kLeafRecStride;
    std::vector<float4> rec(R * n, make_float4(0.f, 0.f, 0.f, 0.f));
    for (size_t k = 0; k < n; ++k) {
      const float* b = m->h_boxes.data() + 6 * k;
      rec[R * k + 0] = make_float4(b[0], b[1], b[2], m->h_weight[k]);
      rec[R * k + 1] = make_float4(b[3], b[4], b[5], 0.0f);
      rec[R * k + 2] = make_float4(m->h_normal[3 * k], m->h_normal[3 * k + 1], m->h_normal[3 * k + 2], 0.0f);
    }
    Q3DR_TRY(m->d_leaf_rec.ensure(R * n));
    Q3DR_CUDA(cudaMemcpy(m->d_leaf_rec.p, rec.data(), R * n * sizeof(float4), cudaMemcpyHostToDevice));
  }
  Q3DR_TRY(m->d_depth.ensure(n));
  Q3DR_CUDA(cudaMemcpy(m->d_depth.p, m->h_depth.data(), n, cudaMemcpyHostToDevice));
  {
    // sorted distinct plane coordinates per axis (flag_viewpoints_kernel)
    std::vector<float> all;
    for (int a = 0; a < 3; ++a) {
      std::vector<float> v(2 * n);
      for (size_t k = 0; k < n; ++k) {
        v[2 * k] = m->h_boxes[6 * k + a];
        v[2 * k + 1] = m->h_boxes[6 * k + 3 + a];
      }
      std::sort(v.begin(), v.end());
      v.erase(std::unique(v.begin(), v.end()), v.end());
      m->n_planes[a] = (uint32_t)v.size();
      all.insert(all.end(), v.begin(), v.end());
      m->h_planes[a] = std::move(v);
    }
    Q3DR_TRY(m->d_planes.ensure(std::max<size_t>(all.size(), 1)));
    Q3DR_CUDA(cudaMemcpy(m->d_planes.p, all.data(), all.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  Q3DR_TRY(m->d_counter.ensure(4));
  Q3DR_CUDA(cudaMemset(m->d_counter.p, 0, 4 * sizeof(uint32_t)));
  Q3DR_TRY(m->h_scalar.ensure(8));
  m->smem_bytes = (size_t)m->tree->top_nodes * sizeof(q3dr::Node4);
  Q3DR_TRY(configure_rays_kernel(m));
  Q3DR_TRY(configure_kernel2(m));
  return Q3DR_OK;
}

// scratch for `slots` viewpoints in flight, `cap` result entries per viewpoint
int ensure_scratch(q3dr_map* m, size_t slots, uint64_t cap) {
  const uint64_t leaf_stride = (m->n + q3dr::kSegLeaves - 1) / q3dr::kSegLeaves * q3dr::kSegLeaves;
  const uint64_t word_stride = leaf_stride / 32;
  const uint32_t n_segs = (uint32_t)(leaf_stride / q3dr::kSegLeaves);
  if (slots > m->slots) {
    m->d_first_pix.release();
    m->d_bitmap.release();
    m->d_seg_u32.release();
    Q3DR_TRY(m->d_seg_u32.ensure(2 * slots * n_segs));
    Q3DR_CUDA(cudaMemsetAsync(m->d_seg_u32.p, 0, m->d_seg_u32.bytes(), m->stream));
    m->n_segs = n_segs;
    Q3DR_TRY(m->d_first_pix.ensure(slots * leaf_stride));
    Q3DR_TRY(m->d_bitmap.ensure(slots * word_stride));
    Q3DR_CUDA(cudaMemsetAsync(m->d_first_pix.p, 0xFF, m->d_first_pix.bytes(), m->stream));
    Q3DR_CUDA(cudaMemsetAsync(m->d_bitmap.p, 0, m->d_bitmap.bytes(), m->stream));
    m->fp_epoch = 0;
    m->fp_tagged_garbage = false;
    Q3DR_TRY(m->d_kept.ensure(slots));
    Q3DR_TRY(m->d_vp_flags.ensure(slots));
    Q3DR_TRY(m->d_vp_list.ensure(slots));
    m->slots = slots;
    m->leaf_stride = leaf_stride;
    m->word_stride = word_stride;
    m->max_blocks = 0;  // the block tables below are laid out with m->slots: lay them out again
  }
  {
    // block tables: [m->slots + 1] blk_off | [m->max_blocks] blk_kept | [m->max_blocks] blk_koff.  Allocation and
    // layout both derive from the stored pair (m->slots, m->max_blocks), which only ever grows together with the buffer.
    const size_t want_blocks = m->slots * ((cap + q3dr::kSegThreads - 1) / q3dr::kSegThreads + 1);
    if (want_blocks > m->max_blocks) {
      m->d_blk_u32.release();
      m->d_seg_f64.release();
      Q3DR_TRY(m->d_blk_u32.ensure(m->slots + 1 + 2 * want_blocks));
      Q3DR_TRY(m->d_seg_f64.ensure(want_blocks));
      m->max_blocks = want_blocks;
    }
    if (m->slots + 1 + 2 * m->max_blocks > m->d_blk_u32.cap) return fail(Q3DR_ERR_INTERNAL, "block table layout exceeds its buffer");
  }
  const size_t need = m->slots * cap;
  if (need > m->t_leaf.cap || cap != m->tmp_cap) {
    Q3DR_TRY(m->t_leaf.ensure(need));
    Q3DR_TRY(m->t_info.ensure(need));
    Q3DR_TRY(m->t_dsq.ensure(need));
    Q3DR_TRY(m->t_pix.ensure(need));
    m->tmp_cap = cap;
  }
  return Q3DR_OK;
}

size_t choose_chunk(const q3dr_map* m, uint32_t tiles_per_vp, size_t n_vp, uint64_t cap) {
  // enough packets in flight that the persistent grid's tail is small ...
  const size_t resident_warps = (size_t)m->grid2_ctas[q3dr::kModeScore] * (size_t)variant_of(m).warps;
  size_t want = (64 * resident_warps + tiles_per_vp - 1) / tiles_per_vp;  // ~64 packets per resident warp
  // ... and enough viewpoints per chunk to amortise the per-chunk host sync and the small scoring kernels, few enough
  // that the D2H copy of a finished chunk overlaps the next one (measured on C2: 150 / 296 / 500 / 1000 viewpoints per
  // chunk -> 26.6 / 26.0 / 28.2 / 29.5 ms end to end)
  want = std::max<size_t>(want, 2 * (size_t)m->sm_count);
  if (const char* e = std::getenv("Q3DR_CHUNK")) {
    want = (size_t)std::min<long>(std::max<long>(1, std::strtol(e, nullptr, 10)), (long)q3dr::kMaxChunkViewpoints);
  }
  // ... bounded by scratch memory (first-pixel tables + temporary lists)
  const uint64_t per_slot = ((m->n + q3dr::kSegLeaves - 1) / q3dr::kSegLeaves * q3dr::kSegLeaves) * 4 + cap * 16;
  const uint64_t budget = 8ull << 30;
  const size_t by_mem = (size_t)std::max<uint64_t>(1, budget / std::max<uint64_t>(per_slot, 1));
  size_t b = std::min(want, by_mem);
  b = std::min(b, n_vp);
  b = std::min<size_t>(b, q3dr::kMaxChunkViewpoints);
  return std::max<size_t>(b, 1);
}

float elapsed_ms(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) {
    cudaGetLastError();  // (an event that was never recorded) -- statistics only, must not leak into the next call's checks
    return 0.f;
  }
  return ms;
}

void fill_raycast_params(const q3dr_map* m, const float K[4], uint32_t x0, uint32_t x1, uint32_t y0, uint32_t y1,
                         float max_range, q3dr::RaycastParams* p) {
  std::memset(p, 0, sizeof(*p));
  p->nodes = m->d_nodes.p;
  p->n_top = m->tree->top_nodes;
  p->fx = K[0];
  p->fy = K[1];
  p->cx = K[2];
  p->cy = K[3];
  p->x0 = x0;
  p->y0 = y0;
  p->w = x1 - x0;
  p->h = y1 - y0;
  const uint32_t tile_w = (uint32_t)variant_of(m).tw, tile_h = (uint32_t)(32 / variant_of(m).tw * variant_of(m).nr);
  p->tiles_x = (p->w + tile_w - 1) / tile_w;
  const uint32_t tiles_y = (p->h + tile_h - 1) / tile_h;
  p->tiles_per_vp = p->tiles_x * tiles_y;
  p->tile_run = m->tile_run;
  p->inv_tiles_x = 1.0 / (double)p->tiles_x;
  p->inv_tiles_per_vp = 1.0 / (double)p->tiles_per_vp;
  p->max_range = max_range;
  p->leaf_depth = m->d_depth.p;
  p->tile_counter = m->d_counter.p;
  p->bounded_map = m->bounded ? 1u : 0u;
}

}  // namespace
int q3dr_detail::check_map(const q3dr_map* m) {
  if (!m) return fail(Q3DR_ERR_INVALID_ARG, "map is NULL");
  if (!m->d_nodes.p) return fail(Q3DR_ERR_INVALID_ARG, "map holds no device tree (failed reset?)");
  return Q3DR_OK;
}
using q3dr_detail::check_map;
namespace {

int check_window(uint32_t x0, uint32_t x1, uint32_t y0, uint32_t y1) {
  if (x1 <= x0 || y1 <= y0) return fail(Q3DR_ERR_INVALID_ARG, "empty pixel window");
  if ((uint64_t)(x1 - x0) * (uint64_t)(y1 - y0) >= 0xFFFFFFFFull) return fail(Q3DR_ERR_INVALID_ARG, "window too large");
  return Q3DR_OK;
}

// One scoring call.  The chunks of a call are software-pipelined on ONE stream so that the SMs never wait for the host:
//
//   stream:  R(0) S(0) P(0) c(0) R(1) | S(1) P(1) c(1) R(2) | S(2) P(2) c(2) ...     R = raycast, S = score, P = pack
//   host  :  ............. issue R(1), then wait for c(0) (the chunk's entry count), hand chunk 0 to its consumers ...
//
// The host needs the entry count of chunk k to know whether the CSR pools were large enough and to tell the consumers
// (D2H copy, peer pushes, callback) which range is final; by then R(k+1) is queued, so the wait costs nothing on the
// device, and the consumers' copies overlap R(k+1).  P(k) is issued BEFORE the count is known, into the pools as they
// are; it skips entries beyond their capacity.  If the count then exceeds the capacity (first calls on a map only) the
// pools are grown and P(k) is issued again -- the chunk's temporary lists are intact until S(k+1) is issued.  Pools
// that grow are replaced, never freed inside the call (q3dr_map::retired): pointers handed to a chunk consumer and D2H
// copies in flight stay valid.  Timing events are read once, after the final synchronisation.
int score_device_impl(q3dr_map* m, const float K[4], uint32_t image_width, uint32_t x0, uint32_t x1, uint32_t y0,
                      uint32_t y1, const float* d_Rt, size_t n_vp, const q3dr_score_opts* opts, cudaStream_t stream,
                      q3dr_result* result, bool stream_to_host, const NormalSource* ns = nullptr,
                      const float* h_Rt = nullptr) {
  Q3DR_CUDA(cudaSetDevice(m->device));
  const int nmode = ns ? ns->mode : (int)q3dr::kNormalVoxel;
  if (nmode != q3dr::kNormalVoxel) {
    if (ns->img_h == 0 || x1 > image_width || y1 > ns->img_h) {
      return fail(Q3DR_ERR_INVALID_ARG, "pixel window exceeds the normal image");
    }
    if (nmode == q3dr::kNormalCode && (!ns->mesh || !ns->mesh->d_tris.p || ns->mesh->device != m->device)) {
      return fail(Q3DR_ERR_INVALID_ARG, "mesh is NULL or lives on another device than the map");
    }
    if (nmode == q3dr::kNormalImage && !ns->images) return fail(Q3DR_ERR_INVALID_ARG, "normal_images is NULL");
  }
  m->stream = stream;
  m->host_result_valid = false;
  m->free_retired();  // the previous call is complete and its consumers have been told so (see q3dr_raycast.h)
  const bool want_pix = (m->result_fields & Q3DR_RESULT_FIRST_PIXEL) != 0;
  const bool want_dsq = (m->result_fields & Q3DR_RESULT_DIST_SQ) != 0;
  uint64_t copied_entries = 0;
  q3dr::RaycastParams rp;
  fill_raycast_params(m, K, x0, x1, y0, y1, opts->raycast_max_range, &rp);
  const uint64_t cap = std::min<uint64_t>((uint64_t)rp.w * rp.h, m->n);
  const size_t B = choose_chunk(m, rp.tiles_per_vp, n_vp, cap);
  if ((uint64_t)B * rp.tiles_per_vp >= 0xFFFFFFFFull) return fail(Q3DR_ERR_INVALID_ARG, "chunk too large");
  Q3DR_TRY(ensure_scratch(m, B, cap));
  if (nmode == q3dr::kNormalCode) Q3DR_TRY(m->t_code.ensure(m->slots * cap));
  if (nmode == q3dr::kNormalImage && !ns->images_on_device) {
    Q3DR_TRY(m->d_nimg.ensure(B * (size_t)image_width * ns->img_h));
  }
  const bool tagged = (uint64_t)rp.w * rp.h <= (uint64_t)q3dr::kPixIndexMask + 1u;
  if (m->scratch_dirty || (!tagged && m->fp_tagged_garbage)) {
    // the kernels leave the dedup tables clean themselves; after a failed call they may not have, and an untagged call
    // cannot read past the tagged entries earlier calls left behind
    Q3DR_CUDA(cudaMemsetAsync(m->d_first_pix.p, 0xFF, m->d_first_pix.bytes(), stream));
    Q3DR_CUDA(cudaMemsetAsync(m->d_bitmap.p, 0, m->d_bitmap.bytes(), stream));
    Q3DR_CUDA(cudaMemsetAsync(m->d_seg_u32.p, 0, m->d_seg_u32.bytes(), stream));
    m->scratch_dirty = false;
    m->fp_epoch = 0;
    m->fp_tagged_garbage = false;
  }
  Q3DR_TRY(m->d_offsets.ensure(n_vp + 1));
  Q3DR_TRY(m->d_total.ensure(std::max<size_t>(n_vp, 1)));
  Q3DR_TRY(m->d_hit_count.ensure(std::max<size_t>(n_vp, 1)));
  Q3DR_CUDA(cudaMemsetAsync(m->d_offsets.p, 0, sizeof(uint64_t), stream));
  const size_t n_chunks = (n_vp + B - 1) / B;
  while (m->ev_chunk.size() < 5 * n_chunks) {
    cudaEvent_t e = nullptr;
    Q3DR_CUDA(cudaEventCreate(&e));
    m->ev_chunk.push_back(e);
  }

  const Derived dv = derive(*opts);
  const uint32_t S = m->n_segs;
  uint32_t* seg_hits = m->d_seg_u32.p;
  uint32_t* seg_off = seg_hits + m->slots * S;
  uint32_t* blk_off = m->d_blk_u32.p;
  uint32_t* blk_kept = blk_off + m->slots + 1;
  uint32_t* blk_koff = blk_kept + m->max_blocks;
  const int score_grid = m->sm_count * 3;  // 79 registers x 256 threads: three CTAs per SM
  const int flat_grid = m->sm_count * 8;
  q3dr::ScoreParams2 sp;
  std::memset(&sp, 0, sizeof(sp));
  sp.leaf_rec = m->d_leaf_rec.p;
  sp.fx = K[0];
  sp.fy = K[1];
  sp.cx = K[2];
  sp.cy = K[3];
  sp.image_width_f = (float)image_width;
  sp.x0 = x0;
  sp.y0 = y0;
  sp.w = rp.w;
  sp.thr = dv.thr;
  sp.kf = dv.kf;
  sp.a_thr = dv.a_thr;
  sp.ka = dv.ka;
  sp.ignore_sign = opts->incidence_ignore_dot_product_sign;
  sp.ignore_zero = opts->ignore_voxels_with_zero_information;
  sp.n_segs = S;
  sp.first_pix = m->d_first_pix.p;
  sp.bitmap = m->d_bitmap.p;
  sp.leaf_stride = m->leaf_stride;
  sp.word_stride = m->word_stride;
  sp.cap = cap;
  sp.tmp_leaf = m->t_leaf.p;
  sp.tmp_info = m->t_info.p;
  sp.tmp_pix = m->t_pix.p;
  sp.tmp_dsq = m->t_dsq.p;
  sp.seg_hits = seg_hits;
  sp.seg_off = seg_off;
  sp.blk_off = blk_off;
  sp.blk_kept = blk_kept;
  sp.blk_koff = blk_koff;
  sp.blk_total = m->d_seg_f64.p;
  sp.kept_count = m->d_kept.p;
  sp.tmp_code = m->t_code.p;
  sp.img_w = image_width;
  sp.img_h = ns ? ns->img_h : 0;
  sp.pix_mask = tagged ? q3dr::kPixIndexMask : 0xFFFFFFFFu;
  sp.reset_fp = tagged ? 0 : 1;
  sp.want_dsq = want_dsq ? 1 : 0;
  q3dr::MeshDev mesh_dev{};
  q3dr::RenderOptsDev ropts_dev{};
  if (nmode == q3dr::kNormalCode) {
    mesh_dev.nodes = ns->mesh->d_nodes.p;
    mesh_dev.tris = ns->mesh->d_tris.p;
    mesh_dev.normals = ns->mesh->d_normals.p;
    mesh_dev.n_tris = (uint32_t)ns->mesh->n_tris;
    mesh_dev.root_ref = ns->mesh->root_ref;
    mesh_dev.pad = ns->mesh->pad;
    ropts_dev.near_plane = ns->ropts.near_plane;
    ropts_dev.far_plane = ns->ropts.far_plane;
    ropts_dev.clear_argb = ns->ropts.clear_argb;
  }

  rp.first_pix = m->d_first_pix.p;
  rp.bitmap = m->d_bitmap.p;
  rp.leaf_stride = m->leaf_stride;
  rp.word_stride = m->word_stride;
  rp.seg_hits = seg_hits;
  rp.n_segs = S;

  q3dr_stats st{};
  st.viewpoints_per_chunk = (uint32_t)B;
  Q3DR_CUDA(cudaEventRecord(m->ev[3], stream));

  // Small batches with host-resident poses (the planner scoring one sampled pose at a time): the "camera centre on a
  // box plane" lookup is three binary searches per viewpoint on the host, which saves the flag kernel and the second
  // raycast launch -- a tenth of such a call's latency.  -1: flags come from flag_viewpoints_kernel (the general case).
  int host_flagged = -1;
  if (h_Rt != nullptr && n_vp <= 64 && n_chunks == 1) {
    host_flagged = 0;
    for (size_t v = 0; v < n_vp && !host_flagged; ++v) {
      for (int a = 0; a < 3; ++a) {
        const float o = h_Rt[12 * v + 4 * a + 3];
        const std::vector<float>& pl = m->h_planes[a];
        if (!(o == o) || std::binary_search(pl.begin(), pl.end(), o)) host_flagged = 1;
      }
    }
  }

  auto issue_raycast = [&](size_t c) -> int {
    const size_t v0 = c * B, nb = std::min(B, n_vp - v0);
    if (tagged) {
      if (m->fp_epoch >= q3dr::kPixEpochs - 1u) {  // tags used up (once per 1022 chunks): start over from a clean table
        Q3DR_CUDA(cudaMemsetAsync(m->d_first_pix.p, 0xFF, m->d_first_pix.bytes(), stream));
        m->fp_epoch = 0;
      }
      ++m->fp_epoch;
      rp.pix_tag = (q3dr::kPixEpochs - m->fp_epoch) << q3dr::kPixTagShift;
      rp.pix_limit = rp.pix_tag | q3dr::kPixIndexMask;
      m->fp_tagged_garbage = true;
    } else {
      rp.pix_tag = 0u;
      rp.pix_limit = 0xFFFFFFFEu;
    }
    rp.rt = d_Rt + 12 * v0;
    rp.total_tiles = (uint32_t)(nb * rp.tiles_per_vp);
    rp.vp_flags = m->d_vp_flags.p;
    rp.vp_list = m->d_vp_list.p;
    rp.n_flagged = m->d_counter.p + 2;
    Q3DR_CUDA(cudaMemsetAsync(m->d_counter.p, 0, 3 * sizeof(uint32_t), stream));
    Q3DR_CUDA(cudaEventRecord(m->ev_chunk[5 * c + 0], stream));
    m->scratch_dirty = true;  // cleared when this chunk's scoring kernels have been issued
    const int wpc0 = variant_of(m).warps;
    const int grid0 = (int)std::min<uint64_t>((uint64_t)m->grid2_ctas[q3dr::kModeScore], ((uint64_t)rp.total_tiles + wpc0 - 1) / wpc0);
    const size_t dyn0 = variant_of(m).smem_top ? m->smem_bytes : 0;
    if (host_flagged >= 0) {
      // one launch: the fast instantiation if no centre lies on a box plane, else the checking one for all viewpoints
      rp.vp_flags = nullptr;
      rp.vp_list = nullptr;
      if (host_flagged) {
        variant_of(m).kernel[q3dr::kModeScore]<<<grid0, wpc0 * 32, dyn0, stream>>>(rp);
      } else {
        variant_of(m).score_fast<<<grid0, wpc0 * 32, dyn0, stream>>>(rp);
      }
      Q3DR_CUDA(cudaGetLastError());
      Q3DR_CUDA(cudaEventRecord(m->ev_chunk[5 * c + 1], stream));
      st.kernel_launches += 1;
      return Q3DR_OK;
    }
    // viewpoints whose centre lies on a box plane go to the checking instantiation, everybody else to the fast one
    q3dr::flag_viewpoints_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, stream>>>(
        rp.rt, (uint32_t)nb, m->d_planes.p, m->d_planes.p + m->n_planes[0], m->d_planes.p + m->n_planes[0] + m->n_planes[1],
        m->n_planes[0], m->n_planes[1], m->n_planes[2], m->d_vp_flags.p, m->d_vp_list.p, m->d_counter.p + 2);
    Q3DR_CUDA(cudaGetLastError());
    const int wpc = variant_of(m).warps;
    const int grid = (int)std::min<uint64_t>((uint64_t)m->grid2_ctas[q3dr::kModeScore], ((uint64_t)rp.total_tiles + wpc - 1) / wpc);
    const size_t dyn_smem = variant_of(m).smem_top ? m->smem_bytes : 0;
    variant_of(m).score_fast<<<grid, wpc * 32, dyn_smem, stream>>>(rp);
    Q3DR_CUDA(cudaGetLastError());
    variant_of(m).kernel[q3dr::kModeScore]<<<grid, wpc * 32, dyn_smem, stream>>>(rp);  // exits at once if none is flagged
    Q3DR_CUDA(cudaGetLastError());
    Q3DR_CUDA(cudaEventRecord(m->ev_chunk[5 * c + 1], stream));
    st.kernel_launches += 3;
    return Q3DR_OK;
  };

  auto issue_score = [&](size_t c) -> int {
    const size_t v0 = c * B, nb = std::min(B, n_vp - v0);
    sp.rt = d_Rt + 12 * v0;
    sp.hit_count = m->d_hit_count.p + v0;
    sp.total = m->d_total.p + v0;
    q3dr::seg_scan_kernel<<<(unsigned)nb, q3dr::kSegThreads, 0, stream>>>(seg_hits, seg_off, S, sp.hit_count);
    Q3DR_CUDA(cudaGetLastError());
    q3dr::emit_list_kernel<<<dim3(S, (unsigned)nb), q3dr::kSegThreads, 0, stream>>>(sp);
    Q3DR_CUDA(cudaGetLastError());
    q3dr::block_offsets_kernel<<<1, q3dr::kSegThreads, 0, stream>>>(sp.hit_count, (uint32_t)nb, blk_off);
    Q3DR_CUDA(cudaGetLastError());
    if (nmode == q3dr::kNormalVoxel) {
      q3dr::score_blocks_kernel<q3dr::kNormalVoxel><<<score_grid, q3dr::kSegThreads, 0, stream>>>(sp, (uint32_t)nb);
    } else if (nmode == q3dr::kNormalCode) {
      q3dr::mesh_kept_normals_kernel<<<flat_grid, q3dr::kSegThreads, 0, stream>>>(sp, (uint32_t)nb, mesh_dev, ropts_dev);
      Q3DR_CUDA(cudaGetLastError());
      q3dr::score_blocks_kernel<q3dr::kNormalCode><<<score_grid, q3dr::kSegThreads, 0, stream>>>(sp, (uint32_t)nb);
      st.kernel_launches += 1;
    } else {
      const size_t img_px = (size_t)image_width * ns->img_h;
      if (ns->images_on_device) {
        sp.normal_images = ns->images + v0 * img_px;
      } else {
        Q3DR_CUDA(cudaMemcpyAsync(m->d_nimg.p, ns->images + v0 * img_px, nb * img_px * sizeof(uint32_t),
                                  cudaMemcpyHostToDevice, stream));
        sp.normal_images = m->d_nimg.p;
      }
      q3dr::score_blocks_kernel<q3dr::kNormalImage><<<score_grid, q3dr::kSegThreads, 0, stream>>>(sp, (uint32_t)nb);
    }
    Q3DR_CUDA(cudaGetLastError());
    q3dr::finalize_viewpoints_kernel<<<(unsigned)nb, q3dr::kSegThreads, 0, stream>>>(sp);
    Q3DR_CUDA(cudaGetLastError());
    m->scratch_dirty = false;  // emit_list cleared the bitmap and the segment counters; first_pix is tagged or was reset
    q3dr::csr_offsets_kernel<<<1, q3dr::kSegThreads, 0, stream>>>(m->d_kept.p, (uint32_t)nb, m->d_offsets.p + v0);
    Q3DR_CUDA(cudaGetLastError());
    Q3DR_CUDA(cudaEventRecord(m->ev_chunk[5 * c + 2], stream));
    Q3DR_CUDA(cudaMemcpyAsync(m->h_scalar.p, m->d_offsets.p + v0 + nb, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
    Q3DR_CUDA(cudaEventRecord(m->ev_count, stream));
    st.kernel_launches += 6;
    return Q3DR_OK;
  };

  std::vector<uint8_t> packed(n_chunks, 0);
  auto issue_pack = [&](size_t c) -> int {
    const size_t v0 = c * B, nb = std::min(B, n_vp - v0);
    packed[c] = 1;
    q3dr::PackParams2 pp;
    pp.offsets = m->d_offsets.p + v0;
    pp.blk_off = blk_off;
    pp.blk_koff = blk_koff;
    pp.hit_count = m->d_hit_count.p + v0;
    pp.n_vp = (uint32_t)nb;
    pp.ignore_zero = sp.ignore_zero;
    pp.cap = cap;
    pp.tmp_leaf = m->t_leaf.p;
    pp.tmp_info = m->t_info.p;
    pp.tmp_pix = m->t_pix.p;
    pp.tmp_dsq = m->t_dsq.p;
    pp.out_leaf = m->r_leaf.p;
    pp.out_info = m->r_info.p;
    pp.out_pix = want_pix ? m->r_pix.p : nullptr;
    pp.out_dsq = want_dsq ? m->r_dsq.p : nullptr;
    pp.pool_cap = m->r_leaf.cap;
    Q3DR_CUDA(cudaEventRecord(m->ev_chunk[5 * c + 3], stream));
    q3dr::pack_blocks_kernel<<<flat_grid, q3dr::kSegThreads, 0, stream>>>(pp);
    Q3DR_CUDA(cudaGetLastError());
    Q3DR_CUDA(cudaEventRecord(m->ev_chunk[5 * c + 4], stream));
    Q3DR_CUDA(cudaEventRecord(m->ev_packed, stream));
    st.kernel_launches += 1;
    return Q3DR_OK;
  };

  uint64_t total_entries = 0;
  // the optional pools always match the (leaf, information) pools
  if (want_pix && m->r_pix.cap < m->r_leaf.cap) Q3DR_TRY(m->r_pix.grow_deferred(m->r_leaf.cap, 0, stream, &m->retired));
  if (want_dsq && m->r_dsq.cap < m->r_leaf.cap) Q3DR_TRY(m->r_dsq.grow_deferred(m->r_leaf.cap, 0, stream, &m->retired));
  if (n_chunks > 0) Q3DR_TRY(issue_raycast(0));
  for (size_t c = 0; c < n_chunks; ++c) {
    const size_t v0 = c * B, nb = std::min(B, n_vp - v0);
    Q3DR_TRY(issue_score(c));
    // packed right away into the pools as they are (entries beyond their capacity are skipped, see below): in the steady
    // state the chunk's result is final BEFORE the next raycast starts, so its D2H / peer pushes overlap that raycast
    const bool packed_early = m->r_leaf.cap > 0;
    if (packed_early) Q3DR_TRY(issue_pack(c));
    if (c + 1 < n_chunks) Q3DR_TRY(issue_raycast(c + 1));  // keeps the SMs busy while the host waits below
    Q3DR_CUDA(cudaEventSynchronize(m->ev_count));
    const uint64_t chunk_first_entry = total_entries;
    total_entries = m->h_scalar.p[0];
    if (total_entries > m->r_leaf.cap || !packed_early) {
      if (total_entries > m->r_leaf.cap) {
        // project the call's total from what the viewpoints so far produced, so that a call grows its pools about once
        const double per_vp = (double)total_entries / (double)(v0 + nb);
        const size_t want = (size_t)std::max<double>((double)total_entries, per_vp * (double)n_vp * 1.15 + 4096.0);
        Q3DR_TRY(m->r_leaf.grow_deferred(want, chunk_first_entry, stream, &m->retired));
        Q3DR_TRY(m->r_info.grow_deferred(want, chunk_first_entry, stream, &m->retired));
        if (want_pix) Q3DR_TRY(m->r_pix.grow_deferred(want, chunk_first_entry, stream, &m->retired));
        if (want_dsq) Q3DR_TRY(m->r_dsq.grow_deferred(want, chunk_first_entry, stream, &m->retired));
      }
      // (first calls only) the chunk's temporary lists are still intact -- S(c + 1) has not been issued: pack again
      if (total_entries > chunk_first_entry) Q3DR_TRY(issue_pack(c));
    }
    st.chunks += 1;
    if (m->chunk_cb) {
      q3dr_result part;
      part.n_viewpoints = v0 + nb;
      part.n_entries = total_entries;
      part.offsets = m->d_offsets.p;
      part.leaf = m->r_leaf.p;
      part.information = m->r_info.p;
      part.first_pixel = want_pix ? m->r_pix.p : nullptr;
      part.dist_sq = want_dsq ? m->r_dsq.p : nullptr;
      part.total = m->d_total.p;
      part.hit_count = m->d_hit_count.p;
      m->chunk_cb(m->chunk_cb_user, v0, nb, chunk_first_entry, total_entries - chunk_first_entry, &part);
    }
    if (m->exchange) {
      Q3DR_TRY(q3dr_detail::exchange_push_chunk(m->exchange, m->ev_packed, m->r_leaf.p, m->r_info.p, chunk_first_entry,
                                                total_entries - chunk_first_entry));
    }
    if (stream_to_host && total_entries > copied_entries) {
      // this chunk's slice of the CSR pool is final: ship it while the next chunk is being cast
      if (total_entries > m->h_leaf.cap) {
        Q3DR_CUDA(cudaStreamSynchronize(m->copy_stream));
        const double per_vp = (double)total_entries / (double)(v0 + nb);
        const size_t want = (size_t)std::max<double>((double)total_entries, per_vp * (double)n_vp * 1.15 + 4096.0);
        adopt_pooled_host_pair(m, want);
        Q3DR_TRY(m->h_leaf.grow_keep(want, copied_entries));
        Q3DR_TRY(m->h_info.grow_keep(want, copied_entries));
      }
      if (want_pix) Q3DR_TRY(m->h_pix.grow_keep(m->h_leaf.cap, copied_entries));
      if (want_dsq) Q3DR_TRY(m->h_dsq.grow_keep(m->h_leaf.cap, copied_entries));
      const size_t c0 = copied_entries, cn = total_entries - copied_entries;
      Q3DR_CUDA(cudaStreamWaitEvent(m->copy_stream, m->ev_packed, 0));
      Q3DR_CUDA(cudaMemcpyAsync(m->h_leaf.p + c0, m->r_leaf.p + c0, cn * sizeof(int32_t), cudaMemcpyDeviceToHost, m->copy_stream));
      Q3DR_CUDA(cudaMemcpyAsync(m->h_info.p + c0, m->r_info.p + c0, cn * sizeof(float), cudaMemcpyDeviceToHost, m->copy_stream));
      if (want_pix) {
        Q3DR_CUDA(cudaMemcpyAsync(m->h_pix.p + c0, m->r_pix.p + c0, cn * sizeof(uint32_t), cudaMemcpyDeviceToHost, m->copy_stream));
      }
      if (want_dsq) {
        Q3DR_CUDA(cudaMemcpyAsync(m->h_dsq.p + c0, m->r_dsq.p + c0, cn * sizeof(float), cudaMemcpyDeviceToHost, m->copy_stream));
      }
      copied_entries = total_entries;
    }
  }
  Q3DR_CUDA(cudaEventRecord(m->ev[2], stream));
  Q3DR_CUDA(cudaStreamSynchronize(stream));
  if (stream_to_host) {
    Q3DR_CUDA(cudaStreamSynchronize(m->copy_stream));
    m->host_result_valid = true;
  }
  for (size_t c = 0; c < n_chunks; ++c) {
    st.raycast_ms += elapsed_ms(m->ev_chunk[5 * c + 0], m->ev_chunk[5 * c + 1]);
    st.score_ms += elapsed_ms(m->ev_chunk[5 * c + 1], m->ev_chunk[5 * c + 2]);
    if (packed[c]) st.score_ms += elapsed_ms(m->ev_chunk[5 * c + 3], m->ev_chunk[5 * c + 4]);
  }
  // S(c) is bracketed by the end of R(c) and its own end event; R(c+1) is issued after that event, so the brackets
  // do not overlap
  if (n_vp > 0) st.total_ms = elapsed_ms(m->ev[3], m->ev[2]);
  st.rays = (uint64_t)n_vp * rp.w * rp.h;
  m->stats = st;
  m->last_n_vp = n_vp;
  m->last_n_entries = total_entries;
  m->stream = nullptr;  // everything is complete: follow-up calls (q3dr_result_to_host, set consumers, overlap queries)
                        // run on the default stream and never touch the caller's stream handle again
  if (result) {
    result->n_viewpoints = n_vp;
    result->n_entries = total_entries;
    result->offsets = m->d_offsets.p;
    result->leaf = m->r_leaf.p;
    result->information = m->r_info.p;
    result->first_pixel = want_pix ? m->r_pix.p : nullptr;
    result->dist_sq = want_dsq ? m->r_dsq.p : nullptr;
    result->total = m->d_total.p;
    result->hit_count = m->d_hit_count.p;
  }
  return Q3DR_OK;
}

}  // namespace

extern "C" {

int q3dr_abi_version(void) { return Q3DR_ABI_VERSION; }

const char* q3dr_status_string(int status) {
  switch (status) {
    case Q3DR_OK: return "ok";
    case Q3DR_ERR_INVALID_ARG: return "invalid argument";
    case Q3DR_ERR_CUDA: return "CUDA error";
    case Q3DR_ERR_NO_DEVICE: return "no usable CUDA device (the engine has no CPU path)";
    case Q3DR_ERR_OUT_OF_MEMORY: return "out of memory";
    case Q3DR_ERR_INTERNAL: return "internal error";
    default: return "unknown status";
  }
}

const char* q3dr_last_error(void) { return q3dr_detail::last_error().c_str(); }

int q3dr_device_count(int* count) {
  if (!count) return fail(Q3DR_ERR_INVALID_ARG, "count is NULL");
  *count = 0;
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(Q3DR_ERR_NO_DEVICE, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
  }
  *count = c;
  return Q3DR_OK;
}

void q3dr_score_opts_default(q3dr_score_opts* o) {
  if (!o) return;
  o->voxel_sensor_size_ratio_threshold = 0.002f;
  o->voxel_sensor_size_ratio_inv_falloff_factor = 0.001f;
  o->incidence_angle_threshold_degrees = 30.0f;
  o->incidence_angle_inv_falloff_factor_degrees = 30.0f;
  o->incidence_ignore_dot_product_sign = 0;
  o->ignore_voxels_with_zero_information = 1;
  o->raycast_min_range = 0.0f;
  o->raycast_max_range = 60.0f;
}

int q3dr_splitmedian_order(const float* boxes, size_t n, uint32_t* order, uint32_t* depth) {
  if (!boxes || !order) return fail(Q3DR_ERR_INVALID_ARG, "boxes/order is NULL");
  if (n == 0 || n >= 0x7FFFFFFFull) return fail(Q3DR_ERR_INVALID_ARG, "leaf count out of range");
  try {
    q3dr::splitmedian_order(boxes, n, order, depth);
  } catch (const std::bad_alloc&) {
    return fail(Q3DR_ERR_OUT_OF_MEMORY, "host allocation failed");
  }
  return Q3DR_OK;
}

int q3dr_voxel_indices(size_t n, uint32_t* out) {
  if (!out) return fail(Q3DR_ERR_INVALID_ARG, "out is NULL");
  if (n == 0 || n >= 0x7FFFFFFFull) return fail(Q3DR_ERR_INVALID_ARG, "leaf count out of range");
  try {
    q3dr::splitmedian_voxel_indices(n, out);
  } catch (const std::bad_alloc&) {
    return fail(Q3DR_ERR_OUT_OF_MEMORY, "host allocation failed");
  }
  return Q3DR_OK;
}

// Eigen's QuaternionBase::toRotationMatrix in binary32, as Pose::getTransformationImageToWorld uses it.
int q3dr_pose_to_rt(const float q[4], const float trans[3], float rt[12]) {
  if (!q || !trans || !rt) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  const float w = q[0], x = q[1], y = q[2], z = q[3];
  const float tx = 2.0f * x, ty = 2.0f * y, tz = 2.0f * z;
  const float twx = tx * w, twy = ty * w, twz = tz * w;
  const float txx = tx * x, txy = ty * x, txz = tz * x;
  const float tyy = ty * y, tyz = tz * y, tzz = tz * z;
  rt[0] = 1.0f - (tyy + tzz);
  rt[1] = txy - twz;
  rt[2] = txz + twy;
  rt[3] = trans[0];
  rt[4] = txy + twz;
  rt[5] = 1.0f - (txx + tzz);
  rt[6] = tyz - twx;
  rt[7] = trans[1];
  rt[8] = txz - twy;
  rt[9] = tyz + twx;
  rt[10] = 1.0f - (txx + tyy);
  rt[11] = trans[2];
  return Q3DR_OK;
}

// Host-only diagnostic: flattens `boxes` exactly like q3dr_map_create and checks the invariants the
// traversal relies on (every leaf referenced once; every child box contains the boxes of all leaves
// below it, bit for bit as a union).  No device needed.
int q3dr_debug_flat_tree(const float* boxes, size_t n, int top_levels, uint64_t* n_nodes, uint32_t* depth,
                         uint32_t* max_stack, uint32_t* top_nodes) {
  if (!boxes || n == 0) return fail(Q3DR_ERR_INVALID_ARG, "boxes is NULL or empty");
  q3dr::FlatTree t;
  try {
    q3dr::build_flat_tree(boxes, n, top_levels, &t);
  } catch (const std::bad_alloc&) {
    return fail(Q3DR_ERR_OUT_OF_MEMORY, "host allocation failed");
  }
  if (n_nodes) *n_nodes = t.nodes.size();
  if (depth) *depth = t.depth;
  if (max_stack) *max_stack = t.max_stack;
  if (top_nodes) *top_nodes = t.top_nodes;
  // bottom-up check: nodes are laid out parents-before-children, so walk the array backwards
  const size_t nn = t.nodes.size();
  std::vector<float> nb(6 * nn);
  std::vector<uint8_t> seen(n, 0);
  std::vector<uint8_t> referenced(nn, 0);
  for (size_t i = nn; i-- > 0;) {
    const q3dr::Node4& nd = t.nodes[i];
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int s = 0; s < 4; ++s) {
      uint32_t c = nd.child[s];
      if (s == 0) {
        uint32_t lm = 0;
        for (int q = 0; q < 4; ++q) lm |= ((nd.child[q] >> 31) & 1u) << q;
        if (((c >> q3dr::kLeafMaskShift) & 0xFu) != lm) return fail(Q3DR_ERR_INTERNAL, "leaf mask mismatch");
      }
      c &= q3dr::kLeafFlag | q3dr::kPayloadMask;  // slot 0 carries the leaf mask, slot 1 the order axis
      if (c == q3dr::kEmptyChild) continue;
      const float* ref;
      if (c & q3dr::kLeafFlag) {
        const uint32_t li = c & q3dr::kPayloadMask;
        if (li >= n || seen[li]) return fail(Q3DR_ERR_INTERNAL, "leaf referenced twice or out of range");
        seen[li] = 1;
        ref = boxes + 6 * (size_t)li;
      } else {
        const uint32_t ci = c & q3dr::kNodeIndexMask;
        if (ci <= i || ci >= nn || referenced[ci]) return fail(Q3DR_ERR_INTERNAL, "bad inner child reference");
        referenced[ci] = 1;
        ref = nb.data() + 6 * (size_t)ci;
      }
      for (int a = 0; a < 3; ++a) {
        if (nd.mn[a][s] != ref[a] || nd.mx[a][s] != ref[3 + a]) return fail(Q3DR_ERR_INTERNAL, "child box mismatch");
        mn[a] = std::min(mn[a], ref[a]);
        mx[a] = std::max(mx[a], ref[3 + a]);
      }
    }
    for (int a = 0; a < 3; ++a) {
      nb[6 * i + a] = mn[a];
      nb[6 * i + 3 + a] = mx[a];
    }
  }
  for (size_t k = 0; k < n; ++k) {
    if (!seen[k]) return fail(Q3DR_ERR_INTERNAL, "leaf missing from the flattened tree");
  }
  for (size_t i = 1; i < nn; ++i) {
    if (!referenced[i]) return fail(Q3DR_ERR_INTERNAL, "unreferenced node");
  }
  return Q3DR_OK;
}

}  // extern "C"

namespace q3dr_detail {
int map_create_impl(const float* boxes, const float* weights, const float* normals, const uint32_t* depth, size_t n,
                    int device, std::shared_ptr<const q3dr::FlatTree> shared_tree, q3dr_map** out) {
  if (!out) return fail(Q3DR_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  if (!boxes) return fail(Q3DR_ERR_INVALID_ARG, "boxes is NULL");
  if (n == 0 || n >= (size_t)q3dr::kPayloadMask) return fail(Q3DR_ERR_INVALID_ARG, "leaf count out of range (< 2^27 - 1)");
  int count = 0;
  Q3DR_TRY(q3dr_device_count(&count));
  if (count <= 0) return fail(Q3DR_ERR_NO_DEVICE, "no CUDA device visible; this engine has no CPU path");
  if (device < 0 || device >= count) return fail(Q3DR_ERR_INVALID_ARG, "device index out of range");
  q3dr_map* m = new (std::nothrow) q3dr_map;
  if (!m) return fail(Q3DR_ERR_OUT_OF_MEMORY, "host allocation failed");
  int st = Q3DR_OK;
  try {
    m->device = device;
    m->n = n;
    m->h_boxes.assign(boxes, boxes + 6 * n);
    m->bounded = true;
    for (size_t i = 0; i < 6 * n; ++i) {
      if (!(std::fabs(boxes[i]) < 1e18f)) m->bounded = false;
    }
    for (int a = 0; a < 3; ++a) {
      m->root_box[a] = FLT_MAX;
      m->root_box[3 + a] = -FLT_MAX;
    }
    for (size_t k = 0; k < n; ++k) {
      for (int a = 0; a < 3; ++a) {
        m->root_box[a] = std::min(m->root_box[a], boxes[6 * k + a]);
        m->root_box[3 + a] = std::max(m->root_box[3 + a], boxes[6 * k + 3 + a]);
      }
    }
    if (weights) {
      m->h_weight.assign(weights, weights + n);
    } else {
      m->h_weight.assign(n, 1.0f);
    }
    if (normals) {
      m->h_normal.assign(normals, normals + 3 * n);
    } else {
      m->h_normal.assign(3 * n, 0.0f);
    }
    m->h_depth.resize(n);
    uint32_t ref_depth = 0;
    for (size_t k = 0; k < n; ++k) {
      const uint32_t d = depth ? depth[k] : q3dr::splitmedian_leaf_depth(n, k);
      m->h_depth[k] = (uint8_t)std::min<uint32_t>(d, 255u);
      ref_depth = std::max(ref_depth, d);
    }
    m->ref_depth = ref_depth;
    if (shared_tree) {
      m->tree = shared_tree;
    } else {
      int top_levels = 4;
      if (const char* e = std::getenv("Q3DR_TOP_LEVELS")) top_levels = std::max(1, std::min(6, atoi(e)));
      std::shared_ptr<q3dr::FlatTree> t = std::make_shared<q3dr::FlatTree>();
      q3dr::build_flat_tree(m->h_boxes.data(), n, top_levels, t.get());
      m->tree = t;
    }
    if (m->tree->max_stack > (uint32_t)q3dr::kStackEntries) {
      st = fail(Q3DR_ERR_INTERNAL, "flattened tree too deep for the traversal stack");
    }
    if (m->tree->nodes.size() >= (size_t)q3dr::kNodeIndexMask) st = fail(Q3DR_ERR_INVALID_ARG, "too many nodes");
  } catch (const std::bad_alloc&) {
    st = fail(Q3DR_ERR_OUT_OF_MEMORY, "host allocation failed");
  }
  if (st == Q3DR_OK) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
      st = fail(Q3DR_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
    } else if (prop.major < 10) {
      st = fail(Q3DR_ERR_NO_DEVICE, "device is not sm_100-class; this library carries sm_100a code only");
    } else {
      m->sm_count = prop.multiProcessorCount;
    }
  }
  if (st == Q3DR_OK) {
    cudaError_t e = cudaSetDevice(device);
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&m->ev[i]);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_packed, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_count, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) st = fail(Q3DR_ERR_CUDA, std::string("event setup: ") + cudaGetErrorString(e));
  }
  if (st == Q3DR_OK) st = upload_map(m);
  if (st != Q3DR_OK) {
    q3dr_map_destroy(m);
    return st;
  }
  *out = m;
  return Q3DR_OK;
}
}  // namespace q3dr_detail
using q3dr_detail::map_create_impl;

extern "C" {

int q3dr_map_create(const float* boxes, const float* weights, const float* normals, const uint32_t* depth, size_t n,
                    int device, q3dr_map** out) {
  return map_create_impl(boxes, weights, normals, depth, n, device, nullptr, out);
}

int q3dr_map_update_weights(q3dr_map* m, const float* weights) {
  Q3DR_TRY(check_map(m));
  if (!weights) return fail(Q3DR_ERR_INVALID_ARG, "weights is NULL");
  Q3DR_CUDA(cudaSetDevice(m->device));
  m->h_weight.assign(weights, weights + m->n);
  Q3DR_CUDA(cudaMemcpy2D(reinterpret_cast<char*>(m->d_leaf_rec.p) + 12, q3dr::kLeafRecStride * sizeof(float4), weights, sizeof(float),
                         sizeof(float), m->n, cudaMemcpyHostToDevice));
  return Q3DR_OK;
}

int q3dr_map_update_normals(q3dr_map* m, const float* normals) {
  Q3DR_TRY(check_map(m));
  if (!normals) return fail(Q3DR_ERR_INVALID_ARG, "normals is NULL");
  Q3DR_CUDA(cudaSetDevice(m->device));
  m->h_normal.assign(normals, normals + 3 * m->n);
  Q3DR_CUDA(cudaMemcpy2D(reinterpret_cast<char*>(m->d_leaf_rec.p) + 32, q3dr::kLeafRecStride * sizeof(float4), normals, 3 * sizeof(float),
                         3 * sizeof(float), m->n, cudaMemcpyHostToDevice));
  return Q3DR_OK;
}

int q3dr_map_get_info(const q3dr_map* m, q3dr_map_info* info) {
  if (!m || !info) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  info->n_leaves = m->n;
  info->n_nodes = m->tree->nodes.size();
  info->tree_depth = m->tree->depth;
  info->ref_depth = m->ref_depth;
  info->node_bytes = m->tree->nodes.size() * sizeof(q3dr::Node4);
  info->device_bytes = m->device_bytes();
  info->device = m->device;
  info->sm_count = m->sm_count;
  return Q3DR_OK;
}

int q3dr_map_check_guards(q3dr_map* m, int* overwritten) {
  if (!m || !overwritten) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  *overwritten = 0;
  if (!guards_enabled()) return fail(Q3DR_ERR_INVALID_ARG, "guards are off: set Q3DR_GUARD=1 before the library allocates");
  Q3DR_CUDA(cudaSetDevice(m->device));
  Q3DR_CUDA(cudaDeviceSynchronize());
  const int bad = m->overwritten_guards();
  if (bad < 0) return fail(Q3DR_ERR_CUDA, "guard words could not be read back");
  *overwritten = bad;
  return Q3DR_OK;
}

int q3dr_map_last_stats(const q3dr_map* m, q3dr_stats* stats) {
  if (!m || !stats) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  *stats = m->stats;
  return Q3DR_OK;
}

int q3dr_map_reset(q3dr_map* m) {
  if (!m) return fail(Q3DR_ERR_INVALID_ARG, "map is NULL");
  cudaSetDevice(m->device);
  cudaDeviceSynchronize();
  cudaGetLastError();
  release_device(m);
  return upload_map(m);
}

int q3dr_map_destroy(q3dr_map* m) {
  if (!m) return Q3DR_OK;
  cudaSetDevice(m->device);
  release_device(m);
  m->h_scalar.release();
  m->h_offsets.release();
  m->h_total.release();
  m->h_info.release();
  m->h_dsq.release();
  m->h_hit_count.release();
  m->h_pix.release();
  m->h_leaf.release();
  m->h_rt.release();
  for (int i = 0; i < 4; ++i) {
    if (m->ev[i]) cudaEventDestroy(m->ev[i]);
  }
  if (m->ev_packed) cudaEventDestroy(m->ev_packed);
  if (m->ev_count) cudaEventDestroy(m->ev_count);
  for (cudaEvent_t e : m->ev_chunk) cudaEventDestroy(e);
  m->free_retired();
  if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
  if (m->own_stream) cudaStreamDestroy(m->own_stream);
  cudaGetLastError();
  delete m;
  return Q3DR_OK;
}

}  // extern "C"

// raycast2_kernel<kModePixels> of one viewpoint whose pose is already in device memory: per-pixel records into
// m->d_pixels, on `s`, no synchronisation (q3dr_queries.cu: real-image down-weighting)
int q3dr_detail::raycast_window_device(q3dr_map* m, const float K[4], const float* d_rt, uint32_t x0, uint32_t x1, uint32_t y0,
                                       uint32_t y1, float max_range, cudaStream_t s) {
  q3dr::RaycastParams rp;
  fill_raycast_params(m, K, x0, x1, y0, y1, max_range, &rp);
  const size_t npix = (size_t)rp.w * rp.h;
  Q3DR_TRY(m->d_pixels.ensure(npix));
  Q3DR_CUDA(cudaMemsetAsync(m->d_counter.p, 0, sizeof(uint32_t), s));
  rp.rt = d_rt;
  rp.total_tiles = rp.tiles_per_vp;
  rp.pixels = m->d_pixels.p;
  const int wpc = variant_of(m).warps;
  const int grid = (int)std::min<uint64_t>((uint64_t)m->grid2_ctas[q3dr::kModePixels], ((uint64_t)rp.total_tiles + wpc - 1) / wpc);
  variant_of(m).kernel[q3dr::kModePixels]<<<grid, wpc * 32, variant_of(m).smem_top ? m->smem_bytes : 0, s>>>(rp);
  Q3DR_CUDA(cudaGetLastError());
  return Q3DR_OK;
}

extern "C" {

int q3dr_raycast_window(q3dr_map* m, const float K[4], const float Rt[12], uint32_t x0, uint32_t x1, uint32_t y0,
                        uint32_t y1, float /*min_range*/, float max_range, q3dr_pixel_hit* out) {
  Q3DR_TRY(check_map(m));
  if (!K || !Rt || !out) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  Q3DR_TRY(check_window(x0, x1, y0, y1));
  Q3DR_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = nullptr;
  q3dr::RaycastParams rp;
  fill_raycast_params(m, K, x0, x1, y0, y1, max_range, &rp);
  const size_t npix = (size_t)rp.w * rp.h;
  Q3DR_TRY(m->d_rt.ensure(12));
  Q3DR_TRY(m->d_pixels.ensure(npix));
  Q3DR_CUDA(cudaMemcpyAsync(m->d_rt.p, Rt, 12 * sizeof(float), cudaMemcpyHostToDevice, s));
  Q3DR_CUDA(cudaMemsetAsync(m->d_counter.p, 0, sizeof(uint32_t), s));
  rp.rt = m->d_rt.p;
  rp.total_tiles = rp.tiles_per_vp;
  rp.pixels = m->d_pixels.p;
  const int wpc = variant_of(m).warps;
  const int grid = (int)std::min<uint64_t>((uint64_t)m->grid2_ctas[q3dr::kModePixels], ((uint64_t)rp.total_tiles + wpc - 1) / wpc);
  Q3DR_CUDA(cudaEventRecord(m->ev[0], s));
  variant_of(m).kernel[q3dr::kModePixels]<<<grid, wpc * 32, variant_of(m).smem_top ? m->smem_bytes : 0, s>>>(rp);
  Q3DR_CUDA(cudaGetLastError());
  Q3DR_CUDA(cudaEventRecord(m->ev[1], s));
  Q3DR_CUDA(cudaMemcpyAsync(out, m->d_pixels.p, npix * sizeof(q3dr_pixel_hit), cudaMemcpyDeviceToHost, s));
  Q3DR_CUDA(cudaStreamSynchronize(s));
  q3dr_stats st{};
  st.rays = npix;
  st.kernel_launches = 1;
  st.raycast_ms = elapsed_ms(m->ev[0], m->ev[1]);
  st.total_ms = st.raycast_ms;
  st.chunks = 1;
  st.viewpoints_per_chunk = 1;
  m->stats = st;
  return Q3DR_OK;
}

int q3dr_intersect_rays(q3dr_map* m, const float* origins, const float* directions, size_t n, float /*min_range*/,
                        float max_range, q3dr_pixel_hit* out) {
  Q3DR_TRY(check_map(m));
  if (!origins || !directions || !out) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  if (n == 0) return Q3DR_OK;
  if (n >= 0xFFFFFFFFull) return fail(Q3DR_ERR_INVALID_ARG, "too many rays for one call");
  Q3DR_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = nullptr;
  Q3DR_TRY(m->d_ray_o.ensure(3 * n));
  Q3DR_TRY(m->d_ray_d.ensure(3 * n));
  Q3DR_TRY(m->d_pixels.ensure(n));
  Q3DR_CUDA(cudaMemcpyAsync(m->d_ray_o.p, origins, 3 * n * sizeof(float), cudaMemcpyHostToDevice, s));
  Q3DR_CUDA(cudaMemcpyAsync(m->d_ray_d.p, directions, 3 * n * sizeof(float), cudaMemcpyHostToDevice, s));
  Q3DR_CUDA(cudaMemsetAsync(m->d_counter.p, 0, sizeof(uint32_t), s));
  q3dr::RaycastParams rp;
  std::memset(&rp, 0, sizeof(rp));
  rp.nodes = m->d_nodes.p;
  rp.n_top = m->tree->top_nodes;
  rp.max_range = max_range;
  rp.leaf_depth = m->d_depth.p;
  rp.tile_counter = m->d_counter.p;
  rp.ray_o = m->d_ray_o.p;
  rp.ray_d = m->d_ray_d.p;
  rp.n_rays = (uint32_t)n;
  rp.total_tiles = (uint32_t)((n + 31) / 32);
  rp.pixels = m->d_pixels.p;
  const int grid = (int)std::min<uint64_t>((uint64_t)m->grid_rays_ctas,
                                          ((uint64_t)rp.total_tiles + q3dr::kWarpsPerCta - 1) / q3dr::kWarpsPerCta);
  q3dr::raycast_rays_kernel<<<grid, q3dr::kThreadsPerCta, m->smem_bytes, s>>>(rp);
  Q3DR_CUDA(cudaGetLastError());
  Q3DR_CUDA(cudaMemcpyAsync(out, m->d_pixels.p, n * sizeof(q3dr_pixel_hit), cudaMemcpyDeviceToHost, s));
  Q3DR_CUDA(cudaStreamSynchronize(s));
  return Q3DR_OK;
}

int q3dr_score_viewpoints_device(q3dr_map* m, const float K[4], uint32_t image_width, uint32_t x0, uint32_t x1,
                                 uint32_t y0, uint32_t y1, const float* d_Rt, size_t n, const q3dr_score_opts* opts,
                                 void* stream, q3dr_result* result) {
  Q3DR_TRY(check_map(m));
  if (!K || (!d_Rt && n) || !opts) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  Q3DR_TRY(check_window(x0, x1, y0, y1));
  if (image_width == 0) return fail(Q3DR_ERR_INVALID_ARG, "image_width is 0");
  if ((reinterpret_cast<uintptr_t>(d_Rt) & 15u) != 0 && n) {
    // the kernel fetches a pose as three 16-byte words: realign a misaligned caller buffer
    Q3DR_CUDA(cudaSetDevice(m->device));
    Q3DR_TRY(m->d_rt.ensure(12 * n));
    Q3DR_CUDA(cudaMemcpyAsync(m->d_rt.p, d_Rt, 12 * n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    d_Rt = m->d_rt.p;
  }
  return score_device_impl(m, K, image_width, x0, x1, y0, y1, d_Rt, n, opts, (cudaStream_t)stream, result, false);
}

int q3dr_map_set_chunk_callback(q3dr_map* m, q3dr_chunk_callback cb, void* user) {
  if (!m) return fail(Q3DR_ERR_INVALID_ARG, "map is NULL");
  m->chunk_cb = cb;
  m->chunk_cb_user = cb ? user : nullptr;
  return Q3DR_OK;
}

int q3dr_result_to_host(q3dr_map* m, q3dr_result* result) {
  Q3DR_TRY(check_map(m));
  if (!result) return fail(Q3DR_ERR_INVALID_ARG, "result is NULL");
  Q3DR_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = m->own_stream;  // the scoring call returned synchronised; nothing of the caller's stream is touched
  const size_t nv = m->last_n_vp, ne = m->last_n_entries;
  const bool want_pix = (m->result_fields & Q3DR_RESULT_FIRST_PIXEL) != 0 && (m->r_pix.p != nullptr || ne == 0);
  const bool want_dsq = (m->result_fields & Q3DR_RESULT_DIST_SQ) != 0 && (m->r_dsq.p != nullptr || ne == 0);
  Q3DR_TRY(m->h_offsets.ensure(nv + 1));
  Q3DR_TRY(m->h_total.ensure(std::max<size_t>(nv, 1)));
  Q3DR_TRY(m->h_hit_count.ensure(std::max<size_t>(nv, 1)));
  adopt_pooled_host_pair(m, std::max<size_t>(ne, 1));
  Q3DR_TRY(m->h_leaf.ensure(std::max<size_t>(ne, 1)));
  Q3DR_TRY(m->h_info.ensure(std::max<size_t>(ne, 1)));
  if (want_pix) Q3DR_TRY(m->h_pix.ensure(std::max<size_t>(ne, 1)));
  if (want_dsq) Q3DR_TRY(m->h_dsq.ensure(std::max<size_t>(ne, 1)));
  Q3DR_CUDA(cudaMemcpyAsync(m->h_offsets.p, m->d_offsets.p, (nv + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
  if (nv) {
    Q3DR_CUDA(cudaMemcpyAsync(m->h_total.p, m->d_total.p, nv * sizeof(float), cudaMemcpyDeviceToHost, s));
    Q3DR_CUDA(cudaMemcpyAsync(m->h_hit_count.p, m->d_hit_count.p, nv * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  }
  if (ne && !m->host_result_valid) {
    Q3DR_CUDA(cudaMemcpyAsync(m->h_leaf.p, m->r_leaf.p, ne * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    Q3DR_CUDA(cudaMemcpyAsync(m->h_info.p, m->r_info.p, ne * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (want_pix) Q3DR_CUDA(cudaMemcpyAsync(m->h_pix.p, m->r_pix.p, ne * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    if (want_dsq) Q3DR_CUDA(cudaMemcpyAsync(m->h_dsq.p, m->r_dsq.p, ne * sizeof(float), cudaMemcpyDeviceToHost, s));
  }
  Q3DR_CUDA(cudaStreamSynchronize(s));
  m->host_result_valid = true;
  result->n_viewpoints = nv;
  result->n_entries = ne;
  result->offsets = m->h_offsets.p;
  result->leaf = m->h_leaf.p;
  result->information = m->h_info.p;
  result->first_pixel = want_pix ? m->h_pix.p : nullptr;
  result->dist_sq = want_dsq ? m->h_dsq.p : nullptr;
  result->total = m->h_total.p;
  result->hit_count = m->h_hit_count.p;
  return Q3DR_OK;
}

int q3dr_result_detach_host(q3dr_map* m, q3dr_host_block* out) {
  Q3DR_TRY(check_map(m));
  if (!out) return fail(Q3DR_ERR_INVALID_ARG, "out is NULL");
  std::memset(out, 0, sizeof(*out));
  if (!m->host_result_valid || m->h_leaf.p == nullptr || m->h_info.p == nullptr) {
    return fail(Q3DR_ERR_INVALID_ARG, "no host result to detach (q3dr_score_viewpoints or q3dr_result_to_host first)");
  }
  out->leaf = m->h_leaf.p;
  out->information = m->h_info.p;
  out->n_entries = m->last_n_entries;
  out->capacity = std::min(m->h_leaf.cap, m->h_info.cap);
  m->h_leaf.p = nullptr;
  m->h_leaf.cap = 0;
  m->h_info.p = nullptr;
  m->h_info.cap = 0;
  m->host_result_valid = false;
  adopt_pooled_host_pair(m, 0);  // the largest released pair, if any: the next call then allocates nothing
  return Q3DR_OK;
}

int q3dr_host_block_release(q3dr_host_block* block) {
  if (!block) return Q3DR_OK;
  if (block->leaf != nullptr && block->information != nullptr && block->capacity > 0) {
    HostPair hp;
    hp.leaf = block->leaf;
    hp.info = block->information;
    hp.cap = (size_t)block->capacity;
    host_pool_give(hp);
  } else {
    if (block->leaf) cudaFreeHost(block->leaf);
    if (block->information) cudaFreeHost(block->information);
    cudaGetLastError();
  }
  std::memset(block, 0, sizeof(*block));
  return Q3DR_OK;
}

int q3dr_host_pool_trim(void) {
  std::vector<HostPair> drop;
  {
    std::lock_guard<std::mutex> lock(host_pool_mutex());
    drop.swap(host_pool());
  }
  for (const HostPair& hp : drop) {
    if (hp.leaf) cudaFreeHost(hp.leaf);
    if (hp.info) cudaFreeHost(hp.info);
  }
  cudaGetLastError();
  return Q3DR_OK;
}

int q3dr_map_set_result_fields(q3dr_map* m, uint32_t fields) {
  if (!m) return fail(Q3DR_ERR_INVALID_ARG, "map is NULL");
  if (fields & ~(uint32_t)(Q3DR_RESULT_FIRST_PIXEL | Q3DR_RESULT_DIST_SQ)) return fail(Q3DR_ERR_INVALID_ARG, "unknown result field");
  m->result_fields = fields;
  return Q3DR_OK;
}

int q3dr_score_viewpoints(q3dr_map* m, const float K[4], uint32_t image_width, uint32_t x0, uint32_t x1, uint32_t y0,
                          uint32_t y1, const float* Rt, size_t n, const q3dr_score_opts* opts, q3dr_result* result) {
  Q3DR_TRY(check_map(m));
  if (!K || (!Rt && n) || !opts || !result) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  Q3DR_TRY(check_window(x0, x1, y0, y1));
  if (image_width == 0) return fail(Q3DR_ERR_INVALID_ARG, "image_width is 0");
  Q3DR_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = m->own_stream;
  Q3DR_TRY(m->d_rt.ensure(std::max<size_t>(12 * n, 12)));
  if (n) Q3DR_CUDA(cudaMemcpyAsync(m->d_rt.p, Rt, 12 * n * sizeof(float), cudaMemcpyHostToDevice, s));
  Q3DR_TRY(score_device_impl(m, K, image_width, x0, x1, y0, y1, m->d_rt.p, n, opts, s, nullptr, true, nullptr, Rt));
  return q3dr_result_to_host(m, result);
}

// ---- image-space normals (row N2) ----

int q3dr_score_viewpoints_mesh_normals_device(q3dr_map* m, const q3dr_mesh* mesh, const q3dr_render_opts* ropts,
                                              const float K[4], uint32_t image_width, uint32_t image_height,
                                              uint32_t x0, uint32_t x1, uint32_t y0, uint32_t y1, const float* d_Rt,
                                              size_t n, const q3dr_score_opts* opts, void* stream,
                                              q3dr_result* result) {
  Q3DR_TRY(check_map(m));
  if (!K || (!d_Rt && n) || !opts || !mesh || !ropts) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  Q3DR_TRY(check_window(x0, x1, y0, y1));
  if (image_width == 0 || image_height == 0) return fail(Q3DR_ERR_INVALID_ARG, "empty image");
  if ((reinterpret_cast<uintptr_t>(d_Rt) & 15u) != 0 && n) {
    Q3DR_CUDA(cudaSetDevice(m->device));
    Q3DR_TRY(m->d_rt.ensure(12 * n));
    Q3DR_CUDA(cudaMemcpyAsync(m->d_rt.p, d_Rt, 12 * n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    d_Rt = m->d_rt.p;
  }
  NormalSource ns;
  ns.mode = q3dr::kNormalCode;
  ns.mesh = mesh;
  ns.ropts = *ropts;
  ns.img_h = image_height;
  return score_device_impl(m, K, image_width, x0, x1, y0, y1, d_Rt, n, opts, (cudaStream_t)stream, result, false, &ns);
}

int q3dr_score_viewpoints_mesh_normals(q3dr_map* m, const q3dr_mesh* mesh, const q3dr_render_opts* ropts,
                                       const float K[4], uint32_t image_width, uint32_t image_height, uint32_t x0,
                                       uint32_t x1, uint32_t y0, uint32_t y1, const float* Rt, size_t n,
                                       const q3dr_score_opts* opts, q3dr_result* result) {
  Q3DR_TRY(check_map(m));
  if (!K || (!Rt && n) || !opts || !result || !mesh || !ropts) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  Q3DR_TRY(check_window(x0, x1, y0, y1));
  if (image_width == 0 || image_height == 0) return fail(Q3DR_ERR_INVALID_ARG, "empty image");
  Q3DR_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = m->own_stream;
  Q3DR_TRY(m->d_rt.ensure(std::max<size_t>(12 * n, 12)));
  if (n) Q3DR_CUDA(cudaMemcpyAsync(m->d_rt.p, Rt, 12 * n * sizeof(float), cudaMemcpyHostToDevice, s));
  NormalSource ns;
  ns.mode = q3dr::kNormalCode;
  ns.mesh = mesh;
  ns.ropts = *ropts;
  ns.img_h = image_height;
  Q3DR_TRY(score_device_impl(m, K, image_width, x0, x1, y0, y1, m->d_rt.p, n, opts, s, nullptr, true, &ns));
  return q3dr_result_to_host(m, result);
}

int q3dr_score_viewpoints_normal_images(q3dr_map* m, const uint32_t* normal_images, const float K[4],
                                        uint32_t image_width, uint32_t image_height, uint32_t x0, uint32_t x1,
                                        uint32_t y0, uint32_t y1, const float* Rt, size_t n,
                                        const q3dr_score_opts* opts, q3dr_result* result) {
  Q3DR_TRY(check_map(m));
  if (!K || (!Rt && n) || !opts || !result || (!normal_images && n)) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  Q3DR_TRY(check_window(x0, x1, y0, y1));
  if (image_width == 0 || image_height == 0) return fail(Q3DR_ERR_INVALID_ARG, "empty image");
  Q3DR_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = m->own_stream;
  Q3DR_TRY(m->d_rt.ensure(std::max<size_t>(12 * n, 12)));
  if (n) Q3DR_CUDA(cudaMemcpyAsync(m->d_rt.p, Rt, 12 * n * sizeof(float), cudaMemcpyHostToDevice, s));
  NormalSource ns;
  ns.mode = q3dr::kNormalImage;
  ns.images = normal_images;
  ns.images_on_device = false;
  ns.img_h = image_height;
  if (n == 0) ns.mode = q3dr::kNormalVoxel;  // nothing to look up
  Q3DR_TRY(score_device_impl(m, K, image_width, x0, x1, y0, y1, m->d_rt.p, n, opts, s, nullptr, true, &ns));
  return q3dr_result_to_host(m, result);
}

int q3dr_raycast_unique(q3dr_map* m, const float K[4], const float Rt[12], uint32_t x0, uint32_t x1, uint32_t y0,
                        uint32_t y1, float min_range, float max_range, q3dr_pixel_hit* out, size_t cap,
                        size_t* count) {
  Q3DR_TRY(check_map(m));
  if (!K || !Rt || !count || (!out && cap)) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  q3dr_score_opts o;
  q3dr_score_opts_default(&o);
  o.ignore_voxels_with_zero_information = 0;
  o.raycast_min_range = min_range;
  o.raycast_max_range = max_range;
  q3dr_result r;
  const uint32_t saved_fields = m->result_fields;
  m->result_fields = Q3DR_RESULT_FIRST_PIXEL;  // the second pass below re-casts exactly the kept pixels
  const int st_score = q3dr_score_viewpoints(m, K, x1 > x0 ? x1 - x0 : 1, x0, x1, y0, y1, Rt, 1, &o, &r);
  m->result_fields = saved_fields;
  Q3DR_TRY(st_score);
  *count = (size_t)r.n_entries;
  if (r.n_entries > cap) return fail(Q3DR_ERR_INVALID_ARG, "output capacity too small");
  if (r.n_entries == 0) return Q3DR_OK;
  // Full records of the kept pixels: re-cast exactly those pixels (device pixel list = first_pixel of the
  // result still resident in HBM).  A kept pixel's ray hits its voxel by construction.
  const size_t ne = (size_t)r.n_entries;
  cudaStream_t s = nullptr;
  Q3DR_TRY(m->d_pixels.ensure(ne));
  Q3DR_CUDA(cudaMemsetAsync(m->d_counter.p, 0, sizeof(uint32_t), s));
  q3dr::RaycastParams rp;
  fill_raycast_params(m, K, x0, x1, y0, y1, max_range, &rp);
  rp.rt = m->d_rt.p;  // the pose uploaded by q3dr_score_viewpoints above
  rp.pix_list = m->r_pix.p;
  rp.n_rays = (uint32_t)ne;
  rp.total_tiles = (uint32_t)((ne + 31) / 32);
  rp.pixels = m->d_pixels.p;
  const int grid = (int)std::min<uint64_t>((uint64_t)m->grid_rays_ctas,
                                          ((uint64_t)rp.total_tiles + q3dr::kWarpsPerCta - 1) / q3dr::kWarpsPerCta);
  q3dr::raycast_rays_kernel<<<grid, q3dr::kThreadsPerCta, m->smem_bytes, s>>>(rp);
  Q3DR_CUDA(cudaGetLastError());
  Q3DR_CUDA(cudaMemcpyAsync(out, m->d_pixels.p, ne * sizeof(q3dr_pixel_hit), cudaMemcpyDeviceToHost, s));
  Q3DR_CUDA(cudaStreamSynchronize(s));
  for (size_t i = 0; i < ne; ++i) {
    if (out[i].leaf != r.leaf[i]) return fail(Q3DR_ERR_INTERNAL, "kept pixel does not re-hit its voxel");
  }
  return Q3DR_OK;
}

}  // extern "C"
