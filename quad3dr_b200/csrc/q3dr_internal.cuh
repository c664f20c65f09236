// q3dr_internal.cuh -- shared by the translation units of libq3dr_raycast.so (not installed): error plumbing, device /
// pinned buffers, the q3dr_map object.
#pragma once

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/q3dr_raycast.h"
#include "bvh_build.h"

namespace q3dr_detail {

std::string& last_error();  // thread-local message of the last failing call (defined in q3dr_api.cu)

inline int fail(int status, const std::string& msg) {
  last_error() = msg;
  return status;
}

}  // namespace q3dr_detail

using q3dr_detail::fail;


#define Q3DR_CUDA(call)                                                                          \
  do {                                                                                           \
    cudaError_t err__ = (call);                                                                  \
    if (err__ != cudaSuccess) {                                                                  \
      int st__ = (err__ == cudaErrorMemoryAllocation) ? Q3DR_ERR_OUT_OF_MEMORY : Q3DR_ERR_CUDA;  \
      if (err__ == cudaErrorNoDevice || err__ == cudaErrorInsufficientDriver) st__ = Q3DR_ERR_NO_DEVICE; \
      return fail(st__, std::string(#call) + ": " + cudaGetErrorString(err__));                  \
    }                                                                                            \
  } while (0)

#define Q3DR_TRY(expr)             \
  do {                             \
    int st__ = (expr);             \
    if (st__ != Q3DR_OK) return st__; \
  } while (0)

// Guard words behind every device buffer (Q3DR_GUARD=1 in the environment): compute-sanitizer is closed on the GPU pool
// this library is developed on, so writes past the end of a buffer are caught by our own means -- every DevBuf is
// allocated with kGuardBytes of a known pattern behind it, and q3dr_map_check_guards() reads the patterns back.
constexpr size_t kGuardBytes = 256;
constexpr unsigned char kGuardPattern = 0xA5;
inline bool guards_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("Q3DR_GUARD");
    return e != nullptr && e[0] != '\0' && e[0] != '0';
  }();
  return on;
}
inline cudaError_t guarded_malloc(void** q, size_t bytes) {
  if (!guards_enabled()) return cudaMalloc(q, bytes);
  cudaError_t e = cudaMalloc(q, bytes + kGuardBytes);
  if (e == cudaSuccess) e = cudaMemset(static_cast<char*>(*q) + bytes, kGuardPattern, kGuardBytes);
  return e;
}
// 0 = intact (or guards off), 1 = overwritten, -1 = could not be read
inline int guard_state(const void* p, size_t bytes) {
  if (!guards_enabled() || p == nullptr) return 0;
  unsigned char h[kGuardBytes];
  if (cudaMemcpy(h, static_cast<const char*>(p) + bytes, kGuardBytes, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  for (size_t i = 0; i < kGuardBytes; ++i) {
    if (h[i] != kGuardPattern) return 1;
  }
  return 0;
}

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  int guard() const { return guard_state(p, cap * sizeof(T)); }
  int ensure(size_t n, bool keep = false, cudaStream_t s = 0) {
    if (n <= cap) return Q3DR_OK;
    size_t want = std::max(n, cap + cap / 2);
    T* q = nullptr;
    Q3DR_CUDA(guarded_malloc(reinterpret_cast<void**>(&q), want * sizeof(T)));
    if (keep && p && cap) {
      cudaError_t e = cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) {
        cudaFree(q);
        return fail(Q3DR_ERR_CUDA, std::string("grow copy: ") + cudaGetErrorString(e));
      }
    }
    if (p) cudaFree(p);
    p = q;
    cap = want;
    return Q3DR_OK;
  }
  // Grows WITHOUT freeing the old block: the first `used` elements are copied on `s` (stream ordered, no host sync) and
  // the old block is handed to `retired`, to be freed when nothing can refer to it any more (q3dr_map::free_retired).
  // Pointers given out earlier -- chunk callbacks, D2H copies in flight on another stream -- stay valid meanwhile.
  int grow_deferred(size_t n, size_t used, cudaStream_t s, std::vector<void*>* retired) {
    if (n <= cap) return Q3DR_OK;
    T* q = nullptr;
    Q3DR_CUDA(guarded_malloc(reinterpret_cast<void**>(&q), n * sizeof(T)));
    if (p && used) {
      cudaError_t e = cudaMemcpyAsync(q, p, used * sizeof(T), cudaMemcpyDeviceToDevice, s);
      if (e != cudaSuccess) {
        cudaFree(q);
        return fail(Q3DR_ERR_CUDA, std::string("grow copy: ") + cudaGetErrorString(e));
      }
    }
    if (p) retired->push_back(p);
    p = q;
    cap = n;
    return Q3DR_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  size_t bytes() const { return cap * sizeof(T); }
};

template <typename T>
struct PinnedBuf {
  T* p = nullptr;
  size_t cap = 0;
  int ensure(size_t n) {
    if (n <= cap) return Q3DR_OK;
    size_t want = std::max(n, cap + cap / 2);
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    Q3DR_CUDA(cudaHostAlloc(&p, want * sizeof(T), cudaHostAllocPortable));
    cap = want;
    return Q3DR_OK;
  }
  // grows keeping the first `used` elements (caller makes sure no copy into the old block is in flight)
  int grow_keep(size_t n, size_t used) {
    if (n <= cap) return Q3DR_OK;
    size_t want = std::max(n, cap + cap / 2);
    T* q = nullptr;
    Q3DR_CUDA(cudaHostAlloc(&q, want * sizeof(T), cudaHostAllocPortable));
    if (p && used) std::memcpy(q, p, used * sizeof(T));
    if (p) cudaFreeHost(p);
    p = q;
    cap = want;
    return Q3DR_OK;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

struct Derived {
  float thr, kf, a_thr, ka;
};

// viewpoint_score.cpp:20-28 / viewpoint_planner.cpp:107-109 of the reference (fp32, same order)
inline Derived derive(const q3dr_score_opts& o) {
  Derived d;
  d.thr = o.voxel_sensor_size_ratio_threshold;
  d.kf = 1.0f / o.voxel_sensor_size_ratio_inv_falloff_factor;
  d.a_thr = o.incidence_angle_threshold_degrees * float(M_PI) / float(180);
  d.ka = 1.0f / (o.incidence_angle_inv_falloff_factor_degrees * float(M_PI) / float(180));
  return d;
}


struct q3dr_exchange;

struct q3dr_map {
  int device = 0;
  int sm_count = 0;
  size_t n = 0;
  uint32_t ref_depth = 0;
  bool bounded = false;  // all box coordinates finite and below 1e18 in magnitude
  float root_box[6] = {0, 0, 0, 0, 0, 0};  // union of the leaf boxes = the reference root node's bounding box
  // host copies (kept for q3dr_map_reset)
  std::vector<float> h_boxes, h_weight, h_normal;
  std::vector<uint8_t> h_depth;
  std::shared_ptr<const q3dr::FlatTree> tree;  // shared between the per-GPU replicas of a q3dr_multi
  // device: the map
  DevBuf<float4> d_nodes;
  DevBuf<float4> d_leaf_rec;  // [n][4]: {min xyz, weight}, {max xyz, 0}, {normal xyz, 0}, pad (64-byte records)
  DevBuf<uint8_t> d_depth;
  DevBuf<uint32_t> d_counter;           // [0] tiles of the main launch, [1] tiles of the checking launch, [2] flagged viewpoints
  DevBuf<float> d_planes;               // sorted distinct box plane coordinates, x | y | z
  uint32_t n_planes[3] = {0, 0, 0};
  std::vector<float> h_planes[3];       // the same on the host (small host-resident batches are flagged without a kernel)
  DevBuf<uint8_t> d_vp_flags;           // [slots] camera centre on a box plane (flag_viewpoints_kernel)
  DevBuf<uint32_t> d_vp_list;           // [slots] the flagged viewpoints of the chunk
  // device: per-chunk scratch
  DevBuf<uint32_t> d_first_pix, d_bitmap, d_kept;
  DevBuf<uint32_t> d_seg_u32;  // [2][slots][n_segs]: hits, off
  DevBuf<uint32_t> d_blk_u32;  // [slots + 1] blk_off, then [2][max blocks]: kept, koff
  DevBuf<double> d_seg_f64;    // [max blocks] blk_total
  size_t max_blocks = 0;
  uint32_t n_segs = 0;
  size_t slots = 0;
  uint64_t leaf_stride = 0, word_stride = 0;
  DevBuf<int32_t> t_leaf;
  DevBuf<float> t_info, t_dsq;
  DevBuf<uint32_t> t_pix;
  DevBuf<uint32_t> t_code;   // image-space normal codes next to the lists (allocated on first use)
  DevBuf<uint32_t> d_nimg;   // one chunk of the caller's normal images
  uint64_t tmp_cap = 0;
  // device: results of the last scoring call
  DevBuf<uint64_t> d_offsets;
  DevBuf<float> d_total;
  DevBuf<uint32_t> d_hit_count;
  DevBuf<int32_t> r_leaf;
  DevBuf<float> r_info, r_dsq;
  DevBuf<uint32_t> r_pix;
  uint64_t last_n_vp = 0, last_n_entries = 0;
  // staging
  DevBuf<float> d_rt, d_ray_o, d_ray_d;
  DevBuf<q3dr_pixel_hit> d_pixels;
  PinnedBuf<uint64_t> h_scalar;
  PinnedBuf<uint64_t> h_offsets;
  PinnedBuf<float> h_total, h_info, h_dsq;
  PinnedBuf<uint32_t> h_hit_count, h_pix;
  PinnedBuf<int32_t> h_leaf;
  PinnedBuf<float> h_rt;
  // set consumers (setops_kernels.cuh)
  DevBuf<float> d_dense;     // [n] scratch: one viewpoint's informations scattered over the leaves (NaN elsewhere)
  DevBuf<float> d_set_out;   // per-viewpoint sums
  DevBuf<uint32_t> d_set_idx;
  DevBuf<double> d_set_part;
  bool dense_ready = false;
  bool scratch_dirty = false;  // a scoring call failed between its raycast and its clean-up: tables need a reset
  // overlap queries (overlap_kernels.cuh)
  DevBuf<float> d_q_boxes;
  DevBuf<uint32_t> d_q_counts;
  DevBuf<uint64_t> d_q_offsets;
  DevBuf<int32_t> d_q_leaf, d_q_sorted;
  DevBuf<uint8_t> d_q_tmp, d_q_valid;
  std::vector<uint64_t> h_q_offsets;
  std::vector<int32_t> h_q_leaf;
  // q3dr_generate_viewpoint_entries: per-candidate CSR offsets / totals / hit counts (host)
  std::vector<uint64_t> cand_offsets;
  std::vector<float> cand_total;
  std::vector<uint32_t> cand_hits;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // D2H of finished chunks overlaps the next chunk's kernels
  cudaStream_t own_stream = nullptr;   // non-blocking; the host-buffer entry points work on it (they return synchronised),
                                       // so two maps driven by two host threads never serialise on the legacy stream
  cudaEvent_t ev_packed = nullptr;
  bool host_result_valid = false;      // the host pools already hold the entries of the last call
  q3dr_chunk_callback chunk_cb = nullptr;
  void* chunk_cb_user = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  // per-chunk timing events of a scoring call (5 per chunk: raycast begin / end, scoring end, pack begin / end); read
  // once after the call's final synchronisation, never waited on inside the chunk loop
  std::vector<cudaEvent_t> ev_chunk;
  cudaEvent_t ev_count = nullptr;      // "the entry count of the chunk is in h_scalar"
  std::vector<void*> retired;          // result pools replaced while growing; freed at the start of the next call
  uint32_t result_fields = 3u;         // Q3DR_RESULT_* the scoring calls deliver besides (leaf, information)
  // first-pixel table: values are (tag << 22 | pixel) with a tag that DEcreases from chunk to chunk, so entries of earlier
  // chunks read as "not hit" and the table never needs a reset pass (windows of up to 2^22 pixels; larger ones run
  // untagged and the scoring kernel resets what it read)
  uint32_t fp_epoch = 0;
  bool fp_tagged_garbage = false;      // the table holds tagged entries of earlier chunks
  q3dr_exchange* exchange = nullptr;   // pushes every finished chunk to the peers (q3dr_exchange_begin_step)
  int grid_rays_ctas = 0;
  int grid2_ctas[2] = {0, 0};
  uint32_t tile_run = 1;               // consecutive tiles per warp and counter visit (Q3DR_TILE_RUN)
  int ray_variant = 0;                 // packet shape of the camera-window kernel (q3dr_api.cu: kRayVariants)
  size_t smem_bytes = 0;
  q3dr_stats stats{};

  void free_retired() {
    for (void* q : retired) cudaFree(q);
    retired.clear();
  }

  // number of device buffers of this map whose guard words were overwritten (Q3DR_GUARD=1), or -1 on a read error
  int overwritten_guards() const {
    const int states[] = {d_nodes.guard(), d_leaf_rec.guard(), d_depth.guard(), d_counter.guard(), d_planes.guard(),
                          d_vp_flags.guard(), d_vp_list.guard(), d_first_pix.guard(), d_bitmap.guard(), d_kept.guard(),
                          d_seg_u32.guard(), d_blk_u32.guard(), d_seg_f64.guard(), t_leaf.guard(), t_info.guard(),
                          t_dsq.guard(), t_pix.guard(), t_code.guard(), d_nimg.guard(), d_offsets.guard(), d_total.guard(),
                          d_hit_count.guard(), r_leaf.guard(), r_info.guard(), r_dsq.guard(), r_pix.guard(), d_rt.guard(),
                          d_ray_o.guard(), d_ray_d.guard(), d_pixels.guard(), d_dense.guard(), d_set_out.guard(),
                          d_set_idx.guard(), d_set_part.guard(), d_q_boxes.guard(), d_q_counts.guard(), d_q_offsets.guard(),
                          d_q_leaf.guard(), d_q_sorted.guard(), d_q_tmp.guard(), d_q_valid.guard()};
    int bad = 0;
    for (int st : states) {
      if (st < 0) return -1;
      bad += st;
    }
    return bad;
  }

  uint64_t device_bytes() const {
    return d_nodes.bytes() + d_leaf_rec.bytes() + d_depth.bytes() + d_seg_u32.bytes() + d_blk_u32.bytes() + d_seg_f64.bytes() +
           d_counter.bytes() + d_first_pix.bytes() + d_bitmap.bytes() + d_kept.bytes() + t_leaf.bytes() +
           t_info.bytes() + t_dsq.bytes() + t_pix.bytes() + t_code.bytes() + d_nimg.bytes() + d_offsets.bytes() + d_total.bytes() +
           d_hit_count.bytes() + r_leaf.bytes() + r_info.bytes() + r_dsq.bytes() + r_pix.bytes() + d_rt.bytes() +
           d_ray_o.bytes() + d_ray_d.bytes() + d_pixels.bytes();
  }
};

// Triangle mesh of the image-space normals (row N2): binary BVH + triangles + vertex normals in HBM.
struct q3dr_mesh {
  int device = 0;
  size_t n_tris = 0;
  size_t n_nodes = 0;
  uint32_t depth = 0;
  int32_t root_ref = 0;
  float pad = 0.f;
  DevBuf<float4> d_nodes;    // [n_nodes][4]
  DevBuf<float4> d_tris;     // [n_tris][3] BVH order
  DevBuf<float4> d_normals;  // [n_tris][3] original order
  DevBuf<float> d_rt;
  DevBuf<uint32_t> d_image;
  float last_render_ms = 0.f;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  uint64_t device_bytes() const { return d_nodes.bytes() + d_tris.bytes() + d_normals.bytes() + d_rt.bytes() + d_image.bytes(); }
};

// normal source of a scoring call (score_device_impl)
struct NormalSource {
  int mode = 0;                          // q3dr::NormalMode
  const q3dr_mesh* mesh = nullptr;       // kNormalCode
  q3dr_render_opts ropts{};              // kNormalCode
  const uint32_t* images = nullptr;      // kNormalImage: n x img_h x img_w ARGB32
  bool images_on_device = false;
  uint32_t img_h = 0;
};

namespace q3dr_detail {
int check_map(const q3dr_map* m);
int raycast_window_device(q3dr_map* m, const float K[4], const float* d_rt, uint32_t x0, uint32_t x1, uint32_t y0,
                          uint32_t y1, float max_range, cudaStream_t s);
// q3dr_exchange.cu: pushes entries [first_entry, first_entry + n_entries) of the chunk just packed (ready when `ready`
// has fired) into this rank's slot on every peer; called by the scoring call for each finished chunk
int exchange_push_chunk(q3dr_exchange* ex, cudaEvent_t ready, const int32_t* d_leaf, const float* d_info,
                        uint64_t first_entry, uint64_t n_entries);
int map_create_impl(const float* boxes, const float* weights, const float* normals, const uint32_t* depth, size_t n,
                    int device, std::shared_ptr<const q3dr::FlatTree> shared_tree, q3dr_map** out);
}  // namespace q3dr_detail
