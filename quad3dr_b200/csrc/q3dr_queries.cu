// q3dr_queries.cu -- the C ABI entry points around the hot path: consumers of the voxel sets,
// box-vs-map overlap queries, octree leaves -> BVH leaves and grid weights, and q3dr_multi (one process, several GPUs).
#include "q3dr_internal.cuh"

#include "setops_kernels.cuh"
#include "overlap_kernels.cuh"
#include "octree_kernels.cuh"
#include "score_kernels.cuh"  // kLeafRecStride, kMaxChunkViewpoints, csr_offsets_kernel

#include "downweight_kernels.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_segmented_radix_sort.cuh>

using q3dr_detail::check_map;
using q3dr_detail::map_create_impl;

extern "C" {

// ---- consumers of the voxel sets ----

}  // extern "C"

struct q3dr_observed {
  q3dr_map* map = nullptr;
  DevBuf<float> d;  // [n_leaves], NaN = absent
};

namespace {

int fill_set_params(const q3dr_map* m, float factor, q3dr::SetParams* sp) {
  if (m->last_n_vp == 0 && m->d_offsets.p == nullptr) return fail(Q3DR_ERR_INVALID_ARG, "no scoring result on this map yet");
  sp->offsets = m->d_offsets.p;
  sp->leaf = m->r_leaf.p;
  sp->info = m->r_info.p;
  sp->leaf_rec = m->d_leaf_rec.p;
  sp->rec_stride = (uint32_t)q3dr::kLeafRecStride;
  sp->factor = factor;
  return Q3DR_OK;
}

int fill_nan(float* p, size_t n, cudaStream_t s) {
  q3dr::fill_f32_kernel<<<256, 256, 0, s>>>(p, n, std::nanf(""));
  Q3DR_CUDA(cudaGetLastError());
  return Q3DR_OK;
}

}  // namespace

extern "C" {

int q3dr_observed_create(q3dr_map* m, q3dr_observed** out) {
  if (!out) return fail(Q3DR_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  Q3DR_TRY(check_map(m));
  Q3DR_CUDA(cudaSetDevice(m->device));
  q3dr_observed* o = new (std::nothrow) q3dr_observed;
  if (!o) return fail(Q3DR_ERR_OUT_OF_MEMORY, "host allocation failed");
  o->map = m;
  int st = o->d.ensure(m->n);
  if (st == Q3DR_OK) st = fill_nan(o->d.p, m->n, m->stream);
  if (st == Q3DR_OK && cudaStreamSynchronize(m->stream) != cudaSuccess) st = fail(Q3DR_ERR_CUDA, "sync failed");
  if (st != Q3DR_OK) {
    o->d.release();
    delete o;
    return st;
  }
  *out = o;
  return Q3DR_OK;
}

int q3dr_observed_clear(q3dr_observed* o) {
  if (!o) return fail(Q3DR_ERR_INVALID_ARG, "obs is NULL");
  Q3DR_CUDA(cudaSetDevice(o->map->device));
  Q3DR_TRY(fill_nan(o->d.p, o->map->n, o->map->stream));
  Q3DR_CUDA(cudaStreamSynchronize(o->map->stream));
  return Q3DR_OK;
}

int q3dr_observed_destroy(q3dr_observed* o) {
  if (!o) return Q3DR_OK;
  cudaSetDevice(o->map->device);
  o->d.release();
  delete o;
  return Q3DR_OK;
}

int q3dr_observed_download(const q3dr_observed* o, float* out) {
  if (!o || !out) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  Q3DR_CUDA(cudaSetDevice(o->map->device));
  Q3DR_CUDA(cudaMemcpy(out, o->d.p, o->map->n * sizeof(float), cudaMemcpyDeviceToHost));
  return Q3DR_OK;
}

int q3dr_new_information(q3dr_map* m, const q3dr_observed* o, float factor, const uint32_t* viewpoints, size_t n,
                         float* out) {
  Q3DR_TRY(check_map(m));
  if (!o || (!out && n)) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  if (o->map != m) return fail(Q3DR_ERR_INVALID_ARG, "observed map belongs to another occupancy map");
  if (n == 0) return Q3DR_OK;
  if (n > m->last_n_vp) return fail(Q3DR_ERR_INVALID_ARG, "more viewpoints than the last result holds");
  Q3DR_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = m->stream;
  q3dr::SetParams sp;
  Q3DR_TRY(fill_set_params(m, factor, &sp));
  Q3DR_TRY(m->d_set_out.ensure(n));
  const uint32_t* d_idx = nullptr;
  if (viewpoints) {
    for (size_t i = 0; i < n; ++i) {
      if (viewpoints[i] >= m->last_n_vp) return fail(Q3DR_ERR_INVALID_ARG, "viewpoint index out of range");
    }
    Q3DR_TRY(m->d_set_idx.ensure(n));
    Q3DR_CUDA(cudaMemcpyAsync(m->d_set_idx.p, viewpoints, n * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    d_idx = m->d_set_idx.p;
  }
  q3dr::new_information_kernel<<<(unsigned)n, q3dr::kSetThreads, 0, s>>>(sp, o->d.p, d_idx, m->d_set_out.p);
  Q3DR_CUDA(cudaGetLastError());
  Q3DR_CUDA(cudaMemcpyAsync(out, m->d_set_out.p, n * sizeof(float), cudaMemcpyDeviceToHost, s));
  Q3DR_CUDA(cudaStreamSynchronize(s));
  return Q3DR_OK;
}

int q3dr_observed_add_viewpoint(q3dr_map* m, q3dr_observed* o, float factor, uint32_t viewpoint, float* new_information) {
  Q3DR_TRY(check_map(m));
  if (!o) return fail(Q3DR_ERR_INVALID_ARG, "obs is NULL");
  if (o->map != m) return fail(Q3DR_ERR_INVALID_ARG, "observed map belongs to another occupancy map");
  if (viewpoint >= m->last_n_vp) return fail(Q3DR_ERR_INVALID_ARG, "viewpoint index out of range");
  Q3DR_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = m->stream;
  q3dr::SetParams sp;
  Q3DR_TRY(fill_set_params(m, factor, &sp));
  const unsigned grid = 64;
  Q3DR_TRY(m->d_set_part.ensure(grid));
  Q3DR_TRY(m->d_set_out.ensure(1));
  q3dr::observed_add_kernel<<<grid, q3dr::kSetThreads, 0, s>>>(sp, o->d.p, viewpoint, m->d_set_part.p);
  Q3DR_CUDA(cudaGetLastError());
  q3dr::sum_parts_kernel<<<1, 32, 0, s>>>(m->d_set_part.p, grid, m->d_set_out.p);
  Q3DR_CUDA(cudaGetLastError());
  float v = 0.f;
  Q3DR_CUDA(cudaMemcpyAsync(&v, m->d_set_out.p, sizeof(float), cudaMemcpyDeviceToHost, s));
  Q3DR_CUDA(cudaStreamSynchronize(s));
  if (new_information) *new_information = v;
  return Q3DR_OK;
}

int q3dr_observed_voxel_set(q3dr_map* m, q3dr_observed* o, float factor, const int32_t* leaf, const float* information,
                            size_t n, int add, float* new_information) {
  Q3DR_TRY(check_map(m));
  if (!o || ((!leaf || !information) && n)) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  if (o->map != m) return fail(Q3DR_ERR_INVALID_ARG, "observed map belongs to another occupancy map");
  if (n >= 0x7FFFFFFFull) return fail(Q3DR_ERR_INVALID_ARG, "voxel set too large");
  for (size_t i = 0; i < n; ++i) {
    if (leaf[i] < 0 || (size_t)leaf[i] >= m->n) return fail(Q3DR_ERR_INVALID_ARG, "leaf index out of range");
  }
  if (new_information) *new_information = 0.0f;
  if (n == 0) return Q3DR_OK;
  Q3DR_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = m->stream;
  q3dr::SetParams sp;
  std::memset(&sp, 0, sizeof(sp));
  sp.leaf_rec = m->d_leaf_rec.p;
  sp.rec_stride = (uint32_t)q3dr::kLeafRecStride;
  sp.factor = factor;
  DevBuf<int32_t> d_leaf;
  DevBuf<float> d_info;
  struct Guard {
    DevBuf<int32_t>* a;
    DevBuf<float>* b;
    ~Guard() { a->release(); b->release(); }
  } guard{&d_leaf, &d_info};
  Q3DR_TRY(d_leaf.ensure(n));
  Q3DR_TRY(d_info.ensure(n));
  Q3DR_CUDA(cudaMemcpyAsync(d_leaf.p, leaf, n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  Q3DR_CUDA(cudaMemcpyAsync(d_info.p, information, n * sizeof(float), cudaMemcpyHostToDevice, s));
  const unsigned grid = 64;
  Q3DR_TRY(m->d_set_part.ensure(grid));
  Q3DR_TRY(m->d_set_out.ensure(1));
  if (add) {
    q3dr::observed_add_list_kernel<<<grid, q3dr::kSetThreads, 0, s>>>(sp, o->d.p, d_leaf.p, d_info.p, (uint32_t)n, m->d_set_part.p);
  } else {
    q3dr::new_information_list_kernel<<<grid, q3dr::kSetThreads, 0, s>>>(sp, o->d.p, d_leaf.p, d_info.p, (uint32_t)n, m->d_set_part.p);
  }
  Q3DR_CUDA(cudaGetLastError());
  q3dr::sum_parts_kernel<<<1, 32, 0, s>>>(m->d_set_part.p, grid, m->d_set_out.p);
  Q3DR_CUDA(cudaGetLastError());
  float v = 0.f;
  Q3DR_CUDA(cudaMemcpyAsync(&v, m->d_set_out.p, sizeof(float), cudaMemcpyDeviceToHost, s));
  Q3DR_CUDA(cudaStreamSynchronize(s));
  if (new_information) *new_information = v;
  return Q3DR_OK;
}

int q3dr_overlap_information(q3dr_map* m, uint32_t ref, const uint32_t* others, size_t n, float* out) {
  Q3DR_TRY(check_map(m));
  if ((!others || !out) && n) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  if (ref >= m->last_n_vp) return fail(Q3DR_ERR_INVALID_ARG, "viewpoint index out of range");
  for (size_t i = 0; i < n; ++i) {
    if (others[i] >= m->last_n_vp) return fail(Q3DR_ERR_INVALID_ARG, "viewpoint index out of range");
  }
  if (n == 0) return Q3DR_OK;
  Q3DR_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = m->stream;
  q3dr::SetParams sp;
  Q3DR_TRY(fill_set_params(m, 1.0f, &sp));
  if (!m->dense_ready || m->d_dense.cap < m->n) {
    Q3DR_TRY(m->d_dense.ensure(m->n));
    Q3DR_TRY(fill_nan(m->d_dense.p, m->n, s));
    m->dense_ready = true;
  }
  Q3DR_TRY(m->d_set_idx.ensure(n));
  Q3DR_TRY(m->d_set_out.ensure(n));
  Q3DR_CUDA(cudaMemcpyAsync(m->d_set_idx.p, others, n * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  q3dr::scatter_set_kernel<<<64, 256, 0, s>>>(sp, ref, m->d_dense.p, 0);
  Q3DR_CUDA(cudaGetLastError());
  q3dr::overlap_information_kernel<<<(unsigned)n, q3dr::kSetThreads, 0, s>>>(sp, m->d_dense.p, m->d_set_idx.p, m->d_set_out.p);
  Q3DR_CUDA(cudaGetLastError());
  q3dr::scatter_set_kernel<<<64, 256, 0, s>>>(sp, ref, m->d_dense.p, 1);
  Q3DR_CUDA(cudaGetLastError());
  Q3DR_CUDA(cudaMemcpyAsync(out, m->d_set_out.p, n * sizeof(float), cudaMemcpyDeviceToHost, s));
  Q3DR_CUDA(cudaStreamSynchronize(s));
  return Q3DR_OK;
}

// ---- box-vs-map overlap queries ----

int q3dr_overlap_boxes(q3dr_map* m, const float* boxes, size_t n, q3dr_overlap_result* result) {
  Q3DR_TRY(check_map(m));
  if (!result || (!boxes && n)) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  if (n >= 0x7FFFFFFFull) return fail(Q3DR_ERR_INVALID_ARG, "too many boxes for one call");
  Q3DR_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = m->stream;
  m->h_q_offsets.assign(n + 1, 0);
  m->h_q_leaf.clear();
  if (n) {
    Q3DR_TRY(m->d_q_boxes.ensure(6 * n));
    Q3DR_TRY(m->d_q_counts.ensure(n));
    Q3DR_TRY(m->d_q_offsets.ensure(n + 1));
    Q3DR_CUDA(cudaMemcpyAsync(m->d_q_boxes.p, boxes, 6 * n * sizeof(float), cudaMemcpyHostToDevice, s));
    const unsigned grid = (unsigned)((n + q3dr::kOverlapThreads - 1) / q3dr::kOverlapThreads);
    q3dr::overlap_boxes_kernel<<<grid, q3dr::kOverlapThreads, 0, s>>>(m->d_nodes.p, m->d_q_boxes.p, (uint32_t)n,
                                                                     m->d_q_counts.p, nullptr, nullptr);
    Q3DR_CUDA(cudaGetLastError());
    Q3DR_CUDA(cudaMemsetAsync(m->d_q_offsets.p, 0, sizeof(uint64_t), s));
    for (size_t v0 = 0; v0 < n; v0 += q3dr::kMaxChunkViewpoints) {  // csr_offsets_kernel continues from offsets[v0]
      const size_t nb = std::min<size_t>(q3dr::kMaxChunkViewpoints, n - v0);
      q3dr::csr_offsets_kernel<<<1, q3dr::kSegThreads, 0, s>>>(m->d_q_counts.p + v0, (uint32_t)nb, m->d_q_offsets.p + v0);
      Q3DR_CUDA(cudaGetLastError());
    }
    Q3DR_CUDA(cudaMemcpyAsync(m->h_q_offsets.data(), m->d_q_offsets.p, (n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    Q3DR_CUDA(cudaStreamSynchronize(s));
    const uint64_t ne = m->h_q_offsets[n];
    if (ne >= 0x7FFFFFFFull) return fail(Q3DR_ERR_INVALID_ARG, "overlap result too large for one call");
    if (ne) {
      Q3DR_TRY(m->d_q_leaf.ensure(ne));
      Q3DR_TRY(m->d_q_sorted.ensure(ne));
      q3dr::overlap_boxes_kernel<<<grid, q3dr::kOverlapThreads, 0, s>>>(m->d_nodes.p, m->d_q_boxes.p, (uint32_t)n,
                                                                       m->d_q_counts.p, m->d_q_offsets.p, m->d_q_leaf.p);
      Q3DR_CUDA(cudaGetLastError());
      // ascending leaf index inside each box's list = the reference's left-to-right visiting order
      size_t tmp_bytes = 0;
      Q3DR_CUDA(cub::DeviceSegmentedRadixSort::SortKeys(nullptr, tmp_bytes, m->d_q_leaf.p, m->d_q_sorted.p, (int)ne, (int)n,
                                                        m->d_q_offsets.p, m->d_q_offsets.p + 1, 0, 32, s));
      Q3DR_TRY(m->d_q_tmp.ensure(tmp_bytes));
      Q3DR_CUDA(cub::DeviceSegmentedRadixSort::SortKeys(m->d_q_tmp.p, tmp_bytes, m->d_q_leaf.p, m->d_q_sorted.p, (int)ne, (int)n,
                                                        m->d_q_offsets.p, m->d_q_offsets.p + 1, 0, 32, s));
      m->h_q_leaf.resize(ne);
      Q3DR_CUDA(cudaMemcpyAsync(m->h_q_leaf.data(), m->d_q_sorted.p, ne * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
      Q3DR_CUDA(cudaStreamSynchronize(s));
    }
  }
  result->n_boxes = n;
  result->n_entries = m->h_q_offsets[n];
  result->offsets = m->h_q_offsets.data();
  result->leaf = m->h_q_leaf.data();
  return Q3DR_OK;
}

int q3dr_valid_object_positions(q3dr_map* m, const float* positions, size_t n, const float object_bbox[6],
                                const float bvh_bbox[6], float obstacle_free_height, uint8_t* valid) {
  Q3DR_TRY(check_map(m));
  if (!object_bbox || !bvh_bbox || ((!positions || !valid) && n)) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  if (n == 0) return Q3DR_OK;
  if (n >= 0x7FFFFFFFull) return fail(Q3DR_ERR_INVALID_ARG, "too many positions for one call");
  Q3DR_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = m->stream;
  q3dr::ValidParams vp;
  for (int i = 0; i < 6; ++i) {
    vp.obj[i] = object_bbox[i];
    vp.bvh_bbox[i] = bvh_bbox[i];
    vp.root_box[i] = m->root_box[i];
  }
  vp.obstacle_free_height = obstacle_free_height;
  Q3DR_TRY(m->d_q_boxes.ensure(3 * n));
  Q3DR_TRY(m->d_q_valid.ensure(n));
  Q3DR_CUDA(cudaMemcpyAsync(m->d_q_boxes.p, positions, 3 * n * sizeof(float), cudaMemcpyHostToDevice, s));
  const unsigned grid = (unsigned)((n + q3dr::kOverlapThreads - 1) / q3dr::kOverlapThreads);
  q3dr::valid_positions_kernel<<<grid, q3dr::kOverlapThreads, 0, s>>>(m->d_nodes.p, m->d_q_boxes.p, (uint32_t)n, vp,
                                                                     m->d_q_valid.p);
  Q3DR_CUDA(cudaGetLastError());
  Q3DR_CUDA(cudaMemcpyAsync(valid, m->d_q_valid.p, n, cudaMemcpyDeviceToHost, s));
  Q3DR_CUDA(cudaStreamSynchronize(s));
  return Q3DR_OK;
}

// ---- occupancy octree -> BVH leaves, grid weights ----

int q3dr_octree_leaves_to_boxes(const float* center, const float* size, const float* occupancy,
                                const uint32_t* observation_count, size_t n, const float bvh_bbox[6],
                                float obstacle_free_height, float occupancy_threshold, uint32_t observation_threshold,
                                int device, float* out_boxes, uint32_t* out_source, size_t* n_out) {
  if (!n_out) return fail(Q3DR_ERR_INVALID_ARG, "n_out is NULL");
  *n_out = 0;
  if (n == 0) return Q3DR_OK;
  if (!center || !size || !occupancy || !observation_count || !bvh_bbox || !out_boxes || !out_source)
    return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  if (n >= 0x7FFFFFFFull) return fail(Q3DR_ERR_INVALID_ARG, "too many octree leaves for one call");
  int count = 0;
  Q3DR_TRY(q3dr_device_count(&count));
  if (count <= 0) return fail(Q3DR_ERR_NO_DEVICE, "no CUDA device visible; this engine has no CPU path");
  if (device < 0 || device >= count) return fail(Q3DR_ERR_INVALID_ARG, "device index out of range");
  Q3DR_CUDA(cudaSetDevice(device));
  DevBuf<float> d_center, d_size, d_occ, d_box, d_out_box;
  DevBuf<uint32_t> d_cnt, d_pos, d_out_src, d_keep;
  DevBuf<uint8_t> d_tmp;
  struct Guard {
    DevBuf<float>*a, *b, *c, *d, *e;
    DevBuf<uint32_t>*f, *g, *h, *i;
    DevBuf<uint8_t>* j;
    ~Guard() {
      a->release(); b->release(); c->release(); d->release(); e->release();
      f->release(); g->release(); h->release(); i->release(); j->release();
    }
  } guard{&d_center, &d_size, &d_occ, &d_box, &d_out_box, &d_cnt, &d_pos, &d_out_src, &d_keep, &d_tmp};
  Q3DR_TRY(d_center.ensure(3 * n));
  Q3DR_TRY(d_size.ensure(n));
  Q3DR_TRY(d_occ.ensure(n));
  Q3DR_TRY(d_cnt.ensure(n));
  Q3DR_TRY(d_keep.ensure(n));
  Q3DR_TRY(d_box.ensure(6 * n));
  Q3DR_TRY(d_pos.ensure(n + 1));
  Q3DR_CUDA(cudaMemcpy(d_center.p, center, 3 * n * sizeof(float), cudaMemcpyHostToDevice));
  Q3DR_CUDA(cudaMemcpy(d_size.p, size, n * sizeof(float), cudaMemcpyHostToDevice));
  Q3DR_CUDA(cudaMemcpy(d_occ.p, occupancy, n * sizeof(float), cudaMemcpyHostToDevice));
  Q3DR_CUDA(cudaMemcpy(d_cnt.p, observation_count, n * sizeof(uint32_t), cudaMemcpyHostToDevice));
  q3dr::LeafRuleParams lp;
  lp.center = d_center.p;
  lp.size = d_size.p;
  lp.occupancy = d_occ.p;
  lp.observation_count = d_cnt.p;
  lp.n = (uint32_t)n;
  for (int i = 0; i < 6; ++i) lp.bvh_bbox[i] = bvh_bbox[i];
  lp.obstacle_free_height = obstacle_free_height;
  lp.occupancy_threshold = occupancy_threshold;
  lp.observation_threshold = observation_threshold;
  const unsigned grid = (unsigned)((n + 255) / 256);
  q3dr::leaf_rule_kernel<<<grid, 256>>>(lp, d_keep.p, d_box.p);
  Q3DR_CUDA(cudaGetLastError());
  size_t tmp_bytes = 0;
  Q3DR_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_keep.p, d_pos.p, (int)n));
  Q3DR_TRY(d_tmp.ensure(tmp_bytes));
  Q3DR_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp.p, tmp_bytes, d_keep.p, d_pos.p, (int)n));
  uint32_t last_pos = 0, last_keep = 0;
  Q3DR_CUDA(cudaMemcpy(&last_pos, d_pos.p + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost));
  Q3DR_CUDA(cudaMemcpy(&last_keep, d_keep.p + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost));
  const size_t kept = (size_t)last_pos + last_keep;
  if (kept) {
    Q3DR_TRY(d_out_box.ensure(6 * kept));
    Q3DR_TRY(d_out_src.ensure(kept));
    q3dr::leaf_compact_kernel<<<grid, 256>>>(d_keep.p, d_pos.p, d_box.p, (uint32_t)n, d_out_box.p, d_out_src.p);
    Q3DR_CUDA(cudaGetLastError());
    Q3DR_CUDA(cudaMemcpy(out_boxes, d_out_box.p, 6 * kept * sizeof(float), cudaMemcpyDeviceToHost));
    Q3DR_CUDA(cudaMemcpy(out_source, d_out_src.p, kept * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  }
  *n_out = kept;
  return Q3DR_OK;
}

int q3dr_map_update_weights_from_grid(q3dr_map* m, const float grid_origin[3], float grid_increment,
                                      const int32_t grid_dim[3], const float* grid_weight,
                                      float voxel_information_lambda, const uint32_t* observation_count,
                                      float* weights_out) {
  Q3DR_TRY(check_map(m));
  if (!grid_origin || !grid_dim || !grid_weight || !observation_count) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  if (!(grid_increment > 0.0f) || grid_dim[0] <= 0 || grid_dim[1] <= 0 || grid_dim[2] <= 0)
    return fail(Q3DR_ERR_INVALID_ARG, "bad grid");
  Q3DR_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = m->stream;
  const size_t cells = (size_t)grid_dim[0] * grid_dim[1] * grid_dim[2];
  DevBuf<float> d_gw, d_w;
  DevBuf<uint32_t> d_cnt;
  struct Guard {
    DevBuf<float>*a, *b;
    DevBuf<uint32_t>* c;
    ~Guard() { a->release(); b->release(); c->release(); }
  } guard{&d_gw, &d_w, &d_cnt};
  Q3DR_TRY(d_gw.ensure(cells));
  Q3DR_TRY(d_w.ensure(m->n));
  Q3DR_TRY(d_cnt.ensure(m->n));
  Q3DR_CUDA(cudaMemcpyAsync(d_gw.p, grid_weight, cells * sizeof(float), cudaMemcpyHostToDevice, s));
  Q3DR_CUDA(cudaMemcpyAsync(d_cnt.p, observation_count, m->n * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  q3dr::GridParams g;
  for (int i = 0; i < 3; ++i) {
    g.origin[i] = grid_origin[i];
    g.dim[i] = grid_dim[i];
  }
  g.increment = grid_increment;
  g.lambda = voxel_information_lambda;
  q3dr::grid_weights_kernel<<<(unsigned)((m->n + 255) / 256), 256, 0, s>>>(m->d_leaf_rec.p, (uint32_t)q3dr::kLeafRecStride,
                                                                          (uint32_t)m->n, g, d_gw.p, d_cnt.p, d_w.p);
  Q3DR_CUDA(cudaGetLastError());
  // the host copy of the weights follows (q3dr_map_reset re-uploads from it)
  Q3DR_CUDA(cudaMemcpyAsync(m->h_weight.data(), d_w.p, m->n * sizeof(float), cudaMemcpyDeviceToHost, s));
  Q3DR_CUDA(cudaStreamSynchronize(s));
  if (weights_out) std::memcpy(weights_out, m->h_weight.data(), m->n * sizeof(float));
  return Q3DR_OK;
}

// ---- real-image down-weighting (row N4) ----

int q3dr_map_downweight_real_viewpoints(q3dr_map* m, const float K[4], uint32_t width, uint32_t height, const float* Rt,
                                        const float* depth_maps, size_t n, const q3dr_score_opts* opts,
                                        float invalid_pixel_observation_factor, const q3dr_mesh* mesh,
                                        const q3dr_render_opts* ropts, const uint32_t* normal_images, float* weights_out,
                                        uint64_t* n_updates) {
  Q3DR_TRY(check_map(m));
  if (!K || !opts || (n && (!Rt || !depth_maps))) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  if (width == 0 || height == 0 || (uint64_t)width * height >= 0xFFFFFFFFull) return fail(Q3DR_ERR_INVALID_ARG, "bad image size");
  if (mesh && normal_images) return fail(Q3DR_ERR_INVALID_ARG, "give a mesh or normal images, not both");
  if (mesh && (!ropts || !mesh->d_tris.p || mesh->device != m->device)) {
    return fail(Q3DR_ERR_INVALID_ARG, "mesh without render options, or on another device than the map");
  }
  Q3DR_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = m->own_stream;
  const uint32_t n_pix = width * height;
  // the reference's default range is numeric_limits<float>::max(): the bound max_range^2 overflows to +inf, i.e. unlimited
  const float max_range = (opts->raycast_max_range >= 1e18f) ? -1.0f : opts->raycast_max_range;
  DevBuf<float> d_depth, d_rt, d_w;
  DevBuf<uint32_t> d_img;
  DevBuf<uint64_t> d_keys, d_sorted;
  DevBuf<uint8_t> d_tmp;
  DevBuf<unsigned long long> d_count;
  struct Cleanup {
    DevBuf<float>&a, &b, &c;
    DevBuf<uint32_t>& d;
    DevBuf<uint64_t>&e, &f;
    DevBuf<uint8_t>& g;
    DevBuf<unsigned long long>& h;
    ~Cleanup() { a.release(); b.release(); c.release(); d.release(); e.release(); f.release(); g.release(); h.release(); }
  } cleanup{d_depth, d_rt, d_w, d_img, d_keys, d_sorted, d_tmp, d_count};
  Q3DR_TRY(d_depth.ensure(n_pix));
  Q3DR_TRY(d_rt.ensure(12));
  Q3DR_TRY(d_keys.ensure(n_pix));
  Q3DR_TRY(d_sorted.ensure(n_pix));
  Q3DR_TRY(d_count.ensure(1));
  if (normal_images) Q3DR_TRY(d_img.ensure(n_pix));
  size_t tmp_bytes = 0;
  Q3DR_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, d_keys.p, d_sorted.p, (int)n_pix, 0, 64, s));
  Q3DR_TRY(d_tmp.ensure(std::max<size_t>(tmp_bytes, 16)));
  Q3DR_CUDA(cudaMemsetAsync(d_count.p, 0, sizeof(unsigned long long), s));

  const Derived dv = derive(*opts);
  q3dr::ScoreParams2 sp;
  std::memset(&sp, 0, sizeof(sp));
  sp.fx = K[0];
  sp.fy = K[1];
  sp.cx = K[2];
  sp.cy = K[3];
  sp.image_width_f = (float)width;
  sp.w = width;
  sp.thr = dv.thr;
  sp.kf = dv.kf;
  sp.a_thr = dv.a_thr;
  sp.ka = dv.ka;
  sp.ignore_sign = opts->incidence_ignore_dot_product_sign;
  sp.img_w = width;
  sp.img_h = height;
  q3dr::MeshDev mesh_dev{};
  q3dr::RenderOptsDev ropts_dev{};
  if (mesh) {
    mesh_dev.nodes = mesh->d_nodes.p;
    mesh_dev.tris = mesh->d_tris.p;
    mesh_dev.normals = mesh->d_normals.p;
    mesh_dev.n_tris = (uint32_t)mesh->n_tris;
    mesh_dev.root_ref = mesh->root_ref;
    mesh_dev.pad = mesh->pad;
    ropts_dev.near_plane = ropts->near_plane;
    ropts_dev.far_plane = ropts->far_plane;
    ropts_dev.clear_argb = ropts->clear_argb;
  }
  q3dr::DownweightParams dp;
  dp.leaf_rec = m->d_leaf_rec.p;
  dp.rt = d_rt.p;
  dp.normal_image = d_img.p;
  dp.n_keys = n_pix;
  dp.factor = invalid_pixel_observation_factor;
  const unsigned blocks = (n_pix + 255) / 256;
  for (size_t v = 0; v < n; ++v) {
    // images follow each other (weights carry over); everything of one image is stream ordered, no host sync
    Q3DR_CUDA(cudaMemcpyAsync(d_rt.p, Rt + 12 * v, 12 * sizeof(float), cudaMemcpyHostToDevice, s));
    Q3DR_CUDA(cudaMemcpyAsync(d_depth.p, depth_maps + v * (size_t)n_pix, n_pix * sizeof(float), cudaMemcpyHostToDevice, s));
    if (normal_images) {
      Q3DR_CUDA(cudaMemcpyAsync(d_img.p, normal_images + v * (size_t)n_pix, n_pix * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    }
    Q3DR_TRY(q3dr_detail::raycast_window_device(m, K, d_rt.p, 0, width, 0, height, max_range, s));
    q3dr::dw_keys_kernel<<<blocks, 256, 0, s>>>(m->d_pixels.p, d_depth.p, n_pix, d_keys.p);
    Q3DR_CUDA(cudaGetLastError());
    Q3DR_CUDA(cub::DeviceRadixSort::SortKeys(d_tmp.p, tmp_bytes, d_keys.p, d_sorted.p, (int)n_pix, 0, 64, s));
    if (mesh) {
      q3dr::dw_apply_kernel<q3dr::kNormalCode><<<blocks, 256, 0, s>>>(sp, dp, d_sorted.p, mesh_dev, ropts_dev, d_count.p);
    } else if (normal_images) {
      q3dr::dw_apply_kernel<q3dr::kNormalImage><<<blocks, 256, 0, s>>>(sp, dp, d_sorted.p, mesh_dev, ropts_dev, d_count.p);
    } else {
      q3dr::dw_apply_kernel<q3dr::kNormalVoxel><<<blocks, 256, 0, s>>>(sp, dp, d_sorted.p, mesh_dev, ropts_dev, d_count.p);
    }
    Q3DR_CUDA(cudaGetLastError());
    // (pageable host sources: the copies above have been staged when cudaMemcpyAsync returns)
  }
  // the host copy of the weights follows (q3dr_map_reset re-uploads from it)
  Q3DR_TRY(d_w.ensure(m->n));
  q3dr::gather_weights_kernel<<<(unsigned)((m->n + 255) / 256), 256, 0, s>>>(m->d_leaf_rec.p, (uint32_t)m->n, d_w.p);
  Q3DR_CUDA(cudaGetLastError());
  Q3DR_CUDA(cudaMemcpyAsync(m->h_weight.data(), d_w.p, m->n * sizeof(float), cudaMemcpyDeviceToHost, s));
  unsigned long long count = 0;
  Q3DR_CUDA(cudaMemcpyAsync(&count, d_count.p, sizeof(count), cudaMemcpyDeviceToHost, s));
  Q3DR_CUDA(cudaStreamSynchronize(s));
  if (weights_out) std::memcpy(weights_out, m->h_weight.data(), m->n * sizeof(float));
  if (n_updates) *n_updates = (uint64_t)count;
  return Q3DR_OK;
}

// ---- one process, several GPUs: the map replicated, viewpoints sharded ----

}  // extern "C"

struct q3dr_multi {
  std::vector<q3dr_map*> maps;
  std::vector<int> devices;
  std::vector<q3dr_result> parts;
  std::vector<uint64_t> first;
  std::vector<q3dr_exchange*> exchanges;  // one per device once q3dr_multi_enable_gather was called
  std::vector<q3dr_mesh*> meshes;  // one per device when image-space normals are on (row N2)
  uint32_t mesh_image_height = 0;
  q3dr_render_opts mesh_ropts{};
};

static void multi_drop_meshes(q3dr_multi* mm) {
  for (q3dr_mesh* me : mm->meshes) q3dr_mesh_destroy(me);
  mm->meshes.clear();
}

extern "C" {

int q3dr_multi_create(const float* boxes, const float* weights, const float* normals, const uint32_t* depth, size_t n,
                      const int* devices, int n_devices, q3dr_multi** out) {
  if (!out) return fail(Q3DR_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  int count = 0;
  Q3DR_TRY(q3dr_device_count(&count));
  if (count <= 0) return fail(Q3DR_ERR_NO_DEVICE, "no CUDA device visible; this engine has no CPU path");
  if (n_devices <= 0) n_devices = count;
  if (n_devices > count && devices == nullptr) return fail(Q3DR_ERR_INVALID_ARG, "more devices requested than visible");
  q3dr_multi* mm = new (std::nothrow) q3dr_multi;
  if (!mm) return fail(Q3DR_ERR_OUT_OF_MEMORY, "host allocation failed");
  int st = Q3DR_OK;
  std::shared_ptr<const q3dr::FlatTree> tree;  // flattened once, uploaded to every device
  for (int i = 0; i < n_devices && st == Q3DR_OK; ++i) {
    const int dev = devices ? devices[i] : i;
    q3dr_map* m = nullptr;
    st = map_create_impl(boxes, weights, normals, depth, n, dev, tree, &m);
    if (st == Q3DR_OK) {
      tree = m->tree;
      mm->maps.push_back(m);
      mm->devices.push_back(dev);
    }
  }
  if (st != Q3DR_OK) {
    const std::string msg = q3dr_detail::last_error();
    for (q3dr_map* m : mm->maps) q3dr_map_destroy(m);
    delete mm;
    return fail(st, msg);
  }
  mm->parts.resize(mm->maps.size());
  mm->first.assign(mm->maps.size() + 1, 0);
  *out = mm;
  return Q3DR_OK;
}

int q3dr_multi_set_normal_mesh(q3dr_multi* mm, const float* tri_vertices, const float* tri_normals, size_t nf,
                               uint32_t image_height, const q3dr_render_opts* ropts) {
  if (!mm) return fail(Q3DR_ERR_INVALID_ARG, "multi is NULL");
  multi_drop_meshes(mm);
  if (!tri_vertices) return Q3DR_OK;
  if (!ropts || image_height == 0) return fail(Q3DR_ERR_INVALID_ARG, "render options / image height missing");
  for (int dev : mm->devices) {
    q3dr_mesh* me = nullptr;
    const int st = q3dr_mesh_create(tri_vertices, tri_normals, nf, dev, &me);
    if (st != Q3DR_OK) {
      const std::string msg = q3dr_detail::last_error();
      multi_drop_meshes(mm);
      return fail(st, msg);
    }
    mm->meshes.push_back(me);
  }
  mm->mesh_image_height = image_height;
  mm->mesh_ropts = *ropts;
  return Q3DR_OK;
}

int q3dr_multi_enable_gather(q3dr_multi* mm, uint64_t max_viewpoints_per_gpu, uint64_t max_entries_per_gpu) {
  if (!mm) return fail(Q3DR_ERR_INVALID_ARG, "multi is NULL");
  for (q3dr_exchange* ex : mm->exchanges) q3dr_exchange_destroy(ex);
  mm->exchanges.clear();
  if (max_viewpoints_per_gpu == 0) return Q3DR_OK;  // switched off
  const uint32_t G = (uint32_t)mm->maps.size();
  int st = Q3DR_OK;
  for (uint32_t g = 0; g < G && st == Q3DR_OK; ++g) {
    q3dr_exchange* ex = nullptr;
    st = q3dr_exchange_create(mm->maps[g], G, g, max_viewpoints_per_gpu, max_entries_per_gpu, &ex);
    if (st == Q3DR_OK) mm->exchanges.push_back(ex);
  }
  if (st == Q3DR_OK) st = q3dr_exchange_connect(mm->exchanges.data(), G);
  if (st != Q3DR_OK) {
    const std::string msg = q3dr_detail::last_error();
    for (q3dr_exchange* ex : mm->exchanges) q3dr_exchange_destroy(ex);
    mm->exchanges.clear();
    return fail(st, msg);
  }
  return Q3DR_OK;
}

int q3dr_multi_gathered(q3dr_multi* mm, uint32_t on, uint32_t src, q3dr_result* out) {
  if (!mm || !out) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  if (mm->exchanges.empty()) return fail(Q3DR_ERR_INVALID_ARG, "q3dr_multi_enable_gather was not called");
  if (on >= mm->exchanges.size()) return fail(Q3DR_ERR_INVALID_ARG, "device index out of range");
  return q3dr_exchange_view(mm->exchanges[on], src, out);
}

int q3dr_multi_destroy(q3dr_multi* mm) {
  if (!mm) return Q3DR_OK;
  for (q3dr_exchange* ex : mm->exchanges) q3dr_exchange_destroy(ex);
  mm->exchanges.clear();
  multi_drop_meshes(mm);
  for (q3dr_map* m : mm->maps) q3dr_map_destroy(m);
  delete mm;
  return Q3DR_OK;
}

int q3dr_multi_device_count(const q3dr_multi* mm) { return mm ? (int)mm->maps.size() : 0; }

q3dr_map* q3dr_multi_map(q3dr_multi* mm, int i) {
  if (!mm || i < 0 || i >= (int)mm->maps.size()) return nullptr;
  return mm->maps[(size_t)i];
}

int q3dr_multi_update_weights(q3dr_multi* mm, const float* weights) {
  if (!mm) return fail(Q3DR_ERR_INVALID_ARG, "multi is NULL");
  for (q3dr_map* m : mm->maps) Q3DR_TRY(q3dr_map_update_weights(m, weights));
  return Q3DR_OK;
}

int q3dr_multi_score_viewpoints(q3dr_multi* mm, const float K[4], uint32_t image_width, uint32_t x0, uint32_t x1,
                                uint32_t y0, uint32_t y1, const float* Rt, size_t n, const q3dr_score_opts* opts,
                                q3dr_multi_result* result) {
  if (!mm || !result || !K || !opts || (!Rt && n)) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  const size_t G = mm->maps.size();
  // contiguous slices, sizes differing by at most one
  for (size_t g = 0; g <= G; ++g) mm->first[g] = (uint64_t)((n / G) * g + std::min<size_t>(g, n % G));
  std::vector<int> status(G, Q3DR_OK);
  std::vector<std::string> message(G);
  std::vector<std::thread> workers;
  for (size_t g = 0; g < G; ++g) {
    workers.emplace_back([&, g]() {
      const size_t lo = (size_t)mm->first[g], hi = (size_t)mm->first[g + 1];
      q3dr_exchange* ex = mm->exchanges.empty() ? nullptr : mm->exchanges[g];
      if (ex && q3dr_exchange_begin_step(ex) != Q3DR_OK) ex = nullptr;  // (cannot fail once connected)
      if (!mm->meshes.empty()) {
        status[g] = q3dr_score_viewpoints_mesh_normals(mm->maps[g], mm->meshes[g], &mm->mesh_ropts, K, image_width,
                                                       mm->mesh_image_height, x0, x1, y0, y1, Rt + 12 * lo, hi - lo, opts,
                                                       &mm->parts[g]);
      } else {
        status[g] = q3dr_score_viewpoints(mm->maps[g], K, image_width, x0, x1, y0, y1, Rt + 12 * lo, hi - lo, opts, &mm->parts[g]);
      }
      if (status[g] != Q3DR_OK) message[g] = q3dr_last_error();  // thread-local: carry it out of the worker
      if (ex) {
        // every worker takes part, failed or not, so that nobody waits for flags that never come
        const int est = q3dr_exchange_end_step(ex, status[g], mm->maps[g]->own_stream);
        if (status[g] == Q3DR_OK && est != Q3DR_OK) {
          status[g] = est;
          message[g] = q3dr_last_error();
        }
      }
    });
  }
  for (std::thread& t : workers) t.join();
  for (size_t g = 0; g < G; ++g) {
    if (status[g] != Q3DR_OK) return fail(status[g], "device " + std::to_string(mm->devices[g]) + ": " + message[g]);
  }
  result->n_parts = (uint32_t)G;
  result->n_viewpoints = n;
  result->first_viewpoint = mm->first.data();
  result->parts = mm->parts.data();
  return Q3DR_OK;
}

}  // extern "C"
