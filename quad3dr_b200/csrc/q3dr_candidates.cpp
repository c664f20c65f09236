// q3dr_candidates.cpp -- batched candidate generation for the viewpoint graph.
//
// The reference adds candidates one at a time (ViewpointPlanner::tryToAddViewpointEntry,
// viewpoint_planner/src/planner/viewpoint_planner_graph.cpp:542-655): valid object position -> distance gate against the
// k nearest viewpoints already in the graph -> raycast + score -> minimum voxel count / minimum information -> add.
// Every accepted candidate enters the graph at once and gates the candidates after it.  Because the accept decision is
// a conjunction, the expensive part can be batched without changing a single decision:
//   1. validity of all candidate positions in one overlap query            (q3dr_valid_object_positions, GPU)
//   2. raycast + score of ALL valid candidates in one engine call          (GPU; rays are cheaper than the gate)
//   3. one sequential pass over the candidates in order on the host: the reference's distance gate against the
//      viewpoints accepted so far (exact k nearest neighbours), then the two thresholds on the precomputed results.
// The reference's gate uses FLANN through bh::ApproximateNearestNeighbor (third party, approximate, absent here); this
// restates it with exact nearest neighbours, ascending squared distance, ties by insertion order.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <unordered_map>
#include <vector>

#include "q3dr_internal.cuh"

namespace {

// ViewpointPlanner::computeExplorationStep (viewpoint_planner_graph.cpp:672-690) with BoundingBox3D::isOutside /
// distanceTo (geometry.h:235-238,257-272)
float exploration_step_at(const float p[3], const q3dr_candidate_opts& o) {
  float step = o.exploration_step;
  bool outside = false;
  for (int i = 0; i < 3; ++i) outside = outside || p[i] < o.exploration_bbox[i] || p[i] > o.exploration_bbox[3 + i];
  if (outside) {
    float dist_sq = 0;
    for (int i = 0; i < 3; ++i) {
      if (p[i] < o.exploration_bbox[i] || p[i] > o.exploration_bbox[3 + i]) {
        const float d = std::min(std::fabs(p[i] - o.exploration_bbox[i]), std::fabs(p[i] - o.exploration_bbox[3 + i]));
        dist_sq += d * d;
      }
    }
    const float distance_to_bbox = std::sqrt(dist_sq);
    const float dilation_distance = distance_to_bbox - o.exploration_dilation_start_bbox_distance;
    if (dilation_distance > 0) {
      if (o.exploration_dilation_squared) {
        step += o.exploration_dilation_speed * dilation_distance * dilation_distance;
      } else {
        step += o.exploration_dilation_speed * dilation_distance;
      }
    }
  }
  if (step > o.exploration_step_max) step = o.exploration_step_max;
  return step;
}

// Eigen::QuaternionBase::angularDistance (3.3): d = a * conj(b); 2 * atan2(|d.vec|, |d.w|).  q = (w, x, y, z).
float angular_distance(const float a[4], const float b[4]) {
  const float bw = b[0], bx = -b[1], by = -b[2], bz = -b[3];
  const float w = a[0] * bw - a[1] * bx - a[2] * by - a[3] * bz;
  const float x = a[0] * bx + a[1] * bw + a[2] * bz - a[3] * by;
  const float y = a[0] * by + a[2] * bw + a[3] * bx - a[1] * bz;
  const float z = a[0] * bz + a[3] * bw + a[1] * by - a[2] * bx;
  return 2.0f * std::atan2(std::sqrt(x * x + (y * y + z * z)), std::fabs(w));
}

// The accepted viewpoints in a uniform grid, for exact k-nearest-neighbour queries inside a radius.
struct PointGrid {
  float cell;
  std::unordered_map<uint64_t, std::vector<uint32_t>> cells;
  std::vector<float> pos;   // 3 per point
  std::vector<float> quat;  // 4 per point
  static int64_t coord(float v, float cell) { return (int64_t)std::floor((double)v / (double)cell); }
  static uint64_t key(int64_t x, int64_t y, int64_t z) {
    return ((uint64_t)(x & 0x1FFFFF) << 42) | ((uint64_t)(y & 0x1FFFFF) << 21) | (uint64_t)(z & 0x1FFFFF);
  }
  void add(const float p[3], const float q[4]) {
    const uint32_t idx = (uint32_t)(pos.size() / 3);
    pos.insert(pos.end(), p, p + 3);
    quat.insert(quat.end(), q, q + 4);
    cells[key(coord(p[0], cell), coord(p[1], cell), coord(p[2], cell))].push_back(idx);
  }
  size_t size() const { return pos.size() / 3; }
  static float dist_sq(const float* a, const float* b) {  // FLANN's L2: sequential accumulation
    float r = 0;
    for (int i = 0; i < 3; ++i) {
      const float d = a[i] - b[i];
      r += d * d;
    }
    return r;
  }
  // all points with dist_sq < radius_sq, as (dist_sq, index)
  void within(const float p[3], float radius_sq, std::vector<std::pair<float, uint32_t>>* out) const {
    out->clear();
    const float radius = std::sqrt(radius_sq);
    const double span = (double)radius / (double)cell;
    if (!(span < 6.0)) {  // huge radius (far outside the exploration box): every point is a candidate
      for (uint32_t i = 0; i < size(); ++i) {
        const float d = dist_sq(p, &pos[3 * i]);
        if (d < radius_sq) out->push_back(std::make_pair(d, i));
      }
      return;
    }
    const int64_t x0 = coord(p[0] - radius, cell), x1 = coord(p[0] + radius, cell);
    const int64_t y0 = coord(p[1] - radius, cell), y1 = coord(p[1] + radius, cell);
    const int64_t z0 = coord(p[2] - radius, cell), z1 = coord(p[2] + radius, cell);
    for (int64_t x = x0; x <= x1; ++x)
      for (int64_t y = y0; y <= y1; ++y)
        for (int64_t z = z0; z <= z1; ++z) {
          auto it = cells.find(key(x, y, z));
          if (it == cells.end()) continue;
          for (uint32_t i : it->second) {
            const float d = dist_sq(p, &pos[3 * i]);
            if (d < radius_sq) out->push_back(std::make_pair(d, i));
          }
        }
  }
};

}  // namespace

extern "C" {

void q3dr_candidate_opts_default(q3dr_candidate_opts* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  for (int i = 0; i < 3; ++i) {
    o->object_bbox[i] = -0.5f;
    o->object_bbox[3 + i] = 0.5f;
    o->bvh_bbox[i] = -FLT_MAX;
    o->bvh_bbox[3 + i] = FLT_MAX;
    o->exploration_bbox[i] = -FLT_MAX;
    o->exploration_bbox[3 + i] = FLT_MAX;
  }
  o->obstacle_free_height = FLT_MAX;
  o->discard_dist_knn = 10;
  o->exploration_step = 3.0f;
  o->exploration_step_max = FLT_MAX;
  o->exploration_dilation_squared = 1;
  o->exploration_dilation_start_bbox_distance = 0.0f;
  o->exploration_dilation_speed = 2.0f;
  o->exploration_angular_dist_threshold_degrees = 30.0f;
  o->min_voxel_count = 100;
  o->min_information = 10.0f;
}

int q3dr_generate_viewpoint_entries(q3dr_map* m, const float K[4], uint32_t image_width, uint32_t x0, uint32_t x1, uint32_t y0,
                                    uint32_t y1, const float* cand_pos, const float* cand_quat, size_t n,
                                    const float* existing_pos, const float* existing_quat, size_t n_existing,
                                    const q3dr_score_opts* opts, const q3dr_candidate_opts* copts, uint8_t* status_out,
                                    q3dr_result* result) {
  Q3DR_TRY(q3dr_detail::check_map(m));
  if (!K || !opts || !copts || !result || (n && (!cand_pos || !cand_quat || !status_out)) ||
      (n_existing && (!existing_pos || !existing_quat))) {
    return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  }
  if (!(copts->exploration_step > 0)) return fail(Q3DR_ERR_INVALID_ARG, "exploration_step must be positive");
  try {
    // 1. isValidObjectPosition for every candidate position
    std::vector<uint8_t> valid(n);
    Q3DR_TRY(q3dr_valid_object_positions(m, cand_pos, n, copts->object_bbox, copts->bvh_bbox, copts->obstacle_free_height,
                                         valid.data()));
    // 2. raycast + score every valid candidate in one call (ignore_voxels_with_zero_information = true at this call
    //    site, viewpoint_planner_graph.cpp:591-593 -- the caller's opts say so); invalid ones get an empty entry
    std::vector<uint32_t> slot(n, 0xFFFFFFFFu);
    std::vector<float> rt;
    rt.reserve(12 * n);
    uint32_t n_cast = 0;
    for (size_t j = 0; j < n; ++j) {
      if (!valid[j]) continue;
      float r[12];
      Q3DR_TRY(q3dr_pose_to_rt(cand_quat + 4 * j, cand_pos + 3 * j, r));  // Pose::createFromImageToWorldTransformation
      rt.insert(rt.end(), r, r + 12);
      slot[j] = n_cast++;
    }
    q3dr_result cast;
    Q3DR_TRY(q3dr_score_viewpoints(m, K, image_width, x0, x1, y0, y1, rt.data(), n_cast, opts, &cast));
    // 3. the reference's sequential decisions, candidate after candidate
    PointGrid grid;
    grid.cell = copts->exploration_step;
    for (size_t e = 0; e < n_existing; ++e) grid.add(existing_pos + 3 * e, existing_quat + 4 * e);
    const float ang_thr = copts->exploration_angular_dist_threshold_degrees * float(M_PI) / float(180.0);
    std::vector<std::pair<float, uint32_t>> near;
    for (size_t j = 0; j < n; ++j) {
      if (!valid[j]) {
        status_out[j] = Q3DR_CANDIDATE_INVALID_POSITION;
        continue;
      }
      const float* p = cand_pos + 3 * j;
      const float* q = cand_quat + 4 * j;
      float step = exploration_step_at(p, *copts);
      // only neighbours closer than sqrt(0.6) * step can discard, the threshold never grows along the list, and the
      // list is in ascending distance: the k nearest INSIDE that radius are the k nearest that matter
      grid.within(p, 0.6f * step * step, &near);
      std::sort(near.begin(), near.end());
      bool discard = false;
      const size_t k = std::min<size_t>(near.size(), copts->discard_dist_knn);
      for (size_t i = 0; i < k && !discard; ++i) {
        const uint32_t nb = near[i].second;
        step = std::min(step, exploration_step_at(&grid.pos[3 * nb], *copts));
        const float thr = 0.6f * step * step;
        const float ang = angular_distance(q, &grid.quat[4 * nb]);
        if (near[i].first < thr && ang < ang_thr) discard = true;
      }
      if (discard) {
        status_out[j] = Q3DR_CANDIDATE_TOO_CLOSE;
        continue;
      }
      const uint32_t s = slot[j];
      const uint64_t voxels = cast.offsets[s + 1] - cast.offsets[s];
      if (voxels < copts->min_voxel_count) {
        status_out[j] = Q3DR_CANDIDATE_TOO_FEW_VOXELS;
        continue;
      }
      if (cast.total[s] < copts->min_information) {
        status_out[j] = Q3DR_CANDIDATE_TOO_LITTLE_INFORMATION;
        continue;
      }
      status_out[j] = Q3DR_CANDIDATE_ACCEPTED;
      grid.add(p, q);
    }
    // the CSR result, re-indexed by candidate: invalid candidates own an empty range
    m->cand_offsets.assign(n + 1, 0);
    m->cand_total.assign(std::max<size_t>(n, 1), 0.0f);
    m->cand_hits.assign(std::max<size_t>(n, 1), 0u);
    uint64_t at = 0;
    for (size_t j = 0; j < n; ++j) {
      m->cand_offsets[j] = at;
      if (slot[j] != 0xFFFFFFFFu) {
        const uint32_t s = slot[j];
        at = cast.offsets[s + 1];
        m->cand_total[j] = cast.total[s];
        m->cand_hits[j] = cast.hit_count[s];
      }
    }
    m->cand_offsets[n] = at;
    *result = cast;
    result->n_viewpoints = n;
    result->offsets = m->cand_offsets.data();
    result->total = m->cand_total.data();
    result->hit_count = m->cand_hits.data();
  } catch (const std::bad_alloc&) {
    return fail(Q3DR_ERR_OUT_OF_MEMORY, "host allocation failed");
  }
  return Q3DR_OK;
}

}  // extern "C"
