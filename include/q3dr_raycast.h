/*
 * q3dr_raycast.h -- C ABI of the B200-native viewpoint raycast + scoring engine.
 *
 * Drop-in boundary for the one data-parallel hot path of Quad3DR.  Plain
 * pointers and sizes only; every entry point returns an int status (Q3DR_OK == 0) and never throws.
 * The library runs on sm_100a CUDA kernels only: there is no CPU fallback, and every compute entry
 * point fails with Q3DR_ERR_NO_DEVICE when no CUDA device is usable.
 *
 * Reference interfaces replaced (paths relative to the reference checkout):
 *   q3dr_map_create / q3dr_map_destroy
 *        bvh::Tree::ensureCudaTreeIsInitialized + CudaTree::createCopyFromHostTree, Tree::clear
 *        viewpoint_planner/src/bvh/bvh.h:949-955,296-314 ; viewpoint_planner/src/bvh/bvh.cu:25-83
 *   q3dr_map_update_weights
 *        ViewpointPlannerData::updateWeights* writing NodeObject::weight
 *        viewpoint_planner/src/planner/viewpoint_planner_data.cpp:476-482,573-577
 *   q3dr_raycast_window
 *        bvh::Tree::raycastCuda / raycastWithScreenCoordinatesCuda
 *        viewpoint_planner/src/bvh/bvh.h:396-451,453-509 ; bvh.cu:583-689
 *   q3dr_intersect_rays
 *        bvh::Tree::intersectsCuda                         bvh.h:512-541 ; bvh.cu:546-579
 *   q3dr_raycast_unique
 *        ViewpointRaycast::getRaycastHitVoxels[WithScreenCoordinates]Cuda (+ removeDuplicateRaycastHitVoxels)
 *        viewpoint_planner/src/planner/viewpoint_raycast.cpp:221-283,305-335
 *   q3dr_score_viewpoints / q3dr_score_viewpoints_device
 *        ViewpointPlanner::getRaycastHitVoxelsWithInformationScore, batched over viewpoints
 *        viewpoint_planner/src/planner/viewpoint_planner_raycast.cpp:89-148,
 *        viewpoint_planner/src/planner/viewpoint_planner_scoring.cpp:11-38,67-73,116-122,
 *        viewpoint_planner/src/planner/viewpoint_score.cpp:45-68
 *   q3dr_splitmedian_order
 *        bvh::Tree::build -> splitMedian (defines the tie-break order)     bvh.h:361-373,764-788
 *   q3dr_pose_to_rt
 *        Pose::getTransformationImageToWorld                               include/bh/pose.h:122-127
 *   q3dr_mesh_create / q3dr_mesh_render_normals / q3dr_decode_normal
 *        ViewpointOffscreenRenderer::uploadPoissonMesh, drawPoissonMeshNormals, computePoissonMeshNormalVector
 *        viewpoint_planner/src/planner/viewpoint_offscreen_renderer.cpp:138-200,276-293,361-397,499-527 ;
 *        viewpoint_planner/src/shaders/triangles_normals.{v,f}.glsl
 *   q3dr_generate_viewpoint_entries
 *        ViewpointPlanner::tryToAddViewpointEntry, batched        viewpoint_planner_graph.cpp:542-655
 *   q3dr_octree_read / q3dr_octree_leaves
 *        octomap::AbstractOcTree::read + the begin_tree() walk of generateBVHTree
 *        viewpoint_planner/src/octree/occupancy_map.hxx:1123-1127 ; viewpoint_planner_data.cpp:820-838
 *   q3dr_map_downweight_real_viewpoints
 *        ViewpointPlannerData::updateWeightsWithRealViewpoints        viewpoint_planner_data.cpp:498-571
 *   q3dr_exchange_* / q3dr_multi_enable_gather
 *        no counterpart (the reference is single-GPU, viewpoint_planner_data.cpp:328-335): the gather of the sharded
 *        candidate set, SURVEY.md section 8e
 *   q3dr_score_viewpoints_mesh_normals / q3dr_score_viewpoints_normal_images
 *        getRaycastHitVoxelsWithInformationScore with options_.enable_opengl = true: computeNormalVector takes the
 *        normal from the mesh-normal image at the voxel's kept pixel
 *        viewpoint_planner/src/planner/viewpoint_planner_scoring.cpp:26-38
 *
 * Numeric contract: strict IEEE-754 binary32 in the reference's operation
 * order, no FMA contraction.  hit(ray) = leaf of least dist_sq among leaves the ray's slab test
 * accepts (dist_sq <= max_range^2; origin inside a leaf => dist_sq 0); equal dist_sq => the leaf
 * with the HIGHER INDEX wins.  Callers therefore pass leaves in the left-to-right leaf order of the
 * reference's bvh::Tree (what a walk of the host tree yields, or q3dr_splitmedian_order computes),
 * so that "higher index" == "visited later by the reference's depth-first recursion".
 */
#ifndef Q3DR_RAYCAST_H_
#define Q3DR_RAYCAST_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define Q3DR_ABI_VERSION 5

enum q3dr_status {
  Q3DR_OK = 0,
  Q3DR_ERR_INVALID_ARG = 1, /* null pointer, empty map, bad window ... (BH_EXCEPTION class) */
  Q3DR_ERR_CUDA = 2,        /* a CUDA call or kernel failed (bh::CudaError / Tree::CudaError class) */
  Q3DR_ERR_NO_DEVICE = 3,   /* no usable CUDA device: the engine has no CPU path */
  Q3DR_ERR_OUT_OF_MEMORY = 4,
  Q3DR_ERR_INTERNAL = 5
};

typedef struct q3dr_map q3dr_map;

/* Same fields, meaning and defaults as ViewpointScore::Options (viewpoint_score.h:17-41) plus the
 * raycast ranges of ViewpointPlanner::Options (viewpoint_planner.h:201-202). */
typedef struct q3dr_score_opts {
  float voxel_sensor_size_ratio_threshold;          /* 0.002 */
  float voxel_sensor_size_ratio_inv_falloff_factor; /* 0.001 */
  float incidence_angle_threshold_degrees;          /* 30 */
  float incidence_angle_inv_falloff_factor_degrees; /* 30 */
  int32_t incidence_ignore_dot_product_sign;        /* 0 */
  int32_t ignore_voxels_with_zero_information;      /* 1 */
  float raycast_min_range;                          /* 0; accepted and ignored like the reference */
  float raycast_max_range;                          /* 60; <= 0 means unlimited */
} q3dr_score_opts;

/* One pixel of raycast*Cuda (bvh::Tree::IntersectionResultWithScreenCoordinates, bvh.h:162-182). */
typedef struct q3dr_pixel_hit {
  float intersection[3]; /* hit point; 0 for a miss */
  int32_t leaf;          /* leaf index, -1 for a miss (reference: node == nullptr) */
  uint32_t depth;        /* depth of the leaf in the reference's bvh::Tree; 0 for a miss */
  float dist_sq;         /* 0 for a miss */
  float screen_x;        /* pixel x, y as floats (viewpoint_raycast.cpp:203) */
  float screen_y;
} q3dr_pixel_hit;

/* Per-viewpoint hit lists + scores in CSR form.  Viewpoint v owns entries [offsets[v], offsets[v+1]).
 * Entries are sorted by ascending leaf index.  Host arrays are owned by the map and stay valid until
 * the next scoring call on the same map or q3dr_map_destroy. */
typedef struct q3dr_result {
  uint64_t n_viewpoints;
  uint64_t n_entries;          /* == offsets[n_viewpoints] */
  const uint64_t* offsets;     /* n_viewpoints + 1 */
  const int32_t* leaf;         /* voxel (leaf index) */
  const float* information;    /* its observation score */
  const uint32_t* first_pixel; /* row-major index inside the window of the first pixel that hit it */
  const float* dist_sq;        /* dist_sq of that kept pixel */
  const float* total;          /* n_viewpoints: sum of information (fp64 accumulate, rounded once) */
  const uint32_t* hit_count;   /* n_viewpoints: unique voxels hit before the information > 0 filter */
} q3dr_result;

/* The two large host arrays of a result (leaf, information) handed over to the caller, see q3dr_result_detach_host. */
typedef struct q3dr_host_block {
  int32_t* leaf;
  float* information;
  uint64_t n_entries; /* valid entries */
  uint64_t capacity;  /* entries allocated */
} q3dr_host_block;

typedef struct q3dr_map_info {
  uint64_t n_leaves;
  uint64_t n_nodes;       /* 4-wide nodes in the flattened tree */
  uint32_t tree_depth;    /* depth of the flattened 4-wide tree */
  uint32_t ref_depth;     /* depth of the reference's splitMedian tree, ceil(log2(n)) */
  uint64_t node_bytes;    /* bytes of the node array in HBM */
  uint64_t device_bytes;  /* all device memory currently held by the map */
  int32_t device;
  int32_t sm_count;
} q3dr_map_info;

/* Kernel statistics of the last scoring / raycast call on a map. */
typedef struct q3dr_stats {
  uint64_t rays;             /* rays cast */
  uint64_t kernel_launches;  /* kernels of this library launched */
  float raycast_ms;          /* device time of the traversal kernel(s), CUDA events on the launch stream */
  float score_ms;            /* device time of the dedup/score/compaction kernels */
  float total_ms;            /* device time of the whole call (first launch to last) */
  uint32_t chunks;
  uint32_t viewpoints_per_chunk;
} q3dr_stats;

int q3dr_abi_version(void);
const char* q3dr_status_string(int status);
/* Message of the last failing call on this thread (empty string if none). */
const char* q3dr_last_error(void);
int q3dr_device_count(int* count);

void q3dr_score_opts_default(q3dr_score_opts* opts);

/* ---- host-side helpers (no device work) ---- */

/* Leaf order of the reference's Tree::build for boxes (n x 6: min xyz, max xyz) given in the order
 * they would be handed to build(): order[k] = input index of the k-th leaf, left to right.
 * depth (nullable): depth[k] = depth of that leaf in the reference tree. */
int q3dr_splitmedian_order(const float* boxes, size_t n, uint32_t* order, uint32_t* depth);

/* out[k] = position of the k-th leaf (left to right) in the reference's all-node iterator (bvh.h:226-235) = the
 * voxel index its viewpoint-graph serialisation stores (viewpoint_planner_serialization.h:88-105). */
int q3dr_voxel_indices(size_t n, uint32_t* out);

/* Quaternion (w, x, y, z) + translation -> row-major 3x4 [R|t], image-to-world. */
int q3dr_pose_to_rt(const float quat_wxyz[4], const float trans[3], float rt_out[12]);

/* Host-only diagnostic (no device): flattens the leaves like q3dr_map_create and verifies the tree's
 * invariants; reports its size.  Any out pointer may be NULL. */
int q3dr_debug_flat_tree(const float* boxes, size_t n, int top_levels, uint64_t* n_nodes, uint32_t* depth,
                         uint32_t* max_stack, uint32_t* top_nodes);

/* ---- occupancy map: flatten + upload ---- */

/* boxes   n x 6 floats, leaf k has tie-break priority k (see header comment).  Boxes should have positive extent on
 *         every axis, as the reference guarantees (generateBVHTree drops isEmpty() boxes): for a zero-thickness box a
 *         ray exactly parallel to an axis AND lying in the box's plane sends 0 * inf = NaN through the reference's
 *         order-dependent min / max, and the reference's own answer then depends on the shape of its tree.
 * weights n floats or NULL (=> 1).   normals n x 3 floats or NULL (=> zero vectors).
 * depth   n uint32 (depth of each leaf in the reference tree, only echoed in q3dr_pixel_hit.depth)
 *         or NULL (=> the splitMedian depth implied by n and k). */
int q3dr_map_create(const float* boxes, const float* weights, const float* normals, const uint32_t* depth,
                    size_t n, int device, q3dr_map** out);
/* (Every scoring / raycast entry point returns with its stream synchronised and keeps no reference to a caller's stream
 * afterwards, so updates and follow-up queries never race with work of an earlier call.) */
int q3dr_map_update_weights(q3dr_map* map, const float* weights);
int q3dr_map_update_normals(q3dr_map* map, const float* normals);
int q3dr_map_get_info(const q3dr_map* map, q3dr_map_info* info);
int q3dr_map_last_stats(const q3dr_map* map, q3dr_stats* stats);
/* Debugging aid.  With Q3DR_GUARD=1 in the environment (read once, before the first allocation) every device buffer of
 * a map is followed by 256 guard bytes of a known pattern; this call synchronises the device and reports how many
 * buffers had their guard overwritten (a write past the end).  Without Q3DR_GUARD it fails with INVALID_ARG. */
int q3dr_map_check_guards(q3dr_map* map, int* overwritten);
/* Frees and re-uploads all device state (the reference's recovery after a CUDA error, bvh.h:443-449). */
int q3dr_map_reset(q3dr_map* map);
int q3dr_map_destroy(q3dr_map* map);

/* ---- compat layer: one viewpoint, per-pixel results ---- */

/* K = {fx, fy, cx, cy}; Rt = row-major 3x4 [R|t]; window [x0,x1) x [y0,y1) in pixels.
 * out: (x1-x0)*(y1-y0) records, row-major, HOST memory. */
int q3dr_raycast_window(q3dr_map* map, const float K[4], const float Rt[12], uint32_t x0, uint32_t x1, uint32_t y0,
                        uint32_t y1, float min_range, float max_range, q3dr_pixel_hit* out);

/* origins, directions: n x 3 floats (directions are normalised like bh::Ray).  out: n records;
 * screen_x / screen_y are 0. */
int q3dr_intersect_rays(q3dr_map* map, const float* origins, const float* directions, size_t n, float min_range,
                        float max_range, q3dr_pixel_hit* out);

/* getRaycastHitVoxelsWithScreenCoordinates(remove_duplicates = true): one record per unique voxel,
 * the data of the first pixel (row-major) that hit it, ascending leaf index.  out capacity `cap`
 * records; *count receives the number of unique voxels (call fails with INVALID_ARG if > cap). */
int q3dr_raycast_unique(q3dr_map* map, const float K[4], const float Rt[12], uint32_t x0, uint32_t x1, uint32_t y0,
                        uint32_t y1, float min_range, float max_range, q3dr_pixel_hit* out, size_t cap,
                        size_t* count);

/* ---- the hot path: batched viewpoint scoring ---- */

/* Rt: n x 12 floats in HOST memory.  image_width is the camera width used by the resolution factor
 * (computeRelativeSizeOnSensorHorizontal); the window is [x0,x1) x [y0,y1).  result: filled in. */
int q3dr_score_viewpoints(q3dr_map* map, const float K[4], uint32_t image_width, uint32_t x0, uint32_t x1,
                          uint32_t y0, uint32_t y1, const float* Rt, size_t n, const q3dr_score_opts* opts,
                          q3dr_result* result);

/* Same with poses already resident in device memory (d_Rt) and results left in device memory
 * (pointers in `result` are DEVICE pointers owned by the map).  stream: a cudaStream_t (0 = default). */
int q3dr_score_viewpoints_device(q3dr_map* map, const float K[4], uint32_t image_width, uint32_t x0, uint32_t x1,
                                 uint32_t y0, uint32_t y1, const float* d_Rt, size_t n,
                                 const q3dr_score_opts* opts, void* stream, q3dr_result* result);

/* Chunk notification.  A scoring call works through its viewpoints in chunks; as soon as the kernels that finish a
 * chunk have been ISSUED on the call's stream, `cb` is invoked on the calling host thread with the viewpoint range and
 * the entry range of the chunk and the device pointers of the partial CSR result (offsets[0 .. first + n], total,
 * hit_count, leaf / information / first_pixel / dist_sq [0 .. first_entry + n_entries) are final once the work issued
 * so far on the stream has run).  A consumer orders its own work after an event recorded on that stream -- e.g.
 * copy-engine pushes of the chunk into the other GPUs' memory while the next chunk is being cast
 * (quad3dr_b200/dist.py: PeerGather).  The pointers may change between calls (pools grow), not between the chunks of
 * a call whose pools already have the needed capacity; a pool that had to grow is REPLACED, and the replaced block
 * stays allocated and intact until the next scoring call on the map begins, so work a consumer enqueued on an earlier
 * chunk's pointers remains valid for the whole call.  cb == NULL removes the callback. */
typedef void (*q3dr_chunk_callback)(void* user, uint64_t first_viewpoint, uint64_t n_viewpoints, uint64_t first_entry,
                                    uint64_t n_entries, const q3dr_result* partial_device_result);
int q3dr_map_set_chunk_callback(q3dr_map* map, q3dr_chunk_callback cb, void* user);

/* Copies the device-resident result of the last q3dr_score_viewpoints_device call to host memory
 * owned by the map. */
int q3dr_result_to_host(q3dr_map* map, q3dr_result* result);

/* Ownership hand-over of the host arrays `leaf` and `information` of the last result (after q3dr_score_viewpoints* with
 * host buffers or q3dr_result_to_host): the map forgets the two pinned arrays -- they stay valid until
 * q3dr_host_block_release, whatever is called on the map meanwhile, and may outlive it -- and takes others for its next
 * call (released blocks are kept in a small process-wide pool, so a caller that keeps a few results alive and drops
 * them again causes no further allocation).  The C++ shim's VoxelWithInformationSet views rest on this: the 8 bytes
 * per voxel the device produced are never copied again on the host.  offsets / total / hit_count of the q3dr_result stay
 * with the map (small).  Replaces the conversion loop of bvh.h:430,487 / viewpoint_planner_raycast.cpp:116-148. */
int q3dr_result_detach_host(q3dr_map* map, q3dr_host_block* out);
int q3dr_host_block_release(q3dr_host_block* block); /* block may be all zeros; it is zeroed on return */
int q3dr_host_pool_trim(void); /* frees the released blocks the pool still holds (at most three); returns Q3DR_OK */

/* Which per-entry arrays the scoring calls on this map deliver besides (leaf, information).  The reference's facade
 * (getRaycastHitVoxelsWithInformationScore -> VoxelWithInformationSet, viewpoint_planner_raycast.cpp:116-148) consumes
 * voxel + information only: with fields == 0 a result entry costs 8 instead of 16 bytes of D2H / peer traffic and the
 * kept pixel's dist_sq is not recomputed.  Arrays that are switched off come back as NULL.  Default: both on. */
#define Q3DR_RESULT_FIRST_PIXEL 1u
#define Q3DR_RESULT_DIST_SQ 2u
int q3dr_map_set_result_fields(q3dr_map* map, uint32_t fields);

/* ---- consumers of the voxel sets: the planner's path search ----
 *
 * They work on the CSR result of the LAST q3dr_score_viewpoints[_device] call on the map, in device memory;
 * `viewpoint` arguments index that result.  A q3dr_observed is one ViewpointPath::observed_voxel_map
 * (viewpoint_planner.h: unordered_map<voxel, accumulated information>) held as a dense per-leaf array on the device.
 * Per-voxel arithmetic is the reference's; sums over a set are fp64 sums rounded once (the reference sums in hash
 * order, fp32). */
typedef struct q3dr_observed q3dr_observed;

int q3dr_observed_create(q3dr_map* map, q3dr_observed** out); /* empty map */
int q3dr_observed_clear(q3dr_observed* obs);
int q3dr_observed_destroy(q3dr_observed* obs);
/* out: n_leaves floats, NaN = voxel not in the map */
int q3dr_observed_download(const q3dr_observed* obs, float* out);

/* ViewpointPlanner::computeNewInformation (viewpoint_planner_scoring.cpp:124-147) for n viewpoints of the result
 * (viewpoints == NULL: viewpoints 0..n-1):  sum over the voxel set of
 *   observed ? min(f * info, weight - observed) : f * info,      f = viewpoint_information_factor (1/3). */
int q3dr_new_information(q3dr_map* map, const q3dr_observed* obs, float viewpoint_information_factor,
                         const uint32_t* viewpoints, size_t n, float* out);

/* addObservedVoxelsToViewpointPath (viewpoint_planner_path.cpp:203-233): folds the voxel set of `viewpoint` into
 * the map (insert f * info, or add it clamped at the voxel weight) and returns the novel information. */
int q3dr_observed_add_viewpoint(q3dr_map* map, q3dr_observed* obs, float viewpoint_information_factor,
                                uint32_t viewpoint, float* new_information);

/* The same two operations for a voxel set handed over explicitly (host arrays, each voxel at most once, any order):
 * a stored ViewpointEntry::voxel_set that is not part of the resident result.  add != 0: fold it into the map
 * (addObservedVoxelsToViewpointPath); add == 0: only evaluate it (computeNewInformation). */
int q3dr_observed_voxel_set(q3dr_map* map, q3dr_observed* obs, float viewpoint_information_factor, const int32_t* leaf,
                            const float* information, size_t n, int add, float* new_information);

/* compute_overlap_information_lambda (viewpoint_planner_path.cpp:838-848) of each viewpoint others[i] (set1)
 * against viewpoint `ref` (set2).  Like the reference's lookup (VoxelWithInformation::operator==,
 * occupied_tree.h:73-79) a voxel counts only if it carries bit-identical information in both sets. */
int q3dr_overlap_information(q3dr_map* map, uint32_t ref, const uint32_t* others, size_t n, float* out);

/* ---- box-vs-map overlap queries ---- */

/* bvh::Tree::intersects(bbox) (bvh.h:544-554,911-939) for n boxes (n x 6: min xyz, max xyz): per box the leaves
 * whose box overlaps it (touching counts, geometry.h:277-285), ascending leaf index = the reference's visiting
 * order.  CSR arrays are host memory owned by the map, valid until the next overlap call on it. */
typedef struct q3dr_overlap_result {
  uint64_t n_boxes;
  uint64_t n_entries;
  const uint64_t* offsets; /* n_boxes + 1 */
  const int32_t* leaf;
} q3dr_overlap_result;

int q3dr_overlap_boxes(q3dr_map* map, const float* boxes, size_t n, q3dr_overlap_result* result);

/* ViewpointPlannerData::isValidObjectPosition (viewpoint_planner_data.cpp:209-235) for n positions (n x 3), without
 * the no-fly-zone polygons (test those on the host first).  object_bbox: the drone's box relative to its position;
 * bvh_bbox: ViewpointPlannerData::bvh_bbox_.  valid[i] = 1 / 0. */
int q3dr_valid_object_positions(q3dr_map* map, const float* positions, size_t n, const float object_bbox[6],
                                const float bvh_bbox[6], float obstacle_free_height, uint8_t* valid);

/* ---- batched candidate generation ----
 *
 * ViewpointPlanner::tryToAddViewpointEntry (viewpoint_planner_graph.cpp:542-655) for n candidate poses IN ORDER, with
 * the decisions of the reference's one-at-a-time loop: a candidate is accepted iff its position is valid
 * (isValidObjectPosition), it is not too close to a viewpoint already in the graph -- the existing ones and the
 * candidates accepted before it -- (k nearest neighbours, squared distance < 0.6 * exploration_step^2 and angular
 * distance below the threshold, :549-583), it sees at least min_voxel_count voxels and min_information information
 * (:599-613).  The expensive parts run batched on the GPU (one overlap query, one scoring call over all valid
 * candidates); only the distance gate is sequential, on the host, with exact nearest neighbours (the reference asks
 * FLANN, an approximate third-party index).  Poses: position (x, y, z) + image-to-world quaternion (w, x, y, z), what
 * Pose::createFromImageToWorldTransformation takes (:664). */
typedef struct q3dr_candidate_opts {
  float object_bbox[6];                           /* drone_bbox_ relative to the position, min xyz max xyz */
  float bvh_bbox[6];                              /* ViewpointPlannerData::bvh_bbox_ */
  float obstacle_free_height;
  uint32_t discard_dist_knn;                      /* viewpoint_discard_dist_knn, 10 */
  float exploration_step;                         /* viewpoint_exploration_step, 3 */
  float exploration_step_max;                     /* FLT_MAX */
  int32_t exploration_dilation_squared;           /* 1 */
  float exploration_dilation_start_bbox_distance; /* 0 */
  float exploration_dilation_speed;               /* 2 */
  float exploration_bbox[6];                      /* getExplorationBbox() */
  float exploration_angular_dist_threshold_degrees; /* 30 */
  uint32_t min_voxel_count;                       /* viewpoint_min_voxel_count, 100 */
  float min_information;                          /* viewpoint_min_information, 10 */
} q3dr_candidate_opts;

enum q3dr_candidate_status {
  Q3DR_CANDIDATE_ACCEPTED = 0,
  Q3DR_CANDIDATE_INVALID_POSITION = 1,
  Q3DR_CANDIDATE_TOO_CLOSE = 2,
  Q3DR_CANDIDATE_TOO_FEW_VOXELS = 3,
  Q3DR_CANDIDATE_TOO_LITTLE_INFORMATION = 4
};

void q3dr_candidate_opts_default(q3dr_candidate_opts* opts);

/* status_out: n q3dr_candidate_status values.  result: CSR over all n candidates (host memory owned by the map, valid
 * until the next scoring call): the voxel set and total of every candidate with a valid position, whatever its status
 * (an invalid position owns an empty range); accepted candidates are the new ViewpointEntries. */
int q3dr_generate_viewpoint_entries(q3dr_map* map, const float K[4], uint32_t image_width, uint32_t x0, uint32_t x1,
                                    uint32_t y0, uint32_t y1, const float* candidate_positions,
                                    const float* candidate_quaternions_wxyz, size_t n, const float* existing_positions,
                                    const float* existing_quaternions_wxyz, size_t n_existing,
                                    const q3dr_score_opts* opts, const q3dr_candidate_opts* candidate_opts,
                                    uint8_t* status_out, q3dr_result* result);

/* ---- occupancy octree -> BVH leaves, grid weights ---- */

/* ViewpointPlannerData::generateBVHTree's per-leaf rule (viewpoint_planner_data.cpp:820-863) for n octree leaves in
 * begin_tree() order: center (n x 3, it.getCoordinate()), size (it.getSize()), occupancy, observation_count.
 * A leaf is dropped if it is free (occupancy <= occupancy_threshold, 0.51) and known (observation_count >=
 * observation_threshold, 1), or if its box -- center -/+ size/2, constrained to bvh_bbox, cropped at
 * obstacle_free_height -- is empty.  out_boxes (capacity n x 6) / out_source (capacity n, index of the octree leaf)
 * receive the surviving leaves in input order; *n_out their number.  Runs on `device`; host buffers in and out. */
int q3dr_octree_leaves_to_boxes(const float* center, const float* size, const float* occupancy,
                                const uint32_t* observation_count, size_t n, const float bvh_bbox[6],
                                float obstacle_free_height, float occupancy_threshold, uint32_t observation_threshold,
                                int device, float* out_boxes, uint32_t* out_source, size_t* n_out);

/* The host half of row N1: OctoMap's ".ot" stream -> the octree's leaves in begin_tree() order, ready for
 * q3dr_octree_leaves_to_boxes.  Replaces octomap::AbstractOcTree::read (viewpoint_planner/src/octree/
 * occupancy_map.hxx:1123-1127) + the begin_tree() walk of generateBVHTree (viewpoint_planner_data.cpp:820-838) for the
 * purpose of feeding the engine; quad3dr_b200/csrc/octree_io.cpp states the format.  No device work. */
typedef struct q3dr_octree q3dr_octree;

enum q3dr_octree_payload {
  Q3DR_OCTREE_AUTO = 0,      /* by the stream's id; for the reference's id, the node type that fits the stream */
  Q3DR_OCTREE_OCCUPANCY = 1, /* OccupancyNode: float occupancy, uint32 observation_count */
  Q3DR_OCTREE_AUGMENTED = 2, /* AugmentedOccupancyNode: ... + float weight + uint64 observation_count_sum */
  Q3DR_OCTREE_LOGODDS = 3    /* stock octomap::OcTree: float log-odds */
};

typedef struct q3dr_octree_info {
  char id[64];         /* tree type of the header, e.g. "Ait_OccupancyMap" */
  double resolution;   /* metres, edge of a finest-level voxel */
  uint64_t n_nodes;    /* "size" of the header = nodes in the stream */
  uint64_t n_leaves;
  uint32_t tree_depth; /* 16 */
  uint32_t payload;    /* q3dr_octree_payload actually used */
} q3dr_octree_info;

int q3dr_octree_read(const char* path, uint32_t payload, q3dr_octree** out);
int q3dr_octree_get_info(const q3dr_octree* octree, q3dr_octree_info* info);
/* Leaves in begin_tree() order; any array may be NULL.  center: n x 3 (it.getCoordinate()), size (it.getSize()),
 * weight: AugmentedOccupancyNode::weight (0 for the other payloads), depth: octree depth of the leaf. */
int q3dr_octree_leaves(const q3dr_octree* octree, float* center, float* size, float* occupancy,
                       uint32_t* observation_count, float* weight, uint32_t* depth);
int q3dr_octree_destroy(q3dr_octree* octree);

/* ViewpointPlannerData::updateWeights (viewpoint_planner_data.cpp:438-482,573-577) on the device: every grid cell
 * (centre grid_origin + grid_increment * (ix, iy, iz), edge grid_increment) writes
 * grid_weight[ix][iy][iz] * exp(-voxel_information_lambda * observation_count) into the leaves its box overlaps,
 * cells in ix, iy, iz order (the last overlapping cell wins); other leaves keep their weight.
 * observation_count: n_leaves, leaf order.  weights_out (nullable): the resulting n_leaves weights. */
int q3dr_map_update_weights_from_grid(q3dr_map* map, const float grid_origin[3], float grid_increment,
                                      const int32_t grid_dim[3], const float* grid_weight,
                                      float voxel_information_lambda, const uint32_t* observation_count,
                                      float* weights_out);

/* ---- image-space normals: the reference's enable_opengl = true branch ----
 *
 * The reference renders the Poisson mesh's vertex normals off screen with OpenGL and scores every voxel with the
 * normal found at the voxel's kept pixel.  Here the image is DEFINED by a ray cast through the pixel centres of the
 * reference's projection matrix against the triangles (closest camera-space depth inside [near, far], equal depth
 * -> lower triangle index, barycentric blend of the vertex normals, the shader's 0.5 n + 0.5 into 8 bits, no hit =
 * clear colour); quad3dr_b200/csrc/mesh_kernels.cuh states the arithmetic.  It equals an OpenGL render except at
 * pixels whose centre is within rounding distance of a triangle edge or a depth tie.  A caller who keeps the
 * reference's renderer hands its images to q3dr_score_viewpoints_normal_images instead. */
typedef struct q3dr_mesh q3dr_mesh;

typedef struct q3dr_render_opts {
  float near_plane;    /* 0.5  (viewpoint_offscreen_renderer.cpp:21) */
  float far_plane;     /* 1e5  (:22) */
  uint32_t clear_argb; /* 0xFFFFFFFF = clear_color_ (1,1,1,1) (:19); it decodes to the normal (1,1,1)/sqrt(3) */
} q3dr_render_opts;

typedef struct q3dr_mesh_info {
  uint64_t n_triangles;
  uint64_t n_nodes;
  uint32_t tree_depth;
  int32_t device;
  uint64_t device_bytes;
  float last_render_ms; /* device time of the last q3dr_mesh_render_normals kernel */
} q3dr_mesh_info;

void q3dr_render_opts_default(q3dr_render_opts* opts);

/* tri_vertices: nf x 9 floats (three corners per triangle), tri_normals: nf x 9 floats (the unit normals
 * uploadPoissonMesh attaches to the corners) or NULL => every corner gets normalize(cross(v2 - v1, v3 - v2))
 * (viewpoint_offscreen_renderer.cpp:162-166). */
int q3dr_mesh_create(const float* tri_vertices, const float* tri_normals, size_t nf, int device, q3dr_mesh** out);
int q3dr_mesh_destroy(q3dr_mesh* mesh);
int q3dr_mesh_get_info(const q3dr_mesh* mesh, q3dr_mesh_info* info);

/* drawPoissonMeshNormals for n poses (Rt: n x 12, host): out_argb receives n x height x width QImage::Format_ARGB32
 * pixels (0xAARRGGBB, row 0 on top) in HOST memory.  K = {fx, fy, cx, cy}. */
int q3dr_mesh_render_normals(q3dr_mesh* mesh, const float K[4], uint32_t width, uint32_t height, const float* Rt,
                             size_t n, const q3dr_render_opts* ropts, uint32_t* out_argb);

/* The pixel decode of computePoissonMeshNormalVector (viewpoint_offscreen_renderer.cpp:519-526); host only. */
int q3dr_decode_normal(uint32_t argb, float normal_out[3]);

/* q3dr_score_viewpoints with image-space normals from `mesh` (same device as the map): only the kept pixel of every
 * hit voxel is cast against the mesh, inside the scoring pass; the result equals scoring with the images
 * q3dr_mesh_render_normals returns.  image_width x image_height is the camera (= render target) size. */
int q3dr_score_viewpoints_mesh_normals(q3dr_map* map, const q3dr_mesh* mesh, const q3dr_render_opts* ropts,
                                       const float K[4], uint32_t image_width, uint32_t image_height, uint32_t x0,
                                       uint32_t x1, uint32_t y0, uint32_t y1, const float* Rt, size_t n,
                                       const q3dr_score_opts* opts, q3dr_result* result);
/* poses and results in device memory, like q3dr_score_viewpoints_device */
int q3dr_score_viewpoints_mesh_normals_device(q3dr_map* map, const q3dr_mesh* mesh, const q3dr_render_opts* ropts,
                                              const float K[4], uint32_t image_width, uint32_t image_height,
                                              uint32_t x0, uint32_t x1, uint32_t y0, uint32_t y1, const float* d_Rt,
                                              size_t n, const q3dr_score_opts* opts, void* stream,
                                              q3dr_result* result);

/* q3dr_score_viewpoints with the caller's normal images (HOST memory, n x image_height x image_width ARGB32, e.g. the
 * QImage bits of drawPoissonMeshNormals): every voxel is scored with the decoded pixel at its kept pixel. */
int q3dr_score_viewpoints_normal_images(q3dr_map* map, const uint32_t* normal_images, const float K[4],
                                        uint32_t image_width, uint32_t image_height, uint32_t x0, uint32_t x1,
                                        uint32_t y0, uint32_t y1, const float* Rt, size_t n,
                                        const q3dr_score_opts* opts, q3dr_result* result);

/* ---- real-image down-weighting ----
 *
 * ViewpointPlannerData::updateWeightsWithRealViewpoints (viewpoint_planner_data.cpp:498-571) for n real images that
 * share one depth camera (K, width x height = the depth map's size; group images by camera): every pixel is cast
 * (remove_duplicates = false, viewpoint_raycast.cpp:166-219), and every hit pixel whose depth-map value is <= 0 or
 * infinite lowers the weight of the voxel it hit, in scan order, image after image:
 *     information = computeViewpointObservationScore(viewpoint, voxel, pixel)      (with the voxel's CURRENT weight)
 *     weight      = max(0, weight - invalid_pixel_observation_factor * information)
 * opts: ViewpointScore options + ranges (raycast_min_range is ignored like everywhere, Q1; the reference's default
 * real_observed_voxels_raycast_max_range = FLT_MAX means unlimited, and so does any value >= 1e18 here).
 * Normals: NodeObject::normal (mesh == NULL && normal_images == NULL: enable_opengl = false), the mesh at the pixel
 * itself (mesh + ropts), or the caller's normal images (n x height x width ARGB32).  depth_maps: n x height x width
 * floats, HOST.  The map's weights are updated in place (device and host copy); weights_out (nullable, n_leaves)
 * receives them, n_updates (nullable) the number of (pixel, voxel) updates applied. */
int q3dr_map_downweight_real_viewpoints(q3dr_map* map, const float K[4], uint32_t width, uint32_t height,
                                        const float* Rt, const float* depth_maps, size_t n,
                                        const q3dr_score_opts* opts, float invalid_pixel_observation_factor,
                                        const q3dr_mesh* mesh, const q3dr_render_opts* ropts,
                                        const uint32_t* normal_images, float* weights_out, uint64_t* n_updates);

/* ---- one process, several GPUs: the map replicated in every GPU's HBM, the viewpoints of a
 * call sharded in contiguous slices, one host thread per GPU, results left per GPU (no merge copy).  For one process
 * per GPU (torchrun / MPI) use q3dr_map_* on each rank instead and gather with NCCL (quad3dr_b200/dist.py). ---- */
typedef struct q3dr_multi q3dr_multi;

typedef struct q3dr_multi_result {
  uint32_t n_parts;                 /* = number of GPUs */
  uint64_t n_viewpoints;
  const uint64_t* first_viewpoint;  /* n_parts + 1: part g holds viewpoints [first[g], first[g+1]) of the call */
  const q3dr_result* parts;         /* n_parts results with LOCAL viewpoint numbering, host memory owned by the maps */
} q3dr_multi_result;

/* devices: n_devices CUDA device indices, or NULL for devices 0 .. n_devices-1 (n_devices <= 0: all visible).
 * The tree is flattened once and uploaded to every device. */
int q3dr_multi_create(const float* boxes, const float* weights, const float* normals, const uint32_t* depth, size_t n,
                      const int* devices, int n_devices, q3dr_multi** out);
int q3dr_multi_destroy(q3dr_multi* multi);
int q3dr_multi_device_count(const q3dr_multi* multi);
q3dr_map* q3dr_multi_map(q3dr_multi* multi, int i); /* the replica on the i-th device (for the single-GPU entry points) */
int q3dr_multi_update_weights(q3dr_multi* multi, const float* weights);
int q3dr_multi_score_viewpoints(q3dr_multi* multi, const float K[4], uint32_t image_width, uint32_t x0, uint32_t x1,
                                uint32_t y0, uint32_t y1, const float* Rt, size_t n, const q3dr_score_opts* opts,
                                q3dr_multi_result* result);
/* Image-space normals (row N2) for q3dr_multi_score_viewpoints: the mesh is uploaded to every device of `multi`
 * (q3dr_mesh_create per GPU) and every following scoring call takes the normals from it at the kept pixels, with
 * render target image_width (of the call) x image_height.  tri_vertices == NULL switches back to the voxel normals. */
int q3dr_multi_set_normal_mesh(q3dr_multi* multi, const float* tri_vertices, const float* tri_normals, size_t nf,
                               uint32_t image_height, const q3dr_render_opts* ropts);

/* ---- result exchange between GPUs ("per-viewpoint hit sets and scores are gathered") ----
 *
 * Every rank (one GPU each; ranks may be threads of one process or separate processes) owns a RECEIVE BUFFER in its
 * own HBM with one slot per source rank.  While a scoring call runs, every finished chunk of its CSR result is pushed
 * into this rank's slot on every rank by the copy engines (cudaMemcpyAsync into peer memory over NVLink / NVSwitch, on
 * one side stream per target rank, no SM involved) while the SMs cast the next chunk; q3dr_exchange_end_step pushes the per-viewpoint
 * headers and a completion flag behind them and then waits -- on the device, in `stream` -- for the flags of all
 * ranks.  After it returns every rank holds the results of ALL ranks (the all-gather of the reference's single-GPU
 * result, viewpoint_planner_data.cpp:328-335 selects one device; here the candidate set is sharded).
 * Two buffers alternate between steps: a view of step s stays valid until this rank's next q3dr_exchange_begin_step.
 *
 * Same process:       q3dr_exchange_create on every map, then q3dr_exchange_connect(all exchanges).
 * Separate processes: q3dr_exchange_create, q3dr_exchange_export (2 x 64-byte CUDA IPC handles), ship the handles of
 *                     all ranks to all ranks by any means (MPI, torch.distributed ...), q3dr_exchange_import. */
typedef struct q3dr_exchange q3dr_exchange;

#define Q3DR_IPC_HANDLE_BYTES 64

/* byte offsets inside one receive buffer (host-only arithmetic; what tests and foreign readers of the buffer need) */
typedef struct q3dr_exchange_layout {
  uint64_t header_off;   /* per slot: {u64 n_viewpoints, u64 n_entries, u64 step, u64 status} */
  uint64_t offsets_off;  /* u64 [max_viewpoints + 1] */
  uint64_t total_off;    /* f32 [max_viewpoints] */
  uint64_t leaf_off;     /* i32 [max_entries] */
  uint64_t info_off;     /* f32 [max_entries] */
  uint64_t slot_bytes;   /* one slot; slot of source rank r starts at r * slot_bytes */
  uint64_t flags_off;    /* u64 [n_ranks] completion flags behind the slots: (step << 1) | failed */
  uint64_t buffer_bytes;
} q3dr_exchange_layout;

int q3dr_exchange_layout_of(uint32_t n_ranks, uint64_t max_viewpoints, uint64_t max_entries, q3dr_exchange_layout* out);

/* max_viewpoints / max_entries: the largest per-rank viewpoint count / result size of any step (same on all ranks). */
int q3dr_exchange_create(q3dr_map* map, uint32_t n_ranks, uint32_t rank, uint64_t max_viewpoints, uint64_t max_entries,
                         q3dr_exchange** out);
/* Unmaps the peers' buffers (no more pushes from this rank).  Across processes: every rank disconnects, the ranks meet
 * at a barrier of their own, then every rank destroys -- so nobody frees memory a peer still has mapped. */
int q3dr_exchange_disconnect(q3dr_exchange* ex);
int q3dr_exchange_destroy(q3dr_exchange* ex); /* disconnects first if need be */
int q3dr_exchange_export(q3dr_exchange* ex, void* handles_out /* 2 x Q3DR_IPC_HANDLE_BYTES */);
int q3dr_exchange_import(q3dr_exchange* ex, const void* all_handles /* n_ranks x 2 x Q3DR_IPC_HANDLE_BYTES, by rank */);
int q3dr_exchange_connect(q3dr_exchange* const* all, uint32_t n_ranks);
/* Arms the exchange: the next scoring call on the map pushes its chunks.  Call it on every rank, every step. */
int q3dr_exchange_begin_step(q3dr_exchange* ex);
/* local_status: status of this rank's scoring call; a non-zero value is forwarded so that all ranks fail together
 * instead of waiting for data that never comes.  stream: a cudaStream_t the wait is enqueued on (and synchronised).
 * The wait is a small kernel polling the flags in this rank's HBM.  One rank per device is the supported deployment:
 * ranks that SHARE a device (only the tests do that) must not make device-synchronising calls (cudaMalloc, cudaFree)
 * while another rank of that device sits in end_step -- such a call waits for the polling kernel, which waits for
 * the caller's flag. */
int q3dr_exchange_end_step(q3dr_exchange* ex, int local_status, void* stream);
/* Result of source rank `src` in this rank's receive buffer (DEVICE pointers; first_pixel, dist_sq, hit_count NULL). */
int q3dr_exchange_view(const q3dr_exchange* ex, uint32_t src, q3dr_result* out);

/* q3dr_multi with the exchange switched on: after q3dr_multi_score_viewpoints every GPU of `multi` holds the results
 * of all GPUs; q3dr_multi_gathered returns the part scored by GPU `src` as GPU `on` holds it (device pointers). */
int q3dr_multi_enable_gather(q3dr_multi* multi, uint64_t max_viewpoints_per_gpu, uint64_t max_entries_per_gpu);
int q3dr_multi_gathered(q3dr_multi* multi, uint32_t on, uint32_t src, q3dr_result* out);

#ifdef __cplusplus
}
#endif
#endif /* Q3DR_RAYCAST_H_ */
