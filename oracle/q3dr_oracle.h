/*
 * q3dr_oracle.h -- CPU ORACLE for the Quad3DR viewpoint raycast + scoring path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.  The
 * product (quad3dr_b200/, libq3dr_raycast.so) never links, imports or calls it.
 *
 * It restates, in dependency-free C++17 with strict IEEE-754 binary32 arithmetic
 * (no FMA contraction, no reassociation: built with -ffp-contract=off, no fast-math),
 * the reference's CPU algorithm.  Citations are relative to the reference checkout:
 *   tree build      viewpoint_planner/src/bvh/bvh.h:361-373,764-788,96-106
 *   ray query       viewpoint_planner/src/bvh/bvh.h:376-391,842-909
 *   box / ray math  include/bh/math/geometry.h:65-66,159-160,223-238,319-351,357-361
 *   camera ray      viewpoint_planner/src/planner/viewpoint.cpp:88-96,
 *                   viewpoint_planner/src/reconstruction/sparse_reconstruction.cpp:132-135,148-154,
 *                   include/bh/pose.h:122-127
 *   pixel loop      viewpoint_planner/src/planner/viewpoint_raycast.cpp:166-219
 *   dedup           viewpoint_planner/src/planner/viewpoint_raycast.cpp:321-335
 *   scoring         viewpoint_planner/src/planner/viewpoint_planner_scoring.cpp:11-38,67-73,116-122,
 *                   viewpoint_planner/src/planner/viewpoint_score.cpp:20-28,45-68,
 *                   include/ait/utilities.h:65-68
 *   aggregation     viewpoint_planner/src/planner/viewpoint_planner_raycast.cpp:116-148
 *
 * PARITY STATUS.  The reference holds no golden vector, known-answer test or fixture for this path
 *  and its CPU path cannot be compiled here (Eigen, Boost, octomap, OMPL, Qt5 absent).
 *   PINNED against outputs of the reference itself: the closest-hit query.  The reference's own device code
 *     (viewpoint_planner/src/bvh/bvh.cu, the same recursion as bvh.h:842-909) is compiled from the reference
 *     checkout by oracle/ref_cuda/Makefile and run on the GPU box; with orc_set_sqnorm_sequential(1) -- the one
 *     arithmetic difference between the reference's device and CPU code (include/bh/cuda_math.h:135-141 vs
 *     Eigen) -- this oracle equals it BIT FOR BIT on leaf, depth, dist_sq and intersection point over five
 *     scenes plus tie / inside-a-box / on-a-face / short-range cases (tests/test_gpu_refpin.py).  That pins
 *     traversal order, tie-breaking, the inside rule, the range rule and the slab arithmetic.
 *   UNPINNED (stated conventions, no reference output available to check them): Eigen's fixed-size-3
 *     reduction order, restated as  a0 + (a1 + a2)  (Eigen 3.3 redux_novec_unroller halves the range;
 *     cmake.sh points at eigen-3.3.1) for squaredNorm / dot / 3x3*vec3; normalized() as componentwise
 *     division by sqrt(squaredNorm()) when that is > 0; and the scoring formulas of viewpoint_score.cpp /
 *     viewpoint_planner_scoring.cpp, which exist only on the reference's CPU side.
 *   UNPINNED, image-space normals (row N2): the reference's normal image comes from an OpenGL rasteriser, which cannot
 *     run here; it is restated as a brute-force ray cast (see the N2 block below).  The pixel decode and the scoring
 *     with the decoded normal follow viewpoint_offscreen_renderer.cpp:519-526 / viewpoint_planner_scoring.cpp:26-38.
 * The reference compiles this code under "#pragma GCC optimize(fast-math)"; its exact bits are therefore
 * compiler dependent.  This oracle pins strict IEEE as THE definition.
 */
#ifndef Q3DR_ORACLE_H_
#define Q3DR_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_tree orc_tree;

/* Scoring options, same meaning and defaults as ViewpointScore::Options
 * (viewpoint_score.h:17-41) / ViewpointPlanner::Options (viewpoint_planner.h:320-329). */
typedef struct orc_score_opts {
  float voxel_sensor_size_ratio_threshold;          /* 0.002 */
  float voxel_sensor_size_ratio_inv_falloff_factor; /* 0.001 */
  float incidence_angle_threshold_degrees;          /* 30 */
  float incidence_angle_inv_falloff_factor_degrees; /* 30 */
  int32_t incidence_ignore_dot_product_sign;        /* 0 */
  int32_t ignore_voxels_with_zero_information;      /* 1 at every graph-building call site */
  float raycast_min_range;                          /* accepted, never read (bvh.h:381) */
  float raycast_max_range;                          /* 60; <= 0 means unlimited */
} orc_score_opts;

/* boxes: n x 6 floats (min xyz, max xyz) in the order handed to Tree::build.  n >= 1. */
orc_tree* orc_tree_build(const float* boxes, size_t n);
void orc_tree_free(orc_tree* t);
size_t orc_tree_num_leaves(const orc_tree* t);
size_t orc_tree_num_nodes(const orc_tree* t);
size_t orc_tree_depth(const orc_tree* t);
/* order[k] = input index of the k-th leaf in left-to-right DFS order (the tie-break order). */
void orc_tree_dfs_order(const orc_tree* t, uint32_t* order);
/* depth_out[i] = BVH depth (root = 0) of the leaf built from input box i. */
void orc_tree_leaf_depths(const orc_tree* t, uint32_t* depth_out);
/* all-node iterator order (bvh.h:226-235): idx_out[i] = voxel index of input leaf i. */
void orc_tree_voxel_index(const orc_tree* t, uint32_t* idx_out);
/* root bounding box (6 floats) */
void orc_tree_root_box(const orc_tree* t, float* box_out);

/* Quaternion (w,x,y,z) + translation -> row-major 3x4 [R|t] as Eigen's
 * QuaternionBase::toRotationMatrix (pose.h:122-127). */
void orc_pose_to_rt(const float quat_wxyz[4], const float trans[3], float rt_out[12]);

/* Per-pixel closest-hit query over the window [x0,x1) x [y0,y1), row-major.
 * K = {fx, fy, cx, cy};  Rt = row-major 3x4 [R|t] (image-to-world).
 * Outputs (any may be NULL), all of length (x1-x0)*(y1-y0):
 *   leaf   input index of the hit leaf, -1 for a miss
 *   dsq    squared distance (bits of the reference's dist_sq), 0 for a miss
 *   xyz    intersection point (3 floats per pixel), 0 for a miss
 *   depth  BVH depth of the hit leaf, 0 for a miss
 * Returns the number of tree nodes visited (for the visits/ray statistic). */
uint64_t orc_raycast_window(const orc_tree* t, const float K[4], const float Rt[12],
                            uint32_t x0, uint32_t x1, uint32_t y0, uint32_t y1,
                            float min_range, float max_range,
                            int32_t* leaf, float* dsq, float* xyz, uint32_t* depth);

/* Same query by brute force over all leaves (structure-independence check):
 * argmin dist_sq, ties to the leaf latest in DFS order. */
void orc_raycast_window_brute(const orc_tree* t, const float K[4], const float Rt[12],
                              uint32_t x0, uint32_t x1, uint32_t y0, uint32_t y1,
                              float min_range, float max_range, int32_t* leaf, float* dsq);

/* Arbitrary ray list (bvh.h:512-541): origins/directions n x 3; directions are normalised
 * like bh::Ray's constructor. */
void orc_intersect_rays(const orc_tree* t, const float* origins, const float* directions, size_t n,
                        float min_range, float max_range, int32_t* leaf, float* dsq, float* xyz,
                        uint32_t* depth);

/* The same query for rays used AS GIVEN (no normalisation; inv = 1 / direction), OpenMP over rays.
 * This is the form of the reference's device entry CudaTree::intersectsRecursive (bvh.cu:546-579). */
void orc_intersect_rays_raw(const orc_tree* t, const float* origins, const float* directions, size_t n,
                            float min_range, float max_range, int32_t* leaf, float* dsq, float* xyz,
                            uint32_t* depth);

/* The camera rays (origin, normalised direction) of a pixel window, row-major, as the query sees them. */
void orc_camera_rays(const float K[4], const float Rt[12], uint32_t x0, uint32_t x1, uint32_t y0, uint32_t y1,
                     float* origins, float* directions);

/* Pin-test switch: 1 = squaredNorm associates sequentially like the reference's DEVICE code
 * (include/bh/cuda_math.h:135-141) instead of Eigen's a0 + (a1 + a2).  Default 0 = the contract. */
void orc_set_sqnorm_sequential(int on);

/* The binary splitMedian tree itself (for handing it to the reference's CudaTree::createCopyFromHostTree):
 * boxes n_nodes x 6, left/right child node index or -1, leaf = input leaf index or -1 for inner nodes. */
void orc_tree_export_nodes(const orc_tree* t, float* boxes, int32_t* left, int32_t* right, int32_t* leaf,
                           int32_t* root);

/* observation score of one voxel seen from camera centre `cam` (scoring.cpp:116-122). */
float orc_voxel_information(const float box[6], float weight, const float normal[3],
                            const float cam[3], float fx, uint32_t image_width,
                            const orc_score_opts* opts);

/* getRaycastHitVoxelsWithInformationScore for one viewpoint over the window.
 * weights: n floats; normals: n x 3 floats (input-leaf order).
 * Outputs, capacity `cap` each (cap >= number of unique hit voxels, at most window size):
 *   out_leaf   ascending input leaf index of every voxel kept in the voxel set
 *   out_info   its information
 *   out_pixel  row-major index (within the window) of the first pixel that hit it
 *   out_dsq    dist_sq of that kept pixel
 * hit_count   = unique voxels hit before the information>0 filter
 * total       = (float) of the fp64 sum of kept informations in ascending leaf order
 * total_f32   = fp32 running sum in ascending leaf order
 * Returns the number of kept voxels (<= cap) or (size_t)-1 if cap is too small. */
size_t orc_score_viewpoint(const orc_tree* t, const float* weights, const float* normals,
                           const float K[4], uint32_t image_width, const float Rt[12],
                           uint32_t x0, uint32_t x1, uint32_t y0, uint32_t y1,
                           const orc_score_opts* opts, size_t cap,
                           int32_t* out_leaf, float* out_info, uint32_t* out_pixel, float* out_dsq,
                           uint32_t* hit_count, float* total, float* total_f32);

/* Batch form used for the CPU baseline: n_vp viewpoints (Rt: n_vp x 12), OpenMP over
 * viewpoints with `threads` threads (1 = the reference as shipped: serial pixel loop).
 * Only counts and totals are returned.  Returns total rays cast. */
uint64_t orc_score_batch(const orc_tree* t, const float* weights, const float* normals,
                         const float K[4], uint32_t width, uint32_t height,
                         const float* Rt, size_t n_vp, const orc_score_opts* opts, int threads,
                         uint32_t* kept_count, uint32_t* hit_count, float* total);

/* ---- occupancy octree -> BVH leaves, grid weights ---- */
/* generateBVHTree's rule for one octree leaf (viewpoint_planner_data.cpp:820-863). */
int orc_octree_leaf_to_box(const float center[3], float size, float occupancy, uint32_t observation_count,
                           const float bvh_bbox[6], float obstacle_free_height, float occupancy_threshold,
                           uint32_t observation_threshold, float box_out[6]);
/* updateWeights' writes into the BVH leaves (viewpoint_planner_data.cpp:438-482,573-577), brute force. */
void orc_update_weights_from_grid(const float* boxes, size_t n, const float grid_origin[3], float grid_increment,
                                  const int32_t grid_dim[3], const float* grid_weight, float lambda,
                                  const uint32_t* observation_count, float* weight);

/* ---- box-vs-map overlap queries ---- */
/* Tree::intersects(bbox) (bvh.h:544-554,911-939): input indices of the overlapping leaves in visiting order. */
size_t orc_overlap_box(const orc_tree* t, const float box[6], size_t cap, int32_t* out_leaf);
/* ViewpointPlannerData::isValidObjectPosition (viewpoint_planner_data.cpp:209-235), no-fly zones excluded. */
int orc_valid_object_position(const orc_tree* t, const float position[3], const float object_bbox[6],
                              const float bvh_bbox[6], float obstacle_free_height);

/* ---- consumers of the voxel sets ----
 * leaf / info: one viewpoint's voxel set, ascending leaf index.  observed: dense over the leaves, NaN = voxel not
 * in ViewpointPath::observed_voxel_map.  Return value: the fp32 running sum in entry order (the reference sums in
 * hash order); *sum_f64 (nullable): the same sum accumulated in fp64. */
float orc_new_information(const int32_t* leaf, const float* info, size_t n, const float* observed, const float* weight,
                          float factor, double* sum_f64);
float orc_observed_add(const int32_t* leaf, const float* info, size_t n, float* observed, const float* weight,
                       float factor, double* sum_f64);
float orc_overlap_information(const int32_t* leaf1, const float* info1, size_t n1, const int32_t* leaf2,
                              const float* info2, size_t n2, double* sum_f64);

/* ---- image-space normals: the enable_opengl = true branch ----
 * PARITY UNPINNED against the reference's OpenGL rasteriser (it cannot run here): the normal image is restated as a
 * brute-force ray cast over all triangles through the pixel centres of the reference's projection
 * (viewpoint_offscreen_renderer.cpp:276-293), closest camera-space depth in [near, far], equal depth -> lower
 * triangle index, barycentric blend of the corner normals, 0.5 n + 0.5 rounded into 8 bits
 * (shaders/triangles_normals.f.glsl), no hit -> clear colour.  The pixel decode (:519-526) and the scoring with the
 * decoded normal (viewpoint_planner_scoring.cpp:26-38) follow the reference.
 * tri_vertices / tri_normals: nf x 9 floats (three corners per triangle). */
void orc_decode_normal(uint32_t argb, float normal_out[3]);
/* normalize(cross(v2 - v1, v3 - v2)) on all three corners (viewpoint_offscreen_renderer.cpp:162-166) */
void orc_mesh_face_normals(const float* tri_vertices, size_t nf, float* tri_normals);
void orc_mesh_render_normals(const float* tri_vertices, const float* tri_normals, size_t nf, const float K[4],
                             uint32_t width, uint32_t height, const float Rt[12], float near_plane, float far_plane,
                             uint32_t clear_argb, uint32_t* out_argb /* height x width, 0xAARRGGBB */);
void orc_mesh_cast_pixels(const float* tri_vertices, const float* tri_normals, size_t nf, const float K[4],
                          uint32_t width, uint32_t height, const float Rt[12], float near_plane, float far_plane,
                          uint32_t clear_argb, const uint32_t* xs, const uint32_t* ys, size_t n, uint32_t* out_argb);
/* orc_score_viewpoint with the normal of every voxel taken from normal_image at the voxel's kept pixel */
size_t orc_score_viewpoint_image_normals(const orc_tree* t, const float* weights, const uint32_t* normal_image,
                                         uint32_t image_height, const float K[4], uint32_t image_width,
                                         const float Rt[12], uint32_t x0, uint32_t x1, uint32_t y0, uint32_t y1,
                                         const orc_score_opts* opts, size_t cap, int32_t* out_leaf, float* out_info,
                                         uint32_t* out_pixel, float* out_dsq, uint32_t* hit_count, float* total);

/* orc_score_batch plus, per viewpoint, an order-independent checksum of its voxel set:
 * sum over the kept entries of splitmix64((uint64)leaf << 32 | first_pixel)  (mod 2^64). */
/* updateWeightsWithRealViewpoints for one real image (viewpoint_planner_data.cpp:498-571): weights updated in place,
 * normal_image nullable (enable_opengl = true: ARGB32 normal image of the pose, height x width).  Returns the number of
 * weight updates (hit pixels with an invalid depth). */
uint64_t orc_downweight_real_viewpoint(const orc_tree* t, float* weights, const float* normals,
                                       const uint32_t* normal_image, const float K[4], uint32_t width, uint32_t height,
                                       const float Rt[12], const float* depth_map, const orc_score_opts* opts,
                                       float invalid_pixel_observation_factor);

uint64_t orc_score_batch_checksum(const orc_tree* t, const float* weights, const float* normals,
                                  const float K[4], uint32_t width, uint32_t height,
                                  const float* Rt, size_t n_vp, const orc_score_opts* opts, int threads,
                                  uint32_t* kept_count, uint32_t* hit_count, float* total, uint64_t* set_checksum);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif /* Q3DR_ORACLE_H_ */
