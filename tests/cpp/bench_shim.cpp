// bench_shim.cpp -- times the reference-facing C++ call surface (include/q3dr/occupied_tree_cuda.hpp) on a workload
// dumped by bench.py: the number a Quad3DR maintainer sees through getRaycastHitVoxelsWithInformationScore*, next to
// the raw C-ABI call it wraps.  Measurement tooling, not product code.
//
//   bench_shim <dir> <steps> <warmup>
// <dir> holds boxes.f32 (n x 6), weight.f32 (n), normal.f32 (n x 3), depth.u32 (n), rt.f32 (nv x 12) and meta.txt
// ("n nv width height fx fy cx cy").  Prints one JSON object.
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "../../include/q3dr/occupied_tree_cuda.hpp"

namespace {

struct Node {  // stands in for bvh::Node: the shim only echoes the pointer
  int dummy;
};
struct Mat44 {
  float m[4][4];
  float operator()(int r, int c) const { return m[r][c]; }
};
struct Mat34 {
  float m[3][4];
  float operator()(int r, int c) const { return m[r][c]; }
};
struct Camera {
  Mat44 K;
  std::size_t w, h;
  const Mat44& intrinsics() const { return K; }
  std::size_t width() const { return w; }
  std::size_t height() const { return h; }
};
struct Pose {
  Mat34 T;
  Mat34 getTransformationImageToWorld() const { return T; }
};
struct Viewpoint {
  const Camera* cam;
  Pose p;
  const Camera& camera() const { return *cam; }
  const Pose& pose() const { return p; }
};

template <typename T>
std::vector<T> slurp(const std::string& path, std::size_t count) {
  std::vector<T> v(count);
  std::ifstream f(path.c_str(), std::ios::binary);
  if (!f || !f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(count * sizeof(T)))) {
    std::fprintf(stderr, "bench_shim: cannot read %s\n", path.c_str());
    std::exit(2);
  }
  return v;
}

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 4) {
    std::fprintf(stderr, "usage: bench_shim <dir> <steps> <warmup>\n");
    return 2;
  }
  const std::string dir = argv[1];
  const int steps = std::atoi(argv[2]), warmup = std::atoi(argv[3]);
  std::size_t n = 0, nv = 0, w = 0, h = 0;
  float fx, fy, cx, cy;
  {
    std::ifstream f((dir + "/meta.txt").c_str());
    if (!(f >> n >> nv >> w >> h >> fx >> fy >> cx >> cy)) {
      std::fprintf(stderr, "bench_shim: bad meta.txt\n");
      return 2;
    }
  }
  const std::vector<float> boxes = slurp<float>(dir + "/boxes.f32", 6 * n), weight = slurp<float>(dir + "/weight.f32", n),
                           normal = slurp<float>(dir + "/normal.f32", 3 * n), rt = slurp<float>(dir + "/rt.f32", 12 * nv);
  const std::vector<std::uint32_t> depth = slurp<std::uint32_t>(dir + "/depth.u32", n);

  std::vector<Node> nodes(n);
  std::vector<q3dr::LeafRecord<Node> > leaves(n);
  for (std::size_t k = 0; k < n; ++k) {
    for (int i = 0; i < 3; ++i) {
      leaves[k].bbox_min[i] = boxes[6 * k + i];
      leaves[k].bbox_max[i] = boxes[6 * k + 3 + i];
      leaves[k].normal[i] = normal[3 * k + i];
    }
    leaves[k].weight = weight[k];
    leaves[k].node = &nodes[k];
    leaves[k].depth = depth[k];
  }
  typedef q3dr::OccupiedTreeCuda<Node> Tree;
  Tree tree;
  try {
    tree.ensureCudaTreeIsInitialized(leaves, 0);
  } catch (const Tree::Error& e) {
    std::fprintf(stderr, "bench_shim: upload failed: %s\n", e.what());
    return 3;
  }
  Camera cam;
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) cam.K.m[r][c] = (r == c) ? 1.0f : 0.0f;
  cam.K.m[0][0] = fx;
  cam.K.m[1][1] = fy;
  cam.K.m[0][2] = cx;
  cam.K.m[1][2] = cy;
  cam.w = w;
  cam.h = h;
  std::vector<Viewpoint> vps(nv);
  std::vector<const Viewpoint*> batch(nv);
  for (std::size_t v = 0; v < nv; ++v) {
    vps[v].cam = &cam;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) vps[v].p.T.m[r][c] = rt[12 * v + 4 * r + c];
    batch[v] = &vps[v];
  }
  q3dr::ViewpointRaycastCuda<Node> raycaster(&tree, 0.0f, 60.0f);
  const float K[4] = {fx, fy, cx, cy};
  q3dr_score_opts opts;
  q3dr_score_opts_default(&opts);

  double ms_abi = 0, ms_view = 0, ms_set = 0, ms_iter = 0;
  std::uint64_t entries = 0;
  double checksum = 0;
  for (int s = 0; s < warmup + steps; ++s) {
    // (1) the raw C-ABI call the shim wraps, (leaf, information) only
    q3dr_map_set_result_fields(tree.handle(), 0u);
    double t0 = now_ms();
    q3dr_result r;
    if (q3dr_score_viewpoints(tree.handle(), K, (std::uint32_t)w, 0, (std::uint32_t)w, 0, (std::uint32_t)h, rt.data(), nv, &opts,
                              &r) != Q3DR_OK) {
      std::fprintf(stderr, "bench_shim: %s\n", q3dr_last_error());
      return 4;
    }
    double t1 = now_ms();
    entries = r.n_entries;
    // (2) the CSR-backed sets
    double t2 = now_ms();
    {
      std::vector<std::pair<q3dr::VoxelWithInformationSetView<Node>, float> > out =
          raycaster.getRaycastHitVoxelsWithInformationScoreBatchView(batch, 0, w, 0, h, true);
      double t3 = now_ms();
      // a consumer's pass over every set (viewpoint_planner_graph.cpp:46-52 iterates each new set once)
      double acc = 0;
      for (std::size_t v = 0; v < out.size(); ++v) {
        float sum = 0;
        for (const q3dr::VoxelWithInformation<Node>& vi : out[v].first) sum += vi.information;
        acc += sum;
      }
      double t4 = now_ms();
      checksum = acc;
      if (s >= warmup) {
        ms_view += t3 - t2;
        ms_iter += t4 - t3;
      }
    }
    // (3) the reference's own container, filled by all host threads
    double t5 = now_ms();
    {
      std::vector<std::pair<q3dr::VoxelWithInformationSet<Node>, float> > out =
          raycaster.getRaycastHitVoxelsWithInformationScoreBatch(batch, 0, w, 0, h, true);
      double t6 = now_ms();
      if (s >= warmup) ms_set += t6 - t5;
      if (out.size() != nv) return 5;
    }  // (destruction of 10^7 hash nodes is the caller's cost, as in the reference; not timed)
    if (s >= warmup) ms_abi += t1 - t0;
  }
  const double rays = (double)nv * (double)w * (double)h;
  int threads = 1;
#ifdef _OPENMP
  threads = omp_get_max_threads();
#endif
  std::printf(
      "{\"viewpoints\": %zu, \"entries\": %llu, \"host_threads\": %d, \"steps\": %d, "
      "\"c_abi_ms\": %.3f, \"batch_view_ms\": %.3f, \"view_iterate_all_ms\": %.3f, \"batch_unordered_set_ms\": %.3f, "
      "\"c_abi_rays_per_sec\": %.6e, \"batch_view_rays_per_sec\": %.6e, \"batch_unordered_set_rays_per_sec\": %.6e, "
      "\"checksum\": %.9e}\n",
      nv, (unsigned long long)entries, threads, steps, ms_abi / steps, ms_view / steps, ms_iter / steps, ms_set / steps,
      rays / (ms_abi / steps * 1e-3), rays / (ms_view / steps * 1e-3), rays / (ms_set / steps * 1e-3), checksum);
  return 0;
}
