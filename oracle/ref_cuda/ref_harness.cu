// ref_harness.cu -- TEST / BASELINE INFRASTRUCTURE, not the product.
//
// Drives the REFERENCE's own device code -- viewpoint_planner/src/bvh/bvh.cu + bvh.cuh of the reference
// checkout, compiled from where they lie by oracle/ref_cuda/Makefile, not one line copied -- through its
// public C++ interface (bvh::CudaTree<float>) and exposes it over a small C ABI so that
//   * tests/test_gpu_refpin.py can pin the CPU oracle against outputs of the reference itself, and
//   * bench.py can time the reference's recursive kernel, recompiled for sm_100a, beside ours.
// Only tests/ and bench.py load the resulting oracle/_ref/*.so; the product never does.
//
// The host tree is handed to CudaTree::createCopyFromHostTree as an array of layout-compatible 48-byte
// nodes, the way the reference itself does it (bvh.h:949-955 reinterpret_casts its bvh::Node array).
//
// Reference defect worked around here, not reproduced: createCopyFromHostTree uploads nodes in batches of
// num_of_nodes / 20 and never flushes the last partial batch (bvh.cu:40-75), so the last (n mod (n/20)) nodes
// in breadth-first order -- always 1..19 of the deepest leaves, n = 2N-1 being odd -- stay uninitialised on
// the device.  With patch_tail != 0 this harness uploads those nodes itself (same breadth-first numbering),
// so that the device tree is the tree the reference meant to upload.
#include <cstdint>
#include <cstring>
#include <deque>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "bvh.cu"  // the reference translation unit itself (-I <reference>/viewpoint_planner/src/bvh)

namespace {

struct HostNode {  // layout of bvh::CudaNode<float>: box min, box max, left, right, ptr
  float mn[3];
  float mx[3];
  HostNode* left;
  HostNode* right;
  void* ptr;
};
static_assert(sizeof(HostNode) == sizeof(bvh::CudaNode<float>), "CudaNode layout changed");
static_assert(sizeof(HostNode) == 48, "CudaNode is expected to be 48 bytes");

thread_local std::string g_err;

}  // namespace

struct refcuda_tree {
  std::vector<HostNode> host;
  bvh::CudaTree<float>* tree = nullptr;
  size_t depth = 0;
  // scratch for the kernel-only timing path
  bvh::CudaTree<float>::CudaRayType* d_rays = nullptr;
  bvh::CudaTree<float>::CudaIntersectionResultWithScreenCoordinates* d_res = nullptr;
  size_t scratch_rays = 0;
};

extern "C" {

const char* refcuda_last_error(void) { return g_err.c_str(); }

// nodes: boxes nn x 6, left/right node index or -1.  `depth` = tree depth (only sizes the iterative stacks).
int refcuda_create(const float* boxes, const int32_t* left, const int32_t* right, int32_t root, size_t nn,
                   size_t depth, int patch_tail, size_t stack_bytes, refcuda_tree** out) {
  *out = nullptr;
  try {
    refcuda_tree* t = new refcuda_tree;
    t->depth = depth;
    t->host.resize(nn);
    for (size_t i = 0; i < nn; ++i) {
      HostNode& h = t->host[i];
      for (int a = 0; a < 3; ++a) {
        h.mn[a] = boxes[6 * i + a];
        h.mx[a] = boxes[6 * i + 3 + a];
      }
      h.left = left[i] >= 0 ? &t->host[left[i]] : nullptr;
      h.right = right[i] >= 0 ? &t->host[right[i]] : nullptr;
      h.ptr = &h;
    }
    // device recursion needs a deep stack: the reference sets 32 KiB per thread process-wide
    // (viewpoint_planner_data.h:121, viewpoint_planner_data.cpp:328-335)
    if (stack_bytes) {
      cudaError_t e = cudaDeviceSetLimit(cudaLimitStackSize, stack_bytes);
      if (e != cudaSuccess) {
        g_err = std::string("cudaDeviceSetLimit: ") + cudaGetErrorString(e);
        delete t;
        return 2;
      }
    }
    std::streambuf* old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());  // createCopyFromHostTree reports progress on std::cout
    try {
      t->tree = bvh::CudaTree<float>::createCopyFromHostTree(reinterpret_cast<bvh::CudaNode<float>*>(&t->host[root]),
                                                             nn, depth);
    } catch (...) {
      std::cout.rdbuf(old);
      throw;
    }
    std::cout.rdbuf(old);
    if (patch_tail) {
      // breadth-first numbering of createCopyFromHostTree (push_front / pop_back queue, bvh.cu:33-62)
      std::vector<HostNode*> order;
      order.reserve(nn);
      std::deque<HostNode*> q;
      q.push_front(&t->host[root]);
      std::vector<int64_t> li(nn, -1), ri(nn, -1);
      size_t counter = 0;
      while (!q.empty()) {
        HostNode* n = q.back();
        q.pop_back();
        order.push_back(n);
        const size_t me = (size_t)(n - t->host.data());
        if (n->left) {
          q.push_front(n->left);
          li[me] = (int64_t)(counter + q.size());
        }
        if (n->right) {
          q.push_front(n->right);
          ri[me] = (int64_t)(counter + q.size());
        }
        ++counter;
      }
      const size_t batch = nn / 20;
      const size_t copied = batch ? (nn / batch) * batch : 0;
      HostNode* d_base = reinterpret_cast<HostNode*>(t->tree->getRoot());
      std::vector<HostNode> tail;
      for (size_t k = copied; k < nn; ++k) {
        HostNode h = *order[k];
        const size_t me = (size_t)(order[k] - t->host.data());
        h.left = li[me] >= 0 ? d_base + li[me] : nullptr;
        h.right = ri[me] >= 0 ? d_base + ri[me] : nullptr;
        h.ptr = order[k];
        tail.push_back(h);
      }
      if (!tail.empty()) {
        cudaError_t e = cudaMemcpy(d_base + copied, tail.data(), tail.size() * sizeof(HostNode), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
          g_err = std::string("tail upload: ") + cudaGetErrorString(e);
          delete t->tree;
          delete t;
          return 2;
        }
      }
    }
    *out = t;
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 2;
  }
}

void refcuda_destroy(refcuda_tree* t) {
  if (!t) return;
  if (t->d_rays) cudaFree(t->d_rays);
  if (t->d_res) cudaFree(t->d_res);
  delete t->tree;
  delete t;
}

// CudaTree::intersectsRecursive (bvh.cu:546-579): rays used as given.  node_index = index of the hit node in
// the `boxes` array of refcuda_create, -1 for a miss.
int refcuda_intersect(refcuda_tree* t, const float* origins, const float* directions, size_t n, float min_range,
                      float max_range, int32_t* node_index, float* dsq, float* xyz, uint32_t* depth) {
  try {
    std::vector<bvh::CudaTree<float>::CudaRayType> rays(n);
    for (size_t i = 0; i < n; ++i) {
      for (int a = 0; a < 3; ++a) {
        rays[i].origin(a) = origins[3 * i + a];
        rays[i].direction(a) = directions[3 * i + a];
      }
    }
    const auto res = t->tree->intersectsRecursive(rays, min_range, max_range);
    for (size_t i = 0; i < n; ++i) {
      const HostNode* hn = static_cast<const HostNode*>(res[i].node);
      const bool hit = hn != nullptr;
      if (node_index) node_index[i] = hit ? (int32_t)(hn - t->host.data()) : -1;
      if (dsq) dsq[i] = hit ? res[i].dist_sq : 0.0f;
      if (depth) depth[i] = hit ? (uint32_t)res[i].depth : 0u;
      if (xyz) {
        for (int a = 0; a < 3; ++a) xyz[3 * i + a] = hit ? res[i].intersection(a) : 0.0f;
      }
    }
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 2;
  }
}

namespace {
void fill_matrices(const float K[4], const float Rt[12], bvh::CudaMatrix4x4<float>* intr, bvh::CudaMatrix3x4<float>* extr) {
  for (size_t r = 0; r < 4; ++r) {
    for (size_t c = 0; c < 4; ++c) (*intr)(r, c) = (r == c) ? 1.0f : 0.0f;
  }
  (*intr)(0, 0) = K[0];
  (*intr)(1, 1) = K[1];
  (*intr)(0, 2) = K[2];
  (*intr)(1, 2) = K[3];
  for (size_t r = 0; r < 3; ++r) {
    for (size_t c = 0; c < 4; ++c) (*extr)(r, c) = Rt[4 * r + c];
  }
}
}  // namespace

// CudaTree::raycastWithScreenCoordinatesRecursive (bvh.cu:647-720): the call bvh::Tree::raycastWithScreenCoordinatesCuda
// makes (bvh.h:453-509); the ray direction is NOT normalised on this path (bvh.cu:335-346).
int refcuda_raycast(refcuda_tree* t, const float K[4], const float Rt[12], uint32_t x0, uint32_t x1, uint32_t y0,
                    uint32_t y1, float min_range, float max_range, int32_t* node_index, float* dsq, float* xyz,
                    uint32_t* depth, float* screen_xy) {
  try {
    bvh::CudaMatrix4x4<float> intr;
    bvh::CudaMatrix3x4<float> extr;
    fill_matrices(K, Rt, &intr, &extr);
    const auto res = t->tree->raycastWithScreenCoordinatesRecursive(intr, extr, x0, x1, y0, y1, min_range, max_range, true);
    for (size_t i = 0; i < res.size(); ++i) {
      const auto& r = res[i].intersection_result;
      const HostNode* hn = static_cast<const HostNode*>(r.node);
      const bool hit = hn != nullptr;
      if (node_index) node_index[i] = hit ? (int32_t)(hn - t->host.data()) : -1;
      if (dsq) dsq[i] = hit ? r.dist_sq : 0.0f;
      if (depth) depth[i] = hit ? (uint32_t)r.depth : 0u;
      if (xyz) {
        for (int a = 0; a < 3; ++a) xyz[3 * i + a] = hit ? r.intersection(a) : 0.0f;
      }
      if (screen_xy) {
        screen_xy[2 * i] = res[i].screen_coordinates(0);
        screen_xy[2 * i + 1] = res[i].screen_coordinates(1);
      }
    }
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 2;
  }
}

// Timing of n_vp viewpoints, one after the other like the planner issues them.
//   api_ms     wall time of the public calls (kernel + cudaDeviceSynchronize + D2H of W*H 56-byte records +
//              std::vector allocation: what bvh::Tree::raycastWithScreenCoordinatesCuda costs its caller)
//   kernel_ms  device time of raycastWithScreenCoordinatesRecursiveCudaKernel<float> alone (CUDA events), launched
//              with the grid / block computation of the public call (bvh.cu:661-662), results left on the device
//   n_hits     pixels with a hit (sanity)
int refcuda_time_raycasts(refcuda_tree* t, const float K[4], const float* Rt, size_t n_vp, uint32_t w, uint32_t h,
                          float max_range, double* api_ms, double* kernel_ms, uint64_t* n_hits) {
  try {
    bvh::CudaMatrix4x4<float> intr;
    bvh::CudaMatrix3x4<float> extr;
    uint64_t hits = 0;
    cudaDeviceSynchronize();
    timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    for (size_t v = 0; v < n_vp; ++v) {
      fill_matrices(K, Rt + 12 * v, &intr, &extr);
      const auto res = t->tree->raycastWithScreenCoordinatesRecursive(intr, extr, 0, w, 0, h, 0.0f, max_range, true);
      for (const auto& r : res) hits += r.intersection_result.node != nullptr;
    }
    clock_gettime(CLOCK_MONOTONIC, &b);
    if (api_ms) *api_ms = (b.tv_sec - a.tv_sec) * 1e3 + (b.tv_nsec - a.tv_nsec) * 1e-6;
    if (n_hits) *n_hits = hits;
    if (kernel_ms) {
      const size_t n = (size_t)w * h;
      if (n > t->scratch_rays) {
        if (t->d_rays) cudaFree(t->d_rays);
        if (t->d_res) cudaFree(t->d_res);
        CUDA_SAFE_CALL(cudaMalloc(&t->d_rays, n * sizeof(*t->d_rays)));
        CUDA_SAFE_CALL(cudaMalloc(&t->d_res, n * sizeof(*t->d_res)));
        t->scratch_rays = n;
      }
      const size_t threads = 1024;  // CudaTree::kThreadsPerBlock
      const size_t grid = bh::CudaUtils::computeGridSize(n, threads);
      const size_t block = bh::CudaUtils::computeBlockSize(n, threads);
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0);
      for (size_t v = 0; v < n_vp; ++v) {
        fill_matrices(K, Rt + 12 * v, &intr, &extr);
        bvh::raycastWithScreenCoordinatesRecursiveCudaKernel<float><<<grid, block>>>(
            intr, extr, t->d_rays, n, 0, w, 0, h, 0.0f, max_range, t->tree->getRoot(), t->d_res);
      }
      cudaEventRecord(e1);
      CUDA_SAFE_CALL(cudaEventSynchronize(e1));
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      *kernel_ms = ms;
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
      CUDA_CHECK_ERROR();
    }
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 2;
  }
}

}  // extern "C"
