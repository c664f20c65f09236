// raycast_kernels.cuh -- sm_100a kernels of the viewpoint raycast + scoring engine.
//
// Numeric contract: every operation that feeds a hit decision is a single
// correctly rounded binary32 operation in the reference's order (explicit __f*_rn intrinsics, the
// translation unit is also compiled with -fmad=false, no fast-math, no FTZ).
//
// raycast_rays_kernel   one warp = one 32-ray packet (ray lists, pixel lists); persistent grid, packet traversal
//                        of the 4-wide node array.  Camera windows use raycast2_kernel (raycast2_kernels.cuh),
//                        scoring lives in score_kernels.cuh.  This header also holds what both share: the
//                        reference's box test and camera ray in exact binary32, the exact fallback traversal.
#pragma once

#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/q3dr_raycast.h"
#include "bvh_build.h"

namespace q3dr {

constexpr int kWarpsPerCta = 8;
constexpr int kThreadsPerCta = kWarpsPerCta * 32;
constexpr int kStackEntries = 64;
constexpr uint32_t kEmptyPixel = 0xFFFFFFFFu;
constexpr uint32_t kPixTagShift = 22;                       // tagged first-pixel entries: windows of up to 2^22 pixels
constexpr uint32_t kPixIndexMask = (1u << kPixTagShift) - 1u;
constexpr uint32_t kPixEpochs = (1u << (32 - kPixTagShift)) - 1u;  // tags 1022 .. 0 are usable, 1023 = never written
constexpr int kScoreThreads = 512;
#ifndef Q3DR_SMEM_TOP
#define Q3DR_SMEM_TOP 1
#endif
constexpr bool kSmemTopLevels = (Q3DR_SMEM_TOP != 0);

enum RaycastMode : int {   // raycast2_kernel<MODE>
  kModeScore = 0,   // dedup into the per-viewpoint first-pixel table + bitmap
  kModePixels = 1   // per-pixel q3dr_pixel_hit records (raycastCuda compat)
};

struct RaycastParams {
  const float4* __restrict__ nodes;  // Node4 = 8 x float4
  uint32_t n_top;                    // nodes staged in shared memory
  const float* __restrict__ rt;      // [n_vp][12] row-major [R|t]
  float fx, fy, cx, cy;
  uint32_t x0, y0, w, h;  // window
  uint32_t tiles_x, tiles_per_vp;
  double inv_tiles_x, inv_tiles_per_vp;  // 1.0 / the two above (raycast2_kernel's tile decoding)
  uint32_t total_tiles;
  uint32_t tile_run;     // raycast2_kernel: consecutive tiles a warp takes per visit to the tile counter (>= 1)
  float max_range;
  // kModeScore
  uint32_t* first_pix;  // [slot][leaf_stride]
  uint32_t* bitmap;     // [slot][word_stride]
  uint64_t leaf_stride;
  uint64_t word_stride;
  uint32_t* seg_hits;   // [slot][n_segs]: set bits per 32768-leaf segment of the bitmap (score_kernels.cuh)
  uint32_t n_segs;
  // first_pix entries are (pix_tag | pixel); the tag decreases from chunk to chunk, so whatever an earlier chunk left
  // behind compares greater than pix_limit = pix_tag | (largest pixel index) and reads as "not hit yet".  Untagged mode
  // (huge windows): pix_tag = 0, pix_limit = 0xFFFFFFFE and the table is kept at 0xFFFFFFFF between chunks.
  uint32_t pix_tag;
  uint32_t pix_limit;
  // Viewpoints whose camera centre lies exactly on a face plane of some box of the map (flag_viewpoints_kernel) take
  // the kernel instantiation that re-checks such nodes with the reference's literal box test; all others run without
  // that check.  vp_flags[v] != 0: v is such a viewpoint.  The checking instantiation walks vp_list[0 .. *n_flagged).
  const uint8_t* __restrict__ vp_flags;
  const uint32_t* __restrict__ vp_list;
  const uint32_t* __restrict__ n_flagged;
  // kModePixels / raycast_rays_kernel
  q3dr_pixel_hit* pixels;
  const uint8_t* __restrict__ leaf_depth;
  const float* __restrict__ ray_o;  // raycast_rays_kernel: [n][3]
  const float* __restrict__ ray_d;
  const uint32_t* __restrict__ pix_list;  // or: window pixel indices of one viewpoint
  uint32_t n_rays;
  uint32_t* tile_counter;
  uint32_t bounded_map;  // every box coordinate is below 1e18 in magnitude
};

// std::min / std::max semantics of the reference (geometry.h:336-343), NaN behaviour included.
__device__ __forceinline__ float smin(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float smax(float a, float b) { return (a < b) ? b : a; }

struct Ray {
  float ox, oy, oz;
  float dx, dy, dz;
  float ix, iy, iz;
};

// Box test of intersectsRecursive (bvh.h:842-895) for one box:
//   outside -> slab test (geometry.h:319-351), dist_sq = |o - (o + d t)|^2 ; inside -> dist_sq = 0.
// Returns whether the box passes the slab test (or contains the origin); dsq and t are valid then.
__device__ __forceinline__ bool box_test(const Ray& r, float mnx, float mny, float mnz, float mxx, float mxy,
                                         float mxz, float& dsq, float& t) {
  const float ax = __fsub_rn(mnx, r.ox), bx = __fsub_rn(mxx, r.ox);
  const float ay = __fsub_rn(mny, r.oy), by = __fsub_rn(mxy, r.oy);
  const float az = __fsub_rn(mnz, r.oz), bz = __fsub_rn(mxz, r.oz);
  // isOutside: o < mn  <=>  mn - o > 0 exactly (gradual underflow keeps the sign of a difference)
  const bool outside = (ax > 0.0f) | (ay > 0.0f) | (az > 0.0f) | (bx < 0.0f) | (by < 0.0f) | (bz < 0.0f);
  const float t0x = __fmul_rn(ax, r.ix), t1x = __fmul_rn(bx, r.ix);
  const float t0y = __fmul_rn(ay, r.iy), t1y = __fmul_rn(by, r.iy);
  const float t0z = __fmul_rn(az, r.iz), t1z = __fmul_rn(bz, r.iz);
  float tmin = -FLT_MAX, tmax = FLT_MAX;
  tmin = smax(tmin, smin(t0x, t1x));
  tmax = smin(tmax, smax(t0x, t1x));
  tmin = smax(tmin, smin(t0y, t1y));
  tmax = smin(tmax, smax(t0y, t1y));
  tmin = smax(tmin, smin(t0z, t1z));
  tmax = smin(tmax, smax(t0z, t1z));
  const float tt = smax(tmin, 0.0f);
  const bool slab = tmax > tt;
  const float px = __fadd_rn(r.ox, __fmul_rn(r.dx, tt));
  const float py = __fadd_rn(r.oy, __fmul_rn(r.dy, tt));
  const float pz = __fadd_rn(r.oz, __fmul_rn(r.dz, tt));
  const float ex = __fsub_rn(r.ox, px), ey = __fsub_rn(r.oy, py), ez = __fsub_rn(r.oz, pz);
  const float d = __fadd_rn(__fmul_rn(ex, ex), __fadd_rn(__fmul_rn(ey, ey), __fmul_rn(ez, ez)));
  dsq = outside ? d : 0.0f;
  t = outside ? tt : 0.0f;
  return outside ? slab : true;
}

// Camera ray of integer pixel (x, y): viewpoint.cpp:88-96, sparse_reconstruction.cpp:148-154,
// bh::Ray normalisation geometry.h:65-66, inverse direction bvh.h:380.
__device__ __forceinline__ void camera_ray(const float* __restrict__ rt, float fx, float fy, float cx, float cy,
                                           uint32_t x, uint32_t y, Ray& r) {
  const float c0 = __fdiv_rn(__fsub_rn((float)x, cx), fx);
  const float c1 = __fdiv_rn(__fsub_rn((float)y, cy), fy);
  const float r00 = __ldg(rt + 0), r01 = __ldg(rt + 1), r02 = __ldg(rt + 2), tx = __ldg(rt + 3);
  const float r10 = __ldg(rt + 4), r11 = __ldg(rt + 5), r12 = __ldg(rt + 6), ty = __ldg(rt + 7);
  const float r20 = __ldg(rt + 8), r21 = __ldg(rt + 9), r22 = __ldg(rt + 10), tz = __ldg(rt + 11);
  const float d0 = __fadd_rn(__fmul_rn(r00, c0), __fadd_rn(__fmul_rn(r01, c1), r02));
  const float d1 = __fadd_rn(__fmul_rn(r10, c0), __fadd_rn(__fmul_rn(r11, c1), r12));
  const float d2 = __fadd_rn(__fmul_rn(r20, c0), __fadd_rn(__fmul_rn(r21, c1), r22));
  const float n2 = __fadd_rn(__fmul_rn(d0, d0), __fadd_rn(__fmul_rn(d1, d1), __fmul_rn(d2, d2)));
  r.ox = tx;
  r.oy = ty;
  r.oz = tz;
  if (n2 > 0.0f) {
    const float n = __fsqrt_rn(n2);
    r.dx = __fdiv_rn(d0, n);
    r.dy = __fdiv_rn(d1, n);
    r.dz = __fdiv_rn(d2, n);
  } else {
    r.dx = d0;
    r.dy = d1;
    r.dz = d2;
  }
  r.ix = __fdiv_rn(1.0f, r.dx);
  r.iy = __fdiv_rn(1.0f, r.dy);
  r.iz = __fdiv_rn(1.0f, r.dz);
}

__device__ __forceinline__ void make_ray(float ox, float oy, float oz, float d0, float d1, float d2, Ray& r) {
  const float n2 = __fadd_rn(__fmul_rn(d0, d0), __fadd_rn(__fmul_rn(d1, d1), __fmul_rn(d2, d2)));
  r.ox = ox;
  r.oy = oy;
  r.oz = oz;
  if (n2 > 0.0f) {
    const float n = __fsqrt_rn(n2);
    r.dx = __fdiv_rn(d0, n);
    r.dy = __fdiv_rn(d1, n);
    r.dz = __fdiv_rn(d2, n);
  } else {
    r.dx = d0;
    r.dy = d1;
    r.dz = d2;
  }
  r.ix = __fdiv_rn(1.0f, r.dx);
  r.iy = __fdiv_rn(1.0f, r.dy);
  r.iz = __fdiv_rn(1.0f, r.dz);
}

struct PacketHit {
  float dsq;     // best dist_sq (the range bound while nothing is hit)
  float t;       // ray parameter of the best hit
  float tU;      // fast path only: a ray parameter with dist_sq_at(tU) > dsq, verified by evaluation
  int32_t leaf;  // best leaf index, -1 = none
};

// dist_sq of the point at ray parameter t, in the reference's arithmetic (geometry.h:345-348, bvh.h:859):
// |o - (o + d t)|^2.  Every operation is monotone in t >= 0, so the computed value is non-decreasing in t.
__device__ __forceinline__ float dist_sq_at(const Ray& r, float t) {
  const float px = __fadd_rn(r.ox, __fmul_rn(r.dx, t));
  const float py = __fadd_rn(r.oy, __fmul_rn(r.dy, t));
  const float pz = __fadd_rn(r.oz, __fmul_rn(r.dz, t));
  const float ex = __fsub_rn(r.ox, px), ey = __fsub_rn(r.oy, py), ez = __fsub_rn(r.oz, pz);
  return __fadd_rn(__fmul_rn(ex, ex), __fadd_rn(__fmul_rn(ey, ey), __fmul_rn(ez, ez)));
}

// A ray parameter u > t whose dist_sq is strictly greater than `dsq` -- checked by evaluating the
// reference formula at u, not by error analysis.  Because dist_sq_at is non-decreasing, every box
// entered at t' >= u has dist_sq > dsq and may be skipped without computing it; boxes in (t, u) are
// simply visited.  +inf when no small margin separates (pruning by t is then off for this lane).
__device__ __forceinline__ float prune_bound(const Ray& r, float t, float dsq) {
  float rel = 1.0f / 2048.0f, ab = 1.0f / 8192.0f;
#pragma unroll 1
  for (int k = 0; k < 3; ++k) {
    const float u = fmaf(t, rel, t) + ab;
    if (dist_sq_at(r, u) > dsq) return u;
    rel *= 16.0f;
    ab *= 16.0f;
  }
  return __int_as_float(0x7F800000);
}

#define Q3DR_FULL_MASK 0xFFFFFFFFu

// Conditional swap of two (key, node, per-lane dsq) candidates; key and node are warp-uniform.
__device__ __forceinline__ void cswap(uint32_t& ka, uint32_t& na, float& da, uint32_t& kb, uint32_t& nb, float& db) {
  if (kb < ka) {
    uint32_t tk = ka; ka = kb; kb = tk;
    uint32_t tn = na; na = nb; nb = tn;
    float td = da; da = db; db = td;
  }
}

// Packet traversal: the warp walks the 4-wide tree as one unit with a shared control flow; each lane
// keeps its own ray, its own best hit and its own pruning distance.  A node is visited while at least
// one lane can still improve below it (warp vote), children are visited nearest first (warp min of the
// entry dist_sq), far children go on a short stack together with every lane's entry dist_sq so that a
// popped entry no lane needs any more is dropped without touching memory.
//
// Correctness does not depend on the tree: a child is skipped by a lane only if its box fails the
// reference's own test (miss, or dist_sq > best), and rounding is monotone, so the box of an ancestor
// never fails when a descendant leaf would pass.  Equal dist_sq: higher leaf
// index wins (= later in the reference's depth-first order).
static __device__ __noinline__ void traverse_packet_exact(const Ray& ray, PacketHit& best, const float4* __restrict__ nodes,
                                                   const float4* smem_top, uint32_t n_top) {
  float stack_d[kStackEntries];
  uint32_t stack_n[kStackEntries];
  int sp = 0;
  uint32_t node = 0;
  while (true) {
    node &= kNodeIndexMask;  // slot 0 carries the node's leaf mask in bits 27..30
    const float4* np = (node < n_top) ? (smem_top + (size_t)node * 8) : (nodes + (size_t)node * 8);
    const float4 mnx = np[0], mny = np[1], mnz = np[2];
    const float4 mxx = np[3], mxy = np[4], mxz = np[5];
    const uint4 ch = *reinterpret_cast<const uint4*>(np + 6);

    float d0, d1, d2, d3, t0, t1, t2, t3;
    const bool h0 = box_test(ray, mnx.x, mny.x, mnz.x, mxx.x, mxy.x, mxz.x, d0, t0);
    const bool h1 = box_test(ray, mnx.y, mny.y, mnz.y, mxx.y, mxy.y, mxz.y, d1, t1);
    const bool h2 = box_test(ray, mnx.z, mny.z, mnz.z, mxx.z, mxy.z, mxz.z, d2, t2);
    const bool h3 = box_test(ray, mnx.w, mny.w, mnz.w, mxx.w, mxy.w, mxz.w, d3, t3);

    // leaves first: they tighten `best` before the inner children are judged
#define Q3DR_LEAF(CH, H, D, T)                                                         \
    if ((CH) & kLeafFlag) {                                                             \
      if (((CH) & kPayloadMask) != kPayloadMask) {                                      \
        const int32_t li = (int32_t)((CH) & kPayloadMask);                              \
        const bool acc = (H) && ((D) < best.dsq || ((D) == best.dsq && li > best.leaf)); \
        if (acc) {                                                                      \
          best.dsq = (D);                                                               \
          best.t = (T);                                                                 \
          best.leaf = li;                                                               \
        }                                                                               \
      }                                                                                 \
    }
    Q3DR_LEAF(ch.x, h0, d0, t0)
    Q3DR_LEAF(ch.y, h1, d1, t1)
    Q3DR_LEAF(ch.z, h2, d2, t2)
    Q3DR_LEAF(ch.w, h3, d3, t3)
#undef Q3DR_LEAF

    // inner children: warp-uniform key = least entry dist_sq over the lanes that still need the child
    uint32_t k0 = 0xFFFFFFFFu, k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu, k3 = 0xFFFFFFFFu;
    float e0 = __int_as_float(0x7F800000), e1 = e0, e2 = e0, e3 = e0;  // +inf = "this lane does not need it"
#define Q3DR_INNER(CH, H, D, K, E)                                                      \
    if (!((CH) & kLeafFlag)) {                                                          \
      const bool ok = (H) && !((D) > best.dsq);                                         \
      if (ok) (E) = (D);                                                                \
      (K) = __reduce_min_sync(Q3DR_FULL_MASK, ok ? __float_as_uint(D) : 0xFFFFFFFFu);   \
    }
    Q3DR_INNER(ch.x, h0, d0, k0, e0)
    Q3DR_INNER(ch.y, h1, d1, k1, e1)
    Q3DR_INNER(ch.z, h2, d2, k2, e2)
    Q3DR_INNER(ch.w, h3, d3, k3, e3)
#undef Q3DR_INNER

    uint32_t n0 = ch.x, n1 = ch.y, n2 = ch.z, n3 = ch.w;
    // 5-comparator sorting network on warp-uniform keys
    cswap(k0, n0, e0, k1, n1, e1);
    cswap(k2, n2, e2, k3, n3, e3);
    cswap(k0, n0, e0, k2, n2, e2);
    cswap(k1, n1, e1, k3, n3, e3);
    cswap(k1, n1, e1, k2, n2, e2);
    // push far ones first so that the nearest is on top
    if (k3 != 0xFFFFFFFFu) { stack_d[sp] = e3; stack_n[sp] = n3; ++sp; }
    if (k2 != 0xFFFFFFFFu) { stack_d[sp] = e2; stack_n[sp] = n2; ++sp; }
    if (k1 != 0xFFFFFFFFu) { stack_d[sp] = e1; stack_n[sp] = n1; ++sp; }
    if (k0 != 0xFFFFFFFFu) {
      node = n0;  // nearest child: descend directly
      continue;
    }
    // pop until an entry some lane still needs
    bool found = false;
    while (sp > 0) {
      --sp;
      const float d = stack_d[sp];
      const uint32_t n = stack_n[sp];
      if (__any_sync(Q3DR_FULL_MASK, !(d > best.dsq))) {
        node = n;
        found = true;
        break;
      }
    }
    if (!found) break;
  }
}


// Fast variant of the same traversal for packets whose rays (a) have finite inverse directions (no NaN can
// arise, so fminf/fmaxf equal the reference's ternaries) and (b) share one direction octant (the near / far
// plane of every slab is then known per packet and loaded directly: lo = (near - o) * inv is bit-identical
// to min(t0, t1)).  Results are bit-identical to traverse_packet_exact:
//   * origin inside a box: the slab test then yields t = 0 and dist_sq = 0 like the reference's inside
//     branch, except when tmax == 0 (origin on an exit face) -- those nodes are redone with box_test;
//   * inner children are judged on the entry parameter t against best.tU (see prune_bound): skipping needs
//     dist_sq > best, which t >= tU implies; visiting more than the reference never changes the result;
//   * leaves compute dist_sq exactly whenever some lane could still accept them.
// Control flow is derived from warp collectives only (redux.or of the per-lane candidate bits, the node's leaf
// mask), so every branch of the loop is a uniform branch: one node visit costs the slab arithmetic plus a short
// scalar epilogue.  Visit order: nearest candidate first (warp minimum of the entry parameter over the lanes that
// still need the child; computed only when more than one inner child survives).  Far candidates go on the stack
// with every lane's own entry parameter; a popped entry that no lane can still improve under (warp vote against
// tU) is dropped without touching memory.
__device__ __forceinline__ uint32_t pick4(uint32_t k, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  return (k == 0u) ? a : ((k == 1u) ? b : ((k == 2u) ? c : d));
}
__device__ __forceinline__ float pick4f(uint32_t k, float a, float b, float c, float d) {
  return (k == 0u) ? a : ((k == 1u) ? b : ((k == 2u) ? c : d));
}

template <bool kSmemTop>
__device__ __forceinline__ void traverse_packet_fast(const Ray& ray, PacketHit& best, const float4* __restrict__ nodes,
                                                     const float4* smem_top, uint32_t n_top,
                                                     uint32_t octant /* bit a set: direction negative on axis a */) {
  float stack_t[kStackEntries];
  uint32_t stack_n[kStackEntries];
  int sp = 0;
  uint32_t ref = 0;
  // plane indices inside a node: mn x,y,z = 0,1,2 ; mx x,y,z = 3,4,5
  const uint32_t nx = (octant & 1u) ? 3u : 0u, ny = (octant & 2u) ? 4u : 1u, nz = (octant & 4u) ? 5u : 2u;
  const uint32_t fx = 3u - nx, fy = 5u - ny, fz = 7u - nz;
  const float kInf = __int_as_float(0x7F800000);
  while (true) {
    const uint32_t base = ref * 8u;
    float4 nrx, nry, nrz, frx, fry, frz;
    uint4 ch;
    if (kSmemTop && ref < n_top) {
      nrx = smem_top[base + nx]; nry = smem_top[base + ny]; nrz = smem_top[base + nz];
      frx = smem_top[base + fx]; fry = smem_top[base + fy]; frz = smem_top[base + fz];
      ch = *reinterpret_cast<const uint4*>(smem_top + base + 6u);
    } else {
      const float4* np = nodes + base;
      nrx = __ldg(np + nx); nry = __ldg(np + ny); nrz = __ldg(np + nz);
      frx = __ldg(np + fx); fry = __ldg(np + fy); frz = __ldg(np + fz);
      ch = __ldg(reinterpret_cast<const uint4*>(np + 6));
    }

    float t0, t1, t2, t3, x0, x1, x2, x3;
#define Q3DR_SLAB(NX, NY, NZ, FX, FY, FZ, T, X)                                                            \
    {                                                                                                       \
      const float lx = __fmul_rn(__fsub_rn(NX, ray.ox), ray.ix), hx = __fmul_rn(__fsub_rn(FX, ray.ox), ray.ix); \
      const float ly = __fmul_rn(__fsub_rn(NY, ray.oy), ray.iy), hy = __fmul_rn(__fsub_rn(FY, ray.oy), ray.iy); \
      const float lz = __fmul_rn(__fsub_rn(NZ, ray.oz), ray.iz), hz = __fmul_rn(__fsub_rn(FZ, ray.oz), ray.iz); \
      (X) = fminf(fminf(fminf(hx, hy), hz), FLT_MAX);                                                       \
      (T) = fmaxf(fmaxf(fmaxf(lx, ly), lz), 0.0f);                                                          \
    }
    Q3DR_SLAB(nrx.x, nry.x, nrz.x, frx.x, fry.x, frz.x, t0, x0)
    Q3DR_SLAB(nrx.y, nry.y, nrz.y, frx.y, fry.y, frz.y, t1, x1)
    Q3DR_SLAB(nrx.z, nry.z, nrz.z, frx.z, fry.z, frz.z, t2, x2)
    Q3DR_SLAB(nrx.w, nry.w, nrz.w, frx.w, fry.w, frz.w, t3, x3)
#undef Q3DR_SLAB
    bool h0 = x0 > t0, h1 = x1 > t1, h2 = x2 > t2, h3 = x3 > t3;
    // tmax == 0 for some child of some lane (exit parameters are >= 0 whenever they matter, so the least
    // magnitude says it): the lane's origin lies on a face plane of that box, where the reference's inside /
    // outside distinction decides -- redo this node with the reference's full box test.  Rare.
    if (__any_sync(Q3DR_FULL_MASK, fminf(fminf(fabsf(x0), fabsf(x1)), fminf(fabsf(x2), fabsf(x3))) == 0.0f)) {
      const float4* np = nodes + base;
      const float4 mnx = np[0], mny = np[1], mnz = np[2], mxx = np[3], mxy = np[4], mxz = np[5];
      float dd;
      h0 = box_test(ray, mnx.x, mny.x, mnz.x, mxx.x, mxy.x, mxz.x, dd, t0);
      h1 = box_test(ray, mnx.y, mny.y, mnz.y, mxx.y, mxy.y, mxz.y, dd, t1);
      h2 = box_test(ray, mnx.z, mny.z, mnz.z, mxx.z, mxy.z, mxz.z, dd, t2);
      h3 = box_test(ray, mnx.w, mny.w, mnz.w, mxx.w, mxy.w, mxz.w, dd, t3);
    }

    // per-lane candidate bits -> warp-wide "some lane still needs child c" mask (uniform)
    uint32_t mine = ((h0 && t0 < best.tU) ? 1u : 0u) | ((h1 && t1 < best.tU) ? 2u : 0u) |
                    ((h2 && t2 < best.tU) ? 4u : 0u) | ((h3 && t3 < best.tU) ? 8u : 0u);
    uint32_t any = __reduce_or_sync(Q3DR_FULL_MASK, mine);
    const uint32_t leaf_mask = (ch.x >> kLeafMaskShift) & 0xFu;

    // leaves first: they tighten the bound before the inner children are judged
    uint32_t todo = any & leaf_mask;
    if (todo) {
      bool tightened = false;
      do {
        const uint32_t k = (uint32_t)__ffs((int)todo) - 1u;
        todo &= todo - 1u;
        const float T = pick4f(k, t0, t1, t2, t3);
        const int32_t li = (int32_t)(pick4(k, ch.x, ch.y, ch.z, ch.w) & kPayloadMask);
        const float d = dist_sq_at(ray, T);
        const bool acc = ((mine >> k) & 1u) && (d < best.dsq || (d == best.dsq && li > best.leaf));
        if (__any_sync(Q3DR_FULL_MASK, acc)) {
          const float u = prune_bound(ray, T, d);
          if (acc) {
            best.dsq = d;
            best.t = T;
            best.leaf = li;
            best.tU = u;
          }
          tightened = true;
        }
      } while (todo);
      if (tightened && (any & ~leaf_mask)) {
        mine = ((h0 && t0 < best.tU) ? 1u : 0u) | ((h1 && t1 < best.tU) ? 2u : 0u) |
               ((h2 && t2 < best.tU) ? 4u : 0u) | ((h3 && t3 < best.tU) ? 8u : 0u);
        any = __reduce_or_sync(Q3DR_FULL_MASK, mine);
      }
    }

    // inner children
    uint32_t inner = any & ~leaf_mask;
    if (inner) {
      if (inner & (inner - 1u)) {
        // more than one: nearest first; the others go on the stack, farthest first, each with every lane's own
        // entry parameter (+inf for lanes that do not need it)
        uint32_t key[4];
        key[0] = (inner & 1u) ? __reduce_min_sync(Q3DR_FULL_MASK, (mine & 1u) ? __float_as_uint(t0) : 0xFFFFFFFFu) : 0xFFFFFFFFu;
        key[1] = (inner & 2u) ? __reduce_min_sync(Q3DR_FULL_MASK, (mine & 2u) ? __float_as_uint(t1) : 0xFFFFFFFFu) : 0xFFFFFFFFu;
        key[2] = (inner & 4u) ? __reduce_min_sync(Q3DR_FULL_MASK, (mine & 4u) ? __float_as_uint(t2) : 0xFFFFFFFFu) : 0xFFFFFFFFu;
        key[3] = (inner & 8u) ? __reduce_min_sync(Q3DR_FULL_MASK, (mine & 8u) ? __float_as_uint(t3) : 0xFFFFFFFFu) : 0xFFFFFFFFu;
        while (inner & (inner - 1u)) {
          // farthest remaining candidate (ties: higher slot)
          uint32_t far = 0u, far_key = 0u;
#pragma unroll
          for (uint32_t c = 0; c < 4u; ++c) {
            if (((inner >> c) & 1u) && key[c] >= far_key) {
              far = c;
              far_key = key[c];
            }
          }
          inner &= ~(1u << far);
          stack_t[sp] = ((mine >> far) & 1u) ? pick4f(far, t0, t1, t2, t3) : kInf;
          stack_n[sp] = pick4(far, ch.x, ch.y, ch.z, ch.w) & kPayloadMask;
          ++sp;
        }
      }
      ref = pick4((uint32_t)__ffs((int)inner) - 1u, ch.x, ch.y, ch.z, ch.w) & kPayloadMask;
      continue;
    }
    bool found = false;
    while (sp > 0) {
      --sp;
      const float d = stack_t[sp];
      const uint32_t n = __shfl_sync(Q3DR_FULL_MASK, stack_n[sp], 0);  // provably warp-uniform (see raycast2_kernels.cuh)
      if (__any_sync(Q3DR_FULL_MASK, d < best.tU)) {
        ref = n;
        found = true;
        break;
      }
    }
    if (!found) break;
  }
}

// Ray lists (q3dr_intersect_rays: origins + directions) and pixel lists of one viewpoint (the second pass of
// q3dr_raycast_unique): 32 rays per warp, one per lane, q3dr_pixel_hit records out.  Camera windows are cast by
// raycast2_kernel (raycast2_kernels.cuh).
static __global__ void __launch_bounds__(kThreadsPerCta, 4) raycast_rays_kernel(const RaycastParams p) {
  extern __shared__ float4 smem_top[];
  __shared__ uint32_t s_tile[kWarpsPerCta];
  for (uint32_t i = threadIdx.x; i < p.n_top * 8; i += blockDim.x) smem_top[i] = p.nodes[i];
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const float range_sq = (p.max_range > 0.0f) ? __fmul_rn(p.max_range, p.max_range) : FLT_MAX;

  while (true) {
    if (lane == 0) s_tile[warp] = atomicAdd(p.tile_counter, 1u);
    __syncwarp();
    const uint32_t tile = s_tile[warp];
    __syncwarp();
    if (tile >= p.total_tiles) break;

    Ray ray;
    PacketHit best;
    best.leaf = -1;
    best.t = 0.0f;
    best.tU = 0.0f;
    uint32_t px = 0, py = 0;
    const uint32_t i = tile * 32u + lane;
    const bool live = i < p.n_rays;
    if (live && p.pix_list != nullptr) {
      const uint32_t wp = __ldg(p.pix_list + i);
      py = wp / p.w;
      px = wp - py * p.w;
      camera_ray(p.rt, p.fx, p.fy, p.cx, p.cy, p.x0 + px, p.y0 + py, ray);
    } else if (live) {
      make_ray(__ldg(p.ray_o + 3 * (size_t)i), __ldg(p.ray_o + 3 * (size_t)i + 1), __ldg(p.ray_o + 3 * (size_t)i + 2),
               __ldg(p.ray_d + 3 * (size_t)i), __ldg(p.ray_d + 3 * (size_t)i + 1), __ldg(p.ray_d + 3 * (size_t)i + 2), ray);
    }
    const uint32_t live_mask = __ballot_sync(Q3DR_FULL_MASK, live);
    if (live_mask == 0u) continue;
    {
      // a dead lane rides along with a copy of a live lane's ray and never accepts anything
      // (dist_sq >= 0 > -1, t >= 0 > -1)
      const int src = __ffs(live_mask) - 1;
      const float ox = __shfl_sync(Q3DR_FULL_MASK, ray.ox, src), oy = __shfl_sync(Q3DR_FULL_MASK, ray.oy, src),
                  oz = __shfl_sync(Q3DR_FULL_MASK, ray.oz, src);
      const float dx = __shfl_sync(Q3DR_FULL_MASK, ray.dx, src), dy = __shfl_sync(Q3DR_FULL_MASK, ray.dy, src),
                  dz = __shfl_sync(Q3DR_FULL_MASK, ray.dz, src);
      const float ix = __shfl_sync(Q3DR_FULL_MASK, ray.ix, src), iy = __shfl_sync(Q3DR_FULL_MASK, ray.iy, src),
                  iz = __shfl_sync(Q3DR_FULL_MASK, ray.iz, src);
      if (!live) {
        ray.ox = ox; ray.oy = oy; ray.oz = oz;
        ray.dx = dx; ray.dy = dy; ray.dz = dz;
        ray.ix = ix; ray.iy = iy; ray.iz = iz;
      }
    }
    best.dsq = live ? range_sq : -1.0f;

    // fast path: every inverse direction finite (no 0 * inf, so no NaN anywhere) and one octant per packet
    const bool finite = (fabsf(ray.ix) <= FLT_MAX) && (fabsf(ray.iy) <= FLT_MAX) && (fabsf(ray.iz) <= FLT_MAX);
    const uint32_t bx = __ballot_sync(Q3DR_FULL_MASK, ray.ix < 0.0f);
    const uint32_t by = __ballot_sync(Q3DR_FULL_MASK, ray.iy < 0.0f);
    const uint32_t bz = __ballot_sync(Q3DR_FULL_MASK, ray.iz < 0.0f);
    const bool uniform_octant = (bx == 0u || bx == Q3DR_FULL_MASK) && (by == 0u || by == Q3DR_FULL_MASK) &&
                                (bz == 0u || bz == Q3DR_FULL_MASK);
    if (__all_sync(Q3DR_FULL_MASK, finite) && uniform_octant) {
      const float kInf = __int_as_float(0x7F800000);
      best.tU = live ? ((p.max_range > 0.0f) ? prune_bound(ray, p.max_range, range_sq) : kInf) : -1.0f;
      const uint32_t octant = (bx ? 1u : 0u) | (by ? 2u : 0u) | (bz ? 4u : 0u);
      traverse_packet_fast<kSmemTopLevels>(ray, best, p.nodes, smem_top, p.n_top, octant);
    } else {
      best.tU = 0.0f;
      traverse_packet_exact(ray, best, p.nodes, smem_top, p.n_top);
    }

    if (live) {
      const bool hit = best.leaf >= 0;
      // intersection = origin + direction * t  (geometry.h:345-348); t = 0 when the origin is inside
      const float hx = hit ? __fadd_rn(ray.ox, __fmul_rn(ray.dx, best.t)) : 0.0f;
      const float hy = hit ? __fadd_rn(ray.oy, __fmul_rn(ray.dy, best.t)) : 0.0f;
      const float hz = hit ? __fadd_rn(ray.oz, __fmul_rn(ray.dz, best.t)) : 0.0f;
      const uint32_t depth = hit ? (uint32_t)__ldg(p.leaf_depth + best.leaf) : 0u;
      const bool has_xy = p.pix_list != nullptr;
      float4* dst = reinterpret_cast<float4*>(p.pixels + (size_t)i);
      dst[0] = make_float4(hx, hy, hz, __int_as_float(best.leaf));
      dst[1] = make_float4(__uint_as_float(depth), hit ? best.dsq : 0.0f, has_xy ? (float)(p.x0 + px) : 0.0f,
                           has_xy ? (float)(p.y0 + py) : 0.0f);
    }
  }
}

static __global__ void fill_u32_kernel(uint32_t* p, size_t n, uint32_t v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace q3dr
