// raycast2_kernels.cuh -- the camera-window traversal kernel: one warp = one 8x8 pixel packet, TWO rays per lane.
//
// Every ray of a viewpoint starts at the camera centre, so the 24 plane-minus-origin differences of a 4-wide node
// are the same for all rays; with two rays per lane they, the node fetch and the whole scalar epilogue of a node
// visit (candidate mask, leaf loop, descent / stack) are paid once per 64 rays instead of once per 32.  The numeric
// contract is the one of raycast_kernels.cuh: every value that feeds a hit decision is a
// single correctly rounded binary32 operation in the reference's order.
//
//   traverse_packet2<false>  all 64 rays share one direction octant: near / far planes are picked by address
//   traverse_packet2<true>   mixed octants (packets straddling an axis): lo = min(t0, t1), hi = max(t0, t1) per
//                            axis like the reference writes it (no NaN can arise with finite inverse directions,
//                            so fminf / fmaxf equal the reference's ternaries)
//   packets with a non-finite inverse direction (exactly axis-parallel rays) take traverse_packet_exact per ray.
#pragma once

#include "raycast_kernels.cuh"

namespace q3dr {

constexpr int kTile2W = 8;
constexpr int kTile2H = 8;

struct Dir2 {
  float dx, dy, dz;
  float ix, iy, iz;
};

struct Best2 {
  float dsq;     // best dist_sq (the range bound while nothing is hit)
  float t;       // ray parameter of the best hit
  float tU;      // every ray parameter >= tU has dist_sq > dsq in the reference's arithmetic (see bound_after)
  int32_t leaf;  // best leaf index, -1 = none
};

__device__ __forceinline__ float dist_sq_at2(float ox, float oy, float oz, const Dir2& r, float t) {
  const float px = __fadd_rn(ox, __fmul_rn(r.dx, t));
  const float py = __fadd_rn(oy, __fmul_rn(r.dy, t));
  const float pz = __fadd_rn(oz, __fmul_rn(r.dz, t));
  const float ex = __fsub_rn(ox, px), ey = __fsub_rn(oy, py), ez = __fsub_rn(oz, pz);
  return __fadd_rn(__fmul_rn(ex, ex), __fadd_rn(__fmul_rn(ey, ey), __fmul_rn(ez, ez)));
}

// ---------------------------------------------------------------------------------------------------------------
// Pruning bound.  best.dsq is the reference's dist_sq of the best hit, computed at ray parameter T.  The traversal
// skips a box on its ENTRY PARAMETER t alone when t >= tU, where tU must guarantee that the reference's computed
// dist_sq at any parameter >= tU is strictly greater than best.dsq (the reference prunes on dist_sq > best only).
//
//   tU = fl(fl(T * (1 + 2^-11)) + A),   A = 2^-13 + 2^-20 * M,   M = max |origin component|
//
// Proof sketch (eps = 2^-24, round to nearest, |dir| = 1 +- 4 eps for a direction normalised from a squared norm in
// the normal range, no overflow: |origin|, t < 2^60).  For parameter t the reference computes
//   e_i = fl(o_i - fl(o_i + fl(d_i t))) = -d_i t + eta_i,  |eta_i| <= eps (|o_i| + 3 t)(1 + 8 eps) =: H(t)
//   dist_sq = fl-sum of e_i^2 = |e|^2 (1 + theta), |theta| <= 3.01 eps                 (all terms non-negative)
// hence  sqrt(dist_sq(t)) in [ L(t), U(t) ],  L(t) = (t (1 - 4 eps) - sqrt(3) H(t)) (1 - 2 eps),
//                                             U(t) = (t (1 + 4 eps) + sqrt(3) H(t)) (1 + 2 eps),  L increasing in t.
// L(u) > U(T) holds as soon as u - T > eps (24 T + 3.5 M) (1 + o(1)) = 2^-19.4 T + 2^-22.2 M; the tU above exceeds
// T by at least 2^-11.01 T + 0.99 A, i.e. by a factor > 4 on either term, and L(tU) > 0 because tU >= A >= 2^-20 M.
// So dist_sq(t') >= L(t')^2 >= L(tU)^2 > U(T)^2 >= best.dsq for every t' >= tU.  The same bound with T = max_range
// covers the initial range limit (best.dsq = fl(max_range^2) <= U(max_range)^2).  The kernel checks the
// preconditions per packet and sends every packet that violates one to traverse_packet_exact.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float bound_after(float t, float A) { return fmaf(t, 1.0f / 2048.0f, t) + A; }

// kMixed = false: all 64 rays share one direction octant, near / far planes are picked by address.
// kMixed = true : lo = min(t0, t1), hi = max(t0, t1) per axis like the reference writes it.
// The reference's clamp of tmax at FLT_MAX is omitted: it can only bind when all three exit parameters overflow,
// which a normalised direction (|inv| <= 2 on some axis) cannot produce over a map bounded by 1e18 -- a packet
// precondition as well.  The exit parameter is kept as x' = min(exit, tU): "hit and still needed" is then entry < x'.
template <bool kMixed>
__device__ __forceinline__ void traverse_packet2(float ox, float oy, float oz, float A, const Dir2& r0, const Dir2& r1,
                                                 Best2& b0, Best2& b1, const float4* __restrict__ nodes, uint32_t octant) {
  float stack_t0[kStackEntries];
  float stack_t1[kStackEntries];
  uint32_t stack_n[kStackEntries];
  int sp = 0;
  uint32_t ref = 0;
  // plane indices inside a node: mn x,y,z = 0,1,2 ; mx x,y,z = 3,4,5
  const uint32_t nx = (!kMixed && (octant & 1u)) ? 3u : 0u, ny = (!kMixed && (octant & 2u)) ? 4u : 1u,
                 nz = (!kMixed && (octant & 4u)) ? 5u : 2u;
  const uint32_t fx = 3u - nx, fy = 5u - ny, fz = 7u - nz;
  const float kInf = __int_as_float(0x7F800000);
  while (true) {
    const float4* np = nodes + (size_t)ref * 8u;
    const float4 nrx = __ldg(np + nx), nry = __ldg(np + ny), nrz = __ldg(np + nz);
    const float4 frx = __ldg(np + fx), fry = __ldg(np + fy), frz = __ldg(np + fz);
    const uint4 ch = __ldg(reinterpret_cast<const uint4*>(np + 6));

    float t00, t01, t02, t03, t10, t11, t12, t13;  // entry parameter, [ray][child]
    float x00, x01, x02, x03, x10, x11, x12, x13;  // min(exit parameter, tU)
#define Q3DR_SLAB1(R, B, TA, XA)                                                                            \
      {                                                                                                     \
        float lx = __fmul_rn(anx, R.ix), hx = __fmul_rn(afx, R.ix);                                         \
        float ly = __fmul_rn(any_, R.iy), hy = __fmul_rn(afy, R.iy);                                        \
        float lz = __fmul_rn(anz, R.iz), hz = __fmul_rn(afz, R.iz);                                         \
        if (kMixed) {                                                                                       \
          const float sx = fminf(lx, hx), sy = fminf(ly, hy), sz = fminf(lz, hz);                           \
          hx = fmaxf(lx, hx); hy = fmaxf(ly, hy); hz = fmaxf(lz, hz);                                       \
          lx = sx; ly = sy; lz = sz;                                                                        \
        }                                                                                                   \
        (XA) = fminf(fminf(fminf(hx, hy), hz), B.tU);                                                       \
        (TA) = fmaxf(fmaxf(fmaxf(lx, ly), lz), 0.0f);                                                       \
      }
#define Q3DR_SLAB2(C, TA, XA, TB, XB)                                                                      \
    {                                                                                                       \
      const float anx = __fsub_rn(nrx.C, ox), afx = __fsub_rn(frx.C, ox);                                   \
      const float any_ = __fsub_rn(nry.C, oy), afy = __fsub_rn(fry.C, oy);                                  \
      const float anz = __fsub_rn(nrz.C, oz), afz = __fsub_rn(frz.C, oz);                                   \
      Q3DR_SLAB1(r0, b0, TA, XA)                                                                            \
      Q3DR_SLAB1(r1, b1, TB, XB)                                                                            \
    }
    Q3DR_SLAB2(x, t00, x00, t10, x10)
    Q3DR_SLAB2(y, t01, x01, t11, x11)
    Q3DR_SLAB2(z, t02, x02, t12, x12)
    Q3DR_SLAB2(w, t03, x03, t13, x13)
#undef Q3DR_SLAB2
#undef Q3DR_SLAB1
    // exit parameter == 0 for some child of some ray (tU > 0, so x' == 0 says exactly that): the ray's origin lies
    // on a face plane of the box, where the reference's inside / outside distinction decides -- redo this node with
    // the reference's full box test and express its verdict in the same form.  Rare.
    {
      const float m0 = fminf(fminf(fabsf(x00), fabsf(x01)), fminf(fabsf(x02), fabsf(x03)));
      const float m1 = fminf(fminf(fabsf(x10), fabsf(x11)), fminf(fabsf(x12), fabsf(x13)));
      if (__any_sync(Q3DR_FULL_MASK, fminf(m0, m1) == 0.0f)) {
        const float4 mnx = np[0], mny = np[1], mnz = np[2], mxx = np[3], mxy = np[4], mxz = np[5];
        Ray q;
        float dd;
        q.ox = ox; q.oy = oy; q.oz = oz;
        q.dx = r0.dx; q.dy = r0.dy; q.dz = r0.dz; q.ix = r0.ix; q.iy = r0.iy; q.iz = r0.iz;
        x00 = box_test(q, mnx.x, mny.x, mnz.x, mxx.x, mxy.x, mxz.x, dd, t00) ? b0.tU : -1.0f;
        x01 = box_test(q, mnx.y, mny.y, mnz.y, mxx.y, mxy.y, mxz.y, dd, t01) ? b0.tU : -1.0f;
        x02 = box_test(q, mnx.z, mny.z, mnz.z, mxx.z, mxy.z, mxz.z, dd, t02) ? b0.tU : -1.0f;
        x03 = box_test(q, mnx.w, mny.w, mnz.w, mxx.w, mxy.w, mxz.w, dd, t03) ? b0.tU : -1.0f;
        q.dx = r1.dx; q.dy = r1.dy; q.dz = r1.dz; q.ix = r1.ix; q.iy = r1.iy; q.iz = r1.iz;
        x10 = box_test(q, mnx.x, mny.x, mnz.x, mxx.x, mxy.x, mxz.x, dd, t10) ? b1.tU : -1.0f;
        x11 = box_test(q, mnx.y, mny.y, mnz.y, mxx.y, mxy.y, mxz.y, dd, t11) ? b1.tU : -1.0f;
        x12 = box_test(q, mnx.z, mny.z, mnz.z, mxx.z, mxy.z, mxz.z, dd, t12) ? b1.tU : -1.0f;
        x13 = box_test(q, mnx.w, mny.w, mnz.w, mxx.w, mxy.w, mxz.w, dd, t13) ? b1.tU : -1.0f;
      }
    }

    // "some ray still needs child c"
    uint32_t any = (__any_sync(Q3DR_FULL_MASK, (t00 < x00) || (t10 < x10)) ? 1u : 0u) |
                   (__any_sync(Q3DR_FULL_MASK, (t01 < x01) || (t11 < x11)) ? 2u : 0u) |
                   (__any_sync(Q3DR_FULL_MASK, (t02 < x02) || (t12 < x12)) ? 4u : 0u) |
                   (__any_sync(Q3DR_FULL_MASK, (t03 < x03) || (t13 < x13)) ? 8u : 0u);
    const uint32_t leaf_mask = (ch.x >> kLeafMaskShift) & 0xFu;

    // leaves first: they tighten the bounds before the inner children are judged.  A ray accepts a leaf on the
    // reference's own terms: the slab test passes and dist_sq < best, or == best with a higher leaf index
    // (x' > t is implied for every ray that can pass the dist_sq comparison, see bound_after).
    const uint32_t todo = any & leaf_mask;
    if (todo) {
      bool tightened = false;
#define Q3DR_LEAF2(BIT, CH, TA, XA, TB, XB)                                                        \
      if (todo & (BIT)) {                                                                           \
        const int32_t li = (int32_t)((CH) & kPayloadMask);                                          \
        const float d0 = dist_sq_at2(ox, oy, oz, r0, (TA));                                         \
        const float d1 = dist_sq_at2(ox, oy, oz, r1, (TB));                                         \
        const bool acc0 = ((XA) > (TA)) && (d0 < b0.dsq || (d0 == b0.dsq && li > b0.leaf));         \
        const bool acc1 = ((XB) > (TB)) && (d1 < b1.dsq || (d1 == b1.dsq && li > b1.leaf));         \
        if (__any_sync(Q3DR_FULL_MASK, acc0 || acc1)) {                                             \
          if (acc0) {                                                                               \
            b0.dsq = d0; b0.t = (TA); b0.leaf = li; b0.tU = bound_after((TA), A);                   \
          }                                                                                         \
          if (acc1) {                                                                               \
            b1.dsq = d1; b1.t = (TB); b1.leaf = li; b1.tU = bound_after((TB), A);                   \
          }                                                                                         \
          tightened = true;                                                                         \
        }                                                                                           \
      }
      Q3DR_LEAF2(1u, ch.x, t00, x00, t10, x10)
      Q3DR_LEAF2(2u, ch.y, t01, x01, t11, x11)
      Q3DR_LEAF2(4u, ch.z, t02, x02, t12, x12)
      Q3DR_LEAF2(8u, ch.w, t03, x03, t13, x13)
#undef Q3DR_LEAF2
      if (tightened && (any & ~leaf_mask)) {
        x00 = fminf(x00, b0.tU); x01 = fminf(x01, b0.tU); x02 = fminf(x02, b0.tU); x03 = fminf(x03, b0.tU);
        x10 = fminf(x10, b1.tU); x11 = fminf(x11, b1.tU); x12 = fminf(x12, b1.tU); x13 = fminf(x13, b1.tU);
        any = (__any_sync(Q3DR_FULL_MASK, (t00 < x00) || (t10 < x10)) ? 1u : 0u) |
              (__any_sync(Q3DR_FULL_MASK, (t01 < x01) || (t11 < x11)) ? 2u : 0u) |
              (__any_sync(Q3DR_FULL_MASK, (t02 < x02) || (t12 < x12)) ? 4u : 0u) |
              (__any_sync(Q3DR_FULL_MASK, (t03 < x03) || (t13 < x13)) ? 8u : 0u);
      }
    }

    // inner children
    uint32_t inner = any & ~leaf_mask;
    if (inner) {
      if (inner & (inner - 1u)) {
        // more than one: nearest first; the others go on the stack, farthest first, with every ray's own entry
        // parameter (+inf for rays that do not need the child)
        const uint32_t kNone = 0xFFFFFFFFu;
#define Q3DR_KEY(C, TA, XA, TB, XB)                                                                       \
        ((inner & (1u << (C)))                                                                            \
             ? __reduce_min_sync(Q3DR_FULL_MASK, min(((TA) < (XA)) ? __float_as_uint(TA) : kNone,          \
                                                    ((TB) < (XB)) ? __float_as_uint(TB) : kNone))          \
             : kNone)
        uint32_t key[4];
        key[0] = Q3DR_KEY(0, t00, x00, t10, x10);
        key[1] = Q3DR_KEY(1, t01, x01, t11, x11);
        key[2] = Q3DR_KEY(2, t02, x02, t12, x12);
        key[3] = Q3DR_KEY(3, t03, x03, t13, x13);
#undef Q3DR_KEY
        while (inner & (inner - 1u)) {
          uint32_t far = 0u, far_key = 0u;  // farthest remaining candidate (ties: higher slot)
#pragma unroll
          for (uint32_t c = 0; c < 4u; ++c) {
            if (((inner >> c) & 1u) && key[c] >= far_key) {
              far = c;
              far_key = key[c];
            }
          }
          inner &= ~(1u << far);
          const float T0 = pick4f(far, t00, t01, t02, t03), X0 = pick4f(far, x00, x01, x02, x03);
          const float T1 = pick4f(far, t10, t11, t12, t13), X1 = pick4f(far, x10, x11, x12, x13);
          stack_t0[sp] = (T0 < X0) ? T0 : kInf;  // tU only shrinks: an entry parameter >= tU now stays >= tU
          stack_t1[sp] = (T1 < X1) ? T1 : kInf;
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
      const float e0 = stack_t0[sp], e1 = stack_t1[sp];
      const uint32_t n = stack_n[sp];
      if (__any_sync(Q3DR_FULL_MASK, e0 < b0.tU || e1 < b1.tU)) {
        ref = n;
        found = true;
        break;
      }
    }
    if (!found) break;
  }
}

// Camera ray direction of pixel column x (c0 precomputed) and row y: see camera_ray in raycast_kernels.cuh.
// Returns whether the squared norm lay in [1e-30, 1e30] (the direction is then normalised to 1 +- 4 eps).
__device__ __forceinline__ bool camera_dir2(const float* __restrict__ R /* 12 floats in registers */, float c0, float fy,
                                            float cy, uint32_t y, Dir2& r) {
  const float c1 = __fdiv_rn(__fsub_rn((float)y, cy), fy);
  const float d0 = __fadd_rn(__fmul_rn(R[0], c0), __fadd_rn(__fmul_rn(R[1], c1), R[2]));
  const float d1 = __fadd_rn(__fmul_rn(R[4], c0), __fadd_rn(__fmul_rn(R[5], c1), R[6]));
  const float d2 = __fadd_rn(__fmul_rn(R[8], c0), __fadd_rn(__fmul_rn(R[9], c1), R[10]));
  const float n2 = __fadd_rn(__fmul_rn(d0, d0), __fadd_rn(__fmul_rn(d1, d1), __fmul_rn(d2, d2)));
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
  return n2 >= 1e-30f && n2 <= 1e30f;
}

__device__ __forceinline__ void copy_dir_from_lane(Dir2& r, const Dir2& src_ray, int src_lane, bool take) {
  const float dx = __shfl_sync(Q3DR_FULL_MASK, src_ray.dx, src_lane), dy = __shfl_sync(Q3DR_FULL_MASK, src_ray.dy, src_lane),
              dz = __shfl_sync(Q3DR_FULL_MASK, src_ray.dz, src_lane);
  const float ix = __shfl_sync(Q3DR_FULL_MASK, src_ray.ix, src_lane), iy = __shfl_sync(Q3DR_FULL_MASK, src_ray.iy, src_lane),
              iz = __shfl_sync(Q3DR_FULL_MASK, src_ray.iz, src_lane);
  if (take) {
    r.dx = dx; r.dy = dy; r.dz = dz;
    r.ix = ix; r.iy = iy; r.iz = iz;
  }
}

// MODE: kModeScore or kModePixels (camera windows only; ray lists use raycast_rays_kernel).
template <int MODE>
__global__ void __launch_bounds__(kThreadsPerCta, 4) raycast2_kernel(const RaycastParams p) {
  __shared__ uint32_t s_tile[kWarpsPerCta];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const float range_sq = (p.max_range > 0.0f) ? __fmul_rn(p.max_range, p.max_range) : FLT_MAX;
  const float kInf = __int_as_float(0x7F800000);

  while (true) {
    if (lane == 0) s_tile[warp] = atomicAdd(p.tile_counter, 1u);
    __syncwarp();
    const uint32_t tile = s_tile[warp];
    __syncwarp();
    if (tile >= p.total_tiles) break;

    const uint32_t vp = tile / p.tiles_per_vp;
    const uint32_t rem = tile - vp * p.tiles_per_vp;
    const uint32_t ty = rem / p.tiles_x;
    const uint32_t tx = rem - ty * p.tiles_x;
    const uint32_t px = tx * kTile2W + (lane & 7);
    const uint32_t py0 = ty * kTile2H + (lane >> 3);
    const uint32_t py1 = py0 + 4u;
    const bool live0 = (px < p.w) && (py0 < p.h);
    const bool live1 = (px < p.w) && (py1 < p.h);
    const uint32_t live_mask = __ballot_sync(Q3DR_FULL_MASK, live0);
    if (live_mask == 0u) continue;  // live1 implies live0

    float R[12];
    {
      const float4* rt4 = reinterpret_cast<const float4*>(p.rt + 12 * (size_t)vp);
      const float4 a = __ldg(rt4), b = __ldg(rt4 + 1), c = __ldg(rt4 + 2);
      R[0] = a.x; R[1] = a.y; R[2] = a.z; R[3] = a.w;
      R[4] = b.x; R[5] = b.y; R[6] = b.z; R[7] = b.w;
      R[8] = c.x; R[9] = c.y; R[10] = c.z; R[11] = c.w;
    }
    const float ox = R[3], oy = R[7], oz = R[11];
    const float c0 = __fdiv_rn(__fsub_rn((float)(p.x0 + px), p.cx), p.fx);
    Dir2 r0, r1;
    const bool unit0 = camera_dir2(R, c0, p.fy, p.cy, p.y0 + py0, r0);
    const bool unit1 = camera_dir2(R, c0, p.fy, p.cy, p.y0 + py1, r1);
    {
      // a dead ray rides along with a copy of a live one and never accepts anything (dist_sq >= 0 > -1, t >= 0 > -1)
      const int src = __ffs(live_mask) - 1;
      if (live_mask != Q3DR_FULL_MASK || !__all_sync(Q3DR_FULL_MASK, live1)) {
        const Dir2 s = r0;
        copy_dir_from_lane(r0, s, src, !live0);
        copy_dir_from_lane(r1, s, src, !live1);
      }
    }
    Best2 b0, b1;
    b0.leaf = b1.leaf = -1;
    b0.t = b1.t = 0.0f;
    b0.dsq = live0 ? range_sq : -1.0f;
    b1.dsq = live1 ? range_sq : -1.0f;

    // preconditions of the fast loops (see bound_after / traverse_packet2); dead rays are copies of live ones
    const float M = fmaxf(fmaxf(fabsf(ox), fabsf(oy)), fabsf(oz));
    const bool finite = (fabsf(r0.ix) <= FLT_MAX) && (fabsf(r0.iy) <= FLT_MAX) && (fabsf(r0.iz) <= FLT_MAX) &&
                        (fabsf(r1.ix) <= FLT_MAX) && (fabsf(r1.iy) <= FLT_MAX) && (fabsf(r1.iz) <= FLT_MAX);
    const bool tame = finite && (unit0 || !live0) && (unit1 || !live1) && p.bounded_map && (M <= 1e18f) &&
                      !(p.max_range > 1e18f);
    if (__all_sync(Q3DR_FULL_MASK, tame)) {
      const float A = fmaf(M, 1.0f / 1048576.0f, 1.0f / 8192.0f);
      const float u = (p.max_range > 0.0f) ? bound_after(p.max_range, A) : kInf;
      b0.tU = live0 ? u : -1.0f;
      b1.tU = live1 ? u : -1.0f;
      // per axis: all rays positive, all negative, or mixed
      const uint32_t bx = __ballot_sync(Q3DR_FULL_MASK, r0.ix < 0.0f), cx = __ballot_sync(Q3DR_FULL_MASK, r1.ix < 0.0f);
      const uint32_t by = __ballot_sync(Q3DR_FULL_MASK, r0.iy < 0.0f), cy = __ballot_sync(Q3DR_FULL_MASK, r1.iy < 0.0f);
      const uint32_t bz = __ballot_sync(Q3DR_FULL_MASK, r0.iz < 0.0f), cz = __ballot_sync(Q3DR_FULL_MASK, r1.iz < 0.0f);
      const bool ux = ((bx | cx) == 0u) || ((bx & cx) == Q3DR_FULL_MASK);
      const bool uy = ((by | cy) == 0u) || ((by & cy) == Q3DR_FULL_MASK);
      const bool uz = ((bz | cz) == 0u) || ((bz & cz) == Q3DR_FULL_MASK);
      if (ux && uy && uz) {
        const uint32_t octant = (bx ? 1u : 0u) | (by ? 2u : 0u) | (bz ? 4u : 0u);
        traverse_packet2<false>(ox, oy, oz, A, r0, r1, b0, b1, p.nodes, octant);
      } else {
        traverse_packet2<true>(ox, oy, oz, A, r0, r1, b0, b1, p.nodes, 0u);
      }
    } else {
      // some ray is exactly parallel to an axis, or the packet is outside the fast loops' proven domain (huge
      // coordinates, degenerate pose): the literal reference box test, one ray after the other
      Ray q;
      PacketHit h;
      q.ox = ox; q.oy = oy; q.oz = oz;
      q.dx = r0.dx; q.dy = r0.dy; q.dz = r0.dz; q.ix = r0.ix; q.iy = r0.iy; q.iz = r0.iz;
      h.dsq = b0.dsq; h.t = 0.0f; h.tU = 0.0f; h.leaf = -1;
      traverse_packet_exact(q, h, p.nodes, p.nodes, 0u);
      b0.dsq = h.dsq; b0.t = h.t; b0.leaf = h.leaf;
      q.dx = r1.dx; q.dy = r1.dy; q.dz = r1.dz; q.ix = r1.ix; q.iy = r1.iy; q.iz = r1.iz;
      h.dsq = b1.dsq; h.t = 0.0f; h.tU = 0.0f; h.leaf = -1;
      traverse_packet_exact(q, h, p.nodes, p.nodes, 0u);
      b1.dsq = h.dsq; b1.t = h.t; b1.leaf = h.leaf;
    }

    if (MODE == kModeScore) {
      // warp-aggregated dedup: within each half of the tile lanes are in row-major pixel order, so the lowest lane
      // of each equal-leaf group owns the group's first pixel; only group leaders touch the tables (atomicMin
      // merges the two halves and the other tiles of the viewpoint).
      uint32_t* fpv = p.first_pix + (size_t)vp * p.leaf_stride;
      uint32_t* bmv = p.bitmap + (size_t)vp * p.word_stride;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int32_t leaf = half ? b1.leaf : b0.leaf;
        const uint32_t pix = (half ? py1 : py0) * p.w + px;
        const uint32_t grp = __match_any_sync(Q3DR_FULL_MASK, leaf);
        const bool leader = (leaf >= 0) && ((__ffs(grp) - 1) == lane);
        if (leader) {
          const uint32_t old = atomicMin(fpv + (uint32_t)leaf, p.pix_tag | pix);
          if (old > p.pix_limit) {  // nothing of THIS chunk was there yet: first toucher
            atomicOr(bmv + ((uint32_t)leaf >> 5), 1u << (leaf & 31));
            atomicAdd(p.seg_hits + (size_t)vp * p.n_segs + ((uint32_t)leaf >> 15), 1u);
          }
        }
      }
    } else {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const bool live = half ? live1 : live0;
        if (!live) continue;
        const Best2& b = half ? b1 : b0;
        const Dir2& r = half ? r1 : r0;
        const uint32_t py = half ? py1 : py0;
        const bool hit = b.leaf >= 0;
        // intersection = origin + direction * t  (geometry.h:345-348); t = 0 when the origin is inside
        const float hx = hit ? __fadd_rn(ox, __fmul_rn(r.dx, b.t)) : 0.0f;
        const float hy = hit ? __fadd_rn(oy, __fmul_rn(r.dy, b.t)) : 0.0f;
        const float hz = hit ? __fadd_rn(oz, __fmul_rn(r.dz, b.t)) : 0.0f;
        const uint32_t depth = hit ? (uint32_t)__ldg(p.leaf_depth + b.leaf) : 0u;
        const size_t oi = (size_t)vp * p.w * p.h + (size_t)py * p.w + px;
        float4* dst = reinterpret_cast<float4*>(p.pixels + oi);
        dst[0] = make_float4(hx, hy, hz, __int_as_float(b.leaf));
        dst[1] = make_float4(__uint_as_float(depth), hit ? b.dsq : 0.0f, (float)(p.x0 + px), (float)(p.y0 + py));
      }
    }
  }
}

}  // namespace q3dr
