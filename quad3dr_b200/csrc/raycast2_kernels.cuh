// raycast2_kernels.cuh -- the camera-window traversal kernel: one warp = one pixel packet, SEVERAL rays per lane.
//
// Every ray of a viewpoint starts at the camera centre, so the 24 plane-minus-origin differences of a 4-wide node
// are the same for all rays; with NR rays per lane they, the node fetch and the whole scalar epilogue of a node
// visit (candidate mask, leaf loop, descent / stack) are paid once per 32 * NR rays instead of once per 32.  The
// numeric contract is the one of raycast_kernels.cuh: every value that feeds a hit decision
// is a single correctly rounded binary32 operation in the reference's order.
//
//   traverse_packetN<NR, false, .>  all rays share one direction octant: near / far planes are picked by address
//   traverse_packetN<NR, true, .>   mixed octants (packets straddling an axis): lo = min(t0, t1), hi = max(t0, t1)
//                                   per axis like the reference writes it (no NaN can arise with finite inverse
//                                   directions, so fminf / fmaxf equal the reference's ternaries)
//   packets with a non-finite inverse direction (exactly axis-parallel rays) take traverse_packet_exact per ray.
#pragma once

#include "raycast_kernels.cuh"

namespace q3dr {


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

// One warp = one packet of 32 * NR rays, NR rays per lane.  Every ray of a viewpoint starts at the camera centre, so the
// 24 plane-minus-origin differences of a 4-wide node are the same for all rays of the packet: they, the node fetch and
// the scalar epilogue of a node visit (candidate mask, leaf loop, descent / stack) are paid once per 32 * NR rays.
//
// kMixed = false: all rays of the packet share one direction octant, near / far planes are picked by address.
// kMixed = true : lo = min(t0, t1), hi = max(t0, t1) per axis like the reference writes it.
// kPacked       : the plane differences and the slab products of two CHILDREN at a time with sm_100's packed binary32
//                 instructions (FADD2 / FMUL2: two independent round-to-nearest-even operations per instruction, so
//                 every product is bit-identical to the scalar __fmul_rn; a - o is computed as a + (-o), which is the
//                 same IEEE operation).  One LDG.128 delivers the same plane of children 0..3 in consecutive registers,
//                 i.e. as two aligned register pairs.
// The reference's clamp of tmax at FLT_MAX is omitted: it can only bind when all three exit parameters overflow,
// which a normalised direction (|inv| <= 2 on some axis) cannot produce over a map bounded by 1e18 -- a packet
// precondition as well.  The exit parameter is kept as x' = min(exit, tU): "hit and still needed" is then entry < x'.
//
// kOnFace       : the packet's origin may lie exactly on a face plane of a box.  Only then can an exit parameter be
//                 exactly 0 (below), where the reference's inside / outside distinction decides; such nodes are redone
//                 with the literal box test.  flag_viewpoints_kernel sorts the viewpoints into the two instantiations.
// kOrderedPush  : the far siblings of the entered child are stacked nearest-on-top (see the push) instead of always in
//                 slot order.  (Rejected: visiting surviving children in the builder's slot order along the node's
//                 order axis, with no reductions at all -- C2 17.7 -> 22.4 ms, C3 24.2 -> 35.5 ms,
//                 profiles/r02_variants_static_order.jsonl: the nearest-first choice is worth far more than it costs.)
// kSmemTop      : the first n_top nodes (the top levels, breadth-first) are read from the CTA's shared-memory copy
//                 instead of L1 / L2 (north_star's "shared-memory staging of the top tree levels"; measured, not the
//                 default: see DESIGN.md section 4).
template <int NR, bool kMixed, bool kPacked, bool kOnFace, bool kSmemTop, bool kOrderedPush>
__device__ __forceinline__ void traverse_packetN(float ox, float oy, float oz, float A, const Dir2 (&r)[NR], Best2 (&b)[NR],
                                                 const float4* __restrict__ nodes, const float4* __restrict__ nodes_v,
                                                 uint32_t octant, const float4* smem_top, uint32_t n_top) {
  float stack_t[NR][kStackEntries];
  uint32_t stack_n[kStackEntries];
  int sp = 0;
  uint32_t ref = 0;
  // plane indices inside a node: mn x,y,z = 0,1,2 ; mx x,y,z = 3,4,5
  const uint32_t nx = (!kMixed && (octant & 1u)) ? 3u : 0u, ny = (!kMixed && (octant & 2u)) ? 4u : 1u,
                 nz = (!kMixed && (octant & 4u)) ? 5u : 2u;
  const uint32_t fx = 3u - nx, fy = 5u - ny, fz = 7u - nz;
  const float kInf = __int_as_float(0x7F800000);
  // kPacked: the per-packet constants as register pairs
  const float2 nox = make_float2(-ox, -ox), noy = make_float2(-oy, -oy), noz = make_float2(-oz, -oz);
  float2 iix[NR], iiy[NR], iiz[NR];
#pragma unroll
  for (int j = 0; j < NR; ++j) {
    iix[j] = make_float2(r[j].ix, r[j].ix);
    iiy[j] = make_float2(r[j].iy, r[j].iy);
    iiz[j] = make_float2(r[j].iz, r[j].iz);
  }
  // kPacked, two rays per lane: the directions of the lane's rays as pairs (leaf distances, below)
  const float2 pox = make_float2(ox, ox), poy = make_float2(oy, oy), poz = make_float2(oz, oz);
  const float2 ddx = make_float2(r[0].dx, r[NR - 1].dx), ddy = make_float2(r[0].dy, r[NR - 1].dy),
               ddz = make_float2(r[0].dz, r[NR - 1].dz);
  while (true) {
    const float4* np = nodes + (size_t)ref * 8u;
    float4 nrx, nry, nrz, frx, fry, frz;
    uint4 ch;
    if (kSmemTop && ref < n_top) {
      const float4* sp4 = smem_top + ref * 8u;
      nrx = sp4[nx]; nry = sp4[ny]; nrz = sp4[nz];
      frx = sp4[fx]; fry = sp4[fy]; frz = sp4[fz];
      ch = *reinterpret_cast<const uint4*>(sp4 + 6);
    } else {
      // plane rows through the lane-opaque base (vector address arithmetic, see the kernel), the child words through
      // the plain one: everything the control flow depends on stays provably warp-uniform
      const float4* nv = nodes_v + (size_t)ref * 8u;
      nrx = __ldg(nv + nx); nry = __ldg(nv + ny); nrz = __ldg(nv + nz);
      frx = __ldg(nv + fx); fry = __ldg(nv + fy); frz = __ldg(nv + fz);
      ch = __ldg(reinterpret_cast<const uint4*>(np + 6));
    }
    const uint32_t chv[4] = {ch.x, ch.y, ch.z, ch.w};

    float t[NR][4];  // entry parameter, [ray][child]
    float x[NR][4];  // min(exit parameter, tU)
    if (!kPacked) {
      const float pnx[4] = {nrx.x, nrx.y, nrx.z, nrx.w}, pny[4] = {nry.x, nry.y, nry.z, nry.w}, pnz[4] = {nrz.x, nrz.y, nrz.z, nrz.w};
      const float pfx[4] = {frx.x, frx.y, frx.z, frx.w}, pfy[4] = {fry.x, fry.y, fry.z, fry.w}, pfz[4] = {frz.x, frz.y, frz.z, frz.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float anx = __fsub_rn(pnx[c], ox), afx = __fsub_rn(pfx[c], ox);
        const float any_ = __fsub_rn(pny[c], oy), afy = __fsub_rn(pfy[c], oy);
        const float anz = __fsub_rn(pnz[c], oz), afz = __fsub_rn(pfz[c], oz);
#pragma unroll
        for (int j = 0; j < NR; ++j) {
          float lx = __fmul_rn(anx, r[j].ix), hx = __fmul_rn(afx, r[j].ix);
          float ly = __fmul_rn(any_, r[j].iy), hy = __fmul_rn(afy, r[j].iy);
          float lz = __fmul_rn(anz, r[j].iz), hz = __fmul_rn(afz, r[j].iz);
          if (kMixed) {
            const float sx = fminf(lx, hx), sy = fminf(ly, hy), sz = fminf(lz, hz);
            hx = fmaxf(lx, hx); hy = fmaxf(ly, hy); hz = fmaxf(lz, hz);
            lx = sx; ly = sy; lz = sz;
          }
          x[j][c] = fminf(fminf(fminf(hx, hy), hz), b[j].tU);
          t[j][c] = fmaxf(fmaxf(fmaxf(lx, ly), lz), 0.0f);
        }
      }
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {  // children 2h, 2h + 1
        const float2 anx = __fadd2_rn(h ? make_float2(nrx.z, nrx.w) : make_float2(nrx.x, nrx.y), nox);
        const float2 afx = __fadd2_rn(h ? make_float2(frx.z, frx.w) : make_float2(frx.x, frx.y), nox);
        const float2 any_ = __fadd2_rn(h ? make_float2(nry.z, nry.w) : make_float2(nry.x, nry.y), noy);
        const float2 afy = __fadd2_rn(h ? make_float2(fry.z, fry.w) : make_float2(fry.x, fry.y), noy);
        const float2 anz = __fadd2_rn(h ? make_float2(nrz.z, nrz.w) : make_float2(nrz.x, nrz.y), noz);
        const float2 afz = __fadd2_rn(h ? make_float2(frz.z, frz.w) : make_float2(frz.x, frz.y), noz);
#pragma unroll
        for (int j = 0; j < NR; ++j) {
          float2 lx = __fmul2_rn(anx, iix[j]), hx = __fmul2_rn(afx, iix[j]);
          float2 ly = __fmul2_rn(any_, iiy[j]), hy = __fmul2_rn(afy, iiy[j]);
          float2 lz = __fmul2_rn(anz, iiz[j]), hz = __fmul2_rn(afz, iiz[j]);
          if (kMixed) {
            const float2 sx = make_float2(fminf(lx.x, hx.x), fminf(lx.y, hx.y));
            const float2 sy = make_float2(fminf(ly.x, hy.x), fminf(ly.y, hy.y));
            const float2 sz = make_float2(fminf(lz.x, hz.x), fminf(lz.y, hz.y));
            hx = make_float2(fmaxf(lx.x, hx.x), fmaxf(lx.y, hx.y));
            hy = make_float2(fmaxf(ly.x, hy.x), fmaxf(ly.y, hy.y));
            hz = make_float2(fmaxf(lz.x, hz.x), fmaxf(lz.y, hz.y));
            lx = sx; ly = sy; lz = sz;
          }
          x[j][2 * h] = fminf(fminf(fminf(hx.x, hy.x), hz.x), b[j].tU);
          t[j][2 * h] = fmaxf(fmaxf(fmaxf(lx.x, ly.x), lz.x), 0.0f);
          x[j][2 * h + 1] = fminf(fminf(fminf(hx.y, hy.y), hz.y), b[j].tU);
          t[j][2 * h + 1] = fmaxf(fmaxf(fmaxf(lx.y, ly.y), lz.y), 0.0f);
        }
      }
    }
    // exit parameter == 0 for some child of some ray (tU > 0, so x' == 0 says exactly that): the ray's origin lies
    // on a face plane of the box, where the reference's inside / outside distinction decides -- redo this node with
    // the reference's full box test and express its verdict in the same form.  An exit parameter is a product
    // (plane - origin) * inv with |inv| >= 1 - 2^-22 (normalised direction), so it is 0 only if plane == origin: a
    // viewpoint whose centre shares no coordinate with any box plane of the map never gets here (kOnFace = false).
    if (kOnFace) {
      float mm = kInf;
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        mm = fminf(mm, fminf(fminf(fabsf(x[j][0]), fabsf(x[j][1])), fminf(fabsf(x[j][2]), fabsf(x[j][3]))));
      }
      if (__any_sync(Q3DR_FULL_MASK, mm == 0.0f)) {
        const float4 mnx = np[0], mny = np[1], mnz = np[2], mxx = np[3], mxy = np[4], mxz = np[5];
        Ray q;
        float dd;
        q.ox = ox; q.oy = oy; q.oz = oz;
#pragma unroll
        for (int j = 0; j < NR; ++j) {
          q.dx = r[j].dx; q.dy = r[j].dy; q.dz = r[j].dz; q.ix = r[j].ix; q.iy = r[j].iy; q.iz = r[j].iz;
          x[j][0] = box_test(q, mnx.x, mny.x, mnz.x, mxx.x, mxy.x, mxz.x, dd, t[j][0]) ? b[j].tU : -1.0f;
          x[j][1] = box_test(q, mnx.y, mny.y, mnz.y, mxx.y, mxy.y, mxz.y, dd, t[j][1]) ? b[j].tU : -1.0f;
          x[j][2] = box_test(q, mnx.z, mny.z, mnz.z, mxx.z, mxy.z, mxz.z, dd, t[j][2]) ? b[j].tU : -1.0f;
          x[j][3] = box_test(q, mnx.w, mny.w, mnz.w, mxx.w, mxy.w, mxz.w, dd, t[j][3]) ? b[j].tU : -1.0f;
        }
      }
    }

    // "some ray still needs child c"
    // (four votes, not one warp OR-reduction of per-lane masks: REDUX measured 0.2 % slower on C2 and C3, same box)
    uint32_t any = 0u;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      bool need = false;
#pragma unroll
      for (int j = 0; j < NR; ++j) need = need || (t[j][c] < x[j][c]);
      any |= __any_sync(Q3DR_FULL_MASK, need) ? (1u << c) : 0u;
    }
    const uint32_t leaf_mask = (ch.x >> kLeafMaskShift) & 0xFu;

    // leaves first: they tighten the bounds before the inner children are judged.  A ray accepts a leaf on the
    // reference's own terms: the slab test passes and dist_sq < best, or == best with a higher leaf index
    // (x' > t is implied for every ray that can pass the dist_sq comparison, see bound_after).
    const uint32_t todo = any & leaf_mask;
    if (todo) {
      bool tightened = false;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (todo & (1u << c)) {
          const int32_t li = (int32_t)(chv[c] & kPayloadMask);
          // branch-free per lane: every lane evaluates dist_sq (lanes that cannot accept compute a value nobody reads),
          // the accept decision is a predicate, the update a select -- no divergent path, nothing recomputed
          float d[NR];
          bool acc[NR];
          bool some = false;
          if (kPacked && NR == 2) {
            // the two rays of the lane side by side: the products and the differences with the packed instructions,
            // the sums that consume a product with scalar adds -- the same nine roundings per ray as dist_sq_at2 in
            // the same order.  (ptxas 12.9 contracts mul.rn.f32x2 feeding add.rn.f32x2 into FFMA2 even under
            // --fmad false, one rounding instead of two; a packed product feeding a scalar add.rn is left alone, and so
            // is every add -> mul chain, which is all the slab test above has.  o - p is computed as o + (-p), the
            // same IEEE operation.)
            const float2 tt = make_float2(t[0][c], t[1][c]);
            const float2 qx = __fmul2_rn(ddx, tt), qy = __fmul2_rn(ddy, tt), qz = __fmul2_rn(ddz, tt);
            const float2 ex = __fadd2_rn(pox, make_float2(-__fadd_rn(ox, qx.x), -__fadd_rn(ox, qx.y)));
            const float2 ey = __fadd2_rn(poy, make_float2(-__fadd_rn(oy, qy.x), -__fadd_rn(oy, qy.y)));
            const float2 ez = __fadd2_rn(poz, make_float2(-__fadd_rn(oz, qz.x), -__fadd_rn(oz, qz.y)));
            const float2 sx = __fmul2_rn(ex, ex), sy = __fmul2_rn(ey, ey), sz = __fmul2_rn(ez, ez);
            d[0] = __fadd_rn(sx.x, __fadd_rn(sy.x, sz.x));
            d[NR - 1] = __fadd_rn(sx.y, __fadd_rn(sy.y, sz.y));
          } else {
#pragma unroll
            for (int j = 0; j < NR; ++j) d[j] = dist_sq_at2(ox, oy, oz, r[j], t[j][c]);
          }
#pragma unroll
          for (int j = 0; j < NR; ++j) {
            acc[j] = (x[j][c] > t[j][c]) & ((d[j] < b[j].dsq) | ((d[j] == b[j].dsq) & (li > b[j].leaf)));
            some = some | acc[j];
          }
          const bool accepted = __any_sync(Q3DR_FULL_MASK, some);  // a vote result: provably uniform
          if (accepted) {
#pragma unroll
            for (int j = 0; j < NR; ++j) {
              const float u = bound_after(t[j][c], A);
              if (acc[j]) {
                b[j].dsq = d[j];
                b[j].t = t[j][c];
                b[j].leaf = li;
                b[j].tU = u;
              }
            }
          }
          tightened = tightened || accepted;
        }
      }
      if (tightened && (any & ~leaf_mask)) {
        any = 0u;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          bool need = false;
#pragma unroll
          for (int j = 0; j < NR; ++j) {
            x[j][c] = fminf(x[j][c], b[j].tU);
            need = need || (t[j][c] < x[j][c]);
          }
          any |= __any_sync(Q3DR_FULL_MASK, need) ? (1u << c) : 0u;
        }
      }
    }

    // inner children
    uint32_t inner = any & ~leaf_mask;
    if (inner) {
      uint32_t near_c = 0u;
      uint32_t near_w = (inner & 1u) ? ch.x : (inner & 2u) ? ch.y : (inner & 4u) ? ch.z : ch.w;  // the only one, if so
      if (inner & (inner - 1u)) {
        // more than one: the nearest is entered now (warp minimum of the entry parameters of the rays that need the
        // child); the others go on the stack with every ray's own entry parameter (+inf for rays that do not need the
        // child).  What order far siblings are popped in changes how early the bounds tighten, never the result.
        const uint32_t kNone = 0xFFFFFFFFu;
        uint32_t key[4];
        uint32_t best_key = kNone;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          key[c] = kNone;
          if (inner & (1u << c)) {
            uint32_t k = kNone;
#pragma unroll
            for (int j = 0; j < NR; ++j) k = min(k, (t[j][c] < x[j][c]) ? __float_as_uint(t[j][c]) : kNone);
            k = __reduce_min_sync(Q3DR_FULL_MASK, k);
            key[c] = k;
            if (k < best_key) {
              best_key = k;
              near_c = (uint32_t)c;
              near_w = chv[c];
            }
          }
        }
#define Q3DR_PUSH_CHILD(C)                                                                                  \
  if ((rest & (1u << (C))) != 0u) {                                                                         \
    _Pragma("unroll") for (int j = 0; j < NR; ++j) {                                                        \
      stack_t[j][sp] = (t[j][C] < x[j][C]) ? t[j][C] : kInf; /* tU only shrinks: an entry >= tU stays >= tU */ \
    }                                                                                                       \
    stack_n[sp] = chv[C] & kPayloadMask;                                                                    \
    ++sp;                                                                                                   \
  }
        const uint32_t rest = inner & ~(1u << near_c);
        if (kOrderedPush) {
          // static child indices, no sorting network: the siblings go on the stack in slot order or in reverse slot
          // order, whichever leaves the nearer of the two outermost ones on top (exact for two siblings)
          const uint32_t lo_c = (uint32_t)__ffs((int)rest) - 1u, hi_c = 31u - (uint32_t)__clz((int)rest);
          const uint32_t k_lo = (lo_c == 0u) ? key[0] : (lo_c == 1u) ? key[1] : (lo_c == 2u) ? key[2] : key[3];
          const uint32_t k_hi = (hi_c == 0u) ? key[0] : (hi_c == 1u) ? key[1] : (hi_c == 2u) ? key[2] : key[3];
          if (k_lo <= k_hi) {
            Q3DR_PUSH_CHILD(3) Q3DR_PUSH_CHILD(2) Q3DR_PUSH_CHILD(1) Q3DR_PUSH_CHILD(0)
          } else {
            Q3DR_PUSH_CHILD(0) Q3DR_PUSH_CHILD(1) Q3DR_PUSH_CHILD(2) Q3DR_PUSH_CHILD(3)
          }
        } else {
          Q3DR_PUSH_CHILD(3) Q3DR_PUSH_CHILD(2) Q3DR_PUSH_CHILD(1) Q3DR_PUSH_CHILD(0)
        }
#undef Q3DR_PUSH_CHILD
      }
      ref = near_w & kPayloadMask;
      continue;
    }
    bool found = false;
    while (sp > 0) {
      --sp;
      bool need = false;
#pragma unroll
      for (int j = 0; j < NR; ++j) need = need || (stack_t[j][sp] < b[j].tU);
      // the stack lives in per-thread local memory: say explicitly that every lane holds the same node reference, so
      // that the branches that depend on node words are compiled as uniform branches (no convergence barriers)
      const uint32_t n = __shfl_sync(Q3DR_FULL_MASK, stack_n[sp], 0);
      if (__any_sync(Q3DR_FULL_MASK, need)) {
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
  r.ix = __frcp_rn(r.dx);  // the correctly rounded reciprocal IS 1.0f / x in binary32 (bvh.h:380 cwiseInverse)
  r.iy = __frcp_rn(r.dy);
  r.iz = __frcp_rn(r.dz);
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

// Packet geometry: a tile is TW pixels wide and (32 / TW) * NR pixels tall; lane l carries column l % TW and the NR
// rows l / TW + (32 / TW) * j, j = 0 .. NR-1.  Within one j the lanes are in row-major pixel order.
template <int NR, int TW>
struct PacketShape {
  static constexpr int kRowsPerSlab = 32 / TW;
  static constexpr int kTileH = kRowsPerSlab * NR;
};

// MODE: kModeScore or kModePixels (camera windows only; ray lists use raycast_rays_kernel).
// NR rays per lane.
// kOnFace = false: viewpoints flagged in p.vp_flags are skipped (the kOnFace = true launch casts them, from p.vp_list).
// kMinCtas: resident CTAs per SM the register allocation is tuned for (4 x 64, 3 x 80 or 2 x 128 registers).
// kWarps: warps per CTA (warps never synchronise with each other; the CTA size only sets the register granularity:
// 7 CTAs of 4 warps leave 72 registers per thread, between the 64 of 4 x 8 warps and the 80 of 3 x 8 warps).
template <int MODE, int NR, int TW, bool kPacked, bool kOnFace, int kMinCtas, int kWarps, bool kSmemTop, bool kOrderedPush = false>
__global__ void __launch_bounds__(kWarps * 32, kMinCtas) raycast2_kernel(const RaycastParams p) {
  extern __shared__ float4 smem_top[];  // kSmemTop: p.n_top nodes of 8 float4 (launch with that much dynamic smem)
  if (kSmemTop) {
    for (uint32_t i = threadIdx.x; i < p.n_top * 8u; i += blockDim.x) smem_top[i] = p.nodes[i];
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  // The node array through a base the compiler cannot prove warp-uniform (a lane-dependent zero is added): node
  // references are uniform by construction, and with a provably uniform base the six row addresses of every node visit
  // would be formed in the uniform datapath and copied into vector registers (four instructions per load instead of two).
  uint32_t lane_zero;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane_zero));
  lane_zero = __umulhi(lane_zero, 0x04000000u);  // laneid < 32: always 0
  const float4* const nodes_v = p.nodes + lane_zero;
  const float range_sq = (p.max_range > 0.0f) ? __fmul_rn(p.max_range, p.max_range) : FLT_MAX;
  const float kInf = __int_as_float(0x7F800000);
  constexpr int kRows = PacketShape<NR, TW>::kRowsPerSlab;
  constexpr int kTileH = PacketShape<NR, TW>::kTileH;

  // the checking instantiation of a scoring launch works through the flagged viewpoints only (usually none)
  const bool listed = kOnFace && (p.vp_list != nullptr);
  const uint32_t total_tiles = listed ? __ldg(p.n_flagged) * p.tiles_per_vp : p.total_tiles;
  uint32_t* counter = p.tile_counter + (listed ? 1 : 0);

  // A warp takes RUNS of p.tile_run consecutive tiles (neighbouring packets of one image row): what the previous packet
  // pulled into this SM's L1 -- the bottom of the tree around its pixels -- is what the next one walks first, and the
  // tile counter sees one atomic per run.
  uint32_t run_next = 0, run_end = 0;
  while (true) {
    if (run_next == run_end) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(counter, p.tile_run);
      run_next = __shfl_sync(Q3DR_FULL_MASK, base, 0);
      run_end = run_next + p.tile_run;
    }
    const uint32_t tile = run_next++;
    if (tile >= total_tiles) break;

    // tile -> (viewpoint slot, tile row, tile column): quotients through the reciprocals the host prepared (a double
    // holds any 32-bit integer exactly; the truncated product is the quotient or one below it) instead of two integer
    // divisions by run-time divisors per packet
    uint32_t slot = (uint32_t)__double2uint_rz((double)tile * p.inv_tiles_per_vp);
    if (slot * p.tiles_per_vp > tile) --slot;  // (only a divisor beyond 2^20 could make the estimate overshoot)
    uint32_t rem = tile - slot * p.tiles_per_vp;
    if (rem >= p.tiles_per_vp) {
      rem -= p.tiles_per_vp;
      ++slot;
    }
    const uint32_t vp = listed ? __ldg(p.vp_list + slot) : slot;
    if (!kOnFace && p.vp_flags != nullptr && __ldg(p.vp_flags + vp) != 0) continue;
    uint32_t ty = (uint32_t)__double2uint_rz((double)rem * p.inv_tiles_x);
    if (ty * p.tiles_x > rem) --ty;
    uint32_t tx = rem - ty * p.tiles_x;
    if (tx >= p.tiles_x) {
      tx -= p.tiles_x;
      ++ty;
    }
    const uint32_t px = tx * TW + (lane % TW);
    uint32_t py[NR];
    bool live[NR];
    bool all_live = true;
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      py[j] = ty * kTileH + (lane / TW) + kRows * j;
      live[j] = (px < p.w) && (py[j] < p.h);
      all_live = all_live && live[j];
    }
    const uint32_t live_mask = __ballot_sync(Q3DR_FULL_MASK, live[0]);
    if (live_mask == 0u) continue;  // live[j] implies live[0]

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
    Dir2 r[NR];
    bool unit[NR];
#pragma unroll
    for (int j = 0; j < NR; ++j) unit[j] = camera_dir2(R, c0, p.fy, p.cy, p.y0 + py[j], r[j]);
    {
      // a dead ray rides along with a copy of a live one and never accepts anything (dist_sq >= 0 > -1, t >= 0 > -1)
      const int src = __ffs(live_mask) - 1;
      if (!__all_sync(Q3DR_FULL_MASK, all_live)) {
        const Dir2 s = r[0];
#pragma unroll
        for (int j = 0; j < NR; ++j) copy_dir_from_lane(r[j], s, src, !live[j]);
      }
    }
    Best2 b[NR];
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      b[j].leaf = -1;
      b[j].t = 0.0f;
      b[j].dsq = live[j] ? range_sq : -1.0f;
    }

    // preconditions of the fast loops (see bound_after / traverse_packetN); dead rays are copies of live ones
    const float M = fmaxf(fmaxf(fabsf(ox), fabsf(oy)), fabsf(oz));
    bool tame = p.bounded_map && (M <= 1e18f) && !(p.max_range > 1e18f);
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      tame = tame && (fabsf(r[j].ix) <= FLT_MAX) && (fabsf(r[j].iy) <= FLT_MAX) && (fabsf(r[j].iz) <= FLT_MAX) &&
             (unit[j] || !live[j]);
    }
    if (__all_sync(Q3DR_FULL_MASK, tame)) {
      const float A = fmaf(M, 1.0f / 1048576.0f, 1.0f / 8192.0f);
      const float u = (p.max_range > 0.0f) ? bound_after(p.max_range, A) : kInf;
      // per axis: all rays positive, all negative, or mixed
      bool nxa = false, nya = false, nza = false, pxa = false, pya = false, pza = false;
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        b[j].tU = live[j] ? u : -1.0f;
        nxa = nxa || (r[j].ix < 0.0f); pxa = pxa || !(r[j].ix < 0.0f);
        nya = nya || (r[j].iy < 0.0f); pya = pya || !(r[j].iy < 0.0f);
        nza = nza || (r[j].iz < 0.0f); pza = pza || !(r[j].iz < 0.0f);
      }
      const bool anx = __any_sync(Q3DR_FULL_MASK, nxa), apx = __any_sync(Q3DR_FULL_MASK, pxa);
      const bool any_ = __any_sync(Q3DR_FULL_MASK, nya), apy = __any_sync(Q3DR_FULL_MASK, pya);
      const bool anz = __any_sync(Q3DR_FULL_MASK, nza), apz = __any_sync(Q3DR_FULL_MASK, pza);
      if (!(anx && apx) && !(any_ && apy) && !(anz && apz)) {
        const uint32_t octant = (anx ? 1u : 0u) | (any_ ? 2u : 0u) | (anz ? 4u : 0u);
        traverse_packetN<NR, false, kPacked, kOnFace, kSmemTop, kOrderedPush>(ox, oy, oz, A, r, b, p.nodes, nodes_v, octant, smem_top, p.n_top);
      } else {
        traverse_packetN<NR, true, kPacked, kOnFace, kSmemTop, kOrderedPush>(ox, oy, oz, A, r, b, p.nodes, nodes_v, 0u, smem_top, p.n_top);
      }
    } else {
      // some ray is exactly parallel to an axis, or the packet is outside the fast loops' proven domain (huge
      // coordinates, degenerate pose): the literal reference box test, one ray after the other
      Ray q;
      PacketHit h;
      q.ox = ox; q.oy = oy; q.oz = oz;
#pragma unroll  // static indices: r[] and b[] stay in registers (the callee is not inlined, so this costs NR calls)
      for (int j = 0; j < NR; ++j) {
        q.dx = r[j].dx; q.dy = r[j].dy; q.dz = r[j].dz; q.ix = r[j].ix; q.iy = r[j].iy; q.iz = r[j].iz;
        h.dsq = b[j].dsq; h.t = 0.0f; h.tU = 0.0f; h.leaf = -1;
        traverse_packet_exact(q, h, p.nodes, p.nodes, 0u);
        b[j].dsq = h.dsq; b[j].t = h.t; b[j].leaf = h.leaf;
      }
    }

    if (MODE == kModeScore) {
      // warp-aggregated dedup: within each slab j of the tile lanes are in row-major pixel order, so the lowest lane
      // of each equal-leaf group owns the group's first pixel; only group leaders touch the tables (atomicMin
      // merges the slabs and the other tiles of the viewpoint).
      uint32_t* fpv = p.first_pix + (size_t)vp * p.leaf_stride;
      uint32_t* bmv = p.bitmap + (size_t)vp * p.word_stride;
      // the pixel coordinates are formed again from the lane index (read again, so that the compiler cannot keep the
      // earlier values alive): three registers fewer across the traversal loop, which runs at the 64-register limit
      uint32_t lane2;
      asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane2));
      const uint32_t px2 = tx * TW + (lane2 % TW);
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        const int32_t leaf = b[j].leaf;
        const uint32_t pix = (ty * kTileH + (lane2 / TW) + kRows * j) * p.w + px2;
        const uint32_t grp = __match_any_sync(Q3DR_FULL_MASK, leaf);
        const bool leader = (leaf >= 0) && ((__ffs(grp) - 1) == (int)lane2);
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
      for (int j = 0; j < NR; ++j) {
        if (!live[j]) continue;
        const bool hit = b[j].leaf >= 0;
        // intersection = origin + direction * t  (geometry.h:345-348); t = 0 when the origin is inside
        const float hx = hit ? __fadd_rn(ox, __fmul_rn(r[j].dx, b[j].t)) : 0.0f;
        const float hy = hit ? __fadd_rn(oy, __fmul_rn(r[j].dy, b[j].t)) : 0.0f;
        const float hz = hit ? __fadd_rn(oz, __fmul_rn(r[j].dz, b[j].t)) : 0.0f;
        const uint32_t depth = hit ? (uint32_t)__ldg(p.leaf_depth + b[j].leaf) : 0u;
        const size_t oi = (size_t)vp * p.w * p.h + (size_t)py[j] * p.w + px;
        float4* dst = reinterpret_cast<float4*>(p.pixels + oi);
        dst[0] = make_float4(hx, hy, hz, __int_as_float(b[j].leaf));
        dst[1] = make_float4(__uint_as_float(depth), hit ? b[j].dsq : 0.0f, (float)(p.x0 + px), (float)(p.y0 + py[j]));
      }
    }
  }
}

// Per viewpoint of a chunk: does the camera centre share a coordinate with a face plane of some box of the map?
// planes[a] = the sorted distinct min / max coordinates of all leaf boxes on axis a (inner boxes are unions of leaf
// boxes, so they add no planes).  Flagged viewpoints are listed for the checking instantiation of the raycast kernel.
static __global__ void flag_viewpoints_kernel(const float* __restrict__ rt, uint32_t n_vp, const float* __restrict__ planes_x,
                                              const float* __restrict__ planes_y, const float* __restrict__ planes_z,
                                              uint32_t nx, uint32_t ny, uint32_t nz, uint8_t* __restrict__ flags,
                                              uint32_t* __restrict__ list, uint32_t* __restrict__ count) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vp) return;
  const float o[3] = {rt[12 * (size_t)v + 3], rt[12 * (size_t)v + 7], rt[12 * (size_t)v + 11]};
  const float* pl[3] = {planes_x, planes_y, planes_z};
  const uint32_t np[3] = {nx, ny, nz};
  bool on = false;
  for (int a = 0; a < 3; ++a) {
    uint32_t lo = 0, hi = np[a];  // first plane >= o[a]
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (pl[a][mid] < o[a]) {
        lo = mid + 1;
      } else {
        hi = mid;
      }
    }
    on = on || (lo < np[a] && pl[a][lo] == o[a]);
    on = on || !(o[a] == o[a]);  // NaN origin: let the checking path have it
  }
  flags[v] = on ? 1 : 0;
  if (on) list[atomicAdd(count, 1u)] = v;
}

}  // namespace q3dr
