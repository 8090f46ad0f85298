// Per-point ray -> height-plane geometry and voxel-index quantisation, bit-exact with the
// reference's torch evaluation.  Restates (in fp32, with every rounding pinned by intrinsics so
// that nvcc can neither contract nor reassociate):
//   LSSFPN.get_geometry        layers/backbones/lss_fpn.py:372-401
//   LSSFPN.height2localtion    layers/backbones/lss_fpn.py:350-370
//   quantisation               layers/backbones/lss_fpn.py:487-488   ((g - lower) / size).int()
// (identical code in layers/backbones/bsm_lss_fpn.py:409-460,552-553).
//
// Work that is invariant along the height-bin axis is hoisted per pixel WITHOUT changing any
// rounding: the first two terms of every row of ida^-1 @ (u, v, z_d, 1) do not depend on d, and
// the virtual-camera ray  pv = Mv @ (10x, 10y, 10, w)  is re-used for consecutive bins whenever
// its inputs are bitwise unchanged (always, for the reference's z-preserving IDA matrices).
#pragma once

#include <string.h>

#include "common.cuh"

namespace sgv3d {
namespace geom {

// Per-camera operands, staged in shared memory by the kernels (52 floats + bda).
struct Camera {
  float A[16];   // ida_mat.inverse()                         lss_fpn.py:392
  float Mv[16];  // sensor2virtual @ inverse(intrin)          lss_fpn.py:361
  float Me[16];  // sensor2ego @ inverse(sensor2virtual)      lss_fpn.py:367
  float Bd[16];  // bda_mat (only read when has_bda)          lss_fpn.py:394-398
  float ref_h;   // reference_heights[b, n]                   lss_fpn.py:352-354
  int has_bda;
  int bda_fast;  // bda is bitwise the identity and row 3 of Me is small: bda @ p == p for finite p
};

struct Grid {
  float lower[3];  // fp32(voxel_coord - voxel_size / 2.0)
  float size[3];   // voxel_size
  int X, Y, Z;
  // Exact z-range test without a division: for t = fp32(g.z - lower.z),
  //   0 <= trunc(RN(t / size.z)) < Z   <=>   !(t <= zt_lo) && !(t >= zt_hi)
  // zt_lo / zt_hi are found on the host by bisection over float bit patterns with the same IEEE
  // division (RN(t/s) is monotone in t).  NaN passes both tests, as cvt.rzi(NaN) = 0 is in range.
  float zt_lo, zt_hi;
  float rcp_size[2];  // RN(1 / size.x), RN(1 / size.y) for the guarded fast quantisation
};

template <int ARITH>
__device__ __forceinline__ float dot4(const float *m, float b0, float b1, float b2, float b3) {
  if (ARITH == SGV3D_ARITH_PAIR) {
    const float lo = __fmaf_rn(m[1], b1, __fmul_rn(m[0], b0));
    const float hi = __fmaf_rn(m[3], b3, __fmul_rn(m[2], b2));
    return __fadd_rn(lo, hi);
  }
  float acc = __fmul_rn(m[0], b0);
  if (ARITH == SGV3D_ARITH_FMA) {
    acc = __fmaf_rn(m[1], b1, acc);
    acc = __fmaf_rn(m[2], b2, acc);
    acc = __fmaf_rn(m[3], b3, acc);
  } else {
    acc = __fadd_rn(acc, __fmul_rn(m[1], b1));
    acc = __fadd_rn(acc, __fmul_rn(m[2], b2));
    acc = __fadd_rn(acc, __fmul_rn(m[3], b3));
  }
  return acc;
}

// first two terms of a row: m0*b0 (+) m1*b1  (identical sub-expression in all three orders' heads)
template <int ARITH>
__device__ __forceinline__ float dot2_head(const float *m, float b0, float b1) {
  float acc = __fmul_rn(m[0], b0);
  if (ARITH == SGV3D_ARITH_SEQ) return __fadd_rn(acc, __fmul_rn(m[1], b1));
  return __fmaf_rn(m[1], b1, acc);
}
// remaining two terms: (+) m2*b2 (+) m3*b3
template <int ARITH>
__device__ __forceinline__ float dot2_tail(float acc, const float *m, float b2, float b3) {
  if (ARITH == SGV3D_ARITH_PAIR) return __fadd_rn(acc, __fmaf_rn(m[3], b3, __fmul_rn(m[2], b2)));
  if (ARITH == SGV3D_ARITH_FMA) {
    acc = __fmaf_rn(m[2], b2, acc);
    return __fmaf_rn(m[3], b3, acc);
  }
  acc = __fadd_rn(acc, __fmul_rn(m[2], b2));
  return __fadd_rn(acc, __fmul_rn(m[3], b3));
}

// State carried by one thread while it walks the height bins of one pixel.
template <int ARITH>
struct PixelRay {
  float head[4];        // d-invariant partial sums of the four rows of  A @ (u, v, z, 1)
  float q0, q1, q3;     // inputs of the cached virtual-camera ray
  float pv0, pv1, pv2;  // cached  Mv @ (q0, q1, 10, q3)  (row 3 is overwritten with 1 downstream)
  bool have_pv;

  __device__ __forceinline__ void init(const Camera &cam, float u, float v) {
#pragma unroll
    for (int r = 0; r < 4; ++r) head[r] = dot2_head<ARITH>(cam.A + 4 * r, u, v);
    have_pv = false;
  }

  // ego-frame point for height bin value z (lss_fpn.py:392 -> :398)
  __device__ __forceinline__ void point(const Camera &cam, float z, float &gx, float &gy,
                                        float &gz) {
    // :392  p0 = ida^-1 @ (u, v, z, 1)
    const float p0x = dot2_tail<ARITH>(head[0], cam.A + 0, z, 1.0f);
    const float p0y = dot2_tail<ARITH>(head[1], cam.A + 4, z, 1.0f);
    const float p0z = dot2_tail<ARITH>(head[2], cam.A + 8, z, 1.0f);
    const float p0w = dot2_tail<ARITH>(head[3], cam.A + 12, z, 1.0f);
    // :354  height = -1 * p0.z + reference_height
    const float hgt = __fadd_rn(__fmul_rn(-1.0f, p0z), cam.ref_h);
    // :356-360  ray through the pixel at virtual depth 10
    const float n0 = __fmul_rn(p0x, 10.0f), n1 = __fmul_rn(p0y, 10.0f), n3 = p0w;
    // :361-362  pv = (sensor2virtual @ K^-1) @ ray ; same bits in => same bits out, so re-use
    if (!have_pv || __float_as_uint(n0) != __float_as_uint(q0) ||
        __float_as_uint(n1) != __float_as_uint(q1) || __float_as_uint(n3) != __float_as_uint(q3)) {
      q0 = n0; q1 = n1; q3 = n3;
      pv0 = dot4<ARITH>(cam.Mv + 0, n0, n1, 10.0f, n3);
      pv1 = dot4<ARITH>(cam.Mv + 4, n0, n1, 10.0f, n3);
      pv2 = dot4<ARITH>(cam.Mv + 8, n0, n1, 10.0f, n3);
      have_pv = true;
    }
    // :363-366  ratio = height / pv.y ; pe = pv * ratio ; pe.w = 1
    const float ratio = __fdiv_rn(hgt, pv1);
    const float e0 = __fmul_rn(pv0, ratio), e1 = __fmul_rn(pv1, ratio), e2 = __fmul_rn(pv2, ratio);
    // :367-369  pg = (sensor2ego @ sensor2virtual^-1) @ pe
    gx = dot4<ARITH>(cam.Me + 0, e0, e1, e2, 1.0f);
    gy = dot4<ARITH>(cam.Me + 4, e0, e1, e2, 1.0f);
    gz = dot4<ARITH>(cam.Me + 8, e0, e1, e2, 1.0f);
    // :394-398  pg = bda @ pg
    if (cam.has_bda) {
      const float gw = dot4<ARITH>(cam.Me + 12, e0, e1, e2, 1.0f);
      const float bx = dot4<ARITH>(cam.Bd + 0, gx, gy, gz, gw);
      const float by = dot4<ARITH>(cam.Bd + 4, gx, gy, gz, gw);
      const float bz = dot4<ARITH>(cam.Bd + 8, gx, gy, gz, gw);
      gx = bx; gy = by; gz = bz;
    }
  }

  // Voxel id (y*X + x, or -1 if dropped) of the point at height bin value z: point() + quantise +
  // range test (lss_fpn.py:487-488, voxel_pooling_forward_cuda.cu:24) with two exact shortcuts that
  // only the index path may take: the z test by thresholds, and skipping an identity bda.
  __device__ __forceinline__ int voxel(const Camera &cam, const Grid &g, float z) {
    const float p0x = dot2_tail<ARITH>(head[0], cam.A + 0, z, 1.0f);
    const float p0y = dot2_tail<ARITH>(head[1], cam.A + 4, z, 1.0f);
    const float p0z = dot2_tail<ARITH>(head[2], cam.A + 8, z, 1.0f);
    const float p0w = dot2_tail<ARITH>(head[3], cam.A + 12, z, 1.0f);
    const float hgt = __fadd_rn(__fmul_rn(-1.0f, p0z), cam.ref_h);
    const float n0 = __fmul_rn(p0x, 10.0f), n1 = __fmul_rn(p0y, 10.0f), n3 = p0w;
    if (!have_pv || __float_as_uint(n0) != __float_as_uint(q0) ||
        __float_as_uint(n1) != __float_as_uint(q1) || __float_as_uint(n3) != __float_as_uint(q3)) {
      q0 = n0; q1 = n1; q3 = n3;
      pv0 = dot4<ARITH>(cam.Mv + 0, n0, n1, 10.0f, n3);
      pv1 = dot4<ARITH>(cam.Mv + 4, n0, n1, 10.0f, n3);
      pv2 = dot4<ARITH>(cam.Mv + 8, n0, n1, 10.0f, n3);
      have_pv = true;
    }
    const float ratio = __fdiv_rn(hgt, pv1);
    const float e0 = __fmul_rn(pv0, ratio), e1 = __fmul_rn(pv1, ratio), e2 = __fmul_rn(pv2, ratio);
    float gx = dot4<ARITH>(cam.Me + 0, e0, e1, e2, 1.0f);
    float gy = dot4<ARITH>(cam.Me + 4, e0, e1, e2, 1.0f);
    float gz = dot4<ARITH>(cam.Me + 8, e0, e1, e2, 1.0f);
    if (cam.has_bda) {
      // identity bda: 1*x + 0*y + 0*z + 0*w == x (up to the sign of a zero, which cannot change an
      // index) as long as every operand is finite and w cannot overflow; otherwise evaluate it.
      const float big = 1e15f;
      const bool tame = fabsf(e0) < big && fabsf(e1) < big && fabsf(e2) < big && fabsf(gx) < big &&
                        fabsf(gy) < big && fabsf(gz) < big;
      if (!(cam.bda_fast && tame)) {
        const float gw = dot4<ARITH>(cam.Me + 12, e0, e1, e2, 1.0f);
        const float bx = dot4<ARITH>(cam.Bd + 0, gx, gy, gz, gw);
        const float by = dot4<ARITH>(cam.Bd + 4, gx, gy, gz, gw);
        const float bz = dot4<ARITH>(cam.Bd + 8, gx, gy, gz, gw);
        gx = bx; gy = by; gz = bz;
      }
    }
    const float tz = __fsub_rn(gz, g.lower[2]);
    if (tz <= g.zt_lo || tz >= g.zt_hi) return -1;
    const int ix = quantize_guarded(gx, g.lower[0], g.size[0], g.rcp_size[0]);
    if ((unsigned)ix >= (unsigned)g.X) return -1;
    const int iy = quantize_guarded(gy, g.lower[1], g.size[1], g.rcp_size[1]);
    if ((unsigned)iy >= (unsigned)g.Y) return -1;
    return iy * g.X + ix;
  }

  // trunc(RN((g - lower) / size)) without the IEEE division in the common case.
  // q = RN(t * RN(1/size)) differs from the true quotient Q by at most |Q| * (2^-23 + 2^-48), and the
  // reference value RN(Q) by at most |Q| * 2^-24.  If no integer lies within |q| * 2^-22 of q, both fall
  // strictly between the same two consecutive integers and truncate identically; otherwise (and for
  // zero / non-finite / integer-valued q) the exact division decides.
  static __device__ __forceinline__ int quantize_guarded(float gcoord, float lower, float size, float rcp) {
    const float t = __fsub_rn(gcoord, lower);
    const float q = __fmul_rn(t, rcp);
    const float k = rintf(q);
    if (fabsf(__fsub_rn(q, k)) > __fmul_rn(fabsf(q), 2.384185791015625e-07f))  // 2^-22
      return __float2int_rz(q);
    return __float2int_rz(__fdiv_rn(t, size));
  }
};

// ---------------------------------------------------------------------------------------------
// Fast index path for the calibration structure the reference's data pipeline produces: an IDA matrix
// that leaves z alone (rows 0, 1, 3 of ida^-1 have a zero in column 2: resize / crop / flip / rotate,
// dataset/nusc_mv_det_dataset.py:133-161) and a BDA that is absent or the exact identity.  Under
// those conditions -- all checked on the device, per camera and per pixel, before the path is taken --
// the values below are BITWISE what the general path computes:
//   * rows 0, 1, 3 of p0 = A @ (u, v, z, 1) do not depend on z.  Row r is  head_r (+) A_r2*z (+) A_r3
//     in one of the three evaluation orders; with A_r2 = +-0 and z finite the product is +-0, and
//     adding +-0 changes nothing unless the other addend is -0.  Required: A_r2 == 0, A_r3 is not -0
//     (PAIR order adds it to the product first), head_r is not -0 (SEQ / FMA orders).
//     => the virtual-camera ray pv = Mv @ (10*p0x, 10*p0y, 10, p0w) is computed once per pixel.
//   * identity BDA: 1*x + 0*y + 0*z + 0*w == x up to the sign of a zero (which cannot change an index)
//     when x, y, z, w are finite.  Required: row 3 of Me is exactly (0, 0, 0, 1), so w == 1 whenever
//     e0..e2 are finite, and e_i non-finite implies gx non-finite; the finiteness of gx, gy, gz is
//     tested per point, and a non-finite point sends the whole pixel chunk to the general kernel.
// Row 2 (the height) stays fully general.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_neg_zero(float x) { return __float_as_uint(x) == 0x80000000u; }

// per-camera qualification (one thread)
__device__ __forceinline__ bool camera_is_fast(const Camera &c) {
  bool ok = true;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    if (r == 2) continue;
    ok = ok && (c.A[4 * r + 2] == 0.0f) && !is_neg_zero(c.A[4 * r + 3]);
  }
  if (c.has_bda)
    ok = ok && c.bda_fast && c.Me[12] == 0.0f && c.Me[13] == 0.0f && c.Me[14] == 0.0f && c.Me[15] == 1.0f;
  return ok;
}

template <int ARITH>
struct FastRay {
  float head2;          // d-invariant part of row 2 of A @ (u, v, z, 1)
  float pv0, pv1, pv2;  // Mv @ (10*p0x, 10*p0y, 10, p0w), fixed for the pixel

  // false: the pixel does not qualify (a -0 partial sum)
  __device__ __forceinline__ bool init(const Camera &cam, float u, float v, float z_any) {
    float head[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) head[r] = dot2_head<ARITH>(cam.A + 4 * r, u, v);
    head2 = head[2];
    const float p0x = dot2_tail<ARITH>(head[0], cam.A + 0, z_any, 1.0f);
    const float p0y = dot2_tail<ARITH>(head[1], cam.A + 4, z_any, 1.0f);
    const float p0w = dot2_tail<ARITH>(head[3], cam.A + 12, z_any, 1.0f);
    const float n0 = __fmul_rn(p0x, 10.0f), n1 = __fmul_rn(p0y, 10.0f);
    pv0 = dot4<ARITH>(cam.Mv + 0, n0, n1, 10.0f, p0w);
    pv1 = dot4<ARITH>(cam.Mv + 4, n0, n1, 10.0f, p0w);
    pv2 = dot4<ARITH>(cam.Mv + 8, n0, n1, 10.0f, p0w);
    return !(is_neg_zero(head[0]) || is_neg_zero(head[1]) || is_neg_zero(head[3]));
  }

  // voxel id (y*X + x, -1 dropped) of the point at bin height z; `bad` is raised for a point whose
  // identity-BDA shortcut is not provably exact.  a2 = row 2 of A, me = rows 0..2 of Me (registers).
  __device__ __forceinline__ int voxel(const float *a2, const float *me, float ref_h,
                                       bool check_finite, const Grid &g, float z, bool &bad) const {
    const float p0z = dot2_tail<ARITH>(head2, a2, z, 1.0f);
    const float hgt = __fadd_rn(__fmul_rn(-1.0f, p0z), ref_h);
    const float ratio = __fdiv_rn(hgt, pv1);
    const float e0 = __fmul_rn(pv0, ratio), e1 = __fmul_rn(pv1, ratio), e2 = __fmul_rn(pv2, ratio);
    const float gx = dot4<ARITH>(me + 0, e0, e1, e2, 1.0f);
    const float gy = dot4<ARITH>(me + 4, e0, e1, e2, 1.0f);
    const float gz = dot4<ARITH>(me + 8, e0, e1, e2, 1.0f);
    if (check_finite) bad = bad || !(__fadd_rn(__fadd_rn(fabsf(gx), fabsf(gy)), fabsf(gz)) < INFINITY);
    const float tz = __fsub_rn(gz, g.lower[2]);
    if (tz <= g.zt_lo || tz >= g.zt_hi) return -1;
    const int ix = PixelRay<ARITH>::quantize_guarded(gx, g.lower[0], g.size[0], g.rcp_size[0]);
    if ((unsigned)ix >= (unsigned)g.X) return -1;
    const int iy = PixelRay<ARITH>::quantize_guarded(gy, g.lower[1], g.size[1], g.rcp_size[1]);
    if ((unsigned)iy >= (unsigned)g.Y) return -1;
    return iy * g.X + ix;
  }
};

// ---------------------------------------------------------------------------------------------
// Guarded linear shortcut on top of FastRay.  For a qualified pixel the real-valued ego coordinates are
// LINEAR in the bin's height  hgt = RN(ref_h - p0z):   G_r(hgt) = kappa_r * hgt + Me_r3,
// kappa_r = (sum_i Me_ri * pv_i) / pv1,  so the real-valued voxel coordinate is  Q(hgt) = K * hgt + C.
// The reference's fp32 chain (ratio, e_i, dot4, subtract, divide: FastRay::voxel) deviates from Q by at
// most  8.3 u (S + |lower|) / size   with u = 2^-24 and  S = sum_i |Me_ri pv_i| rho_max + |Me_r3|
// (standard forward error analysis: 1 rounding in ratio, 1 in each e_i, <= 4 in the dot product, 1 in the
// subtraction, 1 in the division; rho_max bounds |hgt / pv1| over all bins), and  q' = fma(hgt, K, C)
// with K, C rounded from fp64 deviates from Q by at most 3 u (S + |lower|) / size.  With the guard
// band  delta = 32 u (S + |lower|) / size  (>= 2.8x the sum of both): whenever q' is farther than delta
// from every integer, trunc(q') == trunc(q_reference) -- truncation only jumps at integers -- and the
// index is taken from q'.  Otherwise (about 0.1 % of the points), and for NaN / overflow (the comparison
// fails), the exact chain decides.  The z-range test is resolved once per pixel when the whole height
// range maps strictly inside the grid's z extent by more than the same kind of margin; if not, the
// pixel uses the exact chain throughout.  The voxel index can therefore never differ from the exact
// kernel's: it is either proven equal or computed by it.
// ---------------------------------------------------------------------------------------------
struct LinearGuard {
  float kx, cx, dx;  // q'_x = fma(hgt, kx, cx), guard band dx (voxel units)
  float ky, cy, dy;
  bool ok;           // false: use the exact chain for every bin of this pixel

  // pv*: the pixel's virtual-camera ray (FastRay); me: rows 0..2 of Me; p0z_lo / p0z_hi: bounds of p0z over
  // the bins; ref_h: camera height.
  __device__ __forceinline__ void init(float pv0, float pv1, float pv2, const float *me, float ref_h,
                                       float p0z_lo, float p0z_hi, const Grid &g) {
    const double u32 = 1.9073486328125e-06;  // 32 * 2^-24
    const double rp = 1.0 / (double)pv1;
    const double h_lo = (double)ref_h - (double)p0z_hi, h_hi = (double)ref_h - (double)p0z_lo;
    const double h_abs = fmax(fabs(h_lo), fabs(h_hi)) * 1.000001 + 1e-30;
    const double rho = h_abs * fabs(rp);
    double kap[3], S[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double t0 = (double)me[4 * r + 0] * pv0, t1 = (double)me[4 * r + 1] * pv1, t2 = (double)me[4 * r + 2] * pv2;
      kap[r] = (t0 + t1 + t2) * rp;
      S[r] = (fabs(t0) + fabs(t1) + fabs(t2)) * rho + fabs((double)me[4 * r + 3]);
    }
    const double isx = 1.0 / (double)g.size[0], isy = 1.0 / (double)g.size[1];
    kx = (float)(kap[0] * isx); cx = (float)(((double)me[3] - (double)g.lower[0]) * isx);
    ky = (float)(kap[1] * isy); cy = (float)(((double)me[7] - (double)g.lower[1]) * isy);
    const double ddx = u32 * (S[0] + fabs((double)g.lower[0])) * isx + 1e-30;
    const double ddy = u32 * (S[1] + fabs((double)g.lower[1])) * isy + 1e-30;
    dx = (float)(ddx * 1.0000002);
    dy = (float)(ddy * 1.0000002);
    // z: t_z(hgt) = kap2 * hgt + (Me_23 - lower_z) must stay inside (zt_lo, zt_hi) by the margin over the whole
    // height range (linear => the two ends decide)
    const double cz = (double)me[11] - (double)g.lower[2];
    const double dz = u32 * (S[2] + fabs((double)g.lower[2])) + 1e-30;
    const double tz_a = kap[2] * h_lo + cz, tz_b = kap[2] * h_hi + cz;
    const bool z_in = fmin(tz_a, tz_b) - dz > (double)g.zt_lo && fmax(tz_a, tz_b) + dz < (double)g.zt_hi;
    // every intermediate of the exact chain must be comfortably finite (then e_i, g_r are finite and the
    // identity-bda shortcut of FastRay holds as well)
    const double big = fmax(fmax(fabs((double)pv0), fabs((double)pv1)), fabs((double)pv2)) * rho;
    ok = z_in && ddx < 0.25 && ddy < 0.25 && big < 1e30 && S[0] < 1e30 && S[1] < 1e30 && S[2] < 1e30 &&
         fabs((double)pv1) > 1e-30;  // NaN anywhere makes a comparison fail => not ok
  }

  // voxel id (y*X + x, -1 dropped) from the linear form; `safe` = proven equal to the exact chain
  __device__ __forceinline__ int voxel(float hgt, const Grid &g, bool &safe) const {
    const float qx = __fmaf_rn(hgt, kx, cx), qy = __fmaf_rn(hgt, ky, cy);
    const float ex = fabsf(__fsub_rn(qx, rintf(qx))), ey = fabsf(__fsub_rn(qy, rintf(qy)));
    safe = ex > dx && ey > dy;  // false for NaN
    const int ix = __float2int_rz(qx), iy = __float2int_rz(qy);
    return ((unsigned)ix < (unsigned)g.X && (unsigned)iy < (unsigned)g.Y) ? iy * g.X + ix : -1;
  }
};

// ---------------------------------------------------------------------------------------------
// LinearWalk: the same guarded linear form as LinearGuard, arranged so that ONE thread can walk the height bins
// of a pixel at ~14 full-rate instructions per bin (pixel-block pipeline, lift_splat_block.cu).
//   q' = fma(hgt, k, c - 0.5)          real-valued voxel coordinate minus one half (one rounding)
//   t  = q' + 1.5 * 2^23               rounds q' to the nearest integer r (exact for |q'| < 2^22); r = floor(Q) whenever
//                                      frac(Q) is not 0, and consecutive bins fall into the same voxel column iff their
//                                      t are bitwise equal
//   e  = q' - (t - 1.5 * 2^23)         exact; |e| = |frac(Q) - 0.5|: the bin is within delta of an integer (where the
//                                      reference's fp32 chain may land on the other side) iff |e| > 0.5 - delta
// Truncation toward zero (the reference's `.int()`) differs from floor only for Q in (-1, 0): that column is index
// 0, so r = -1 is mapped to 0 when the index is formed (a spurious change at the 0 boundary then compares equal).
// mode 0: the exact chain decides every bin; 1: linear form with per-bin guard, z provably inside the grid for the
// whole pixel; 3: the same plus a per-bin z test (t_z = G_z - lower_z is linear in hgt as well: the bin is kept /
// dropped by it unless t_z lies within its own guard band of a z threshold, where the exact chain decides --
// Rope3D / SGV3D height bins reach 3.5 m, above the grid's z range [-5, 3]); 2: every bin of the pixel is provably
// dropped (a coordinate stays outside the grid by more than the guard band over the whole height range).
// ---------------------------------------------------------------------------------------------
struct LinearWalk {
  float kx, cx, thx;  // q'_x = fma(hgt, kx, cx); unsafe iff |e_x| > thx
  float ky, cy, thy;
  float kz, cz;            // t_z = fma(hgt, kz, cz)
  float z_in_lo, z_in_hi;  // kept for sure:    z_in_lo < t_z < z_in_hi
  float z_out_lo, z_out_hi;  // dropped for sure: t_z < z_out_lo or t_z > z_out_hi
  int mode;

  // (not inlined: the fp64 set-up must run once per pixel -- inlined, the compiler re-materialises the constants
  // inside the per-bin loop to save registers)
  __device__ __noinline__ void init(float pv0, float pv1, float pv2, const float *me, float ref_h,
                                    float p0z_lo, float p0z_hi, const Grid &g) {
    const double u32 = 1.9073486328125e-06;  // 32 * 2^-24
    const double rp = 1.0 / (double)pv1;
    const double h_lo = (double)ref_h - (double)p0z_hi, h_hi = (double)ref_h - (double)p0z_lo;
    const double h_abs = fmax(fabs(h_lo), fabs(h_hi)) * 1.000001 + 1e-30;
    const double rho = h_abs * fabs(rp);
    double kap[3], S[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double t0 = (double)me[4 * r + 0] * pv0, t1 = (double)me[4 * r + 1] * pv1, t2 = (double)me[4 * r + 2] * pv2;
      kap[r] = (t0 + t1 + t2) * rp;
      S[r] = (fabs(t0) + fabs(t1) + fabs(t2)) * rho + fabs((double)me[4 * r + 3]);
    }
    const double isx = 1.0 / (double)g.size[0], isy = 1.0 / (double)g.size[1];
    const double kxd = kap[0] * isx, cxd = ((double)me[3] - (double)g.lower[0]) * isx;
    const double kyd = kap[1] * isy, cyd = ((double)me[7] - (double)g.lower[1]) * isy;
    kx = (float)kxd; cx = (float)(cxd - 0.5);
    ky = (float)kyd; cy = (float)(cyd - 0.5);
    // guard band: 32 u (S + |lower| + size) / size (the extra `size` covers the rounding of c - 0.5)
    const double ddx = u32 * ((S[0] + fabs((double)g.lower[0])) * isx + 1.0) + 1e-30;
    const double ddy = u32 * ((S[1] + fabs((double)g.lower[1])) * isy + 1.0) + 1e-30;
    thx = (float)((0.5 - ddx) * 0.9999998);
    thy = (float)((0.5 - ddy) * 0.9999998);
    // ranges of the real-valued coordinates over the pixel's height range (linear => the two ends decide)
    const double qxa = kxd * h_lo + cxd, qxb = kxd * h_hi + cxd;
    const double qya = kyd * h_lo + cyd, qyb = kyd * h_hi + cyd;
    const double czd = (double)me[11] - (double)g.lower[2];
    const double dz = u32 * (S[2] + fabs((double)g.lower[2])) + 1e-30;
    const double tz_a = kap[2] * h_lo + czd, tz_b = kap[2] * h_hi + czd;
    const bool z_in = fmin(tz_a, tz_b) - dz > (double)g.zt_lo && fmax(tz_a, tz_b) + dz < (double)g.zt_hi;
    const bool z_out = fmax(tz_a, tz_b) + dz < (double)g.zt_lo || fmin(tz_a, tz_b) - dz > (double)g.zt_hi;
    const double big = fmax(fmax(fabs((double)pv0), fabs((double)pv1)), fabs((double)pv2)) * rho;
    // every intermediate of the exact chain comfortably finite (then the error bounds hold); NaN fails a comparison
    const bool tame = big < 1e30 && S[0] < 1e30 && S[1] < 1e30 && S[2] < 1e30 && fabs((double)pv1) > 1e-30;
    const bool out = tame && (z_out || fmin(qxa, qxb) - ddx > (double)g.X || fmax(qxa, qxb) + ddx < -1.0 ||
                              fmin(qya, qyb) - ddy > (double)g.Y || fmax(qya, qyb) + ddy < -1.0);
    const double qmax = fmax(fmax(fabs(qxa), fabs(qxb)), fmax(fabs(qya), fabs(qyb)));
    const bool lin = tame && ddx < 0.25 && ddy < 0.25 && qmax < 2.0e6;
    // per-bin z test: one more rounding in fma(hgt, kz, cz) and in the thresholds themselves, covered by 2 dz
    kz = (float)kap[2]; cz = (float)czd;
    const double dz2 = 2.0 * dz + 1e-6 * dz;
    z_in_lo = (float)((double)g.zt_lo + dz2); z_in_hi = (float)((double)g.zt_hi - dz2);
    z_out_lo = (float)((double)g.zt_lo - dz2); z_out_hi = (float)((double)g.zt_hi + dz2);
    const bool z_lin = dz2 < 0.25 * (double)g.size[2] && fabs(tz_a) < 1e30 && fabs(tz_b) < 1e30;
    mode = out ? 2 : (lin && z_in ? 1 : (lin && z_lin ? 3 : 0));
  }
};

constexpr unsigned kWalkMagicBits = 0x4B400000u;  // 1.5 * 2^23

// exact chain for one bin, kept out of line: it runs for ~0.5 % of the bins (guard band) and for the pixels the
// linear form does not cover
template <int ARITH>
__device__ __noinline__ int fast_ray_voxel_noinline(const FastRay<ARITH> *ray, const float *a2, const float *me, float ref_h,
                                                    int check_finite, const Grid *g, float z, int *bad) {
  bool b = *bad != 0;
  const int v = ray->voxel(a2, me, ref_h, check_finite != 0, *g, z, b);
  *bad = b ? 1 : 0;
  return v;
}

// Walk the D height bins of one pixel of a camera that qualifies for the fast path (camera_is_fast) and hand every
// bin whose voxel differs from its predecessor's to `emit(d, voxel)` (consecutive equal voxels are reported once or
// more often -- the caller's run builder compares again).  zs: bin values, hs: per-camera height table
// hgt_d = RN(ref_h - p0z_d) valid iff h_uniform (row 2 of ida^-1 ignores u, v); both in shared memory.
// Returns false when a point left the domain in which the shortcuts are proven exact (caller redoes the pixel's
// block with the general chain).
template <int ARITH, typename Emit>
__device__ __forceinline__ bool walk_fast(const Camera &cam, const Grid &grid, const float *zs, const float *hs,
                                          bool h_uniform, int D, float u, float v, float zmin, float zmax, Emit emit) {
  FastRay<ARITH> ray;
  int bad = ray.init(cam, u, v, zs[0]) ? 0 : 1;
  float a2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) a2[i] = cam.A[8 + i];
  const float *mer = cam.Me;
  const float rh = cam.ref_h;
  const int check_finite = cam.has_bda;
  LinearWalk lw;
  {
    const float pa = dot2_tail<ARITH>(ray.head2, a2, zmin, 1.0f);
    const float pb = dot2_tail<ARITH>(ray.head2, a2, zmax, 1.0f);
    lw.init(ray.pv0, ray.pv1, ray.pv2, mer, rh, fminf(pa, pb), fmaxf(pa, pb), grid);
  }
  const int mode = bad ? 0 : lw.mode;
  if (mode == 1 || mode == 3) {
    const float kx = lw.kx, cx = lw.cx, thx = lw.thx, ky = lw.ky, cy = lw.cy, thy = lw.thy;
    const float magic = __uint_as_float(kWalkMagicBits);
    const bool zchk = mode == 3;
    const float kz = lw.kz, cz = lw.cz, zil = lw.z_in_lo, zih = lw.z_in_hi, zol = lw.z_out_lo, zoh = lw.z_out_hi;
    float ptx = -1.0f, pty = -1.0f;   // t values are >= 2^22: -1 never matches
    auto bin = [&](int d, float hgt) {
      const float qx = __fmaf_rn(hgt, kx, cx), qy = __fmaf_rn(hgt, ky, cy);
      const float tx = __fadd_rn(qx, magic), ty = __fadd_rn(qy, magic);
      const float ex = __fsub_rn(qx, __fsub_rn(tx, magic)), ey = __fsub_rn(qy, __fsub_rn(ty, magic));
      bool z_unsafe = false;
      if (zchk) {
        const float tz = __fmaf_rn(hgt, kz, cz);
        const bool zin = tz > zil && tz < zih, zout = tz < zol || tz > zoh;
        z_unsafe = !(zin || zout);
        if (zout) {   // dropped for sure, whatever x / y say (an x / y guard-band hit cannot bring the point back)
          emit(d, -1);
          ptx = pty = -1.0f;
          return;
        }
      }
      if (fabsf(ex) > thx || fabsf(ey) > thy || z_unsafe) {
        // within the guard band of a voxel boundary: the reference's own fp32 chain decides this bin
        emit(d, fast_ray_voxel_noinline<ARITH>(&ray, a2, mer, rh, check_finite, &grid, zs[d], &bad));
        ptx = pty = -1.0f;
      } else if (tx != ptx || ty != pty) {
        ptx = tx; pty = ty;
        int ix = (int)(__float_as_uint(tx) - kWalkMagicBits), iy = (int)(__float_as_uint(ty) - kWalkMagicBits);
        ix = ix == -1 ? 0 : ix;   // truncation toward zero: (-1, 0) belongs to index 0
        iy = iy == -1 ? 0 : iy;
        emit(d, ((unsigned)ix < (unsigned)grid.X && (unsigned)iy < (unsigned)grid.Y) ? iy * grid.X + ix : -1);
      }
    };
    // the per-camera height table is this pixel's too iff its (u, v) part of row 2 is a zero
    if (h_uniform && ray.head2 == 0.0f) {
      for (int d = 0; d < D; ++d) bin(d, hs[d]);
    } else {
      const float head2 = ray.head2;
      for (int d = 0; d < D; ++d) {
        const float p0z = dot2_tail<ARITH>(head2, a2, zs[d], 1.0f);
        bin(d, __fadd_rn(__fmul_rn(-1.0f, p0z), rh));
      }
    }
  } else if (mode == 0) {
    for (int d = 0; d < D; ++d)
      emit(d, fast_ray_voxel_noinline<ARITH>(&ray, a2, mer, rh, check_finite, &grid, zs[d], &bad));
  }
  return bad == 0;
}

// :487-488  ((g - lower) / size).int() -- fp32 subtract, IEEE divide, cvt.rzi.s32.f32
// (truncation toward zero, saturating, NaN -> 0: what `.int()` does on a CUDA tensor).
__device__ __forceinline__ int quantize1(float g, float lower, float size) {
  return __float2int_rz(__fdiv_rn(__fsub_rn(g, lower), size));
}

// voxel id y*X + x of a kept point, -1 for a dropped one (voxel_pooling_forward_cuda.cu:24)
__device__ __forceinline__ int voxel_of(const Grid &g, int ix, int iy, int iz) {
  const bool kept = (unsigned)ix < (unsigned)g.X && (unsigned)iy < (unsigned)g.Y &&
                    (unsigned)iz < (unsigned)g.Z;
  return kept ? iy * g.X + ix : -1;
}

// cooperative load of one camera's operands into shared memory
__device__ __forceinline__ void load_camera(Camera *s, const float *ida_inv, const float *mv,
                                            const float *me, const float *bda, const float *ref_h,
                                            int bn, int b) {
  const int t = threadIdx.x;
  if (t < 16) {
    s->A[t] = ida_inv[16 * (size_t)bn + t];
    s->Mv[t] = mv[16 * (size_t)bn + t];
    s->Me[t] = me[16 * (size_t)bn + t];
    s->Bd[t] = bda ? bda[16 * (size_t)b + t] : 0.0f;
  }
  if (t == 0) {
    s->ref_h = ref_h[bn];
    s->has_bda = bda != nullptr;
    int fast = bda != nullptr;
    if (bda) {
      for (int k = 0; k < 16; ++k) {
        const float want = (k % 5 == 0) ? 1.0f : 0.0f;
        fast = fast && (__float_as_uint(bda[16 * (size_t)b + k]) == __float_as_uint(want));
      }
      for (int k = 12; k < 16; ++k) fast = fast && (fabsf(me[16 * (size_t)bn + k]) < 1e15f);
    }
    s->bda_fast = fast;
  }
}

// ---- host: exact thresholds for the z-range test ------------------------------------------------
inline int host_quantize1(float t, float size) {
  const float q = t / size;  // IEEE single division on the host as well (no fast-math)
  if (q != q) return 0;
  if (q >= 2147483648.0f) return 2147483647;
  if (q <= -2147483648.0f) return (int)0x80000000;
  return (int)q;
}
inline float ordered_to_float(long long k) {  // monotone bijection int <-> float (finite range)
  const unsigned int u = k >= 0 ? (unsigned int)k : 0x80000000u | (unsigned int)(-k);
  float f;
  memcpy(&f, &u, 4);
  return f;
}
// largest t with trunc(t/size) <= -1  and  smallest t with trunc(t/size) >= Z  (size > 0)
inline void z_thresholds(float size, int Z, float *lo, float *hi) {
  const long long kmax = 0x7f7fffff;  // FLT_MAX
  long long a = -kmax, b = kmax;       // hi: first k in [a,b] with quantize >= Z
  while (a < b) {
    const long long m = a + (b - a) / 2;
    if (host_quantize1(ordered_to_float(m), size) >= Z) b = m; else a = m + 1;
  }
  *hi = ordered_to_float(a);
  a = -kmax; b = kmax;                 // lo: last k with quantize <= -1
  while (a < b) {
    const long long m = a + (b - a + 1) / 2;
    if (host_quantize1(ordered_to_float(m), size) <= -1) a = m; else b = m - 1;
  }
  *lo = ordered_to_float(a);
}

}  // namespace geom
}  // namespace sgv3d
