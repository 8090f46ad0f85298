/*
 * sgv3d_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C CPU restatement of the SGV3D / BEVHeight image->BEV lift-splat hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * leg may load this library; the product path (sgv3d_b200/) must never import it.
 *
 * Every function cites the reference file:line (relative to /root/reference) that it
 * restates.  Parity pin: the reference ships no tests / golden vectors, so
 * this restatement is pinned against the reference's own Python executed in the build
 * container (tests/golden/make_golden.py imports layers/backbones/lss_fpn.py unmodified)
 * -- the committed fixtures under tests/golden/ carry those outputs.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -fopenmp (see oracle/Makefile).
 * -ffp-contract=off is REQUIRED: the whole point is to control where fp32 roundings occur.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <limits.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* Arithmetic orders for the 4-term dot products of the per-point 4x4 mat-vec products
 * (lss_fpn.py:361-362,367-369,392,398 -- torch.matmul -> bmm):
 *   ORACLE_ARITH_SEQ : ((a0*b0 + a1*b1) + a2*b2) + a3*b3, every mul and add rounded
 *                      (what torch CPU bmm does for these shapes -- SURVEY.md §7 hard part 1)
 *   ORACLE_ARITH_FMA : fma(a3,b3, fma(a2,b2, fma(a1,b1, a0*b0)))  (k-ascending FMA chain)
 *   ORACLE_ARITH_PAIR: fma(a1,b1, a0*b0) + fma(a3,b3, a2*b2)      (pairwise FMA: what torch's CUDA
 *                      bmm / cuBLAS does for these shapes on B200 -- 0 bit mismatches over
 *                      4 x 3.7M outputs per stage, tests/probe_arith.py, profiles/arith_probe_r01.json)
 */
#define ORACLE_ARITH_SEQ 0
#define ORACLE_ARITH_FMA 1
#define ORACLE_ARITH_PAIR 2

static inline float dot4(int mode, const float *a, float b0, float b1, float b2, float b3) {
  if (mode == ORACLE_ARITH_PAIR) {
    const float lo = fmaf(a[1], b1, a[0] * b0);
    const float hi = fmaf(a[3], b3, a[2] * b2);
    return lo + hi;
  }
  if (mode == ORACLE_ARITH_FMA) {
    float acc = a[0] * b0;
    acc = fmaf(a[1], b1, acc);
    acc = fmaf(a[2], b2, acc);
    acc = fmaf(a[3], b3, acc);
    return acc;
  }
  float acc = a[0] * b0;
  acc = acc + a[1] * b1;
  acc = acc + a[2] * b2;
  acc = acc + a[3] * b3;
  return acc;
}

/* float -> int32 conversion with the semantics the reference sees on the GPU
 * (`.int()` on a CUDA tensor, lss_fpn.py:487-488 => cvt.rzi.s32.f32): truncate toward zero,
 * saturate, NaN -> 0.  (x86's cvttss2si would give INT_MIN for NaN/overflow instead;
 * SURVEY.md §7 hard part 2.) */
static inline int32_t f2i_rz_sat(float v) {
  if (v != v) return 0;
  if (v >= 2147483648.0f) return INT32_MAX;
  if (v <= -2147483648.0f) return INT32_MIN;
  return (int32_t)v;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/*
 * Per-point ray -> height-plane geometry.
 * Restates LSSFPN.get_geometry (layers/backbones/lss_fpn.py:372-401) and
 * LSSFPN.height2localtion (lss_fpn.py:350-370); identical code in
 * layers/backbones/bsm_lss_fpn.py:409-460.
 *
 * Inputs (all fp32, row-major 4x4):
 *   u_tab[fW], v_tab[fH], z_tab[D] : the three axes of the frustum buffer built by
 *                                    create_frustum (lss_fpn.py:325-348)
 *   ida_inv[BN][16] : ida_mat.inverse()                      (lss_fpn.py:391-392)
 *   mv[BN][16]      : sensor2virtual @ inverse(intrin)       (lss_fpn.py:361)
 *   me[BN][16]      : sensor2ego @ inverse(sensor2virtual)   (lss_fpn.py:367)
 *   bda[B][16] or NULL                                       (lss_fpn.py:394-398)
 *   ref_h[BN]                                                (lss_fpn.py:352-354)
 * Output: xyz[BN][D][fH][fW][3] fp32 (== get_geometry(...)[..., :3]).
 */
void oracle_geometry(int mode, int B, int Nc, int D, int fH, int fW, const float *u_tab,
                     const float *v_tab, const float *z_tab, const float *ida_inv,
                     const float *mv, const float *me, const float *bda, const float *ref_h,
                     float *xyz) {
  const long plane = (long)fH * fW;
#pragma omp parallel for collapse(2) schedule(static)
  for (int bn = 0; bn < B * Nc; ++bn) {
    for (int d = 0; d < D; ++d) {
      const float *A = ida_inv + 16 * bn;
      const float *Mv = mv + 16 * bn;
      const float *Me = me + 16 * bn;
      const float *Bd = bda ? bda + 16 * (bn / Nc) : 0;
      const float rh = ref_h[bn];
      const float z = z_tab[d];
      float *o = xyz + ((long)bn * D + d) * plane * 3;
      for (int h = 0; h < fH; ++h) {
        const float v = v_tab[h];
        for (int w = 0; w < fW; ++w) {
          const float u = u_tab[w];
          /* lss_fpn.py:392  points = ida_mat.inverse().matmul(frustum) */
          float p0[4];
          for (int r = 0; r < 4; ++r) p0[r] = dot4(mode, A + 4 * r, u, v, z, 1.0f);
          /* lss_fpn.py:354  height = -1 * points.z + reference_heights */
          const float hgt = (-1.0f * p0[2]) + rh;
          /* lss_fpn.py:356-360  z := 10, then (x,y) *= z */
          const float q0 = p0[0] * 10.0f, q1 = p0[1] * 10.0f, q2 = 10.0f, q3 = p0[3];
          /* lss_fpn.py:361-362  points_virtual = (sensor2virtual @ K^-1) @ q */
          float pv[4];
          for (int r = 0; r < 4; ++r) pv[r] = dot4(mode, Mv + 4 * r, q0, q1, q2, q3);
          /* lss_fpn.py:363-366  ratio = height / pv.y ; points = pv * ratio ; w := 1 */
          const float ratio = hgt / pv[1];
          const float e0 = pv[0] * ratio, e1 = pv[1] * ratio, e2 = pv[2] * ratio, e3 = 1.0f;
          /* lss_fpn.py:367-369  points = (sensor2ego @ sensor2virtual^-1) @ points */
          float pg[4];
          for (int r = 0; r < 4; ++r) pg[r] = dot4(mode, Me + 4 * r, e0, e1, e2, e3);
          /* lss_fpn.py:394-398  optional bda */
          if (Bd) {
            float pb[4];
            for (int r = 0; r < 4; ++r) pb[r] = dot4(mode, Bd + 4 * r, pg[0], pg[1], pg[2], pg[3]);
            pg[0] = pb[0]; pg[1] = pb[1]; pg[2] = pb[2];
          }
          float *oo = o + ((long)h * fW + w) * 3;
          oo[0] = pg[0]; oo[1] = pg[1]; oo[2] = pg[2];
        }
      }
    }
  }
}

/*
 * Voxel-index quantisation. Restates lss_fpn.py:487-488 (bsm_lss_fpn.py:552-553):
 *   ((geom - (voxel_coord - voxel_size / 2.0)) / voxel_size).int()
 * `lower` = fp32(voxel_coord - voxel_size/2.0) is computed by the caller exactly as the
 * reference does (fp32 tensor ops); fp32 subtraction, fp32 true division, truncation.
 */
void oracle_quantize(long n_points, const float *xyz, const float *lower, const float *size,
                     int32_t *idx) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n_points; ++i) {
    for (int k = 0; k < 3; ++k) {
      const float t = xyz[3 * i + k] - lower[k];
      idx[3 * i + k] = f2i_rz_sat(t / size[k]);
    }
  }
}

/*
 * voxel_pooling forward. Restates voxel_pooling_forward_kernel
 * (ops/voxel_pooling/src/voxel_pooling_forward_cuda.cu:9-36): bounds test incl. z (:24),
 * pos_memo = (b, y, x) (:27-29), out[((b*Y+y)*X+x)*C+c] += feat[pt*C+c] (:30-34).
 * pos_memo must be pre-filled with -1 and out with 0 by the caller
 * (ops/voxel_pooling/voxel_pooling.py:37-40).  Sequential point order => deterministic
 * (the reference's atomics are not).  acc64 != 0 accumulates in double (tolerance anchor,
 * SURVEY.md §7 hard part 7); out64 then receives the double sums.
 */
void oracle_voxel_pooling_forward(int B, int N, int C, int X, int Y, int Z, const int32_t *geom,
                                  const float *feat, float *out, int32_t *pos_memo,
                                  double *out64) {
#pragma omp parallel for schedule(static)
  for (int b = 0; b < B; ++b) {
    for (long p = (long)b * N; p < (long)(b + 1) * N; ++p) {
      const int x = geom[3 * p], y = geom[3 * p + 1], z = geom[3 * p + 2];
      if (x < 0 || x >= X || y < 0 || y >= Y || z < 0 || z >= Z) continue;
      if (pos_memo) {
        pos_memo[3 * p] = b; pos_memo[3 * p + 1] = y; pos_memo[3 * p + 2] = x;
      }
      const long o = (((long)b * Y + y) * X + x) * C;
      const float *f = feat + p * C;
      if (out64) {
        for (int c = 0; c < C; ++c) out64[o + c] += (double)f[c];
      } else {
        for (int c = 0; c < C; ++c) out[o + c] = out[o + c] + f[c];
      }
    }
  }
}

/*
 * voxel_pooling backward. Restates VoxelPooling.backward
 * (ops/voxel_pooling/voxel_pooling.py:57-69): kept = pos_memo[...,0] != -1;
 * grad_feat[p,:] = grad_out[b,:,y,x] for kept p, 0 otherwise.
 * grad_out is (B, C, Y, X) contiguous; grad_feat is (B, N, C).
 */
void oracle_voxel_pooling_backward(int B, int N, int C, int X, int Y, const float *grad_out,
                                   const int32_t *pos_memo, float *grad_feat) {
#pragma omp parallel for schedule(static)
  for (long p = 0; p < (long)B * N; ++p) {
    float *g = grad_feat + p * C;
    if (pos_memo[3 * p] == -1) {
      memset(g, 0, sizeof(float) * C);
      continue;
    }
    const int b = pos_memo[3 * p], y = pos_memo[3 * p + 1], x = pos_memo[3 * p + 2];
    for (int c = 0; c < C; ++c) g[c] = grad_out[(((long)b * C + c) * Y + y) * X + x];
  }
}

/*
 * Whole lift-splat forward without materialising the frustum tensor, accumulated in double:
 *   BEV[b,c,y,x] = sum_{n,d,h,w -> (x,y) kept} height[bn,d,h,w] * ctx[bn,c,h,w]
 * Restates lss_fpn.py:462-495 (outer product :464-466, permute :486, quantised indices
 * supplied by the caller from oracle_geometry + oracle_quantize, voxel_pooling :490-491,
 * final (B,C,Y,X) contiguous :494-495).  The per-point product is rounded to fp32 first
 * (the reference materialises height*ctx in fp32, :464) and then summed in double.
 */
void oracle_lift_splat_forward64(int B, int Nc, int D, int fH, int fW, int C, int X, int Y, int Z,
                                 const int32_t *idx, const float *height, const float *ctx,
                                 double *bev) {
  const long P = (long)fH * fW;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b) {
    for (int c = 0; c < C; ++c) {
      double *o = bev + ((long)b * C + c) * Y * X;
      for (int n = 0; n < Nc; ++n) {
        const long bn = (long)b * Nc + n;
        const float *cx = ctx + (bn * C + c) * P;
        for (int d = 0; d < D; ++d) {
          const float *hh = height + (bn * D + d) * P;
          const int32_t *ii = idx + (bn * D + d) * P * 3;
          for (long p = 0; p < P; ++p) {
            const int x = ii[3 * p], y = ii[3 * p + 1], z = ii[3 * p + 2];
            if (x < 0 || x >= X || y < 0 || y >= Y || z < 0 || z >= Z) continue;
            const float prod = hh[p] * cx[p];
            o[(long)y * X + x] += (double)prod;
          }
        }
      }
    }
  }
}

/*
 * Lift-splat backward in double: given grad_bev (B,C,Y,X) returns
 *   g_height[bn,d,h,w] = sum_c grad_bev[b,c,y,x] * ctx[bn,c,h,w]     (kept points, else 0)
 *   g_ctx[bn,c,h,w]    = sum_d grad_bev[b,c,y,x] * height[bn,d,h,w]  (kept points)
 * i.e. VoxelPooling.backward (voxel_pooling.py:57-69) followed by autograd of the outer
 * product at lss_fpn.py:464-466.
 */
void oracle_lift_splat_backward64(int B, int Nc, int D, int fH, int fW, int C, int X, int Y,
                                  int Z, const int32_t *idx, const float *height,
                                  const float *ctx, const float *grad_bev, double *g_height,
                                  double *g_ctx) {
  const long P = (long)fH * fW;
#pragma omp parallel for schedule(static)
  for (long bn = 0; bn < (long)B * Nc; ++bn) {
    const int b = (int)(bn / Nc);
    for (int d = 0; d < D; ++d) {
      const float *hh = height + (bn * D + d) * P;
      const int32_t *ii = idx + (bn * D + d) * P * 3;
      double *gh = g_height + (bn * D + d) * P;
      for (long p = 0; p < P; ++p) {
        const int x = ii[3 * p], y = ii[3 * p + 1], z = ii[3 * p + 2];
        if (x < 0 || x >= X || y < 0 || y >= Y || z < 0 || z >= Z) {
          gh[p] = 0.0;
          continue;
        }
        double acc = 0.0;
        for (int c = 0; c < C; ++c) {
          const double g = (double)grad_bev[(((long)b * C + c) * Y + y) * X + x];
          acc += g * (double)ctx[(bn * C + c) * P + p];
          g_ctx[(bn * C + c) * P + p] += g * (double)hh[p];
        }
        gh[p] = acc;
      }
    }
  }
}
