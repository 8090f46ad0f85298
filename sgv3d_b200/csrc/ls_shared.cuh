// Pieces shared by the two fused lift-splat pipelines (lift_splat.cu: voxel-tile pipeline; lift_splat_block.cu:
// pixel-block pipeline): problem dimensions, argument validation, cp.async / exp / packed-FMA device helpers.
#pragma once

#include <cuda_bf16.h>

#include "common.cuh"

namespace sgv3d {

constexpr int kChunk = 128;  // pixels per plan chunk of the voxel-tile pipeline == threads per plan CTA
constexpr int kMaxTiles = 4096;  // V <= 262144 voxels per frame (512 x 512)

struct Dims {
  int B, Nc, D, fH, fW, C, X, Y, Z;
  int P;        // fH*fW pixels per camera
  int cpc;      // chunks per camera
  int nchunks;  // chunks per frame = Nc*cpc
  int V;        // X*Y voxels per frame
  int Cpad;     // padded row length of the channels-last copies (elements) = G*(4*NV + NS)
  int G, NV;    // row layout: G lanes per row, NV 4-element vectors per lane (transpose.cuh)
  int esize;    // bytes per context element (4 fp32, 2 bf16)
  int cap;      // max runs per frame (= ELL slots per frame)
  int ntiles;   // ceil(V / 64) reduce tiles per frame
  int logits;             // height tensor holds raw logits (softmax over D fused)
  int cl;                 // BEV map / its gradient are channels-last in memory: (b, y, x, c) (desc.reserved[1] bit 1)
  long long hs, cs;       // element strides between consecutive cameras of height / context
  long long ghs, gcs;     // same for grad_height / grad_context
};


// Row layout by channel count: a G-lane group owns a whole channels-last row, NV 4-element vectors per
// lane (Cpad = 4*G*NV = 16*G*NV bytes in fp32: rows start on 64-byte boundaries, no padding at C = 80).
inline void pick_row_cfg(int C, int *G, int *NV) {
  if (C <= 192) { *G = 8; *NV = ceil_div(C, 32); }
  else { *G = 16; *NV = ceil_div(C, 64); }
}

inline Dims make_dims(const sgv3d_lift_splat_desc *d) {
  Dims m;
  m.B = d->B; m.Nc = d->Nc; m.D = d->D; m.fH = d->fH; m.fW = d->fW; m.C = d->C;
  m.X = d->X; m.Y = d->Y; m.Z = d->Z;
  m.P = m.fH * m.fW;
  m.cpc = ceil_div(m.P, kChunk);
  m.nchunks = m.Nc * m.cpc;
  m.V = m.X * m.Y;
  pick_row_cfg(m.C, &m.G, &m.NV);
  m.Cpad = 4 * m.G * m.NV;
  m.esize = d->ctx_dtype == SGV3D_DTYPE_BF16 ? 2 : 4;
  m.cap = m.nchunks * kChunk * m.D;
  m.ntiles = ceil_div(m.V, 64);
  m.logits = d->height_is_logits;
  m.cl = (d->reserved[1] & 2) ? 1 : 0;
  m.hs = d->height_batch_stride ? d->height_batch_stride : (long long)m.D * m.P;
  m.cs = d->ctx_batch_stride ? d->ctx_batch_stride : (long long)m.C * m.P;
  m.ghs = d->grad_height_batch_stride ? d->grad_height_batch_stride : (long long)m.D * m.P;
  m.gcs = d->grad_ctx_batch_stride ? d->grad_ctx_batch_stride : (long long)m.C * m.P;
  return m;
}


// ---- cp.async: global -> shared copies that do not pass through registers ------------------------
__device__ __forceinline__ void cp_async_4(float *smem_dst, const float *gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_16(float *smem_dst, const float *gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.wait_all;" ::: "memory");
}

__device__ __forceinline__ void cp_async_8(void *smem_dst, const void *gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}


// ---- bulk-async (TMA engine) row copies: cp.async.bulk global -> shared, completion on an mbarrier ----------------
// One instruction moves a whole contiguous row (a multiple of 16 bytes, 16-byte aligned on both sides) without
// touching registers or the LSU; the issuing thread only names source, destination, size and the barrier.  Usage:
//   thread 0: mbar_init(bar, 1); fence; __syncthreads();
//   thread 0: mbar_arrive_expect_tx(bar, total_bytes);   any thread(s): bulk_copy_g2s(dst, src, bytes, bar) ...
//   every consumer: mbar_wait(bar, phase) -- returns once all `total_bytes` have landed (acquire: the data is visible).
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // visible to the async proxy
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *smem_dst, const void *gsrc, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "MBAR_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra MBAR_WAIT_%=;\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// rows of `row_bytes` (multiple of 16, both sides 16-byte aligned): row r from gsrc + r * src_stride_bytes to
// smem_dst + r * dst_stride_bytes.  Called by all threads of the CTA after the barrier was initialised; returns without
// waiting (mbar_wait(bar, phase) does).
__device__ __forceinline__ void bulk_stage_rows(void *smem_dst, unsigned dst_stride_bytes, const void *gsrc,
                                                size_t src_stride_bytes, int nrows, unsigned row_bytes,
                                                unsigned long long *bar) {
  if (threadIdx.x == 0) mbar_arrive_expect_tx(bar, (unsigned)nrows * row_bytes);
  for (int r = threadIdx.x; r < nrows; r += blockDim.x)
    bulk_copy_g2s(static_cast<char *>(smem_dst) + (size_t)r * dst_stride_bytes,
                  static_cast<const char *>(gsrc) + (size_t)r * src_stride_bytes, row_bytes, bar);
}

// exp(x) through the hardware 2^t unit (MUFU.EX2) with a compensated argument: t = x * log2(e) is formed as
// t_hi + t_lo (t_hi the rounded leading product, t_lo its exact residual plus the low part of log2(e)), and
// 2^(t_hi + t_lo) = 2^t_hi * (1 + t_lo ln 2) to first order (|t_lo| < 2^-20 for |x| < 100).  Error ~2 ulp
// (ex2.approx's own 2^-22.5 bound), i.e. libm expf's accuracy class at a third of its instructions.
__device__ __forceinline__ float exp_ex2(float x) {
  const float kHi = 1.44269502162933349609375f, kLo = 1.925963033500011e-8f, kLn2 = 0.693147182464599609375f;
  const float t_hi = __fmul_rn(x, kHi);
  const float t_lo = __fmaf_rn(x, kLo, __fmaf_rn(x, kHi, -t_hi));
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t_hi));
  return x < -104.0f ? 0.0f : __fmaf_rn(r, __fmul_rn(t_lo, kLn2), r);  // exp(-inf) = 0; NaN propagates
}


// acc.{0,1} = w * {x0,x1} + acc.{0,1}: one packed FFMA2 (sm_100 fma.rn.f32x2); each half rounds
// exactly like a scalar fma.rn.
__device__ __forceinline__ void fma2(float &a0, float &a1, float w, float x0, float x1) {
  unsigned long long A, X, W;
  asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(X) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(W) : "f"(w), "f"(w));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(A) : "l"(W), "l"(X), "l"(A));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(A));
}
__device__ __forceinline__ void sts_f32(unsigned addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}


inline int validate(const sgv3d_lift_splat_desc *d, const char *who) {
  SGV3D_REQUIRE(d != nullptr, "%s: desc is null", who);
  SGV3D_REQUIRE(d->B >= 0 && d->Nc > 0 && d->D > 0 && d->fH > 0 && d->fW > 0 && d->C > 0 && d->X > 0 &&
                    d->Y > 0 && d->Z > 0, "%s: bad sizes", who);
  SGV3D_REQUIRE(d->B <= 65535, "%s: B > 65535", who);
  SGV3D_REQUIRE(d->D <= 400, "%s: D=%d > 400 unsupported (height columns are staged in shared memory)", who, d->D);
  SGV3D_REQUIRE(d->C <= 256, "%s: C=%d > 256 unsupported by the fused path", who, d->C);
  SGV3D_REQUIRE((long long)d->X * d->Y <= (long long)kMaxTiles * 64,
                "%s: X*Y exceeds %d voxels per frame", who, kMaxTiles * 64);
  SGV3D_REQUIRE(d->arith >= SGV3D_ARITH_SEQ && d->arith <= SGV3D_ARITH_PAIR, "%s: bad arith", who);
  SGV3D_REQUIRE(d->ctx_dtype == SGV3D_DTYPE_F32 || d->ctx_dtype == SGV3D_DTYPE_BF16, "%s: bad ctx_dtype", who);
  SGV3D_REQUIRE(d->height_is_logits == 0 || d->height_is_logits == 1, "%s: bad height_is_logits", who);
  SGV3D_REQUIRE(d->height_batch_stride >= 0 && d->ctx_batch_stride >= 0 && d->grad_height_batch_stride >= 0 &&
                    d->grad_ctx_batch_stride >= 0, "%s: negative batch stride", who);
  const long long slots = (long long)d->Nc * ceil_div(d->fH * d->fW, kChunk) * kChunk * d->D;
  SGV3D_REQUIRE(slots < (1ll << 31), "%s: more than 2^31 height-bin slots per frame", who);
  // a frame's channels-last context rows are addressed by 32-bit byte offsets, pixel rows by 26 bits
  SGV3D_REQUIRE((long long)d->Nc * d->fH * d->fW < (1ll << 26) &&
                    (long long)d->Nc * d->fH * d->fW * 4 * 256 < (1ll << 32),
                "%s: more than 2^22 pixels per frame", who);
  return SGV3D_OK;
}


// 16-byte cp.async is legal when every (camera, bin, chunk) row start is 16-byte aligned
inline bool columns_vec16(const float *base, long long batch_stride, int P) {
  return (reinterpret_cast<uintptr_t>(base) % 16 == 0) && (batch_stride % 4 == 0) && (P % 4 == 0);
}

template <typename K>
inline int set_smem(K kernel, size_t bytes) {
  if (bytes > 40 * 1024)  // dynamic + static shared memory beyond 48 KB needs the opt-in
    SGV3D_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return SGV3D_OK;
}


}  // namespace sgv3d
