/*
 * sgv3d_b200.h -- C ABI of the B200-native (sm_100a) lift-splat library, libsgv3d_b200.so.
 *
 * Drop-in boundary for the image->BEV view transform of yanglei18/SGV3D (BEVHeight lineage).
 * Every entry point works on raw DEVICE pointers owned by the caller, launches on the given
 * CUDA stream, never allocates, never synchronises the host, never calls exit(); it returns 0
 * on success and a non-zero status otherwise (text via sgv3d_last_error()).
 * Citations are relative to the reference repository root.
 *
 *   reference interface                                             replaced by
 *   ------------------------------------------------------------    -----------------------------
 *   voxel_pooling_forward_kernel_launcher(...)                      sgv3d_voxel_pooling_forward
 *     ops/voxel_pooling/src/voxel_pooling_forward_cuda.cu:38-56
 *   voxel_pooling_forward_wrapper(...) (pybind11)                   (python: ctypes shim over the above)
 *     ops/voxel_pooling/src/voxel_pooling_forward.cpp:26-43
 *   VoxelPooling.backward (ATen index/index_put chain)              sgv3d_voxel_pooling_backward
 *     ops/voxel_pooling/voxel_pooling.py:57-69
 *   LSSFPN.get_geometry + height2localtion + quantise (ATen)        sgv3d_geometry_quantize
 *     layers/backbones/lss_fpn.py:350-401,487-488
 *   softmax(height) (x) context -> permute -> voxel_pooling         sgv3d_lift_splat_plan / _forward /
 *     layers/backbones/lss_fpn.py:462-495,                          _backward (frustum tensor never
 *     layers/backbones/bsm_lss_fpn.py:523-559                       materialised)
 */
#ifndef SGV3D_B200_H_
#define SGV3D_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGV3D_ABI_VERSION 2

#if defined(__GNUC__)
#define SGV3D_API __attribute__((visibility("default")))
#else
#define SGV3D_API
#endif

/* status codes */
#define SGV3D_OK 0
#define SGV3D_ERR_INVALID_ARGUMENT 1
#define SGV3D_ERR_WORKSPACE_TOO_SMALL 2
#define SGV3D_ERR_CUDA 3
#define SGV3D_ERR_UNSUPPORTED 4

/* Evaluation order of the 4-term dot products in the per-point 4x4 mat-vec products
 * (torch.matmul -> bmm at lss_fpn.py:361-362,367-369,392,398):
 *   SEQ: ((a0*b0 + a1*b1) + a2*b2) + a3*b3, every product and sum rounded to fp32
 *        (what torch's CPU bmm does; bit-exact with the reference run on CPU)
 *   FMA: fma(a3,b3, fma(a2,b2, fma(a1,b1, a0*b0)))  (k-ascending FMA chain)
 *   PAIR: fma(a1,b1, a0*b0) + fma(a3,b3, a2*b2)     (pairwise FMA: what torch's CUDA bmm -> cuBLAS
 *        does for these shapes on B200, i.e. bit-exact with the reference run on the GPU; measured
 *        by tests/probe_arith.py, see profiles/arith_probe_r01.json) */
#define SGV3D_ARITH_SEQ 0
#define SGV3D_ARITH_FMA 1
#define SGV3D_ARITH_PAIR 2

/* element type of the context tensor handed to sgv3d_lift_splat_forward/backward */
#define SGV3D_DTYPE_F32 0
#define SGV3D_DTYPE_BF16 1

typedef void *sgv3d_stream_t; /* cudaStream_t */

SGV3D_API int sgv3d_abi_version(void);
/* Thread-local text of the last non-zero status returned on this thread. */
SGV3D_API const char *sgv3d_last_error(void);

/* ------------------------------------------------------------------------------------------
 * (1) op-level drop-in: voxel_pooling(geom_xyz, input_features, voxel_num)
 *     geom_xyz  int32 [B, N, 3]  (x, y, z) voxel index per point
 *     features  fp32  [B, N, C]
 *     out       fp32  [B, Y, X, C]   fully written by the call (no pre-zeroing required)
 *     pos_memo  int32 [B, N, 3]  (b, y, x) of every kept point, -1 -1 -1 for dropped ones;
 *                                fully written by the call; may be NULL
 * Kept  <=>  0<=x<X && 0<=y<Y && 0<=z<Z  (voxel_pooling_forward_cuda.cu:24).
 * Deterministic: points are stably radix-sorted by voxel and summed in point order per voxel,
 * no floating-point atomics.  `workspace` must hold sgv3d_voxel_pooling_workspace_bytes().
 * ---------------------------------------------------------------------------------------- */
SGV3D_API size_t sgv3d_voxel_pooling_workspace_bytes(int B, int N, int C, int X, int Y, int Z);

SGV3D_API int sgv3d_voxel_pooling_forward(int B, int N, int C, int X, int Y, int Z, const int32_t *geom_xyz,
                                const float *features, float *out, int32_t *pos_memo,
                                void *workspace, size_t workspace_bytes, sgv3d_stream_t stream);

/* grad_features[p, :] = grad_out[b, :, y, x] for kept p, 0 otherwise
 * (ops/voxel_pooling/voxel_pooling.py:57-69).  grad_out is addressed through element strides
 * (sb, sc, sy, sx) so that both a contiguous (B,C,Y,X) tensor and a permuted view of a
 * (B,Y,X,C) buffer are accepted.  When sc != 1 the gradient is first transposed to
 * channels-last inside `workspace` (sgv3d_voxel_pooling_backward_workspace_bytes()); with
 * sc == 1 the workspace may be NULL. */
SGV3D_API size_t sgv3d_voxel_pooling_backward_workspace_bytes(int B, int C, int X, int Y);

SGV3D_API int sgv3d_voxel_pooling_backward(int B, int N, int C, int X, int Y, const float *grad_out,
                                 int64_t sb, int64_t sc, int64_t sy, int64_t sx,
                                 const int32_t *pos_memo, float *grad_features, void *workspace,
                                 size_t workspace_bytes, sgv3d_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (2) geometry + quantisation: get_geometry / height2localtion / ((geom - lower) / size).int()
 *     u_tab[fW], v_tab[fH], z_tab[D]   the three axes of the module's `frustum` buffer
 *                                      (lss_fpn.py:325-348): frustum[0,0,:,0], [0,:,0,1], [:,0,0,2]
 *     ida_inv, m_virtual, m_ego        fp32 [B*Nc, 16] row-major: ida.inverse(),
 *                                      sensor2virtual @ inverse(intrin),
 *                                      sensor2ego @ inverse(sensor2virtual)  (lss_fpn.py:392,361,367)
 *     bda                              fp32 [B, 16] or NULL                  (lss_fpn.py:394-398)
 *     ref_heights                      fp32 [B*Nc]
 *     lower3 / size3                   HOST pointers to 3 floats: fp32(voxel_coord - voxel_size/2)
 *                                      and voxel_size (lss_fpn.py:487-488)
 *     idx_out  int32 [B, Nc, D, fH, fW, 3]   (may be NULL)
 *     xyz_out  fp32  [B, Nc, D, fH, fW, 3]   (may be NULL; un-quantised ego coordinates)
 * ---------------------------------------------------------------------------------------- */
SGV3D_API int sgv3d_geometry_quantize(int arith, int B, int Nc, int D, int fH, int fW, const float *u_tab,
                            const float *v_tab, const float *z_tab, const float *ida_inv,
                            const float *m_virtual, const float *m_ego, const float *bda,
                            const float *ref_heights, const float *lower3, const float *size3,
                            int32_t *idx_out, float *xyz_out, sgv3d_stream_t stream);

/* Batched 4x4 inverse, bit-identical to torch.inverse on a CUDA fp32 tensor (what the reference calls at
 * lss_fpn.py:361,367,392; cuBLAS batched LU + solves: reciprocal-scaled LU with partial pivoting and FMA
 * updates, FMA forward / backward substitution, one division by the diagonal).  Up to three sets of n
 * row-major matrices in one launch (ida, intrin, sensor2virtual); a1/inv1 and a2/inv2 may be NULL.
 * Singular or non-finite matrices give unspecified non-finite output (as torch.linalg.inv_ex does). */
SGV3D_API int sgv3d_inverse4x4(int n, const float *a0, const float *a1, const float *a2, float *inv0,
                     float *inv1, float *inv2, sgv3d_stream_t stream);

/* The whole per-camera 4x4 prep of lss_fpn.py:361,367,392 in one launch (n cameras, row-major fp32):
 *   ida_inv   = inverse(ida)
 *   m_virtual = sensor2virtual @ inverse(intrin)
 *   m_ego     = sensor2ego @ inverse(sensor2virtual)
 * inverses as sgv3d_inverse4x4; the products in the rounding order torch's CUDA matmul uses for these
 * batches (tools/probe_matmul.py): product_arith = SGV3D_ARITH_SEQ for a single matrix, SGV3D_ARITH_FMA
 * (k-ascending FMA chain) for two or more.  The Python host verifies both against torch before relying on them. */
SGV3D_API int sgv3d_camera_prep(int n, int product_arith, const float *ida, const float *intrin,
                      const float *sensor2virtual, const float *sensor2ego, float *ida_inv,
                      float *m_virtual, float *m_ego, sgv3d_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (3) fused lift-splat.  The plan (index) depends only on calibration + grid; forward/backward
 *     (values) depend on the activations.  Static roadside cameras can build the plan once.
 *
 *     height   fp32 [B*Nc, D, fH, fW]   softmax-ed height-bin probabilities (lss_fpn.py:462), or the
 *                                       raw logits when desc.height_is_logits = 1
 *     context  fp32|bf16 [B*Nc, C, fH, fW]
 *     bev      fp32 [B, C, Y, X]        contiguous, fully written (lss_fpn.py:494-495)
 *     BEV[b,c,y,x] = sum over kept points (n,d,h,w)->(x,y) of height[bn,d,h,w]*context[bn,c,h,w]
 * ---------------------------------------------------------------------------------------- */
typedef struct sgv3d_lift_splat_desc {
  int32_t B, Nc, D, fH, fW, C; /* frames, cameras per frame, height bins, feature map, channels */
  int32_t X, Y, Z;             /* voxel grid (voxel_num) */
  int32_t arith;               /* SGV3D_ARITH_* */
  int32_t ctx_dtype;           /* SGV3D_DTYPE_* */
  int32_t height_is_logits;    /* 1: `height` holds raw height-net logits; the softmax over D
                                  (lss_fpn.py:462) and its backward run inside the kernels */
  /* Element strides between consecutive cameras (0 => densely packed).  They let `height` and
   * `context` (and their gradients) be views into the height net's (B*Nc, D + C, fH, fW) output,
   * lss_fpn.py:461-466, without a copy.  Within one camera the [D|C][fH][fW] block is dense. */
  int64_t height_batch_stride, ctx_batch_stride;
  int64_t grad_height_batch_stride, grad_ctx_batch_stride;
  /* reserved[0]: kernel pipeline -- 0 / 1 the voxel-tile pipeline (lift_splat.cu), 2 the pixel-block pipeline
   * (lift_splat_block.cu; C <= 96, D <= 255, error otherwise; the Python wrapper's AUTO policy picks it for small
   * inference batches).  Must be the same for every call that shares a workspace.
   * reserved[1]: memory layout of the BEV map -- 0: (B, C, Y, X) contiguous, what lss_fpn.py:494-495 returns; 2: the same
   * logical tensor in channels-last memory order (b, y, x, c), i.e. torch.channels_last strides, for a BEV trunk that runs
   * its convolutions in that format: `bev` of sgv3d_lift_splat_forward is written and `grad_bev` of
   * sgv3d_lift_splat_backward is read in that order (no transposed copies on either side).  Voxel-tile pipeline,
   * C in {16, 32, ..., 96}, 16-byte aligned pointers; error otherwise.  reserved[2..3]: 0. */
  int32_t reserved[4];
} sgv3d_lift_splat_desc;

/* 1 when the calls for this descriptor run on the pixel-block pipeline, 0 for the voxel-tile pipeline. */
SGV3D_API int sgv3d_lift_splat_uses_block_pipeline(const sgv3d_lift_splat_desc *desc);

SGV3D_API size_t sgv3d_lift_splat_workspace_bytes(const sgv3d_lift_splat_desc *desc);

/* Geometry -> per-pixel voxel runs along D -> stable radix sort by voxel.  Same pointer
 * arguments as sgv3d_geometry_quantize.  The result lives in `workspace`. */
SGV3D_API int sgv3d_lift_splat_plan(const sgv3d_lift_splat_desc *desc, const float *u_tab,
                          const float *v_tab, const float *z_tab, const float *ida_inv,
                          const float *m_virtual, const float *m_ego, const float *bda,
                          const float *ref_heights, const float *lower3, const float *size3,
                          void *workspace, size_t workspace_bytes, sgv3d_stream_t stream);

SGV3D_API int sgv3d_lift_splat_forward(const sgv3d_lift_splat_desc *desc, const float *height,
                             const void *context, float *bev, void *workspace,
                             size_t workspace_bytes, sgv3d_stream_t stream);

/* BSMLSSFPN call site (bsm_lss_fpn.py:523-559), inference: the context assembly of bsm_lss_fpn.py:524-529
 *   semantic = softmax(semantic_logits, channel axis)
 *   tran_feat = cat(context, semantic) * (1 - (semantic[:, 0] > background_threshold))
 * happens inside the forward's context pass; the C = (C - Cs) + Cs channel tensor is never materialised.
 *     context          fp32 [B*Nc, C - Cs, fH, fW]   (desc->ctx_batch_stride: elements between cameras, 0 = dense)
 *     semantic_logits  fp32 [B*Nc, Cs, fH, fW]       (semantic_batch_stride likewise)
 * desc->C is the BEV channel count (87 = 80 + 7 in exps/sgv3d/...:40,88); desc->ctx_dtype must be F32.
 * The softmax restates torch's CUDA softmax over a channel axis (per pixel: max, sum of exp(x - max) in
 * channel order, exp(x - max) / sum), so the background mask is bit-identical to the reference's. */
SGV3D_API int sgv3d_lift_splat_forward_bsm(const sgv3d_lift_splat_desc *desc, const float *height,
                                 const float *context, const float *semantic_logits,
                                 int semantic_channels, int64_t semantic_batch_stride,
                                 float background_threshold, float *bev, void *workspace,
                                 size_t workspace_bytes, sgv3d_stream_t stream);

/* grad_height fp32 [B*Nc, D, fH, fW], grad_context fp32 [B*Nc, C, fH, fW]; both fully written.
 * grad_bev fp32 [B, C, Y, X] contiguous. */
SGV3D_API int sgv3d_lift_splat_backward(const sgv3d_lift_splat_desc *desc, const float *grad_bev,
                              const float *height, const void *context, float *grad_height,
                              float *grad_context, void *workspace, size_t workspace_bytes,
                              sgv3d_stream_t stream);

/* Backward of the BSMLSSFPN call site with the context assembly fused (bsm_lss_fpn.py:523-541 under autograd): the
 * kernel re-assembles cat(context, softmax(semantic_logits)) * (1 - background mask) per pixel chunk in shared
 * memory, so neither pass ever materialises the C-channel tensor, and pixels masked as background are skipped (all
 * their gradients are exact zeros).  grad_context fp32 [B*Nc, C - Cs, fH, fW], grad_semantic fp32 [B*Nc, Cs, fH, fW]
 * (gradient w.r.t. the semantic LOGITS: softmax backward; the mask itself is not differentiable), grad_height as in
 * sgv3d_lift_splat_backward.  Cs <= 8, C <= 96, fp32 context, voxel-tile pipeline. */
SGV3D_API int sgv3d_lift_splat_backward_bsm(const sgv3d_lift_splat_desc *desc, const float *grad_bev,
                                  const float *height, const float *context, const float *semantic_logits,
                                  int semantic_channels, int64_t semantic_batch_stride,
                                  float background_threshold, float *grad_height, float *grad_context,
                                  float *grad_semantic, void *workspace, size_t workspace_bytes,
                                  sgv3d_stream_t stream);

/* Debug / parity: expand the plan back to one voxel id per point.
 * vox_out int32 [B, Nc, D, fH, fW]: y*X + x of the voxel the point falls in, -1 if dropped. */
SGV3D_API int sgv3d_lift_splat_plan_expand(const sgv3d_lift_splat_desc *desc, int32_t *vox_out,
                                 void *workspace, size_t workspace_bytes, sgv3d_stream_t stream);

/* Number of kernels this library has launched on this thread since the last reset (bench.py's
 * `gpu_launches`). */
SGV3D_API int64_t sgv3d_launch_count(int reset);

/* Measurement aid (bench.py's roofline): while enabled, every kernel this library launches on
 * the calling thread is bracketed by CUDA events on its launch stream.  sgv3d_profile_report()
 * waits for them, writes one "kernel_name,launches,total_ms" line per kernel into buf, clears
 * the accumulators and returns the text length.  Off by default; zero cost when off. */
SGV3D_API int sgv3d_profile_enable(int on);
SGV3D_API long sgv3d_profile_report(char *buf, size_t buflen);

#ifdef __cplusplus
}
#endif
#endif /* SGV3D_B200_H_ */
