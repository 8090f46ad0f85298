// Batched 2-D transpose with optional row padding, dtype conversion and channel permutation:
//   out[b][s][r] = in[b][r*in_ld + s]   for r < R, s < S;   out[b][s][r] = 0 for R <= r < out_ld
// Used to turn the NCHW context / BEV-gradient planes into channels-last rows (one 64-byte
// aligned row per pixel / voxel) so that the gather kernels read whole rows with 128-bit loads,
// and to turn channels-last gradients back into NCHW.
//
// Channel permutation (RowPerm): inside a channels-last row, channel c sits at element pos(c), chosen
// so that lane l of a G-lane group owns channels {l + G*t} yet still reads them as contiguous
// 128-bit vectors, and so that the lanes of a quarter-warp write 8 different rows of the reduce
// kernel's swizzled [channel][voxel] shared-memory tile (lift_splat.cu).
#pragma once

#include <cuda_bf16.h>

#include "common.cuh"

namespace sgv3d {

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Row layout: G lanes per row (16 or 32), NV 4-element vectors and NS scalars per lane.
//   vector part  [0, 4*G*NV):   element 4*(k*G + l) + e  holds channel  l + G*(4*k + e)
//   scalar part  [4*G*NV, G*(4*NV + NS)):  identity
struct RowPerm {
  int g_shift;  // log2(G); 0 => identity layout
  int nvec;     // 4*G*NV
  __host__ __device__ __forceinline__ int pos(int c) const {  // channel -> element
    if (g_shift == 0 || c >= nvec) return c;
    const int l = c & ((1 << g_shift) - 1), t = c >> g_shift;
    return 4 * (((t >> 2) << g_shift) + l) + (t & 3);
  }
  __host__ __device__ __forceinline__ int chan(int p) const {  // element -> channel
    if (g_shift == 0 || p >= nvec) return p;
    const int e = p & 3, q = p >> 2;
    const int l = q & ((1 << g_shift) - 1), k = q >> g_shift;
    return l + ((4 * k + e) << g_shift);
  }
};

// PERM 0: none.  PERM 1: the OUTPUT minor axis (r) is a permuted channel row: out[s][p] = in[chan(p)][s].
// PERM 2: the INPUT minor axis (s) is a permuted channel row:           out[c][r] = in[r][pos(c)].
// grid: (ceil(S/32), ceil(out_ld/32), batch); block (32, 8)
template <typename Tin, typename Tout, int PERM>
__global__ void __launch_bounds__(256)
transpose_pad_kernel(const Tin *__restrict__ in, Tout *__restrict__ out, int R, int S, int in_ld,
                     size_t in_batch_stride, int out_ld, size_t out_batch_stride, RowPerm perm) {
  __shared__ float tile[32][33];
  const Tin *src = in + (size_t)blockIdx.z * in_batch_stride;
  Tout *dst = out + (size_t)blockIdx.z * out_batch_stride;
  const int s0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    int r = r0 + ty + k, s = s0 + tx;
    const bool in_s = s < S;
    if (PERM == 1) r = perm.chan(r);
    if (PERM == 2) s = perm.pos(s);
    tile[ty + k][tx] = (r < R && in_s) ? to_f32<Tin>(src[(size_t)r * in_ld + s]) : 0.0f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int s = s0 + ty + k, r = r0 + tx;
    if (s < S && r < out_ld) dst[(size_t)s * out_ld + r] = from_f32<Tout>(tile[tx][ty + k]);
  }
}

template <typename Tin, typename Tout, int PERM = 0>
inline void launch_transpose_pad(const Tin *in, Tout *out, int batch, int R, int S, int in_ld,
                                 size_t in_batch_stride, int out_ld, size_t out_batch_stride,
                                 cudaStream_t stream, RowPerm perm = RowPerm{0, 0}) {
  dim3 grid((S + 31) / 32, (out_ld + 31) / 32, batch), block(32, 8);
  transpose_pad_kernel<Tin, Tout, PERM><<<grid, block, 0, stream>>>(in, out, R, S, in_ld, in_batch_stride,
                                                                   out_ld, out_batch_stride, perm);
}

}  // namespace sgv3d
