// Batched 2-D transpose with optional row padding and dtype conversion:
//   out[b][s][r] = in[b][r*in_ld + s]   for r < R, s < S;   out[b][s][r] = 0 for R <= r < out_ld
// Used to turn the NCHW context / BEV-gradient planes into channels-last rows (one 16-byte
// aligned row per pixel / voxel) so that the gather kernels read whole rows with 128-bit loads,
// and to turn channels-last gradients back into NCHW.
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

// grid: (ceil(S/32), ceil(out_ld/32), batch); block (32, 8)
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(256)
transpose_pad_kernel(const Tin *__restrict__ in, Tout *__restrict__ out, int R, int S, int in_ld,
                     size_t in_batch_stride, int out_ld, size_t out_batch_stride) {
  __shared__ float tile[32][33];
  const Tin *src = in + (size_t)blockIdx.z * in_batch_stride;
  Tout *dst = out + (size_t)blockIdx.z * out_batch_stride;
  const int s0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int r = r0 + ty + k, s = s0 + tx;
    tile[ty + k][tx] = (r < R && s < S) ? to_f32<Tin>(src[(size_t)r * in_ld + s]) : 0.0f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int s = s0 + ty + k, r = r0 + tx;
    if (s < S && r < out_ld) dst[(size_t)s * out_ld + r] = from_f32<Tout>(tile[tx][ty + k]);
  }
}

template <typename Tin, typename Tout>
inline void launch_transpose_pad(const Tin *in, Tout *out, int batch, int R, int S, int in_ld,
                                 size_t in_batch_stride, int out_ld, size_t out_batch_stride,
                                 cudaStream_t stream) {
  dim3 grid((S + 31) / 32, (out_ld + 31) / 32, batch), block(32, 8);
  transpose_pad_kernel<Tin, Tout><<<grid, block, 0, stream>>>(in, out, R, S, in_ld, in_batch_stride, out_ld,
                                                             out_batch_stride);
}

}  // namespace sgv3d
