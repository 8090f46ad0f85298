// Op-level drop-in for ops/voxel_pooling (reference: voxel_pooling_forward_cuda.cu:9-56 forward,
// voxel_pooling.py:57-69 backward), re-designed for B200:
//
//   forward  = prepare (kept test, pos_memo, voxel key, digit histogram)
//              -> stable 2-pass radix sort of the points by voxel (sort.cuh)
//              -> per-voxel gather-sum of whole feature rows (one warp per voxel, 128-bit loads,
//                 coalesced channels-last output, zero rows for empty voxels)
//   backward = [transpose grad to channels-last] -> per-point row copy (128-bit loads/stores)
//
// No floating-point atomics: the per-voxel sum runs in ascending point order, so the result is
// bitwise reproducible run to run (the reference's atomicAdd order is not).
// Both kernels are HBM-bound: the B x N x C feature tensor is an API input here and has to be
// read (forward) or written (backward) exactly once -- 4*N*C bytes per frame.
#include "sort.cuh"
#include "transpose.cuh"

namespace sgv3d {
namespace {

struct VpWorkspace {
  int *keys0, *keys1, *pay1, *keys2, *pay2, *hist1, *hist2, *row_ptr;
  int nblk, bins2, V;
  size_t bytes;
};

VpWorkspace carve_vp(void *ws, int B, int N, int X, int Y) {
  VpWorkspace w;
  w.V = X * Y;
  w.nblk = ceil_div(N, sort::kItemsPerBlock);
  w.bins2 = (w.V >> sort::kLowBits) + 1;
  Carver c(ws);
  const size_t bn = (size_t)B * N;
  w.keys0 = c.take<int>(bn);
  w.keys1 = c.take<int>(bn);
  w.pay1 = c.take<int>(bn);
  w.keys2 = c.take<int>(bn);
  w.pay2 = c.take<int>(bn);
  w.hist1 = c.take<int>((size_t)B * sort::kLowBins * w.nblk);
  w.hist2 = c.take<int>((size_t)B * w.bins2 * w.nblk);
  w.row_ptr = c.take<int>((size_t)B * (w.V + 1));
  w.bytes = c.used();
  return w;
}

// ---- prepare: one thread per point ------------------------------------------------------------
// key = y*X + x for kept points, V (sorts behind every real voxel) for dropped ones.
__global__ void __launch_bounds__(sort::kThreads)
vp_prepare_kernel(int N, int X, int Y, int Z, const int32_t *__restrict__ geom,
                  int32_t *__restrict__ pos_memo, int *__restrict__ keys0, int nblk,
                  int *__restrict__ hist1) {
  __shared__ int s_hist[sort::kLowBins];
  const int frame = blockIdx.y, blk = blockIdx.x;
  for (int i = threadIdx.x; i < sort::kLowBins; i += sort::kThreads) s_hist[i] = 0;
  __syncthreads();
  const int V = X * Y;
  const int begin = blk * sort::kItemsPerBlock, end = min(N, begin + sort::kItemsPerBlock);
  const size_t fb = (size_t)frame * N;
  for (int i = begin + threadIdx.x; i < end; i += sort::kThreads) {
    const int32_t *g = geom + (fb + i) * 3;
    const int x = g[0], y = g[1], z = g[2];
    const bool kept = (unsigned)x < (unsigned)X && (unsigned)y < (unsigned)Y &&
                      (unsigned)z < (unsigned)Z;  // voxel_pooling_forward_cuda.cu:24
    const int key = kept ? y * X + x : V;
    keys0[fb + i] = key;
    if (pos_memo) {  // voxel_pooling_forward_cuda.cu:27-29; -1 where the reference leaves its fill
      int32_t *pm = pos_memo + (fb + i) * 3;
      pm[0] = kept ? frame : -1;
      pm[1] = kept ? y : -1;
      pm[2] = kept ? x : -1;
    }
    atomicAdd(&s_hist[key & (sort::kLowBins - 1)], 1);
  }
  __syncthreads();
  int *h = hist1 + (size_t)frame * sort::kLowBins * nblk;
  for (int i = threadIdx.x; i < sort::kLowBins; i += sort::kThreads)
    h[(size_t)i * nblk + blk] = s_hist[i];
}

// ---- reduce: one warp per voxel ---------------------------------------------------------------
// VEC = 4: C % 4 == 0, lanes own float4 slices (row base is 16-byte aligned).  VEC = 1: any C.
// NCH = number of VEC-wide slices per lane = ceil(C / (32*VEC)).
template <int VEC, int NCH>
__global__ void __launch_bounds__(256)
vp_reduce_kernel(int N, int C, int V, const float *__restrict__ feat,
                 const int *__restrict__ row_ptr, const int *__restrict__ pay2,
                 float *__restrict__ out) {
  const int frame = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (v >= V) return;
  const int *rp = row_ptr + (size_t)frame * (V + 1);
  const int lo = rp[v], hi = rp[v + 1];
  const float *fbase = feat + (size_t)frame * N * C;
  const int *perm = pay2 + (size_t)frame * N;
  float acc[NCH][VEC];
#pragma unroll
  for (int k = 0; k < NCH; ++k)
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[k][e] = 0.0f;

  for (int j0 = lo; j0 < hi; j0 += 32) {
    const int cnt = min(32, hi - j0);
    const int mine = (lane < cnt) ? perm[j0 + lane] : 0;
    int i = 0;
    // kInFlight rows in flight per warp before the dependent adds: a warp owns one voxel, the most populated voxels
    // (855 points at DAIR-R50) set the kernel's tail at small batches, and their time is rows / kInFlight x latency
    constexpr int kInFlight = NCH == 1 ? 8 : 4;
    for (; i + kInFlight <= cnt; i += kInFlight) {
      float r[kInFlight][NCH][VEC];
#pragma unroll
      for (int u = 0; u < kInFlight; ++u) {
        const float *row = fbase + (size_t)__shfl_sync(0xffffffffu, mine, i + u) * C;
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
          const int c = (k * 32 + lane) * VEC;
          if (VEC == 4) {
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < C) t = ldg_stream_f4(reinterpret_cast<const float4 *>(row + c));
            r[u][k][0] = t.x; r[u][k][1 % VEC] = t.y; r[u][k][2 % VEC] = t.z; r[u][k][3 % VEC] = t.w;
          } else {
            r[u][k][0] = (c < C) ? ldg_stream_f1(row + c) : 0.0f;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kInFlight; ++u)
#pragma unroll
        for (int k = 0; k < NCH; ++k)
#pragma unroll
          for (int e = 0; e < VEC; ++e) acc[k][e] = __fadd_rn(acc[k][e], r[u][k][e]);
    }
    for (; i < cnt; ++i) {
      const float *row = fbase + (size_t)__shfl_sync(0xffffffffu, mine, i) * C;
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        const int c = (k * 32 + lane) * VEC;
        if (VEC == 4) {
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c < C) t = ldg_stream_f4(reinterpret_cast<const float4 *>(row + c));
          acc[k][0] = __fadd_rn(acc[k][0], t.x); acc[k][1 % VEC] = __fadd_rn(acc[k][1 % VEC], t.y);
          acc[k][2 % VEC] = __fadd_rn(acc[k][2 % VEC], t.z); acc[k][3 % VEC] = __fadd_rn(acc[k][3 % VEC], t.w);
        } else if (c < C) {
          acc[k][0] = __fadd_rn(acc[k][0], ldg_stream_f1(row + c));
        }
      }
    }
  }
  float *o = out + ((size_t)frame * V + v) * C;
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int c = (k * 32 + lane) * VEC;
    if (c < C) {
      if (VEC == 4)
        stg_stream_f4(reinterpret_cast<float4 *>(o + c),
                      make_float4(acc[k][0], acc[k][1 % VEC], acc[k][2 % VEC], acc[k][3 % VEC]));
      else
        stg_stream_f1(o + c, acc[k][0]);
    }
  }
}

template <int VEC, int NCH>
void launch_reduce(int B, int N, int C, int V, const float *feat, const int *row_ptr,
                   const int *pay2, float *out, cudaStream_t s) {
  dim3 grid(ceil_div(V, 8), B);
  vp_reduce_kernel<VEC, NCH><<<grid, 256, 0, s>>>(N, C, V, feat, row_ptr, pay2, out);
}

// ---- backward: one warp per point, whole-row copy ------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256)
vp_backward_kernel(long long total_points, int C, int X, int Y, const float *__restrict__ g_cl,
                   long long sb, long long sy, long long sx, const int32_t *__restrict__ pos_memo,
                   float *__restrict__ grad_feat) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long p = warp; p < total_points; p += nwarps) {
    const int b = pos_memo[p * 3];
    float *dst = grad_feat + p * C;
    const bool kept = b != -1;  // voxel_pooling.py:60
    const float *src = nullptr;
    if (kept) {
      const int y = pos_memo[p * 3 + 1], x = pos_memo[p * 3 + 2];
      src = g_cl + b * sb + y * sy + x * sx;
    }
    for (int c = lane * VEC; c < C; c += 32 * VEC) {
      if (VEC == 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kept) v = *reinterpret_cast<const float4 *>(src + c);
        stg_stream_f4(reinterpret_cast<float4 *>(dst + c), v);
      } else {
        stg_stream_f1(dst + c, kept ? src[c] : 0.0f);
      }
    }
  }
}

}  // namespace
}  // namespace sgv3d

using namespace sgv3d;

extern "C" size_t sgv3d_voxel_pooling_workspace_bytes(int B, int N, int C, int X, int Y, int Z) {
  (void)C; (void)Z;
  if (B <= 0 || N <= 0 || X <= 0 || Y <= 0) return 0;
  return carve_vp(nullptr, B, N, X, Y).bytes;
}

extern "C" int sgv3d_voxel_pooling_forward(int B, int N, int C, int X, int Y, int Z,
                                           const int32_t *geom_xyz, const float *features,
                                           float *out, int32_t *pos_memo, void *workspace,
                                           size_t workspace_bytes, sgv3d_stream_t stream) {
  SGV3D_REQUIRE(B >= 0 && N >= 0 && C > 0 && X > 0 && Y > 0 && Z > 0, "voxel_pooling_forward: bad sizes");
  SGV3D_REQUIRE(B <= 65535, "voxel_pooling_forward: B > 65535");
  SGV3D_REQUIRE((long long)X * Y < (long long)sort::kMaxHighBins << sort::kLowBits,
                "voxel_pooling_forward: X*Y=%lld exceeds %d voxels per frame", (long long)X * Y,
                sort::kMaxHighBins << sort::kLowBits);
  SGV3D_REQUIRE(C <= 1024, "voxel_pooling_forward: C=%d > 1024 unsupported", C);
  SGV3D_REQUIRE(out != nullptr, "voxel_pooling_forward: out is null");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  prof_begin(s);
  if (B == 0) return SGV3D_OK;
  const int V = X * Y;
  if (N == 0) {
    SGV3D_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)B * V * C, s));
    return SGV3D_OK;
  }
  SGV3D_REQUIRE(geom_xyz && features, "voxel_pooling_forward: null input");
  VpWorkspace w = carve_vp(workspace, B, N, X, Y);
  if (!workspace || workspace_bytes < w.bytes) {
    set_error("voxel_pooling_forward: workspace %zu < required %zu bytes", workspace_bytes, w.bytes);
    return SGV3D_ERR_WORKSPACE_TOO_SMALL;
  }
  dim3 gb(w.nblk, B);
  vp_prepare_kernel<<<gb, sort::kThreads, 0, s>>>(N, X, Y, Z, geom_xyz, pos_memo, w.keys0, w.nblk, w.hist1);
  SGV3D_CHECK_LAUNCH("vp_prepare_kernel");
  sort::scan_hist_kernel<<<B, sort::kScanThreads, 0, s>>>(w.hist1, sort::kLowBins, w.nblk, nullptr,
                                                          sort::kItemsPerBlock, nullptr);
  SGV3D_CHECK_LAUNCH("scan_hist_kernel(1)");
  sort::scatter_contiguous_kernel<0, sort::kLowBins - 1, sort::PlaceKeyPayloadFactory>
      <<<gb, sort::kThreads, sizeof(int) * sort::kWarps * sort::kLowBins, s>>>(
          w.keys0, nullptr, (size_t)N, nullptr, N, sort::kLowBins, w.hist1, w.nblk,
          sort::PlaceKeyPayloadFactory{w.keys1, w.pay1, (size_t)N});
  SGV3D_CHECK_LAUNCH("scatter_contiguous_kernel(1)");
  sort::hist_contiguous_kernel<sort::kLowBits><<<gb, sort::kThreads, sizeof(int) * w.bins2, s>>>(
      w.keys1, (size_t)N, nullptr, N, w.bins2, w.nblk, w.hist2);
  SGV3D_CHECK_LAUNCH("hist_contiguous_kernel");
  sort::scan_hist_kernel<<<B, sort::kScanThreads, 0, s>>>(w.hist2, w.bins2, w.nblk, nullptr,
                                                          sort::kItemsPerBlock, nullptr);
  SGV3D_CHECK_LAUNCH("scan_hist_kernel(2)");
  sort::scatter_contiguous_kernel<sort::kLowBits, 0xFFFFFF, sort::PlaceKeyPayloadFactory>
      <<<gb, sort::kThreads, sizeof(int) * sort::kWarps * w.bins2, s>>>(
          w.keys1, w.pay1, (size_t)N, nullptr, N, w.bins2, w.hist2, w.nblk,
          sort::PlaceKeyPayloadFactory{w.keys2, w.pay2, (size_t)N});
  SGV3D_CHECK_LAUNCH("scatter_contiguous_kernel(2)");
  sort::row_ptr_kernel<<<dim3(ceil_div(N, 256) < 512 ? ceil_div(N, 256) : 512, B), 256, 0, s>>>(
      w.keys2, (size_t)N, nullptr, N, V, w.row_ptr);
  SGV3D_CHECK_LAUNCH("row_ptr_kernel");
  const bool vec = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(features) | reinterpret_cast<uintptr_t>(out)) % 16 == 0);
  if (vec) {
    const int nch = ceil_div(C, 128);
    switch (nch) {
      case 1: launch_reduce<4, 1>(B, N, C, V, features, w.row_ptr, w.pay2, out, s); break;
      case 2: launch_reduce<4, 2>(B, N, C, V, features, w.row_ptr, w.pay2, out, s); break;
      case 3: launch_reduce<4, 3>(B, N, C, V, features, w.row_ptr, w.pay2, out, s); break;
      case 4: launch_reduce<4, 4>(B, N, C, V, features, w.row_ptr, w.pay2, out, s); break;
      default: launch_reduce<4, 8>(B, N, C, V, features, w.row_ptr, w.pay2, out, s); break;
    }
  } else {
    const int nch = ceil_div(C, 32);
    if (nch <= 1) launch_reduce<1, 1>(B, N, C, V, features, w.row_ptr, w.pay2, out, s);
    else if (nch <= 2) launch_reduce<1, 2>(B, N, C, V, features, w.row_ptr, w.pay2, out, s);
    else if (nch <= 3) launch_reduce<1, 3>(B, N, C, V, features, w.row_ptr, w.pay2, out, s);
    else if (nch <= 4) launch_reduce<1, 4>(B, N, C, V, features, w.row_ptr, w.pay2, out, s);
    else if (nch <= 8) launch_reduce<1, 8>(B, N, C, V, features, w.row_ptr, w.pay2, out, s);
    else if (nch <= 16) launch_reduce<1, 16>(B, N, C, V, features, w.row_ptr, w.pay2, out, s);
    else launch_reduce<1, 32>(B, N, C, V, features, w.row_ptr, w.pay2, out, s);
  }
  SGV3D_CHECK_LAUNCH("vp_reduce_kernel");
  return SGV3D_OK;
}

extern "C" size_t sgv3d_voxel_pooling_backward_workspace_bytes(int B, int C, int X, int Y) {
  if (B <= 0 || C <= 0 || X <= 0 || Y <= 0) return 0;
  return align_up(sizeof(float) * (size_t)B * X * Y * C, 256);
}

extern "C" int sgv3d_voxel_pooling_backward(int B, int N, int C, int X, int Y, const float *grad_out,
                                            int64_t sb, int64_t sc, int64_t sy, int64_t sx,
                                            const int32_t *pos_memo, float *grad_features,
                                            void *workspace, size_t workspace_bytes,
                                            sgv3d_stream_t stream) {
  SGV3D_REQUIRE(B >= 0 && N >= 0 && C > 0 && X > 0 && Y > 0, "voxel_pooling_backward: bad sizes");
  if (B == 0 || N == 0) return SGV3D_OK;
  SGV3D_REQUIRE(grad_out && pos_memo && grad_features, "voxel_pooling_backward: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  prof_begin(s);
  const float *g_cl = grad_out;
  long long lsb = sb, lsy = sy, lsx = sx;
  if (sc != 1) {
    // planar (B,C,Y,X) contiguous gradient -> channels-last copy in the workspace
    SGV3D_REQUIRE(sx == 1 && sy == X && sc == (int64_t)X * Y && sb == (int64_t)C * X * Y,
                  "voxel_pooling_backward: grad_out must be (B,C,Y,X)-contiguous or channels-last");
    const size_t need = sgv3d_voxel_pooling_backward_workspace_bytes(B, C, X, Y);
    if (!workspace || workspace_bytes < need) {
      set_error("voxel_pooling_backward: workspace %zu < required %zu bytes", workspace_bytes, need);
      return SGV3D_ERR_WORKSPACE_TOO_SMALL;
    }
    float *t = static_cast<float *>(workspace);
    const int V = X * Y;
    launch_transpose_pad<float, float>(grad_out, t, B, C, V, V, (size_t)C * V, C, (size_t)V * C, s);
    SGV3D_CHECK_LAUNCH("transpose_pad_kernel");
    g_cl = t;
    lsb = (long long)V * C; lsy = (long long)X * C; lsx = C;
  }
  const long long total = (long long)B * N;
  const int blocks = (int)((total + 7) / 8 < (long long)kNumSMs * 16 ? (total + 7) / 8 : (long long)kNumSMs * 16);
  const bool vec = (C % 4 == 0) && (lsb % 4 == 0) && (lsy % 4 == 0) && (lsx % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(g_cl) | reinterpret_cast<uintptr_t>(grad_features)) % 16 == 0);
  if (vec)
    vp_backward_kernel<4><<<blocks, 256, 0, s>>>(total, C, X, Y, g_cl, lsb, lsy, lsx, pos_memo, grad_features);
  else
    vp_backward_kernel<1><<<blocks, 256, 0, s>>>(total, C, X, Y, g_cl, lsb, lsy, lsx, pos_memo, grad_features);
  SGV3D_CHECK_LAUNCH("vp_backward_kernel");
  return SGV3D_OK;
}
