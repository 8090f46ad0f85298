// Stable per-frame LSD radix sort by voxel id (two passes: low 8 bits, then the remaining high
// bits), hand-written for sm_100a.  Determinism is the point: ties keep the canonical order in
// which the producer emitted them, so the per-voxel summation order downstream is fixed and no
// floating-point atomics are needed (north star: "pre-sorted by voxel rank ... deterministic").
//
// Building blocks (all integer work, HBM/L2-bound):
//   * producers write a per-block digit histogram  hist[frame][bin][block]   (smem int atomics)
//   * scan_hist_kernel      exclusive scan of that table in bin-major order, one CTA per frame
//   * radix_scatter_kernel  stable scatter: warp-private running counters + __match_any_sync
//                           ranking; canonical order inside a block is (warp, iteration, lane)
#pragma once

#include "common.cuh"

namespace sgv3d {
namespace sort {

constexpr int kLowBits = 8;
constexpr int kLowBins = 1 << kLowBits;  // 256
constexpr int kThreads = 256;            // threads per sort block (8 warps)
constexpr int kWarps = kThreads / kWarp;
constexpr int kItemsPerBlock = 2048;     // contiguous items per sort block (256 per warp)
constexpr int kMaxHighBins = 1024;       // (V >> 8) + 1 <= 1024  =>  V <= 261888 voxels per frame
constexpr int kScanThreads = 1024;

// ---------------------------------------------------------------------------------------------
// Exclusive scan of hist[frame][bin][blk] over (bin-major, blk-minor) for blk < nblk_eff.
// nblk_eff = nblk_max when n_items == nullptr, else ceil(n_items[frame] / items_per_block).
// Writes the exclusive prefix in place and the grand total to total_out[frame] (if non-null).
// ---------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(kScanThreads)
scan_hist_kernel(int *__restrict__ hist, int bins, int nblk_max, const int *__restrict__ n_items,
                 int items_per_block, int *__restrict__ total_out) {
  // One CTA per frame, row-wise: a bin's per-block counts are contiguous (hist[bin][blk]), so a warp reads a whole
  // row with coalesced loads.  1. row totals (warp w owns the rows w, w + 32, ...), 2. exclusive scan of the bin
  // totals (bins <= 1024: one value per thread), 3. every row rewritten with its exclusive prefixes, 32 blocks per
  // step through a warp scan.  (Walking the bin-major sequence with one CTA cost 57 dependent steps, walking it with
  // a contiguous range per thread 175 k uncoalesced sector requests from one SM: ~55 us per frame either way.)
  __shared__ int row_tot[kMaxHighBins];
  __shared__ int warp_tot[kScanThreads / kWarp];
  const int frame = blockIdx.x;
  int nblk = nblk_max;
  if (n_items) nblk = (n_items[frame] + items_per_block - 1) / items_per_block;
  int *h = hist + (size_t)frame * bins * nblk_max;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  constexpr int kWarps_ = kScanThreads / kWarp;
  for (int row = wid; row < bins; row += kWarps_) {
    const int *r = h + (size_t)row * nblk_max;
    int sum = 0;
    int j = lane;
    for (; j + 96 < nblk; j += 128) sum += (r[j] + r[j + 32]) + (r[j + 64] + r[j + 96]);
    for (; j < nblk; j += 32) sum += r[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) row_tot[row] = sum;
  }
  __syncthreads();
  // exclusive scan of the bin totals
  const int mine = t < bins ? row_tot[t] : 0;
  int x = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_tot[wid] = x;
  __syncthreads();
  if (wid == 0) {
    const int w = warp_tot[lane];
    int xs = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, xs, o);
      if (lane >= o) xs += y;
    }
    warp_tot[lane] = xs - w;
    if (lane == 31 && total_out) total_out[frame] = xs;
  }
  __syncthreads();
  if (t < bins) row_tot[t] = warp_tot[wid] + x - mine;   // first position of the bin
  __syncthreads();
  for (int row = wid; row < bins; row += kWarps_) {
    int *r = h + (size_t)row * nblk_max;
    int carry = row_tot[row];
    for (int j0 = 0; j0 < nblk; j0 += 128) {   // four 32-block steps per round trip
      int v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (j0 + 32 * u + lane < nblk) ? r[j0 + 32 * u + lane] : 0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int xx = v[u];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int y = __shfl_up_sync(0xffffffffu, xx, o);
          if (lane >= o) xx += y;
        }
        if (j0 + 32 * u + lane < nblk) r[j0 + 32 * u + lane] = carry + xx - v[u];
        carry += __shfl_sync(0xffffffffu, xx, 31);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Per-block digit histogram of a contiguous key array (used for the second pass, whose input
// order only exists after the first scatter).
// ---------------------------------------------------------------------------------------------
// Blocks are visited with a grid-stride loop (gridDim.x may be far smaller than nblk_max: the item
// count lives on the device, so the launch cannot be sized to it without a host sync).
template <int SHIFT>
__global__ void __launch_bounds__(kThreads)
hist_contiguous_kernel(const int *__restrict__ keys, size_t frame_stride,
                       const int *__restrict__ n_items, int n_fixed, int bins, int nblk_max,
                       int *__restrict__ hist) {
  extern __shared__ int s_hist[];
  const int frame = blockIdx.y;
  const int n = n_items ? n_items[frame] : n_fixed;
  const int *k = keys + (size_t)frame * frame_stride;
  int *h = hist + (size_t)frame * bins * nblk_max;
  for (int blk = blockIdx.x; blk * kItemsPerBlock < n; blk += gridDim.x) {
    const int begin = blk * kItemsPerBlock;
    for (int i = threadIdx.x; i < bins; i += kThreads) s_hist[i] = 0;
    __syncthreads();
    const int end = min(n, begin + kItemsPerBlock);
    int kk[kItemsPerBlock / kThreads];
#pragma unroll
    for (int u = 0; u < kItemsPerBlock / kThreads; ++u) {
      const int i = begin + u * kThreads + threadIdx.x;
      kk[u] = (i < end) ? k[i] : -1;
    }
#pragma unroll
    for (int u = 0; u < kItemsPerBlock / kThreads; ++u)
      if (kk[u] >= 0) atomicAdd(&s_hist[kk[u] >> SHIFT], 1);
    __syncthreads();
    for (int i = threadIdx.x; i < bins; i += kThreads) h[(size_t)i * nblk_max + blk] = s_hist[i];
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// Stable scatter of one sort block.
//   Input accessor `In` yields, for warp w / iteration it / lane, the item (key, payload, valid)
//   in canonical order; n_iters(warp) iterations per warp.
//   gbase = scanned histogram (exclusive, bin-major) for this frame.
// Phase A: per-warp digit counts (smem int atomics: counts are order independent).
// Phase B: cross-warp exclusive scan per digit, then match-any ranking with warp-private running
//          counters => position is a pure function of the canonical order.
// ---------------------------------------------------------------------------------------------
// Default placement: write (key, payload) at the item's sorted position.
struct PlaceKeyPayload {
  int *keys, *payload;
  __device__ __forceinline__ void operator()(int pos, int key, int pay) const {
    keys[pos] = key;
    payload[pos] = pay;
  }
};

// base_of(d) = global position of the block's first item with digit d (exclusive scan of the
// per-block digit histograms).  BinMajorBase reads it from a scan_hist_kernel table.
struct BinMajorBase {
  const int *gbase_frame;
  int nblk_max, blk;
  __device__ __forceinline__ int operator()(int d) const { return gbase_frame[(size_t)d * nblk_max + blk]; }
};

template <int NWARPS, typename In, typename DigitFn, typename BaseOf, typename Place>
__device__ __forceinline__ void stable_scatter_block(const In &in, DigitFn digit_of, int bins,
                                                     const BaseOf &base_of, int *s_cnt /*[NWARPS][bins]*/,
                                                     const Place &place) {
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  for (int i = t; i < NWARPS * bins; i += NWARPS * kWarp) s_cnt[i] = 0;
  __syncthreads();
  int *my = s_cnt + wid * bins;
  const int iters = in.iters(wid);
  for (int it = 0; it < iters; it += 4) {  // 4 independent loads in flight per lane
    int key[4], pay[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) ok[u] = (it + u < iters) && in.load(wid, it + u, lane, key[u], pay[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (ok[u]) atomicAdd(&my[digit_of(key[u])], 1);
  }
  __syncthreads();
  for (int d = t; d < bins; d += NWARPS * kWarp) {
    int base = base_of(d);
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) {
      const int c = s_cnt[w * bins + d];
      s_cnt[w * bins + d] = base;
      base += c;
    }
  }
  __syncthreads();
  const unsigned lt = lanemask_lt();
  for (int it0 = 0; it0 < iters; it0 += 4) {
    int key4[4], pay4[4];
    bool ok4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      key4[u] = 0; pay4[u] = 0;
      ok4[u] = (it0 + u < iters) && in.load(wid, it0 + u, lane, key4[u], pay4[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (it0 + u >= iters) break;  // warp-uniform
      const bool valid = ok4[u];
      const int key = key4[u], pay = pay4[u];
      // invalid lanes get a private pseudo-digit so that they never match a real one
      const int dig = valid ? digit_of(key) : (bins + lane);
      const unsigned peers = __match_any_sync(0xffffffffu, dig);
      const int leader = __ffs(peers) - 1;
      const int rank = __popc(peers & lt);
      int base = 0;
      if (valid && lane == leader) {
        base = my[dig];
        my[dig] = base + __popc(peers);
      }
      base = __shfl_sync(0xffffffffu, base, leader);
      if (valid) place(base + rank, key, pay);
      __syncwarp();
    }
  }
}

// Contiguous input: block `blk` owns items [blk*2048, min(n, (blk+1)*2048)); warp w owns the
// w-th run of 256 items; payload is either explicit or the item index itself.
struct ContiguousInput {
  const int *keys;
  const int *payload;  // nullptr => payload = item index
  int begin, n;
  __device__ __forceinline__ int iters(int) const { return kItemsPerBlock / kWarps / kWarp; }
  __device__ __forceinline__ bool load(int w, int it, int lane, int &key, int &pay) const {
    const int i = begin + w * (kItemsPerBlock / kWarps) + it * kWarp + lane;
    if (i >= n) return false;
    key = keys[i];
    pay = payload ? payload[i] : i;
    return true;
  }
};

template <int SHIFT, int MASK>
struct DigitOf {
  __device__ __forceinline__ int operator()(int key) const { return (key >> SHIFT) & MASK; }
};

// PlaceFactory(frame) -> Place functor for that frame.
template <int SHIFT, int MASK, typename PlaceFactory>
__global__ void __launch_bounds__(kThreads)
scatter_contiguous_kernel(const int *__restrict__ keys_in, const int *__restrict__ payload_in,
                          size_t frame_stride_in, const int *__restrict__ n_items, int n_fixed,
                          int bins, const int *__restrict__ gbase, int nblk_max, PlaceFactory make_place) {
  extern __shared__ int s_cnt[];
  const int frame = blockIdx.y;
  const int n = n_items ? n_items[frame] : n_fixed;
  const auto place = make_place(frame);
  for (int blk = blockIdx.x; blk * kItemsPerBlock < n; blk += gridDim.x) {
    ContiguousInput in{keys_in + (size_t)frame * frame_stride_in,
                       payload_in ? payload_in + (size_t)frame * frame_stride_in : nullptr,
                       blk * kItemsPerBlock, n};
    stable_scatter_block<kWarps>(in, DigitOf<SHIFT, MASK>(), bins,
                                 BinMajorBase{gbase + (size_t)frame * bins * nblk_max, nblk_max, blk}, s_cnt, place);
    __syncthreads();
  }
}

struct PlaceKeyPayloadFactory {
  int *keys, *payload;
  size_t frame_stride;
  __device__ __forceinline__ PlaceKeyPayload operator()(int frame) const {
    return PlaceKeyPayload{keys + (size_t)frame * frame_stride, payload + (size_t)frame * frame_stride};
  }
};

// ---------------------------------------------------------------------------------------------
// row_ptr[v] = first sorted position whose key >= v, for v in [0, V]; keys sorted ascending.
__global__ void __launch_bounds__(256) static
row_ptr_kernel(const int *__restrict__ keys, size_t frame_stride, const int *__restrict__ n_items,
               int n_fixed, int V, int *__restrict__ row_ptr /*[frame][V+1]*/) {
  const int frame = blockIdx.y;
  const int n = n_items ? n_items[frame] : n_fixed;
  const int *k = keys + (size_t)frame * frame_stride;
  int *rp = row_ptr + (size_t)frame * (V + 1);
  const int stride = gridDim.x * blockDim.x;
  const int j0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (n == 0) {
    for (int v = j0; v <= V; v += stride) rp[v] = 0;
    return;
  }
  for (int j = j0; j < n; j += stride) {
    const int key = min(k[j], V);
    const int prev = (j > 0) ? min(k[j - 1], V) : -1;
    for (int v = prev + 1; v <= key; ++v) rp[v] = j;
    if (j == n - 1)
      for (int v = key + 1; v <= V; ++v) rp[v] = n;
  }
}

}  // namespace sort
}  // namespace sgv3d
