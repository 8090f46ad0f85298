// Fused lift-splat for B200 (sm_100a): the B x Nc x D x fH x fW x C frustum tensor of the reference
// (layers/backbones/lss_fpn.py:464-466,486,490) is never materialised.
//
// Key identity: along the height-bin axis D consecutive bins of one pixel fall into the same
// voxel in runs, and  sum_{d in run} p_d * ctx_c = (sum_{d in run} p_d) * ctx_c.  So the scatter
//   BEV[b,c,voxel] += height[d,pixel] * ctx[c,pixel]          (N*C multiply-adds per frame)
// becomes a sparse (voxel x pixel) weighted gather of whole context rows with
// nnz = #runs (3-7x fewer than points; SURVEY.md §7 hard part 3).
//
//   PLAN   (index; depends on calibration + grid only)
//     ls_plan_runs_*_kernel   thread/pixel walks D: bit-exact geometry -> voxel id -> run-length
//                             encoding, emitted pixel-major in an ELL layout; per-chunk histogram of the
//                             runs over 64-voxel tiles
//     ls_scatter_tiles_kernel stable scatter of the runs into per-tile buckets (MSD pass: one digit = the tile);
//                             for small grids the scan of the histograms is done here, per CTA
//                             (ls_scan_tiles_kernel otherwise)
//     ls_finish_tiles_kernel  per tile: stable counting sort of its bucket by voxel-in-tile, sorted entries,
//                             inverse permutation (run -> sorted slot)
//     Ties keep the canonical (chunk, warp, run index, lane) order => the per-voxel summation order is a
//     pure function of the calibration: deterministic, no floating-point atomics.
//   FORWARD (values)
//     ls_lift_prep_kernel     two kinds of CTA in one launch: [softmax over D fused] w[run] = sum_{d in run}
//                             p[d,pixel] from cp.async-staged height columns, written to the run's sorted entry;
//                             context NCHW -> one channels-last row of whole 128-byte lines per pixel
//     ls_reduce_kernel        CTA = 64-voxel tile, entry list cut into equal slices per 8-lane stream, rows
//                             gathered as full 128-byte lines, sums in a fixed order, NCHW tile written once
//   BACKWARD (values; pixel-major, no sort needed)
//     ls_grad_rows_kernel     grad_bev NCHW -> one row per voxel (touched tiles only)
//     ls_backward_chunk_kernel per 64-pixel chunk: softmax, run weights, g_ctx_row = sum_r w_r*G[voxel_r],
//                             gw_r = <ctx_row, G[voxel_r]>, softmax backward, g_height / g_ctx NCHW writes
//     (C > 96: ls_lift_prep_kernel, transpose_pad_kernel, ls_backward_gather_kernel, ls_expand_kernel)
//
// No floating-point atomics anywhere; every sum has a fixed order => bitwise reproducible.
#include <cuda.h>
#include <cuda_bf16.h>

#include <stdlib.h>

#include <algorithm>

#include "geometry.cuh"
#include "ls_block.cuh"
#include "ls_shared.cuh"
#include "sort.cuh"
#include "transpose.cuh"

namespace sgv3d {
namespace {


// Sorted (voxel-major) entry: frame-local pixel row (n*P + p) << 6 | the voxel's index inside its 64-voxel
// reduce tile (written by the plan), and the run weight (written by the forward weights pass through
// run_dst, the plan's ELL slot -> sorted position map).
struct __align__(8) Entry {
  unsigned off;
  float w;
};

// Run in its tile bucket (after the MSD pass, before the per-tile finish).
struct __align__(8) BucketEnt {
  unsigned key;  // frame-local pixel row (n*P + p) << 6 | voxel index inside the tile
  int slot;      // frame-local ELL slot of the run
};

struct Workspace {
  int *run_cnt, *run_vox, *run_d, *run_dst, *hist, *tile_ptr, *chunk_done;
  BucketEnt *bucket;
  Entry *vm_ent;  // sorted (row offset, weight) pairs
  float *w_pm, *gw_pm, *gT, *gctxT;
  void *ctxT;
  size_t bytes;
};

// (channels-last BEV map: the context rows keep the natural channel order -- the reduce writes them out as they are)
RowPerm row_perm(const Dims &m) { return RowPerm{m.cl ? 0 : (m.G == 4 ? 2 : (m.G == 8 ? 3 : 4)), m.Cpad}; }

Workspace carve(void *ws, const Dims &m, int ctx_dtype) {
  Workspace w;
  Carver c(ws);
  const size_t B = m.B, slots = (size_t)m.B * m.cap;
  w.chunk_done = c.take<int>(B * m.nchunks);
  w.run_cnt = c.take<int>(B * m.nchunks * kChunk);
  w.run_vox = c.take<int>(slots);
  w.run_d = c.take<int>(slots);
  w.run_dst = c.take<int>(slots);
  w.hist = c.take<int>(B * m.nchunks * m.ntiles);
  w.tile_ptr = c.take<int>(B * (m.ntiles + 1));
  w.bucket = reinterpret_cast<BucketEnt *>(c.take<int2>(slots));
  w.vm_ent = reinterpret_cast<Entry *>(c.take<int2>(slots));
  w.w_pm = c.take<float>(slots);
  w.gw_pm = c.take<float>(slots);
  const size_t rows = B * m.Nc * m.P;
  if (ctx_dtype == SGV3D_DTYPE_BF16) w.ctxT = c.take<__nv_bfloat16>(rows * m.Cpad);
  else w.ctxT = c.take<float>(rows * m.Cpad);
  const int gpad = m.Cpad;
  w.gT = c.take<float>(B * m.V * gpad);
  w.gctxT = c.take<float>(rows * gpad);
  w.bytes = c.used();
  return w;
}

__device__ __forceinline__ size_t ell_slot(int frame_chunk, int D, int r, int t) {
  return ((size_t)frame_chunk * D + r) * kChunk + t;
}

// ---------------------------------------------------------------------------------------------
// PLAN 1/4: geometry -> voxel id per height bin -> runs.  grid (nchunks, B), 128 threads.
// Pure ALU work (no global reads besides three tiny tables); two bins per iteration for ILP.
// ---------------------------------------------------------------------------------------------
template <int ARITH>
__global__ void __launch_bounds__(kChunk, 5)
ls_plan_runs_kernel(Dims m, const float *__restrict__ u_tab, const float *__restrict__ v_tab,
                    const float *__restrict__ z_tab, const float *__restrict__ ida_inv,
                    const float *__restrict__ mv, const float *__restrict__ me,
                    const float *__restrict__ bda, const float *__restrict__ ref_h, geom::Grid grid,
                    int *__restrict__ run_cnt, int *__restrict__ run_vox, int *__restrict__ run_d,
                    int *__restrict__ hist, const int *__restrict__ chunk_done) {
  __shared__ geom::Camera cam;
  extern __shared__ float z_s[];                           // [D] height-bin values, then
  int *s_hist = reinterpret_cast<int *>(z_s + m.D);        // [ntiles] runs per reduce tile
  // Persistent grid over the (frame, chunk) list: normally the fast kernel has produced every chunk, so the
  // whole launch is one bulk look at the flags (a full-size grid of early exits costs ~6 us per step).
  const int total = m.nchunks * m.B;
  {
    int todo = 0;
    for (int i = blockIdx.x + (int)threadIdx.x * (int)gridDim.x; i < total; i += (int)gridDim.x * kChunk)
      todo |= (chunk_done[i] == 0);
    if (!__syncthreads_or(todo)) return;
  }
  for (int fc = blockIdx.x; fc < total; fc += gridDim.x) {
  if (chunk_done[fc]) continue;  // the fast kernel already produced this chunk (block-uniform)
  const int b = fc / m.nchunks, chunk = fc - b * m.nchunks;
  const int n = chunk / m.cpc, ci = chunk - n * m.cpc;
  const int bn = b * m.Nc + n;
  const int t = threadIdx.x;
  geom::load_camera(&cam, ida_inv, mv, me, bda, ref_h, bn, b);
  for (int d = t; d < m.D; d += kChunk) z_s[d] = z_tab[d];
  for (int i = t; i < m.ntiles; i += kChunk) s_hist[i] = 0;
  __syncthreads();
  const int frame_chunk = b * m.nchunks + chunk;
  const int p = ci * kChunk + t;
  int r = 0;
  if (p < m.P) {
    const int h = p / m.fW, w = p - h * m.fW;
    geom::PixelRay<ARITH> ray;
    ray.init(cam, u_tab[w], v_tab[h]);
    int cur = -1, d0 = 0;
    auto step = [&](int d, int vox) {
      if (vox != cur) {
        if (cur >= 0) {
          const size_t s = ell_slot(frame_chunk, m.D, r, t);
          run_vox[s] = cur;
          run_d[s] = d0 | (d << 16);
          atomicAdd(&s_hist[cur >> 6], 1);
          ++r;
        }
        cur = vox;
        d0 = d;
      }
    };
    int d = 0;
    for (; d + 2 <= m.D; d += 2) {
      const int v0 = ray.voxel(cam, grid, z_s[d]);
      const int v1 = ray.voxel(cam, grid, z_s[d + 1]);
      step(d, v0);
      step(d + 1, v1);
    }
    if (d < m.D) {
      step(d, ray.voxel(cam, grid, z_s[d]));
      ++d;
    }
    step(m.D, -2);  // sentinel closes the last run
  }
  run_cnt[(size_t)frame_chunk * kChunk + t] = r;
  __syncthreads();
  int *hh = hist + (size_t)frame_chunk * m.ntiles;
  for (int i = t; i < m.ntiles; i += kChunk) hh[i] = s_hist[i];
  __syncthreads();  // cam / z_s / s_hist are rewritten by the next chunk
  }
}

// ---------------------------------------------------------------------------------------------
// PLAN 1/4, fast variant (geometry.cuh: FastRay): same outputs as ls_plan_runs_kernel for pixel chunks
// whose camera / pixels qualify; chunk_done[frame_chunk] tells the general kernel to skip them.
// ~half the instructions per point and few enough registers for every CTA of a batch to be resident
// at once (one wave).  grid (nchunks, B), 128 threads.
// ---------------------------------------------------------------------------------------------
template <int ARITH>
__global__ void __launch_bounds__(kChunk, 9)
ls_plan_runs_fast_kernel(Dims m, const float *__restrict__ u_tab, const float *__restrict__ v_tab,
                         const float *__restrict__ z_tab, const float *__restrict__ ida_inv,
                         const float *__restrict__ mv, const float *__restrict__ me,
                         const float *__restrict__ bda, const float *__restrict__ ref_h, geom::Grid grid,
                         int *__restrict__ run_cnt, int *__restrict__ run_vox, int *__restrict__ run_d,
                         int *__restrict__ hist, int *__restrict__ chunk_done) {
  __shared__ geom::Camera cam;
  __shared__ int s_fast, s_zmin, s_zmax;  // z range of the bins as order-preserving ints
  extern __shared__ float z_s[];
  float *h_s = z_s + m.D;
  int *s_hist = reinterpret_cast<int *>(z_s + 2 * m.D);
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int n = chunk / m.cpc, ci = chunk - n * m.cpc;
  const int bn = b * m.Nc + n;
  const int t = threadIdx.x;
  const int frame_chunk = b * m.nchunks + chunk;
  geom::load_camera(&cam, ida_inv, mv, me, bda, ref_h, bn, b);
  if (t == 0) { s_zmin = 0x7fffffff; s_zmax = (int)0x80000000; }
  __syncthreads();
  bool z_ok = true;
  for (int d = t; d < m.D; d += kChunk) {
    const float z = z_tab[d];
    z_s[d] = z;
    z_ok = z_ok && (fabsf(z) < INFINITY);
    int zi = __float_as_int(z);
    zi ^= (zi >> 31) & 0x7fffffff;  // monotone float -> int map
    atomicMin(&s_zmin, zi);
    atomicMax(&s_zmax, zi);
  }
  for (int i = t; i < m.ntiles; i += kChunk) s_hist[i] = 0;
  if (!__syncthreads_and(z_ok)) {  // (also publishes cam / z_s / s_hist)
    if (t == 0) chunk_done[frame_chunk] = 0;
    return;
  }
  if (t == 0) s_fast = geom::camera_is_fast(cam) ? 1 : 0;
  __syncthreads();
  if (!s_fast) {
    if (t == 0) chunk_done[frame_chunk] = 0;
    return;
  }
  // per-camera table of the bins' heights above their planes (valid when row 2 of ida^-1 ignores u, v)
  const bool h_uniform = cam.A[8] == 0.0f && cam.A[9] == 0.0f;
  if (h_uniform) {
    for (int d = t; d < m.D; d += kChunk) {
      const float p0z = geom::dot2_tail<ARITH>(0.0f, cam.A + 8, z_s[d], 1.0f);
      h_s[d] = __fadd_rn(__fmul_rn(-1.0f, p0z), cam.ref_h);
    }
  }
  __syncthreads();
  const int p = ci * kChunk + t;
  int r = 0;
  bool bad = false;
  if (p < m.P) {
    const int h = p / m.fW, w = p - h * m.fW;
    int za = s_zmin, zb = s_zmax;
    za ^= (za >> 31) & 0x7fffffff;
    zb ^= (zb >> 31) & 0x7fffffff;
    int cur = -1, d0 = 0;
    // this chunk's block of the ELL table; run r of this pixel sits at element r * kChunk + t
    int *const rv = run_vox + ell_slot(frame_chunk, m.D, 0, 0);
    int *const rd = run_d + ell_slot(frame_chunk, m.D, 0, 0);
    int eo = t;
    auto step = [&](int d, int vox) {
      if (vox != cur) {
        if (cur >= 0) {
          rv[eo] = cur;
          rd[eo] = d0 | (d << 16);
          eo += kChunk;
          atomicAdd(&s_hist[cur >> 6], 1);
          ++r;
        }
        cur = vox;
        d0 = d;
      }
    };
    // guarded linear walk (geometry.cuh: walk_fast / LinearWalk): ~16 instructions per bin, the exact chain only inside
    // the guard band of a voxel boundary
    bad = !geom::walk_fast<ARITH>(cam, grid, z_s, h_s, h_uniform, m.D, u_tab[w], v_tab[h], __int_as_float(za),
                                  __int_as_float(zb), step);
    step(m.D, -2);  // sentinel closes the last run
  }
  run_cnt[(size_t)frame_chunk * kChunk + t] = r;
  const int any_bad = __syncthreads_or(bad);
  if (t == 0) chunk_done[frame_chunk] = any_bad ? 0 : 1;
  int *hh = hist + (size_t)frame_chunk * m.ntiles;
  for (int i = t; i < m.ntiles; i += kChunk) hh[i] = s_hist[i];
}

// ---------------------------------------------------------------------------------------------
// PLAN 2/4: per frame, turn the per-chunk tile histograms hist[chunk][tile] into scatter bases:
//   hist[chunk][tile] <- number of runs of this tile in chunks < chunk   (exclusive, in place)
//   tile_ptr[tile]    <- number of runs in tiles < tile;  tile_ptr[ntiles] = runs of the frame
// grid (B), 1024 threads.  Thread = tile for the chunk walk (coalesced rows, 8 loads in flight).
// ---------------------------------------------------------------------------------------------
constexpr int kScanThreads = 1024;

__global__ void __launch_bounds__(kScanThreads)
ls_scan_tiles_kernel(Dims m, int *__restrict__ hist, int *__restrict__ tile_ptr) {
  __shared__ int s_tot[kMaxTiles];
  __shared__ int s_part[kMaxTiles];  // [group][bin] partial sums of one sweep (groups * width <= 1024 per sweep)
  __shared__ int s_warp[kScanThreads / 32];
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, wid = t >> 5;
  int *h = hist + (size_t)b * m.nchunks * m.ntiles;
  // `width` threads cover the bins of one sweep; the 1024 / width thread groups split the chunk axis
  int width = 32;
  while (width < m.ntiles && width < kScanThreads) width <<= 1;
  const int groups = kScanThreads / width;
  const int g = t / width, lb = t - g * width;
  const int cpg = (m.nchunks + groups - 1) / groups;
  const int c_lo = min(g * cpg, m.nchunks), c_hi = min(c_lo + cpg, m.nchunks);
  for (int bin0 = 0; bin0 < kMaxTiles; bin0 += width) {
    if (bin0 >= m.ntiles) {  // (block-uniform) tiles beyond the grid hold no runs
      for (int i = bin0 + t; i < kMaxTiles; i += kScanThreads) s_tot[i] = 0;
      break;
    }
    const int bin = bin0 + lb;
    const bool live = bin < m.ntiles;
    // sweep 1: this group's share of the bin's runs
    int part = 0;
    if (live) {
      int c = c_lo;
      for (; c + 8 <= c_hi; c += 8) {
        int v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = h[(size_t)(c + u) * m.ntiles + bin];
#pragma unroll
        for (int u = 0; u < 8; ++u) part += v[u];
      }
      for (; c < c_hi; ++c) part += h[(size_t)c * m.ntiles + bin];
    }
    __syncthreads();  // s_part of the previous sweep has been consumed
    s_part[g * width + lb] = part;
    __syncthreads();
    // sweep 2: exclusive prefix over the chunks, starting from the groups before this one
    int run = 0;
    for (int gg = 0; gg < g; ++gg) run += s_part[gg * width + lb];
    if (g == 0) {
      int tot = 0;
      for (int gg = 0; gg < groups; ++gg) tot += s_part[gg * width + lb];
      s_tot[bin] = live ? tot : 0;
    }
    if (live) {
      int c = c_lo;
      for (; c + 8 <= c_hi; c += 8) {
        int v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = h[(size_t)(c + u) * m.ntiles + bin];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          h[(size_t)(c + u) * m.ntiles + bin] = run;
          run += v[u];
        }
      }
      for (; c < c_hi; ++c) {
        const int v = h[(size_t)c * m.ntiles + bin];
        h[(size_t)c * m.ntiles + bin] = run;
        run += v;
      }
    }
  }
  __syncthreads();
  // exclusive scan over the tiles: thread t owns tiles [4t, 4t + 4)
  int v[4], sum = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[k] = s_tot[4 * t + k];
    sum += v[k];
  }
  int x = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_warp[wid] = x;
  __syncthreads();
  if (wid == 0) {
    const int w = s_warp[lane];
    int xs = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, xs, o);
      if (lane >= o) xs += y;
    }
    s_warp[lane] = xs - w;
  }
  __syncthreads();
  int excl = s_warp[wid] + x - sum;
  int *tp = tile_ptr + (size_t)b * (m.ntiles + 1);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (4 * t + k <= m.ntiles) tp[4 * t + k] = excl;  // tiles >= ntiles hold 0 runs: entry ntiles = total
    excl += v[k];
  }
}

// ---------------------------------------------------------------------------------------------
// PLAN 3/4: stable scatter of the runs into per-tile buckets, straight out of the ELL layout.
// Block = chunk, warp w owns the pixels t = 32w + lane, iteration = run index r.
// ---------------------------------------------------------------------------------------------
// The chunk's run keys (voxel ids), staged in shared memory: keys[r][t] for r < kScatRows; runs beyond
// that (very long rays) are read from the ELL table directly.
constexpr int kScatRows = 32;
struct EllInput {
  const int *keys_s;   // shared: [kScatRows][kChunk]
  const int *run_vox;  // global ELL table
  int frame_chunk, D;
  int cnt;     // runs of this thread's pixel
  int warp_max;
  __device__ __forceinline__ int iters(int) const { return warp_max; }
  // key = voxel id of run `it` of this thread's pixel; payload = the run index (each thread loads and places
  // only its own pixel's runs, so everything else about the run is thread-local state of the placer)
  __device__ __forceinline__ bool load(int w, int it, int lane, int &key, int &pay) const {
    if (it >= cnt) return false;
    const int t = w * 32 + lane;
    key = it < kScatRows ? keys_s[it * kChunk + t] : run_vox[ell_slot(frame_chunk, D, it, t)];
    pay = it;
    return true;
  }
};

struct TileOf {
  __device__ __forceinline__ int operator()(int vox) const { return vox >> 6; }
};
struct TileBase {
  const int *tile_ptr, *hist_chunk;
  __device__ __forceinline__ int operator()(int tile) const { return tile_ptr[tile] + hist_chunk[tile]; }
};
struct PlaceBucket {
  BucketEnt *bucket;
  unsigned row6;   // this thread's frame-local pixel row (n*P + p) << 6
  int slot0;       // frame-local ELL slot of this pixel's run 0; run r sits kChunk further per run
  __device__ __forceinline__ void operator()(int pos, int vox, int run) const {
    BucketEnt e;
    e.key = row6 | (unsigned)(vox & 63);
    e.slot = slot0 + run * kChunk;
    bucket[pos] = e;
  }
};

struct SmemBase {
  const int *base;
  __device__ __forceinline__ int operator()(int tile) const { return base[tile]; }
};

// FUSED: the scan of ls_scan_tiles_kernel is done here, redundantly per CTA, straight from the raw per-chunk
// histograms (small grids: nchunks * ntiles ints stay L2-resident and a CTA reads them in ~1 us, less than the
// separate single-CTA-per-frame scan kernel and its launch cost).  The CTA of chunk 0 publishes tile_ptr.
constexpr int kFusedScanMaxTiles = 1024, kFusedScanMaxCells = 16384;

template <bool FUSED>
__global__ void __launch_bounds__(kChunk)
ls_scatter_tiles_kernel(Dims m, const int *__restrict__ run_cnt, const int *__restrict__ run_vox,
                        const int *__restrict__ hist, int *__restrict__ tile_ptr,
                        BucketEnt *__restrict__ bucket) {
  extern __shared__ int s_cnt[];  // [kChunk / 32][ntiles], then the staged keys [kScatRows][kChunk], then [ntiles] bases
  int *keys_s = s_cnt + (kChunk / 32) * m.ntiles;
  int *s_base = keys_s + kScatRows * kChunk;
  __shared__ int s_wtot[kChunk / 32];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int frame_chunk = b * m.nchunks + chunk;
  EllInput in;
  in.keys_s = keys_s;
  in.run_vox = run_vox;
  in.frame_chunk = frame_chunk;
  in.D = m.D;
  in.cnt = run_cnt[(size_t)frame_chunk * kChunk + threadIdx.x];
  in.warp_max = __reduce_max_sync(0xffffffffu, in.cnt);
  // every key of the chunk in flight at once (one round trip instead of one per batch of runs)
  const int nst = min(in.cnt, kScatRows);
  for (int r = 0; r < nst; ++r)
    cp_async_4(reinterpret_cast<float *>(keys_s + r * kChunk + threadIdx.x),
               reinterpret_cast<const float *>(run_vox + ell_slot(frame_chunk, m.D, r, threadIdx.x)));
  if (FUSED) {
    constexpr int kTpt = kFusedScanMaxTiles / kChunk;  // tiles per thread, at most
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int tpt = (m.ntiles + kChunk - 1) / kChunk;
    const int t0 = t * tpt;
    const int *h = hist + (size_t)b * m.nchunks * m.ntiles;
    int pre[kTpt], tot[kTpt];
#pragma unroll
    for (int k = 0; k < kTpt; ++k) { pre[k] = 0; tot[k] = 0; }
    auto sweep = [&](int c_lo, int c_hi, int (&acc)[kTpt]) {
      int c = c_lo;
      for (; c + 4 <= c_hi; c += 4) {
        int v[4][kTpt];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int k = 0; k < kTpt; ++k)
            v[u][k] = (k < tpt && t0 + k < m.ntiles) ? __ldg(h + (size_t)(c + u) * m.ntiles + t0 + k) : 0;
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int k = 0; k < kTpt; ++k) acc[k] += v[u][k];
      }
      for (; c < c_hi; ++c)
#pragma unroll
        for (int k = 0; k < kTpt; ++k)
          if (k < tpt && t0 + k < m.ntiles) acc[k] += __ldg(h + (size_t)c * m.ntiles + t0 + k);
    };
    sweep(0, chunk, pre);            // runs of each tile in the chunks before this one
    sweep(chunk, m.nchunks, tot);
    int sum = 0;
#pragma unroll
    for (int k = 0; k < kTpt; ++k) { tot[k] += pre[k]; sum += tot[k]; }
    int x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_wtot[wid] = x;
    __syncthreads();
    int excl = x - sum;
    for (int w = 0; w < wid; ++w) excl += s_wtot[w];
    int *tp = tile_ptr + (size_t)b * (m.ntiles + 1);
#pragma unroll
    for (int k = 0; k < kTpt; ++k) {
      if (k < tpt && t0 + k < m.ntiles) {
        s_base[t0 + k] = excl + pre[k];
        if (chunk == 0) tp[t0 + k] = excl;
      }
      excl += tot[k];
    }
    if (chunk == 0 && t == kChunk - 1) tp[m.ntiles] = excl;  // runs of the frame
  }
  cp_async_wait_all();  // each thread reads back only what it copied itself
  if (FUSED)
    sort::stable_scatter_block<kChunk / 32>(
        in, TileOf(), m.ntiles, SmemBase{s_base}, s_cnt,
        PlaceBucket{bucket + (size_t)b * m.cap,
                    (unsigned)((chunk / m.cpc) * m.P + (chunk % m.cpc) * kChunk + (int)threadIdx.x) << 6,
                    chunk * m.D * kChunk + (int)threadIdx.x});
  else
    sort::stable_scatter_block<kChunk / 32>(
        in, TileOf(), m.ntiles,
        TileBase{tile_ptr + (size_t)b * (m.ntiles + 1), hist + (size_t)frame_chunk * m.ntiles}, s_cnt,
        PlaceBucket{bucket + (size_t)b * m.cap,
                    (unsigned)((chunk / m.cpc) * m.P + (chunk % m.cpc) * kChunk + (int)threadIdx.x) << 6,
                    chunk * m.D * kChunk + (int)threadIdx.x});
}

// ---------------------------------------------------------------------------------------------
// PLAN 4/4: per tile, stable counting sort of the bucket by voxel-in-tile (6 bits): the sorted entries
// (pixel row | voxel-in-tile) and the inverse permutation run_dst (ELL slot -> sorted position) that the
// forward weights pass scatters through.
// ---------------------------------------------------------------------------------------------
template <int kFinWarps>
__global__ void __launch_bounds__(kFinWarps * 32)
ls_finish_tiles_kernel(Dims m, const int *__restrict__ tile_ptr, const BucketEnt *__restrict__ bucket,
                       Entry *__restrict__ vm_ent, int *__restrict__ run_dst) {
  __shared__ int s_cnt[kFinWarps][64];
  const int b = blockIdx.y, tile = blockIdx.x;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int *tp = tile_ptr + (size_t)b * (m.ntiles + 1);
  const int lo = tp[tile], hi = tp[tile + 1];
  if (hi == lo) return;
  const int n = hi - lo;
  const int per = (n + kFinWarps - 1) / kFinWarps;
  const int wb = lo + min(wid * per, n), we = lo + min((wid + 1) * per, n);
  const BucketEnt *bk = bucket + (size_t)b * m.cap;
  int *my = s_cnt[wid];
  // the warp's first kFinRegs * 32 entries stay in registers for both passes (all loads in flight at once)
  constexpr int kFinRegs = kFinWarps >= 16 ? 4 : 8;
  uint2 er[kFinRegs];
#pragma unroll
  for (int k = 0; k < kFinRegs; ++k) {
    const int i = wb + k * 32 + lane;
    er[k] = make_uint2(0u, 0u);
    if (i < we) er[k] = __ldg(reinterpret_cast<const uint2 *>(bk + i));
  }
  for (int i = t; i < kFinWarps * 64; i += kFinWarps * 32) (&s_cnt[0][0])[i] = 0;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kFinRegs; ++k)
    if (wb + k * 32 + lane < we) atomicAdd(&my[er[k].x & 63u], 1);
  for (int i = wb + kFinRegs * 32 + lane; i < we; i += 32) atomicAdd(&my[bk[i].key & 63u], 1);
  __syncthreads();
  if (wid == 0) {  // 64 voxels, two per lane: totals -> exclusive scan -> per-warp bases
    int c[2][kFinWarps], tot[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      tot[k] = 0;
#pragma unroll
      for (int w = 0; w < kFinWarps; ++w) {
        c[k][w] = s_cnt[w][2 * lane + k];
        tot[k] += c[k][w];
      }
    }
    const int sum = tot[0] + tot[1];
    int x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    int base = lo + x - sum;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
#pragma unroll
      for (int w = 0; w < kFinWarps; ++w) {
        s_cnt[w][2 * lane + k] = base;
        base += c[k][w];
      }
    }
  }
  __syncthreads();
  const unsigned lt = lanemask_lt();
  Entry *ve = vm_ent + (size_t)b * m.cap;
  int *rd = run_dst + (size_t)b * m.cap;
  auto place = [&](bool valid, unsigned key, int slot) {
    // invalid lanes get a private pseudo-digit so that they never match a real one
    const int dig = valid ? (int)(key & 63u) : 64 + lane;
    const unsigned peers = __match_any_sync(0xffffffffu, dig);
    const int leader = __ffs(peers) - 1;
    const int rank = __popc(peers & lt);
    int base = 0;
    if (valid && lane == leader) {
      base = my[dig];
      my[dig] = base + __popc(peers);
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    if (valid) {
      ve[base + rank].off = key;
      rd[slot] = base + rank;
    }
    __syncwarp();
  };
#pragma unroll
  for (int k = 0; k < kFinRegs; ++k) {
    if (wb + k * 32 >= we) break;  // warp-uniform
    place(wb + k * 32 + lane < we, er[k].x, (int)er[k].y);
  }
  for (int i0 = wb + kFinRegs * 32; i0 < we; i0 += 32) {
    const int i = i0 + lane;
    const bool valid = i < we;
    BucketEnt e;
    e.key = 0; e.slot = 0;
    if (valid) e = bk[i];
    place(valid, e.key, e.slot);
  }
}

// ---------------------------------------------------------------------------------------------
// Column staging: the D x 128 block of height values (or logits) of one pixel chunk goes to shared
// memory with cp.async -- every load of the CTA is in flight at once, rows are read coalesced and
// exactly once.  col[d * kChunk + t].
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_columns(float *col, const float *__restrict__ src /*camera base*/,
                                              int D, int P, int p0, bool vec16) {
  const int tid = threadIdx.x;
  const int npx = min(kChunk, P - p0);
  if (vec16 && npx == kChunk) {
    // bulk-async copies (TMA engine): one instruction per 512-byte row, completion on an mbarrier
    __shared__ __align__(8) unsigned long long s_bar;
    if (tid == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    bulk_stage_rows(col, kChunk * sizeof(float), src + p0, (size_t)P * sizeof(float), D, kChunk * sizeof(float), &s_bar);
    mbar_wait(&s_bar, 0);
    return;
  } else {
    const int t = tid & (kChunk - 1), q = tid / kChunk, nq = blockDim.x / kChunk;
    if (t < npx)
      for (int d = q; d < D; d += nq) cp_async_4(col + d * kChunk + t, src + (size_t)d * P + p0 + t);
  }
  cp_async_wait_all();
  __syncthreads();
}

// Softmax over D of this thread's staged column (torch.softmax within fp32 rounding:
// exp(x - max) / sum).  Leaves the un-normalised exponentials in col and returns 1 / sum, so the
// normalisation costs one multiply per run / per output instead of a pass over the column.
// Four independent max / sum chains (fixed interleaving => still a deterministic order).
__device__ __forceinline__ float softmax_column(float *col, int D, int t) {
  float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  int d = 0;
  for (; d + 4 <= D; d += 4) {
#pragma unroll
    for (int k = 0; k < 4; ++k) mx[k] = fmaxf(mx[k], col[(d + k) * kChunk + t]);
  }
  for (; d < D; ++d) mx[0] = fmaxf(mx[0], col[d * kChunk + t]);
  const float m = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
  float s[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  d = 0;
  for (; d + 4 <= D; d += 4) {
    float e[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) e[k] = exp_ex2(__fsub_rn(col[(d + k) * kChunk + t], m));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      col[(d + k) * kChunk + t] = e[k];
      s[k] = __fadd_rn(s[k], e[k]);
    }
  }
  for (; d < D; ++d) {
    const float e = exp_ex2(__fsub_rn(col[d * kChunk + t], m));
    col[d * kChunk + t] = e;
    s[0] = __fadd_rn(s[0], e);
  }
  return __fdiv_rn(1.0f, __fadd_rn(__fadd_rn(s[0], s[1]), __fadd_rn(s[2], s[3])));
}

// ---------------------------------------------------------------------------------------------
// FORWARD / BACKWARD: run weights  w = sum_{d in run} p[d, pixel]  (ascending d, fixed order), written
// pixel-major (ELL slot of the run; coalesced).  The forward reduce gathers them through the slot
// index its sorted entries carry.
// ---------------------------------------------------------------------------------------------
constexpr int kPrepThreads = 2 * kChunk;  // two threads per pixel column (halves of D, alternate runs)

// Partial softmax statistics of one half-column: max, then exp(x - m) in place and its sum.
// Four independent chains over d = lo + 4i + k (fixed interleaving => deterministic).
__device__ __forceinline__ float column_max(const float *col, int lo, int hi, int t) {
  float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  int d = lo;
  for (; d + 4 <= hi; d += 4) {
#pragma unroll
    for (int k = 0; k < 4; ++k) mx[k] = fmaxf(mx[k], col[(d + k) * kChunk + t]);
  }
  for (; d < hi; ++d) mx[0] = fmaxf(mx[0], col[d * kChunk + t]);
  return fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
}
__device__ __forceinline__ float column_exp_sum(float *col, int lo, int hi, int t, float m) {
  float s[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  int d = lo;
  for (; d + 4 <= hi; d += 4) {
    float e[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) e[k] = exp_ex2(__fsub_rn(col[(d + k) * kChunk + t], m));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      col[(d + k) * kChunk + t] = e[k];
      s[k] = __fadd_rn(s[k], e[k]);
    }
  }
  for (; d < hi; ++d) {
    const float e = exp_ex2(__fsub_rn(col[d * kChunk + t], m));
    col[d * kChunk + t] = e;
    s[0] = __fadd_rn(s[0], e);
  }
  return __fadd_rn(__fadd_rn(s[0], s[1]), __fadd_rn(s[2], s[3]));
}

// BSMLSSFPN context assembly (bsm_lss_fpn.py:524-529), fused into the context-rows pass when `sem` is set:
//   semantic = softmax(sem, channel axis);  rows = cat(context, semantic) * (1 - (semantic[0] > thr))
// The last Cs of the C row channels are the semantic probabilities; `context` holds the other C - Cs.
struct BsmAssembly {
  const float *sem;      // [B*Nc, Cs, fH, fW] semantic logits, or nullptr
  long long sem_stride;  // elements between consecutive cameras
  int Cs;
  float thr;
};


// vm_ent_out != nullptr: forward, the weight goes to the run's voxel-major entry (through run_dst);
// else backward (unfused path), pixel-major w_pm_out.  256 threads: thread (t, h) owns half h of pixel t's
// column for the softmax statistics and the runs r = h mod 2.
template <bool BSM>
__device__ __forceinline__ void weights_role(const Dims &m, const float *__restrict__ height, int vec16,
                                             const int *__restrict__ run_cnt, const int *__restrict__ run_d,
                                             const int *__restrict__ run_dst, float *__restrict__ w_pm_out,
                                             Entry *__restrict__ vm_ent_out, float *col, int b, int chunk,
                                             const BsmAssembly &bsm) {
  __shared__ float s_max[2][kChunk], s_sum[2][kChunk];
  const int n = chunk / m.cpc, ci = chunk - n * m.cpc;
  const int frame_chunk = b * m.nchunks + chunk;
  const int t = threadIdx.x & (kChunk - 1), h = threadIdx.x / kChunk;
  const int p0 = ci * kChunk;
  stage_columns(col, height + (size_t)(b * m.Nc + n) * m.hs, m.D, m.P, p0, vec16 != 0);
  const bool live = p0 + t < m.P;
  const int cnt = live ? run_cnt[(size_t)frame_chunk * kChunk + t] : 0;
  // BSMLSSFPN: a background pixel (semantic[0] > thr, bsm_lss_fpn.py:528-529) has an all-zero context row.  Its runs
  // get the weight -0.0f: the reduce recognises the bit pattern and does not gather the row (w * 0 = 0 either way;
  // real roadside frames are mostly background).  Same arithmetic as the context-rows role.
  bool masked = false;
  if (BSM && bsm.sem && live && vm_ent_out) {
    const float *ss = bsm.sem + (size_t)(b * m.Nc + n) * bsm.sem_stride + p0 + t;
    float mx = ss[0];
    for (int k = 1; k < bsm.Cs; ++k) mx = fmaxf(mx, ss[(size_t)k * m.P]);
    float sum = 0.0f;
    for (int k = 0; k < bsm.Cs; ++k) sum = __fadd_rn(sum, expf(__fsub_rn(ss[(size_t)k * m.P], mx)));
    masked = __fdiv_rn(expf(__fsub_rn(ss[0], mx)), sum) > bsm.thr;
  }
  // the first run descriptors are requested before the softmax so that their latency hides behind it
  constexpr int kPre = 6;
  int pre_d[kPre], pre_dst[kPre];
#pragma unroll
  for (int u = 0; u < kPre; ++u) {
    pre_d[u] = 0; pre_dst[u] = 0;
    const int r = h + 2 * u;
    if (r < cnt) {
      const size_t sl = ell_slot(frame_chunk, m.D, r, t);
      pre_d[u] = run_d[sl];
      if (vm_ent_out) pre_dst[u] = run_dst[sl];
    }
  }
  float scale = 1.0f;
  if (m.logits) {  // block-uniform
    const int half = (m.D + 1) >> 1;
    const int lo = h * half, hi = min(m.D, lo + half);
    const bool work = live && cnt > 0 && !(BSM && masked);   // (a background pixel's weights are -0.0f whatever its heights)
    s_max[h][t] = work ? column_max(col, lo, hi, t) : 0.0f;
    __syncthreads();
    const float mx = fmaxf(s_max[0][t], s_max[1][t]);
    s_sum[h][t] = work ? column_exp_sum(col, lo, hi, t, mx) : 1.0f;
    __syncthreads();  // both halves of every column now hold exp(x - max)
    scale = __fdiv_rn(1.0f, __fadd_rn(s_sum[0][t], s_sum[1][t]));
  }
  auto emit = [&](int r, int packed, int dst) {
    const int d0 = packed & 0xffff, d1 = packed >> 16;
    float acc = 0.0f;
    if (!(BSM && masked))
      for (int d = d0; d < d1; ++d) acc = __fadd_rn(acc, col[d * kChunk + t]);
    const float wgt = (BSM && masked) ? -0.0f : (m.logits ? __fmul_rn(acc, scale) : acc);
    if (vm_ent_out) vm_ent_out[(size_t)b * m.cap + dst].w = wgt;
    else w_pm_out[ell_slot(frame_chunk, m.D, r, t)] = wgt;
  };
#pragma unroll
  for (int u = 0; u < kPre; ++u)
    if (h + 2 * u < cnt) emit(h + 2 * u, pre_d[u], pre_dst[u]);
  // the rest in batches of 4 (strided descriptors)
  for (int r0 = h + 2 * kPre; r0 < cnt; r0 += 8) {
    int packed[4], dst[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      packed[u] = 0; dst[u] = 0;
      if (r0 + 2 * u < cnt) {
        const size_t sl = ell_slot(frame_chunk, m.D, r0 + 2 * u, t);
        packed[u] = run_d[sl];
        if (vm_ent_out) dst[u] = run_dst[sl];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (r0 + 2 * u < cnt) emit(r0 + 2 * u, packed[u], dst[u]);
  }
}

// Context rows: the C x 128 NCHW block of one pixel chunk -> 128 channels-last rows (permuted channel
// order, transpose.cuh: RowPerm).  Reads are coalesced 512-byte channel rows (cp.async, all in flight at
// once), writes are coalesced 128-byte pieces of the pixel rows; the smem tile is [C][129] (conflict-free
// both ways: consecutive pixels on the way in, 32 distinct channels of one pixel on the way out).
template <typename CT>
__device__ __forceinline__ void context_rows_role(const Dims &m, const CT *__restrict__ context,
                                                  CT *__restrict__ ctxT, RowPerm perm, float *smem, int b,
                                                  int chunk, const BsmAssembly &bsm) {
  constexpr int kLd = kChunk + 1;
  constexpr int kWarps = kPrepThreads / 32, kQ = kPrepThreads / kChunk;
  __shared__ float s_keep[kChunk];  // BSM: 0 for background pixels, else 1
  const int n = chunk / m.cpc, ci = chunk - n * m.cpc;
  const int t = threadIdx.x & (kChunk - 1), q = threadIdx.x / kChunk;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int p0 = ci * kChunk;
  const int npx = min(kChunk, m.P - p0);
  const int Cc = m.C - (bsm.sem ? bsm.Cs : 0);  // channels that come from `context`
  const CT *src = context + (size_t)(b * m.Nc + n) * m.cs + p0 + t + (size_t)q * m.P;
  if (t < npx) {
    float *sp = smem + q * kLd + t;
    if (sizeof(CT) == 4) {
#pragma unroll 4
      for (int c = q; c < Cc; c += kQ, sp += kQ * kLd, src += (size_t)kQ * m.P)
        cp_async_4(sp, reinterpret_cast<const float *>(src));
      if (bsm.sem) {
        const float *ss = bsm.sem + (size_t)(b * m.Nc + n) * bsm.sem_stride + p0 + t;
        for (int k = q; k < bsm.Cs; k += kQ) cp_async_4(smem + (Cc + k) * kLd + t, ss + (size_t)k * m.P);
      }
    } else {
#pragma unroll 4
      for (int c = q; c < Cc; c += kQ, sp += kQ * kLd, src += (size_t)kQ * m.P) *sp = to_f32<CT>(*src);
    }
  }
  cp_async_wait_all();
  __syncthreads();
  if (bsm.sem) {  // block-uniform
    if (q == 0 && t < npx) {
      // torch.softmax over the channel axis of an NCHW tensor (spatial softmax kernel: one thread per pixel,
      // sequential over the channels): max, sum += exp(x - max) in channel order, exp(x - max) / sum
      float *col = smem + Cc * kLd + t;
      float mx = col[0];
      for (int k = 1; k < bsm.Cs; ++k) mx = fmaxf(mx, col[k * kLd]);
      float sum = 0.0f;
      for (int k = 0; k < bsm.Cs; ++k) sum = __fadd_rn(sum, expf(__fsub_rn(col[k * kLd], mx)));
      for (int k = 0; k < bsm.Cs; ++k) col[k * kLd] = __fdiv_rn(expf(__fsub_rn(col[k * kLd], mx)), sum);
      s_keep[t] = col[0] > bsm.thr ? 0.0f : 1.0f;  // (1 - mask.int()) of bsm_lss_fpn.py:528-529
    }
    __syncthreads();
  }
  CT *dst = ctxT + ((size_t)(b * m.Nc + n) * m.P + p0) * m.Cpad;
  // lane <-> element of the row (3 pieces of 32 elements cover Cpad <= 96; loop for wider rows); everything
  // that depends on the element only is hoisted out of the pixel loop
  for (int e0 = 0; e0 < m.Cpad; e0 += 32) {
    const int e = e0 + lane;
    if (e >= m.Cpad) continue;
    const int c = perm.chan(e);
    CT *dp = dst + (size_t)wid * m.Cpad + e;
    const size_t dstep = (size_t)kWarps * m.Cpad;
    if (c < m.C) {
      const float *sp = smem + c * kLd + wid;
      if (bsm.sem) {
        for (int px = wid; px < npx; px += kWarps, sp += kWarps, dp += dstep)
          *dp = from_f32<CT>(__fmul_rn(*sp, s_keep[px]));
      } else {
#pragma unroll 4
        for (int px = wid; px < npx; px += kWarps, sp += kWarps, dp += dstep) *dp = from_f32<CT>(*sp);
      }
    } else {
      for (int px = wid; px < npx; px += kWarps, dp += dstep) *dp = from_f32<CT>(0.0f);
    }
  }
}

// The same with the TMA engine on the way in (fp32 context, LSSFPN call site): the chunk's four {32 pixels, C planes}
// boxes of the (P, C, B*Nc) view of `context` arrive with four 3-D tensor loads (SWIZZLE_128B, one mbarrier) instead of
// C * 128 four-byte cp.async copies; a thread then moves 16-byte chunks (4 pixels of one channel) into 4 rows.
__device__ __forceinline__ void context_rows_tma_role(const Dims &m, float *__restrict__ ctxT, RowPerm perm,
                                                      unsigned char *smem, int b, int chunk, const CUtensorMap *ctx_map) {
  __shared__ __align__(8) unsigned long long s_bar;
  const int n = chunk / m.cpc, ci = chunk - n * m.cpc;
  const int p0 = ci * kChunk;
  const int npx = min(kChunk, m.P - p0);
  // boxes on 1 KB boundaries: the swizzle uses shared-memory address bits 7-9 as the row index
  unsigned char *box0 = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
  const unsigned box_bytes = (((unsigned)m.C + 7u) >> 3) * 1024u;
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    mbar_arrive_expect_tx(&s_bar, (kChunk / 32) * (unsigned)m.C * 128u);   // (pixels past P: zero-filled and counted)
    const unsigned long long mp = reinterpret_cast<unsigned long long>(ctx_map);
    for (int k = 0; k < kChunk / 32; ++k)
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
              smem_u32(box0) + k * box_bytes),
          "l"(mp), "r"(p0 + 32 * k), "r"(0), "r"(b * m.Nc + n), "r"(smem_u32(&s_bar))
          : "memory");
  }
  __syncthreads();   // the barrier is initialised before anybody waits on it
  mbar_wait(&s_bar, 0);
  float *dst = ctxT + ((size_t)(b * m.Nc + n) * m.P + p0) * m.Cpad;
  for (int idx = threadIdx.x; idx < (kChunk / 4) * m.Cpad; idx += kPrepThreads) {
    const int jq = idx / m.Cpad, e = idx - jq * m.Cpad;   // pixel quad, element of the row
    const int c = perm.chan(e);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < m.C)
      v = *reinterpret_cast<const float4 *>(box0 + (jq >> 3) * box_bytes + c * 128 + (((jq & 7) ^ (c & 7)) << 4));
    float *dp = dst + (size_t)(4 * jq) * m.Cpad + e;
    if (4 * jq + 0 < npx) dp[0] = v.x;
    if (4 * jq + 1 < npx) dp[m.Cpad] = v.y;
    if (4 * jq + 2 < npx) dp[2 * m.Cpad] = v.z;
    if (4 * jq + 3 < npx) dp[3 * m.Cpad] = v.w;
  }
}

// One launch, two kinds of CTA (blockIdx.z): run weights of a pixel chunk (ALU / latency bound) and
// channels-last context rows of a pixel chunk (bandwidth bound) -- they overlap on every SM.
template <typename CT, bool BSM>
__global__ void __launch_bounds__(kPrepThreads)
ls_lift_prep_kernel(Dims m, const float *__restrict__ height, int vec16, const int *__restrict__ run_cnt,
                    const int *__restrict__ run_d, const int *__restrict__ run_dst,
                    float *__restrict__ w_pm_out, Entry *__restrict__ vm_ent_out,
                    const CT *__restrict__ context, CT *__restrict__ ctxT, RowPerm perm, BsmAssembly bsm, int ctx_tma,
                    const __grid_constant__ CUtensorMap ctx_map) {
  extern __shared__ __align__(128) float lift_smem[];
  if (blockIdx.z == 0)
    weights_role<BSM>(m, height, vec16, run_cnt, run_d, run_dst, w_pm_out, vm_ent_out, lift_smem, blockIdx.y, blockIdx.x, bsm);
  else if (!BSM && sizeof(CT) == 4 && ctx_tma)   // (block-uniform)
    context_rows_tma_role(m, reinterpret_cast<float *>(ctxT), perm, reinterpret_cast<unsigned char *>(lift_smem), blockIdx.y,
                          blockIdx.x, &ctx_map);
  else context_rows_role<CT>(m, context, ctxT, perm, lift_smem, blockIdx.y, blockIdx.x, bsm);
}

// ---------------------------------------------------------------------------------------------
// FORWARD: per-voxel weighted gather of context rows (sparse voxel x pixel times dense pixel x C).
// ---------------------------------------------------------------------------------------------
template <typename CT>
struct RowLoad;
template <>
struct RowLoad<float> {
  // `row` may already include the lane's slice offset; `slice` is an additional 4-element slice index
  static __device__ __forceinline__ void load(const float *row, int slice, float (&v)[4]) {
    const float4 t = __ldg(reinterpret_cast<const float4 *>(row) + slice);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};
template <>
struct RowLoad<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16 *row, int slice, float (&v)[4]) {
    const uint2 t = __ldg(reinterpret_cast<const uint2 *>(row) + slice);
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
  }
};

// ---- reduce: tile geometry ----------------------------------------------------------------------
constexpr int kTileV = 64;     // voxels per reduce CTA: two 32-voxel boxes = 2 x 128-byte rows per channel
// Entries of a tile are staged in shared memory when they fit (stage_cap, chosen per launch from the
// expected tile population); larger tiles read them from the plan in global memory.
static_assert(kTileV == 64, "Entry::off carries the voxel-in-tile index in its low 6 bits");

// 4-element row vector widened to fp32 (byte address).
template <typename CT>
struct RowLd;
template <>
struct RowLd<float> {
  static __device__ __forceinline__ void vec(const unsigned char *p, float (&v)[4]) {
    const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};
template <>
struct RowLd<__nv_bfloat16> {
  static __device__ __forceinline__ void vec(const unsigned char *p, float (&v)[4]) {
    const uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
  }
};

// Byte offset of the 16-byte chunk j (4 voxels) of channel row r inside one 32-voxel box of the
// tile: 128-byte rows, chunk index XOR-ed with (row & 7) -- the TMA SWIZZLE_128B pattern -- so that
// lanes holding different rows of the same voxel column hit different bank groups.
__device__ __forceinline__ int tile_chunk(int r, int j) { return r * 128 + ((j ^ (r & 7)) << 4); }

// Per-lane accumulator of one stream: lane l of a G-lane group owns channels l + G*j, j < 4*NV, of every
// gathered row (vector k, element e <-> j = 4k + e: the permuted channels-last row layout, transpose.cuh).
template <int G, int NV>
struct StreamAcc {
  float a[NV][4];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e) a[k][e] = 0.0f;
  }
  // write the sums to voxel column vt of the swizzled [channel][voxel] tile, then clear
  __device__ __forceinline__ void flush_tile(unsigned tile_lane /*smem addr of row l, chunk 0*/, int l,
                                             unsigned box_bytes, int vt) {
    // row r = l + G*j: (r & 7) = l & 7 for G >= 8;  for G = 4 it is (l + 4*(j & 1)): bit 2 toggles with j
    const unsigned chunk = (((unsigned)vt >> 2) & 7u) ^ ((unsigned)l & 7u);
    const unsigned base = tile_lane + ((unsigned)vt >> 5) * box_bytes + (((unsigned)vt & 3u) << 2);
    const unsigned a0 = base + (chunk << 4), a1 = base + ((chunk ^ 4u) << 4);
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = 4 * k + e;
        sts_f32(((G == 4 && (j & 1)) ? a1 : a0) + j * G * 128, a[k][e]);
        a[k][e] = 0.0f;
      }
  }
  // partial sum of a voxel that an earlier stream started: parked per stream, added by that stream later
  // channels-last output: the sums ARE the voxel's row (natural channel order: vector k of lane l = channels
  // 4*(k*G + l) ..+3) of the [voxel][C] shared-memory tile; the G lanes of a vector cover G*16 contiguous bytes.  Then clear.
  __device__ __forceinline__ void flush_row(float *row, int l, int C) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c4 = 4 * (k * G + l);
      if (c4 < C) *reinterpret_cast<float4 *>(row + c4) = make_float4(a[k][0], a[k][1], a[k][2], a[k][3]);
#pragma unroll
      for (int e = 0; e < 4; ++e) a[k][e] = 0.0f;
    }
  }
  __device__ __forceinline__ void store_head(float *head /*[4*G*NV] of this stream*/, int l) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      *reinterpret_cast<float4 *>(head + 4 * (k * G + l)) = make_float4(a[k][0], a[k][1], a[k][2], a[k][3]);
#pragma unroll
      for (int e = 0; e < 4; ++e) a[k][e] = 0.0f;
    }
  }
  __device__ __forceinline__ void add_head(const float *head, int l) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const float4 h = *reinterpret_cast<const float4 *>(head + 4 * (k * G + l));
      a[k][0] = __fadd_rn(a[k][0], h.x); a[k][1] = __fadd_rn(a[k][1], h.y);
      a[k][2] = __fadd_rn(a[k][2], h.z); a[k][3] = __fadd_rn(a[k][3], h.w);
    }
  }
};

// entry j of the tile's sorted list: from the staged copy in shared memory, or (huge tiles) from the plan
template <bool STAGED>
__device__ __forceinline__ Entry load_entry(const Entry *__restrict__ ent, int j) {
  if (STAGED) return ent[j];
  const uint2 t = __ldg(reinterpret_cast<const uint2 *>(ent + j));
  Entry e;
  e.off = t.x;
  e.w = __uint_as_float(t.y);
  return e;
}

// One stream = one G-lane group walking the slice [j0, j1) of the tile's sorted entries, two entries in
// flight.
// CL: channels-last output -- a finished voxel is a row of the [voxel][C] shared-memory tile (which then leaves as ONE
// contiguous block: the rows of a tile's voxels are adjacent in a (b, y, x, c) map); otherwise a column of the swizzled
// [channel][voxel] tile.
struct ClOut {
  float *rows;  // row of voxel 0 of the tile (shared memory)
  int C;
};
template <int G, int NV, bool CL>
__device__ __forceinline__ void flush_voxel(StreamAcc<G, NV> &acc, unsigned tile_lane, int l, unsigned box_bytes, int vt,
                                            const ClOut &cl) {
  if (CL) {
    acc.flush_row(cl.rows + vt * cl.C, l, cl.C);
  } else {
    acc.flush_tile(tile_lane, l, box_bytes, vt);
  }
}

// SKIP: entries whose weight is -0.0f (background pixels of the BSM call site: weights_role) do not gather their row.
template <typename CT, int G, int NV, bool STAGED, bool CL, bool SKIP>
__device__ __forceinline__ void stream_loop(StreamAcc<G, NV> &acc, int &cur, bool &head,
                                            const Entry *__restrict__ ent, int tile_lo, int j0, int j1,
                                            const unsigned char *__restrict__ lane_rows, unsigned tile_lane,
                                            float *my_head, int l, unsigned box_bytes, const ClOut &cl) {
  constexpr unsigned kRowBytes = 4 * G * NV * sizeof(CT);
  auto load_row = [&](const Entry &en, float (&r)[NV][4]) {
    if (SKIP && __float_as_uint(en.w) == 0x80000000u) {
#pragma unroll
      for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e) r[k][e] = 0.0f;
      return;
    }
    const unsigned char *row = lane_rows + (size_t)(en.off >> 6) * kRowBytes;
#pragma unroll
    for (int k = 0; k < NV; ++k) RowLd<CT>::vec(row + k * G * 4 * sizeof(CT), r[k]);
  };
  auto accumulate = [&](const Entry &en, const float (&r)[NV][4]) {
    const int vt = (int)(en.off & 63u);
    if (vt != cur) {
      if (cur >= 0) {
        if (head) acc.store_head(my_head, l);
        else flush_voxel<G, NV, CL>(acc, tile_lane, l, box_bytes, cur, cl);
        head = false;
      }
      cur = vt;
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      fma2(acc.a[k][0], acc.a[k][1], en.w, r[k][0], r[k][1]);
      fma2(acc.a[k][2], acc.a[k][3], en.w, r[k][2], r[k][3]);
    }
  };
  if (j0 >= j1) return;
  // the slice continues a voxel of the previous slice iff the entry before it carries the same voxel
  if (j0 > tile_lo)
    head = ((load_entry<STAGED>(ent, j0 - 1).off ^ load_entry<STAGED>(ent, j0).off) & 63u) == 0;
  int j = j0;
#pragma unroll 1
  for (; j + 2 <= j1; j += 2) {
    float ra[NV][4], rb[NV][4];
    const Entry ea = load_entry<STAGED>(ent, j), eb = load_entry<STAGED>(ent, j + 1);
    load_row(ea, ra);
    load_row(eb, rb);
    accumulate(ea, ra);
    accumulate(eb, rb);
  }
  if (j < j1) {
    float ra[NV][4];
    const Entry ea = load_entry<STAGED>(ent, j);
    load_row(ea, ra);
    accumulate(ea, ra);
  }
}

// ---------------------------------------------------------------------------------------------
// grid (ceil(V / 64), B), 32 * NSTR * G / 32 threads.  CTA = tile of 64 consecutive voxels:
//   1. the tile's CSR offsets and its sorted entries are staged in shared memory with coalesced loads;
//      the run weights are gathered through the entries' slot index on the way (one round trip for the
//      whole tile); the [channel][voxel] tile is zeroed;
//   2. the tile's entry list is cut into NSTR slices of EQUAL length (+-1), one per G-lane group
//      ("stream"): perfectly regular work whatever the points-per-voxel distribution looks like.  A group
//      owns whole context rows (NV 128-bit loads per lane and entry), packed FFMA2, 2 entries per stream
//      in flight.  A voxel cut by a slice boundary is summed in slice order: every later slice parks its
//      partial ("head") in shared memory, and the slice that started the voxel adds the heads in order
//      after a barrier => one fixed summation order per voxel, no atomics, bitwise reproducible;
//   3. finished voxel sums go to a swizzled [channel][voxel] tile, and the tile leaves as 128-byte row
//      segments of the NCHW planes (fully coalesced, every output byte written exactly once).
// ---------------------------------------------------------------------------------------------
//   CL (channels-last BEV map, desc.reserved[1] bit 1): the tile is [voxel][C] rows instead, a finished voxel's sums
//      are one row of it, and the tile -- 64 adjacent rows of the (b, y, x, c) map, one contiguous block of memory --
//      leaves with a single bulk-async (TMA) shared -> global copy.
template <typename CT, int G, int NV, int NSTR, bool CL, bool SKIP>
__global__ void __launch_bounds__(NSTR * G)
ls_reduce_kernel(Dims m, const CT *__restrict__ ctxT, const int *__restrict__ tile_ptr,
                 const Entry *__restrict__ vm_ent, float *__restrict__ bev, int vec_out, int stage_cap,
                 const __grid_constant__ CUtensorMap bev_map) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int kThreads = NSTR * G;
  constexpr int kRows = 4 * G * NV;                  // = Cpad rows (rows >= C are scratch)
  constexpr unsigned kBoxBytes = kRows * 128;
  unsigned char *tile = smem_raw;                                          // 2 boxes x kRows x 128 B
  float *s_head = reinterpret_cast<float *>(smem_raw + 2 * kBoxBytes);     // NSTR x kRows partial sums
  Entry *s_ent = reinterpret_cast<Entry *>(s_head + NSTR * kRows);         // stage_cap entries
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  const int l = tid & (G - 1);
  const int s = tid / G;  // stream index: consecutive streams sit in one warp
  const int v0 = blockIdx.x * kTileV;
  const int nv = min(kTileV, m.V - v0);
  const int *tp = tile_ptr + (size_t)b * (m.ntiles + 1) + blockIdx.x;
  const int tile_lo = __ldg(tp), tile_hi = __ldg(tp + 1);
  float *out = bev + (size_t)b * m.C * m.V + (CL ? (size_t)v0 * m.C : (size_t)v0);
  ClOut cl;
  cl.rows = reinterpret_cast<float *>(tile); cl.C = m.C;   // (64 * C floats <= the 2 * Cpad * 32 of the NCHW tile)

  // thread <-> (channel row c0 + kThreads/16 * i, 16-byte chunk q) of the tile for the copy-out loops
  constexpr int kRowStep = kThreads / 16;
  static_assert(kRowStep % 8 == 0, "copy-out relies on (row & 7) being loop invariant");
  const int q = tid & 15, c0 = tid >> 4;
  if (tile_hi == tile_lo) {  // no point falls into this tile: zero fill
    if (CL) {  // the tile's rows are one contiguous range of nv * C floats (C % 4 == 0, 16-byte aligned)
      for (int i = tid; i < nv * (m.C >> 2); i += kThreads)
        stg_stream_f4(reinterpret_cast<float4 *>(out) + i, make_float4(0.f, 0.f, 0.f, 0.f));
    } else if (vec_out) {
      if (4 * q < nv) {
        float4 *o4 = reinterpret_cast<float4 *>(out + (size_t)c0 * m.V) + q;
        const size_t step = (size_t)(kRowStep / 4) * m.V;  // kRowStep channel rows, in float4 units
        for (int c = c0; c < m.C; c += kRowStep, o4 += step) stg_stream_f4(o4, make_float4(0.f, 0.f, 0.f, 0.f));
      }
    } else {
      for (int i = tid; i < m.C * kTileV; i += kThreads) {
        const int c = i >> 6, j = i & 63;
        if (j < nv) stg_stream_f1(out + (size_t)c * m.V + j, 0.0f);
      }
    }
    return;
  }

  const int n_t = tile_hi - tile_lo;
  const bool staged = n_t <= stage_cap;
  const Entry *ent = vm_ent + (size_t)b * m.cap;
  if (staged)
    for (int i = tid; i < n_t; i += kThreads) cp_async_8(s_ent + i, ent + tile_lo + i);
  for (int i = tid; i < (CL ? kTileV * m.C / 4 : 2 * (int)kBoxBytes / 16); i += kThreads)
    reinterpret_cast<float4 *>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  cp_async_wait_all();
  __syncthreads();

  // stream s owns the entries [j0, j1) of the tile: equal slices
  const int E = (n_t + NSTR - 1) / NSTR;
  const int j0 = tile_lo + min(s * E, n_t), j1 = tile_lo + min((s + 1) * E, n_t);
  const unsigned char *lane_rows = reinterpret_cast<const unsigned char *>(ctxT) +
                                   ((size_t)b * m.Nc * m.P * m.Cpad + 4 * l) * sizeof(CT);
  const unsigned tile_lane = (unsigned)__cvta_generic_to_shared(tile) + l * 128;
  float *my_head = s_head + s * kRows;

  StreamAcc<G, NV> acc;
  acc.clear();
  int cur = -1;         // voxel (in tile) being accumulated
  bool head = false;    // the segment being accumulated continues a voxel started by an earlier stream
  const Entry *se = s_ent - tile_lo;  // staged entries, addressed like the plan's
  if (staged)
    stream_loop<CT, G, NV, true, CL, SKIP>(acc, cur, head, se, tile_lo, j0, j1, lane_rows, tile_lane, my_head, l, kBoxBytes, cl);
  else
    stream_loop<CT, G, NV, false, CL, SKIP>(acc, cur, head, ent, tile_lo, j0, j1, lane_rows, tile_lane, my_head, l, kBoxBytes, cl);
  // voxel (in tile) of entry j
  auto vox_at = [&](int j) -> int {
    return (int)((staged ? load_entry<true>(se, j) : load_entry<false>(ent, j)).off & 63u);
  };
  // the last segment of the slice: complete iff the slice ends on the voxel's last entry
  bool tail = false;
  if (cur >= 0) {
    if (head) acc.store_head(my_head, l);                 // the whole slice lies inside one earlier voxel
    else if (j1 == tile_hi || vox_at(j1) != cur) flush_voxel<G, NV, CL>(acc, tile_lane, l, kBoxBytes, cur, cl);
    else tail = true;                                      // later slices continue this voxel
  }
  __syncthreads();
  if (tail) {
    // slices s+1, s+2, ... that start inside this voxel parked their partial sums: add them in order
    for (int k = s + 1; k < NSTR; ++k) {
      const int jk = tile_lo + k * E;
      if (jk >= tile_hi || vox_at(jk) != cur) break;
      acc.add_head(s_head + k * kRows, l);
    }
    flush_voxel<G, NV, CL>(acc, tile_lane, l, kBoxBytes, cur, cl);
  }
  __syncthreads();
  if (CL) {
    // one bulk-async copy moves the whole tile (nv * C * 4 bytes, a multiple of 16): the generic-proxy writes above are
    // ordered before the async proxy's reads by the fence, and the CTA may retire once the engine has READ shared memory
    if (tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out), "r"(smem_u32(tile)),
                   "r"((unsigned)(nv * m.C * (int)sizeof(float)))
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    return;
  }

  // tile -> global.  vec_out == 2: the two 32-voxel boxes of the tile ARE TMA boxes (128-byte rows, SWIZZLE_128B):
  // one thread hands each to the TMA engine as a 2-D tensor store ({32 voxels, C channel planes} at (v0, b * C) of
  // the (B * C) x V view of the map) and the CTA retires once the engine has read shared memory -- no LDS / STG.
  if (vec_out == 2) {
    if (tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const unsigned long long mp = reinterpret_cast<unsigned long long>(&bev_map);
      const int y = b * m.C;
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(mp), "r"(v0), "r"(y),
                   "r"(smem_u32(tile))
                   : "memory");
      if (nv > 32)
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(mp), "r"(v0 + 32),
                     "r"(y), "r"(smem_u32(tile) + kBoxBytes)
                     : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    return;
  }
  // vec_out == 1: thread = (channel row, 16-byte chunk); 8 consecutive lanes write one 128-byte line.
  // Rows advance by a multiple of 8 per step, so (row & 7) and with it the swizzled chunk position never change.
  if (vec_out) {
    if (4 * q < nv) {
      const unsigned char *t4 = tile + (q >> 3) * kBoxBytes + tile_chunk(c0, q & 7);
      float4 *o4 = reinterpret_cast<float4 *>(out + (size_t)c0 * m.V) + q;
      const size_t step = (size_t)(kRowStep / 4) * m.V;
      for (int c = c0; c < m.C; c += kRowStep, o4 += step, t4 += kRowStep * 128)
        stg_stream_f4(o4, *reinterpret_cast<const float4 *>(t4));
    }
  } else {
    for (int i = tid; i < m.C * kTileV; i += kThreads) {
      const int c = i >> 6, j = i & 63, qq = j >> 2;
      if (j < nv)
        stg_stream_f1(out + (size_t)c * m.V + j,
                      *reinterpret_cast<const float *>(tile + (qq >> 3) * kBoxBytes + tile_chunk(c, qq & 7) + 4 * (j & 3)));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// BACKWARD: one warp per pixel.  Lanes hold 4-channel slices of the pixel's context row;
// for every run: G row gather (128-bit), g_ctx += w*G, gw = <ctx, G> (fixed butterfly order).
// Four runs are processed together so that four row loads and four butterflies overlap.
// ---------------------------------------------------------------------------------------------
template <typename CT, int NCH>
__global__ void __launch_bounds__(256)
ls_backward_gather_kernel(Dims m, int gpad, const CT *__restrict__ ctxT, const float *__restrict__ gT,
                          const int *__restrict__ run_cnt, const int *__restrict__ run_vox,
                          const float *__restrict__ w_pm, float *__restrict__ gw_pm,
                          float *__restrict__ gctxT) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int fp = blockIdx.x * 8 + (threadIdx.x >> 5);  // frame-local pixel n*P + p
  if (fp >= m.Nc * m.P) return;
  const int n = fp / m.P, p = fp - n * m.P;
  const int ci = p / kChunk, t = p - ci * kChunk;
  const int frame_chunk = b * m.nchunks + n * m.cpc + ci;
  const int cnt = run_cnt[(size_t)frame_chunk * kChunk + t];
  const size_t row_id = (size_t)b * m.Nc * m.P + fp;
  const int nslices_c = m.Cpad / 4, nslices_g = gpad / 4;
  float cx[NCH][4], acc[NCH][4];
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int sl = k * 32 + lane;
#pragma unroll
    for (int e = 0; e < 4; ++e) { cx[k][e] = 0.0f; acc[k][e] = 0.0f; }
    if (sl < nslices_c && cnt > 0) RowLoad<CT>::load(ctxT + row_id * m.Cpad, sl, cx[k]);
  }
  const float *gb = gT + (size_t)b * m.V * gpad;
  for (int r0 = 0; r0 < cnt; r0 += 32) {
    const int nr = min(32, cnt - r0);
    int my_vox = 0;
    float my_w = 0.0f;
    if (lane < nr) {
      const size_t s = ell_slot(frame_chunk, m.D, r0 + lane, t);
      my_vox = run_vox[s];
      my_w = w_pm[s];
    }
    float my_gw = 0.0f;
    for (int q = 0; q < nr; q += 4) {
      float4 g[4][NCH];
      float w4[4], dot[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int src = min(q + u, nr - 1);
        const float *grow = gb + (size_t)__shfl_sync(0xffffffffu, my_vox, src) * gpad;
        w4[u] = __shfl_sync(0xffffffffu, my_w, src);
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
          const int sl = k * 32 + lane;
          g[u][k] = (sl < nslices_g && q + u < nr) ? __ldg(reinterpret_cast<const float4 *>(grow) + sl)
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        dot[u] = 0.0f;
        if (q + u < nr) {
#pragma unroll
          for (int k = 0; k < NCH; ++k) {
            if (k * 32 + lane < nslices_g) {
              acc[k][0] = __fmaf_rn(w4[u], g[u][k].x, acc[k][0]);
              acc[k][1] = __fmaf_rn(w4[u], g[u][k].y, acc[k][1]);
              acc[k][2] = __fmaf_rn(w4[u], g[u][k].z, acc[k][2]);
              acc[k][3] = __fmaf_rn(w4[u], g[u][k].w, acc[k][3]);
              dot[u] = __fmaf_rn(cx[k][0], g[u][k].x, dot[u]);
              dot[u] = __fmaf_rn(cx[k][1], g[u][k].y, dot[u]);
              dot[u] = __fmaf_rn(cx[k][2], g[u][k].z, dot[u]);
              dot[u] = __fmaf_rn(cx[k][3], g[u][k].w, dot[u]);
            }
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int u = 0; u < 4; ++u) dot[u] = __fadd_rn(dot[u], __shfl_xor_sync(0xffffffffu, dot[u], o));
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (lane == q + u) my_gw = dot[u];
    }
    if (lane < nr) gw_pm[ell_slot(frame_chunk, m.D, r0 + lane, t)] = my_gw;
  }
  float *dst = gctxT + row_id * gpad;
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int sl = k * 32 + lane;
    if (sl < nslices_g)
      *(reinterpret_cast<float4 *>(dst) + sl) = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
  }
}

// ---------------------------------------------------------------------------------------------
// BACKWARD: grad_bev (B, C, Y, X) -> one channels-last row per voxel (permuted layout), only for the
// 64-voxel tiles that received at least one run (the rows of untouched voxels are never read).
// grid (ntiles, B), 256 threads; reads 256-byte pieces of the channel planes, writes whole rows.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ls_grad_rows_kernel(Dims m, const float *__restrict__ grad_bev, const int *__restrict__ tile_ptr,
                    float *__restrict__ gT, RowPerm perm) {
  extern __shared__ float gsm[];  // [C][65]
  constexpr int kLd = kTileV + 1;
  const int b = blockIdx.y, tile = blockIdx.x;
  const int *tp = tile_ptr + (size_t)b * (m.ntiles + 1) + tile;
  if (__ldg(tp) == __ldg(tp + 1)) return;
  const int v0 = tile * kTileV;
  const int nv = min(kTileV, m.V - v0);
  const int t = threadIdx.x & 63, q = threadIdx.x >> 6;
  if (t < nv) {
    const float *src = grad_bev + (size_t)b * m.C * m.V + v0 + t;
    for (int c = q; c < m.C; c += 4) cp_async_4(gsm + c * kLd + t, src + (size_t)c * m.V);
  }
  cp_async_wait_all();
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float *dst = gT + ((size_t)b * m.V + v0) * m.Cpad;
  for (int e0 = 0; e0 < m.Cpad; e0 += 32) {
    const int e = e0 + lane;
    if (e >= m.Cpad) continue;
    const int c = perm.chan(e);
    float *dp = dst + (size_t)wid * m.Cpad + e;
    const size_t dstep = (size_t)8 * m.Cpad;
    if (c < m.C) {
      const float *sp = gsm + c * kLd + wid;
#pragma unroll 4
      for (int v = wid; v < nv; v += 8, sp += 8, dp += dstep) *dp = *sp;
    } else {
      for (int v = wid; v < nv; v += 8, dp += dstep) *dp = 0.0f;
    }
  }
}

// The same with the TMA engine on the way in: the tile's two {32 voxels, C planes} boxes of the (B*C) x V view of grad_bev
// arrive with two 2-D tensor loads (SWIZZLE_128B, completion on an mbarrier) instead of 4-byte cp.async copies; a thread
// then moves 16-byte chunks (4 voxels of one channel: conflict-free per quarter warp) into 4 rows.
__global__ void __launch_bounds__(256)
ls_grad_rows_tma_kernel(Dims m, const int *__restrict__ tile_ptr, float *__restrict__ gT, RowPerm perm,
                        const __grid_constant__ CUtensorMap g_map) {
  extern __shared__ __align__(1024) unsigned char gbox[];   // 2 boxes x C rows x 128 B, each box on a 1 KB boundary
  __shared__ __align__(8) unsigned long long s_bar;
  const int b = blockIdx.y, tile = blockIdx.x;
  const int *tp = tile_ptr + (size_t)b * (m.ntiles + 1) + tile;
  if (__ldg(tp) == __ldg(tp + 1)) return;
  const int v0 = tile * kTileV;
  const int nv = min(kTileV, m.V - v0);
  // (the 128-byte swizzle works on shared-memory ADDRESS bits 7-9: a box must start on a 1 KB boundary for "row & 7" to be them)
  const unsigned box_bytes = (((unsigned)m.C + 7u) >> 3) * 1024u;
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    mbar_arrive_expect_tx(&s_bar, 2u * (unsigned)m.C * 128u);   // (out-of-range voxels of the last tile are zero-filled and counted)
    const unsigned long long mp = reinterpret_cast<unsigned long long>(&g_map);
    for (int k = 0; k < 2; ++k)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                       smem_u32(gbox) + k * box_bytes),
                   "l"(mp), "r"(v0 + 32 * k), "r"(b * m.C), "r"(smem_u32(&s_bar))
                   : "memory");
  }
  __syncthreads();   // the barrier is initialised before anybody waits on it
  mbar_wait(&s_bar, 0);
  float *dst = gT + ((size_t)b * m.V + v0) * m.Cpad;
  for (int idx = threadIdx.x; idx < 16 * m.Cpad; idx += 256) {
    const int jq = idx / m.Cpad, e = idx - jq * m.Cpad;   // voxel quad, element of the row
    const int c = perm.chan(e);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < m.C)
      v = *reinterpret_cast<const float4 *>(gbox + (jq >> 3) * box_bytes + c * 128 + (((jq & 7) ^ (c & 7)) << 4));
    float *dp = dst + (size_t)(4 * jq) * m.Cpad + e;
    if (4 * jq + 0 < nv) dp[0] = v.x;
    if (4 * jq + 1 < nv) dp[m.Cpad] = v.y;
    if (4 * jq + 2 < nv) dp[2 * m.Cpad] = v.z;
    if (4 * jq + 3 < nv) dp[3 * m.Cpad] = v.w;
  }
}

// ---------------------------------------------------------------------------------------------
// BACKWARD, fused per pixel chunk (rows of <= 96 channels): one CTA = 64 pixels x 4 lanes.
//   1. the chunk's D x 64 height block and C x 64 context block are staged with cp.async (coalesced,
//      all loads in flight at once);
//   2. a 4-lane group owns one pixel: softmax over D (lane l takes bins d = l mod 4: the same four
//      summation chains, combined in the same order, as the forward weights pass), run weights
//      w_r = sum_{d in run} p_d (lane r mod 4, ascending d);
//   3. per run: 128-bit gather of the voxel's gradient row G (channels-last, permuted layout: lane l holds
//      channels l + 4j), g_ctx += w_r * G (packed FFMA2), gw_r = <ctx, G> (butterfly over the 4 lanes,
//      fixed order); S = sum_r w_r gw_r for the softmax backward;
//   4. g_height[d] = gw[run(d)] (0 for dropped bins), or p_d (gw - S) when the softmax is fused;
//      g_ctx leaves through the shared-memory tile as coalesced NCHW rows.
// No atomics, fixed summation orders => bitwise reproducible.  Replaces weights + gather + expand +
// the g_ctx transpose of the unfused path (four passes over the run table) with one.
// ---------------------------------------------------------------------------------------------
constexpr int kBwdPix = 64;    // pixels per CTA (half a plan chunk)
constexpr int kBwdLd = kBwdPix + 1;

template <typename CT, int NV, int OCC, bool BSM, bool NAT>
__global__ void __launch_bounds__(kBwdPix * 4, OCC)
ls_backward_chunk_kernel(Dims m, const float *__restrict__ height, int vec16, int vec16_out,
                         const CT *__restrict__ context, BsmAssembly bsm, float *__restrict__ g_semantic,
                         const float *__restrict__ gT, const int *__restrict__ run_cnt,
                         const int *__restrict__ run_vox, const int *__restrict__ run_d,
                         float *__restrict__ w_pm, float *__restrict__ gw_pm, float *__restrict__ g_height,
                         float *__restrict__ g_context) {
  constexpr int G = 4, NJ = 4 * NV;   // NJ channels per lane
  constexpr int kRowF = 4 * G * NV;   // floats per gradient row (= Cpad)
  extern __shared__ __align__(128) float bwd_smem[];
  __shared__ __align__(8) unsigned long long s_bar;  // completion of the bulk-async height block
  __shared__ float s_keep[BSM ? kBwdPix : 1];        // BSM: 0 for background pixels
  __shared__ float s_semp[BSM ? 8 : 1][kBwdPix];     // BSM: semantic probabilities (Cs <= 8)
  float *col = bwd_smem;                     // [D][64]   height bins, then exp(x - max)
  float *tile = bwd_smem + m.D * kBwdPix;    // [Cpad][65] context in, g_ctx out
  const int b = blockIdx.y;
  const int chunk = blockIdx.x >> 1, half = blockIdx.x & 1;
  const int n = chunk / m.cpc, ci = chunk - n * m.cpc;
  const int frame_chunk = b * m.nchunks + chunk;
  const int tid = threadIdx.x;
  const int p0 = ci * kChunk + half * kBwdPix;
  if (p0 >= m.P) return;  // second half of a ragged last chunk (block-uniform)
  const int npx = min(kBwdPix, m.P - p0);
  const int bn = b * m.Nc + n;

  // ---- 1. stage ------------------------------------------------------------------------------------
  {
    const int t = tid & (kBwdPix - 1), q = tid >> 6;  // 4 quarter-blocks of threads, each walks a share of the rows
    const bool bulk = vec16 && npx == kBwdPix;   // block-uniform
    if (bulk) {
      // bulk-async copies (TMA engine): one instruction per 256-byte row of the height block
      if (tid == 0) mbar_init(&s_bar, 1);
      __syncthreads();
      bulk_stage_rows(col, kBwdPix * sizeof(float), height + (size_t)bn * m.hs + p0, (size_t)m.P * sizeof(float), m.D,
                      kBwdPix * sizeof(float), &s_bar);
    } else if (t < npx) {
      const float *hs = height + (size_t)bn * m.hs + p0 + t;
      for (int d = q; d < m.D; d += 4) cp_async_4(col + d * kBwdPix + t, hs + (size_t)d * m.P);
    }
    const int Cc = m.C - (BSM ? bsm.Cs : 0);   // channels that come from `context`
    if (t < npx) {
      const CT *cs = context + (size_t)bn * m.cs + p0 + t;
      if (sizeof(CT) == 4) {
        for (int c = q; c < Cc; c += 4)
          cp_async_4(tile + c * kBwdLd + t, reinterpret_cast<const float *>(cs) + (size_t)c * m.P);
        if (BSM) {
          const float *ss = bsm.sem + (size_t)bn * bsm.sem_stride + p0 + t;
          for (int k = q; k < bsm.Cs; k += 4) cp_async_4(tile + (Cc + k) * kBwdLd + t, ss + (size_t)k * m.P);
        }
      } else {
        for (int c = q; c < Cc; c += 4) tile[c * kBwdLd + t] = to_f32<CT>(cs[(size_t)c * m.P]);
      }
    }
    cp_async_wait_all();
    if (bulk) mbar_wait(&s_bar, 0);
    __syncthreads();
    if (BSM) {  // block-uniform
      // BSMLSSFPN context assembly (bsm_lss_fpn.py:524-529), as in the forward's context pass: semantic =
      // softmax(logits) in torch's channel-softmax order, rows = cat(context, semantic) * (1 - (semantic[0] > thr)).
      // The unmasked probabilities stay in s_semp for the softmax backward at the end.
      if (tid < kBwdPix) {
        float keep = 1.0f;
        if (tid < npx) {
          float *colp = tile + Cc * kBwdLd + tid;
          float mx = colp[0];
          for (int k = 1; k < bsm.Cs; ++k) mx = fmaxf(mx, colp[k * kBwdLd]);
          float sum = 0.0f;
          for (int k = 0; k < bsm.Cs; ++k) sum = __fadd_rn(sum, expf(__fsub_rn(colp[k * kBwdLd], mx)));
          for (int k = 0; k < bsm.Cs; ++k) {
            const float pk = __fdiv_rn(expf(__fsub_rn(colp[k * kBwdLd], mx)), sum);
            s_semp[k][tid] = pk;
            colp[k * kBwdLd] = pk;
          }
          keep = colp[0] > bsm.thr ? 0.0f : 1.0f;
        }
        s_keep[tid] = keep;
      }
      __syncthreads();
      // (masked pixels are skipped below: every gradient of theirs is an exact zero)
    }
  }

  const int px = tid >> 2, l = tid & 3;       // group = pixel, lane inside the group
  const bool live = px < npx;
  const int tch = half * kBwdPix + px;        // thread index of this pixel inside the plan chunk
  const bool masked = BSM && s_keep[px] == 0.0f;
  const int cnt = (live && !masked) ? run_cnt[(size_t)frame_chunk * kChunk + tch] : 0;

  // ---- 2. softmax over D -----------------------------------------------------------------------------
  float scale = 1.0f;
  if (m.logits) {
    float mx = -INFINITY;
    if (live)
      for (int d = l; d < m.D; d += 4) mx = fmaxf(mx, col[d * kBwdPix + px]);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float sm = 0.0f;
    if (live)
      for (int d = l; d < m.D; d += 4) {
        const float e = exp_ex2(__fsub_rn(col[d * kBwdPix + px], mx));
        col[d * kBwdPix + px] = e;
        sm = __fadd_rn(sm, e);
      }
    sm = __fadd_rn(sm, __shfl_xor_sync(0xffffffffu, sm, 1));   // (s0 + s1), (s2 + s3)
    sm = __fadd_rn(sm, __shfl_xor_sync(0xffffffffu, sm, 2));   // (s0 + s1) + (s2 + s3)
    scale = __fdiv_rn(1.0f, sm);
  }
  // context row of the pixel -> registers (channel l + 4j)
  // channel held in slot j of lane l: l + 4j in the permuted rows ls_grad_rows_kernel writes; 16 (j / 4) + 4 l + j % 4
  // when the rows are the caller's channels-last gradient itself (NAT: natural order, row stride C = 16 NV)
  auto chan = [&](int j) { return NAT ? 16 * (j >> 2) + 4 * l + (j & 3) : l + 4 * j; };
  float cx[NJ], acc[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = chan(j);
    cx[j] = (live && (NAT || c < m.C)) ? tile[c * kBwdLd + px] : 0.0f;   // (NAT: C = 16 NV, every slot is a channel)
    acc[j] = 0.0f;
  }
  __syncthreads();  // every thread has its context row (and the exponentials are visible to the group)
  // Per-bin gradients gbin[d][pixel] live in the context tile while it is free (between the context rows moving to
  // registers and the g_ctx rows coming back): the run loop drops gw_r into the bins of run r, phase 4a reads them back
  // in one uniform pass over D -- no second walk over the run table, no round trip of gw through global memory.
  // Needs D rows of 65 floats in the tile (the host sizes it for max(Cpad, D) rows when D <= 96) and pays off when the
  // pixels have more than a few runs: taken when at least half of the chunk's pixels have 8 runs or more (block-uniform;
  // measured on B200: thresholds 0 .. 18 are within 1 % of each other -- DAIR-R50 chunk kernel 373 -> 356 us, Rope3D-R50
  // 346 -> 313 us, SGV3D-BSM-R50 662 -> 595 us), else the run table is walked again.
  const bool bins_in_tile = (m.D <= kRowF || (vec16_out & 4)) && __syncthreads_count(cnt >= 8) >= 2 * kBwdPix;
  float *gbin = tile;
  if (bins_in_tile) {
    float4 *z4 = reinterpret_cast<float4 *>(tile);
    for (int i = tid; i < (m.D * kBwdLd + 3) / 4; i += kBwdPix * 4) z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  // ---- 3a. run weights: lane l takes the runs r = l mod 4 (ascending d inside a run) ----------------------
  const size_t ell0 = ell_slot(frame_chunk, m.D, 0, tch);
  for (int r = l; r < cnt; r += 4) {
    const size_t sl = ell0 + (size_t)r * kChunk;
    const int packed = run_d[sl];
    const int d0 = packed & 0xffff, d1 = packed >> 16;
    float wr = 0.0f;
    for (int d = d0; d < d1; ++d) wr = __fadd_rn(wr, col[d * kBwdPix + px]);
    w_pm[sl] = m.logits ? __fmul_rn(wr, scale) : wr;
  }
  __syncthreads();  // every lane of the group reads all of the pixel's weights below

  // ---- 3b. runs --------------------------------------------------------------------------------------
  // Run descriptors are fetched four ahead and two gradient rows are in flight per group, so that the
  // descriptor -> row dependency and the row latency overlap with the arithmetic of earlier runs.
  // The trip count is warp-uniform (longest pixel of the warp) so that the butterflies use the full mask.
  constexpr int g_stride = kRowF;   // (NAT: the caller's rows have C = 16 NV = kRowF floats as well)
  const float *gb = gT + (size_t)b * m.V * g_stride + 4 * l;
  const int cnt_w = __reduce_max_sync(0xffffffffu, cnt);
  float S = 0.0f;
  auto load_g = [&](int vox, float (&g)[NV][4]) {
    const float *grow = gb + (size_t)vox * g_stride;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const float4 t4 = __ldg(reinterpret_cast<const float4 *>(grow + 16 * k));
      g[k][0] = t4.x; g[k][1] = t4.y; g[k][2] = t4.z; g[k][3] = t4.w;
    }
  };
  auto consume = [&](int r, float wr, int packed, const float (&g)[NV][4]) {
    float dot = 0.0f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      fma2(acc[4 * k + 0], acc[4 * k + 1], wr, g[k][0], g[k][1]);
      fma2(acc[4 * k + 2], acc[4 * k + 3], wr, g[k][2], g[k][3]);
#pragma unroll
      for (int e = 0; e < 4; ++e) dot = __fmaf_rn(cx[4 * k + e], g[k][e], dot);
    }
    dot = __fadd_rn(dot, __shfl_xor_sync(0xffffffffu, dot, 1));
    dot = __fadd_rn(dot, __shfl_xor_sync(0xffffffffu, dot, 2));
    if (bins_in_tile) {
      if (r < cnt)
        for (int d = (packed & 0xffff) + l; d < (packed >> 16); d += 4) gbin[d * kBwdLd + px] = dot;
    } else if (l == 0 && r < cnt) {
      gw_pm[ell0 + (size_t)r * kChunk] = dot;
    }
    S = __fmaf_rn(wr, dot, S);   // wr == 0 beyond the pixel's last run
  };
  for (int r0 = 0; r0 < cnt_w; r0 += 4) {
    int vox[4], pk[4];
    float wv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      vox[u] = 0; wv[u] = 0.0f; pk[u] = 0;
      if (r0 + u < cnt) {
        const size_t sl = ell0 + (size_t)(r0 + u) * kChunk;
        vox[u] = run_vox[sl];
        wv[u] = w_pm[sl];
        if (bins_in_tile) pk[u] = run_d[sl];
      }
    }
#pragma unroll
    for (int h = 0; h < 4; h += 2) {
      if (r0 + h < cnt_w) {  // warp-uniform
        float ga[NV][4], gb2[NV][4];
#pragma unroll
        for (int k = 0; k < NV; ++k)
#pragma unroll
          for (int e = 0; e < 4; ++e) { ga[k][e] = 0.0f; gb2[k][e] = 0.0f; }
        if (r0 + h < cnt) load_g(vox[h], ga);
        if (r0 + h + 1 < cnt) load_g(vox[h + 1], gb2);
        consume(r0 + h, wv[h], pk[h], ga);
        consume(r0 + h + 1, wv[h + 1], pk[h + 1], gb2);
      }
    }
  }
  auto park_g_ctx = [&]() {   // g_ctx row -> tile column (the context values are in registers now)
    if (live) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int c = chan(j);
        if (NAT || c < m.C) tile[c * kBwdLd + px] = acc[j];
      }
    }
  };
  if (bins_in_tile) {
    __syncwarp();   // the four lanes of a pixel wrote its bins
    if (live) {
      // ---- 4a. g_height[d] = gw[run(d)] (0 for dropped bins), or p_d (gw - S) with the fused softmax ----------
      for (int d = l; d < m.D; d += 4) {
        const float gv = gbin[d * kBwdLd + px];
        col[d * kBwdPix + px] = m.logits ? __fmul_rn(__fmul_rn(col[d * kBwdPix + px], scale), __fsub_rn(gv, S)) : gv;
      }
    }
    __syncthreads();  // every pixel is done with its bins: the tile takes the g_ctx rows
    park_g_ctx();
  } else {
    park_g_ctx();
  }
  __syncthreads();  // tile complete; gw_pm writes of this CTA are visible to it

  // ---- 4a. g_height: lane l computes the bins d = l mod 4 in place of the staged column, then the block leaves as
  //          whole 256-byte rows (128-bit stores) ----------------------------------------------------------------
  if (live && !bins_in_tile) {
    float *cp = col + px;
    auto put = [&](int d, float gv) {
      float v = gv;
      if (m.logits) v = __fmul_rn(__fmul_rn(cp[d * kBwdPix], scale), __fsub_rn(gv, S));
      cp[d * kBwdPix] = v;
    };
    int dc = l;  // next bin of this lane
    for (int r = 0; r < cnt; ++r) {
      const size_t sl = ell0 + (size_t)r * kChunk;
      const int packed = run_d[sl];
      const float gv = gw_pm[sl];
      const int d0 = packed & 0xffff, d1 = packed >> 16;
      for (; dc < d0; dc += 4) put(dc, 0.0f);
      for (; dc < d1; dc += 4) put(dc, gv);
    }
    for (; dc < m.D; dc += 4) put(dc, 0.0f);
  }
  __syncthreads();
  if ((vec16_out & 1) && npx == kBwdPix) {
    const int t4 = tid & 15;
    float *gh = g_height + (size_t)bn * m.ghs + p0 + 4 * t4;
    for (int d = tid >> 4; d < m.D; d += (kBwdPix * 4) >> 4)
      stg_stream_f4(reinterpret_cast<float4 *>(gh + (size_t)d * m.P), *reinterpret_cast<const float4 *>(col + d * kBwdPix + 4 * t4));
  } else {
    const int t = tid & (kBwdPix - 1), q = tid >> 6;
    if (t < npx) {
      float *gh = g_height + (size_t)bn * m.ghs + p0 + t;
      for (int d = q; d < m.D; d += 4) stg_stream_f1(gh + (size_t)d * m.P, col[d * kBwdPix + t]);
    }
  }
  // ---- 4b. g_ctx tile -> NCHW rows (256-byte segments) ------------------------------------------------
  {
    const int t = tid & (kBwdPix - 1), q = tid >> 6;
    const int Cc = m.C - (BSM ? bsm.Cs : 0);
    if (t < npx) {
      float *gc = g_context + (size_t)bn * m.gcs + p0 + t;
      for (int c = q; c < Cc; c += 4) stg_stream_f1(gc + (size_t)c * m.P, tile[c * kBwdLd + t]);
      if (BSM && q == 0) {
        // backward of semantic = softmax(logits) (the mask is not differentiable; masked pixels hold zeros):
        // g_logit_k = p_k (g_k - sum_j p_j g_j), channel order
        float dot = 0.0f;
        for (int k = 0; k < bsm.Cs; ++k) dot = __fmaf_rn(s_semp[k][t], tile[(Cc + k) * kBwdLd + t], dot);
        float *gs = g_semantic + (size_t)bn * bsm.Cs * m.P + p0 + t;
        for (int k = 0; k < bsm.Cs; ++k)
          stg_stream_f1(gs + (size_t)k * m.P, __fmul_rn(s_semp[k][t], __fsub_rn(tile[(Cc + k) * kBwdLd + t], dot)));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// g_height[d, pixel] = gw[run containing d], 0 for dropped bins; thread per pixel, coalesced rows.
// MODE 0: plain.  MODE 1: softmax backward fused (height holds logits):
//   g_logit[d] = p[d] * (g_p[d] - sum_d' p[d'] g_p[d'])   with  sum_d' p g_p = sum_runs w_run * gw_run.
// MODE 2: write the voxel id instead (plan_expand debug entry point).
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kChunk)
ls_expand_kernel(Dims m, const float *__restrict__ height, int vec16, const int *__restrict__ run_cnt,
                 const int *__restrict__ run_d, const int *__restrict__ run_vox,
                 const float *__restrict__ w_pm, const float *__restrict__ gw_pm,
                 float *__restrict__ g_height, int *__restrict__ vox_out) {
  extern __shared__ __align__(128) float col[];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int n = chunk / m.cpc, ci = chunk - n * m.cpc;
  const int frame_chunk = b * m.nchunks + chunk;
  const int t = threadIdx.x;
  const int p0 = ci * kChunk, p = p0 + t;
  if (MODE == 1) stage_columns(col, height + (size_t)(b * m.Nc + n) * m.hs, m.D, m.P, p0, vec16 != 0);
  if (p >= m.P) return;
  const int cnt = run_cnt[(size_t)frame_chunk * kChunk + t];
  float S = 0.0f, scale = 1.0f;
  if (MODE == 1) {
    scale = softmax_column(col, m.D, t);
    for (int r = 0; r < cnt; ++r) {
      const size_t s = ell_slot(frame_chunk, m.D, r, t);
      S = __fmaf_rn(w_pm[s], gw_pm[s], S);
    }
  }
  const size_t base = (MODE == 2 ? (size_t)(b * m.Nc + n) * m.D * m.P : (size_t)(b * m.Nc + n) * m.ghs) + p;
  int r = 0, d0 = m.D, d1 = m.D;
  float gv = 0.0f;
  int vv = -1;
  if (cnt > 0) {
    const size_t s = ell_slot(frame_chunk, m.D, 0, t);
    const int packed = run_d[s];
    d0 = packed & 0xffff; d1 = packed >> 16;
    if (MODE == 2) vv = run_vox[s]; else gv = gw_pm[s];
  }
  for (int d = 0; d < m.D; ++d) {
    const bool in = d >= d0 && d < d1;
    if (MODE == 2) vox_out[base + (size_t)d * m.P] = in ? vv : -1;
    else if (MODE == 1)
      stg_stream_f1(g_height + base + (size_t)d * m.P,
                    __fmul_rn(__fmul_rn(col[d * kChunk + t], scale), __fsub_rn(in ? gv : 0.0f, S)));
    else stg_stream_f1(g_height + base + (size_t)d * m.P, in ? gv : 0.0f);
    if (d + 1 == d1) {
      if (++r < cnt) {
        const size_t s = ell_slot(frame_chunk, m.D, r, t);
        const int packed = run_d[s];
        d0 = packed & 0xffff; d1 = packed >> 16;
        if (MODE == 2) vv = run_vox[s]; else gv = gw_pm[s];
      } else {
        d0 = d1 = m.D + 1;
      }
    }
  }
}

int check_ws(const Workspace &w, void *ws, size_t bytes, const char *who) {
  if (!ws || bytes < w.bytes) {
    set_error("%s: workspace %zu < required %zu bytes", who, bytes, w.bytes);
    return SGV3D_ERR_WORKSPACE_TOO_SMALL;
  }
  return SGV3D_OK;
}

// 2-D tensor map of the contiguous (B, C, Y, X) map viewed as (B * C) rows of V voxels, box = {32 voxels, C rows},
// SWIZZLE_128B (the reduce tile's layout).  The driver entry point is looked up once; without it the reduce keeps
// its LDS / STG copy-out.
bool make_bev_map(const Dims &m, float *bev, CUtensorMap *map) {
  typedef CUresult (*Encode)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static const Encode encode = [] {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (getenv("SGV3D_NO_TMA_STORE")) return (Encode) nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      fn = nullptr;
    return reinterpret_cast<Encode>(fn);
  }();
  if (!encode || m.V % 4 != 0 || m.C > 256 || reinterpret_cast<uintptr_t>(bev) % 16 != 0) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)m.V, (cuuint64_t)m.B * m.C};
  const cuuint64_t strides[1] = {(cuuint64_t)m.V * sizeof(float)};
  const cuuint32_t box[2] = {32u, (cuuint32_t)m.C};
  const cuuint32_t estr[2] = {1u, 1u};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, bev, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename CT, int G, int NV, int NSTR>
int launch_reduce_cfg(const Dims &m, const Workspace &w, float *bev, cudaStream_t s, bool skip) {
  dim3 grid(ceil_div(m.V, kTileV), m.B);
  // expected entries per touched tile ~ pixels * (runs per pixel ~ 25) / (touched tiles ~ 170 per 128 x 128
  // grid); stage up to ~2x that, within 1.5 K .. 6 K entries (12 .. 48 KB)
  long long expect = (long long)m.Nc * m.P * 25 / (m.ntiles * 2 / 3 + 1);
  int stage_cap = 1536;
  while (stage_cap < 2 * expect && stage_cap < 6144) stage_cap += 1536;
  const size_t smem = (size_t)2 * m.Cpad * 128 + sizeof(float) * NSTR * m.Cpad + sizeof(Entry) * stage_cap;
  // 128-bit output stores need 16-byte aligned voxel quads in every channel plane
  int vec_out = (m.V % 4 == 0) && (reinterpret_cast<uintptr_t>(bev) % 16 == 0);
  CUtensorMap bev_map;
  memset(&bev_map, 0, sizeof(bev_map));
  if (!m.cl && vec_out && make_bev_map(m, bev, &bev_map)) vec_out = 2;
  if (m.cl) {
    SGV3D_REQUIRE(reinterpret_cast<uintptr_t>(bev) % 16 == 0, "lift_splat_forward: channels-last BEV map must be 16-byte aligned");
    if (int rc = set_smem(ls_reduce_kernel<CT, G, NV, NSTR, true, false>, smem)) return rc;
    ls_reduce_kernel<CT, G, NV, NSTR, true, false><<<grid, NSTR * G, smem, s>>>(
        m, static_cast<const CT *>(w.ctxT), w.tile_ptr, w.vm_ent, bev, vec_out, stage_cap, bev_map);
  } else if (skip && sizeof(CT) == 4) {   // BSM call site (fp32 context): background pixels carry the weight -0.0f
    if (int rc = set_smem(ls_reduce_kernel<float, G, NV, NSTR, false, true>, smem)) return rc;
    ls_reduce_kernel<float, G, NV, NSTR, false, true><<<grid, NSTR * G, smem, s>>>(
        m, static_cast<const float *>(w.ctxT), w.tile_ptr, w.vm_ent, bev, vec_out, stage_cap, bev_map);
  } else {
    if (int rc = set_smem(ls_reduce_kernel<CT, G, NV, NSTR, false, false>, smem)) return rc;
    ls_reduce_kernel<CT, G, NV, NSTR, false, false><<<grid, NSTR * G, smem, s>>>(
        m, static_cast<const CT *>(w.ctxT), w.tile_ptr, w.vm_ent, bev, vec_out, stage_cap, bev_map);
  }
  SGV3D_CHECK_LAUNCH("ls_reduce_kernel");
  return SGV3D_OK;
}

template <typename CT, int G, int NV>
int launch_reduce_nstr(const Dims &m, const Workspace &w, float *bev, cudaStream_t s, bool skip) {
  // streams per tile by the expected tile population (pixels * ~25 runs / ~2/3 of the tiles touched): a tile is
  // one CTA, so the most populated tiles of a dense feature map (stride 8) set the kernel's tail
  const long long expect = (long long)m.Nc * m.P * 25 / (m.ntiles * 2 / 3 + 1);
  // (small batches are bound by the tail: twice the streams again; measured on DAIR-R50 and BSM-R50)
  const int nstr = expect > 6000 ? 128 : (expect > 1500 ? (m.B <= 8 ? 128 : 64) : (m.B <= 2 ? 64 : 32));
  if (nstr == 128) return launch_reduce_cfg<CT, G, NV, 128>(m, w, bev, s, skip);
  if (nstr == 64) return launch_reduce_cfg<CT, G, NV, 64>(m, w, bev, s, skip);
  return launch_reduce_cfg<CT, G, NV, 32>(m, w, bev, s, skip);
}

template <typename CT>
int launch_reduce(const Dims &m, const Workspace &w, float *bev, cudaStream_t s, bool skip = false) {
  if (m.G == 8) {
    switch (m.NV) {
      case 1: return launch_reduce_nstr<CT, 8, 1>(m, w, bev, s, skip);
      case 2: return launch_reduce_nstr<CT, 8, 2>(m, w, bev, s, skip);
      case 3: return launch_reduce_nstr<CT, 8, 3>(m, w, bev, s, skip);
      case 4: return launch_reduce_nstr<CT, 8, 4>(m, w, bev, s, skip);
      case 5: return launch_reduce_nstr<CT, 8, 5>(m, w, bev, s, skip);
      default: return launch_reduce_nstr<CT, 8, 6>(m, w, bev, s, skip);
    }
  }
  return launch_reduce_cfg<CT, 16, 4, 16>(m, w, bev, s, skip);
}

template <typename CT, int NV>
int launch_backward_chunk_cfg(const Dims &m, const Workspace &w, const float *height, const void *context,
                              float *grad_height, float *grad_context, cudaStream_t s, BsmAssembly bsm,
                              float *grad_semantic) {
  // the context tile doubles as D per-bin gradient rows of 65 floats: a few more rows when D <= 96 (D = 180 would cost a resident CTA)
  const int tile_rows = m.D <= 96 ? std::max(m.Cpad, m.D) : m.Cpad;
  const size_t smem = sizeof(float) * ((size_t)m.D * kBwdPix + (size_t)tile_rows * kBwdLd);
  // CTAs per SM: the kernel is latency bound (dependent gather -> FMA -> shuffle chains), so a third resident CTA pays
  // for the ~14 words of spill it costs at <= 80 channels (DAIR-R50: 440 -> 385 us at 64 frames); with 96-float rows
  // the spills dominate (SGV3D-BSM-R50: 737 -> 1117 us), so those keep two.  SGV3D_BWD_OCC overrides (experiments).
  const int vec16 = columns_vec16(height, m.hs, m.P) ? 1 : 0;
  const int vec16_out = (columns_vec16(grad_height, m.ghs, m.P) ? 1 : 0) | (tile_rows >= m.D ? 4 : 0);
  static const int occ_env = getenv("SGV3D_BWD_OCC") ? atoi(getenv("SGV3D_BWD_OCC")) : 0;
  const int occ = occ_env ? occ_env : (NV <= 3 ? 4 : (NV <= 5 ? 3 : 2));
#define SGV3D_BWD_CHUNK(OCC)                                                                                   \
  do {                                                                                                         \
    if (bsm.sem) {                                                                                             \
      if (int rc = set_smem(ls_backward_chunk_kernel<CT, NV, OCC, true, false>, smem)) return rc;              \
      ls_backward_chunk_kernel<CT, NV, OCC, true, false><<<dim3(2 * m.nchunks, m.B), kBwdPix * 4, smem, s>>>(  \
          m, height, vec16, vec16_out, static_cast<const CT *>(context), bsm, grad_semantic, w.gT, w.run_cnt,  \
          w.run_vox, w.run_d, w.w_pm, w.gw_pm, grad_height, grad_context);                                     \
      break;                                                                                                   \
    }                                                                                                          \
    if (m.cl) {                                                                                                \
      if (int rc = set_smem(ls_backward_chunk_kernel<CT, NV, OCC, false, true>, smem)) return rc;              \
      ls_backward_chunk_kernel<CT, NV, OCC, false, true><<<dim3(2 * m.nchunks, m.B), kBwdPix * 4, smem, s>>>(  \
          m, height, vec16, vec16_out, static_cast<const CT *>(context), bsm, grad_semantic, w.gT, w.run_cnt,  \
          w.run_vox, w.run_d, w.w_pm, w.gw_pm, grad_height, grad_context);                                     \
      break;                                                                                                   \
    }                                                                                                          \
    if (int rc = set_smem(ls_backward_chunk_kernel<CT, NV, OCC, false, false>, smem)) return rc;               \
    ls_backward_chunk_kernel<CT, NV, OCC, false, false><<<dim3(2 * m.nchunks, m.B), kBwdPix * 4, smem, s>>>(   \
        m, height, vec16, vec16_out, static_cast<const CT *>(context), bsm, grad_semantic, w.gT, w.run_cnt,    \
        w.run_vox, w.run_d, w.w_pm, w.gw_pm, grad_height, grad_context);                                       \
  } while (0)
  if (occ >= 4) SGV3D_BWD_CHUNK(4);
  else if (occ == 3) SGV3D_BWD_CHUNK(3);
  else SGV3D_BWD_CHUNK(2);
#undef SGV3D_BWD_CHUNK
  SGV3D_CHECK_LAUNCH("ls_backward_chunk_kernel");
  return SGV3D_OK;
}

template <typename CT>
int launch_backward_chunk(const Dims &m, const Workspace &w, const float *height, const void *context,
                          float *grad_height, float *grad_context, cudaStream_t s,
                          BsmAssembly bsm = BsmAssembly{nullptr, 0, 0, 0.0f}, float *grad_semantic = nullptr) {
  switch (m.NV) {
    case 1: return launch_backward_chunk_cfg<CT, 1>(m, w, height, context, grad_height, grad_context, s, bsm, grad_semantic);
    case 2: return launch_backward_chunk_cfg<CT, 2>(m, w, height, context, grad_height, grad_context, s, bsm, grad_semantic);
    case 3: return launch_backward_chunk_cfg<CT, 3>(m, w, height, context, grad_height, grad_context, s, bsm, grad_semantic);
    case 4: return launch_backward_chunk_cfg<CT, 4>(m, w, height, context, grad_height, grad_context, s, bsm, grad_semantic);
    case 5: return launch_backward_chunk_cfg<CT, 5>(m, w, height, context, grad_height, grad_context, s, bsm, grad_semantic);
    default: return launch_backward_chunk_cfg<CT, 6>(m, w, height, context, grad_height, grad_context, s, bsm, grad_semantic);
  }
}

template <typename CT>
int launch_backward_gather(const Dims &m, const Workspace &w, int gpad, cudaStream_t s) {
  dim3 grid(ceil_div(m.Nc * m.P, 8), m.B);
  const CT *ctxT = static_cast<const CT *>(w.ctxT);
  if (m.Cpad <= 128)
    ls_backward_gather_kernel<CT, 1><<<grid, 256, 0, s>>>(m, gpad, ctxT, w.gT, w.run_cnt, w.run_vox,
                                                          w.w_pm, w.gw_pm, w.gctxT);
  else
    ls_backward_gather_kernel<CT, 2><<<grid, 256, 0, s>>>(m, gpad, ctxT, w.gT, w.run_cnt, w.run_vox,
                                                          w.w_pm, w.gw_pm, w.gctxT);
  SGV3D_CHECK_LAUNCH("ls_backward_gather_kernel");
  return SGV3D_OK;
}

// 3-D tensor map of the fp32 context viewed as (B*Nc) cameras x C planes x P pixels (camera stride cs elements),
// box = {32 pixels, C planes, 1 camera}, SWIZZLE_128B.
bool make_context_map(const Dims &m, const void *context, CUtensorMap *map) {
  typedef CUresult (*Encode)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static const Encode encode = [] {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (getenv("SGV3D_NO_TMA_LOAD")) return (Encode) nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      fn = nullptr;
    return reinterpret_cast<Encode>(fn);
  }();
  if (!encode || m.P % 4 != 0 || m.cs % 4 != 0 || m.C > 256 || reinterpret_cast<uintptr_t>(context) % 16 != 0) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)m.P, (cuuint64_t)m.C, (cuuint64_t)m.B * m.Nc};
  const cuuint64_t strides[2] = {(cuuint64_t)m.P * sizeof(float), (cuuint64_t)m.cs * sizeof(float)};
  const cuuint32_t box[3] = {32u, (cuuint32_t)m.C, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(context), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// weights + context rows in one launch (forward and backward need both)
int launch_lift_prep(const Dims &m, const Workspace &w, int ctx_dtype, const float *height, const void *context,
                     bool forward, cudaStream_t s, BsmAssembly bsm = BsmAssembly{nullptr, 0, 0, 0.0f}) {
  Entry *vm_out = forward ? w.vm_ent : nullptr;
  dim3 grid(m.nchunks, m.B, 2);
  const size_t smem = sizeof(float) * (size_t)(m.D > m.C ? m.D * kChunk : m.C * (kChunk + 1));
  size_t smem2 = sizeof(float) * (size_t)kChunk * m.D > smem ? sizeof(float) * (size_t)kChunk * m.D : smem;
  const int vec16 = columns_vec16(height, m.hs, m.P) ? 1 : 0;
  // fp32 context of the LSSFPN call site: the context role takes its tiles through TMA tensor loads (four 1 KB-aligned
  // boxes of ceil(C / 8) KB each, plus the slack to align the first)
  CUtensorMap ctx_map;
  memset(&ctx_map, 0, sizeof(ctx_map));
  int ctx_tma = 0;
  // (small launches are latency bound and keep the cp.async role: one frame 15.5 vs 17.3 us; 64 frames 98.6 vs 92.8 us)
  if (ctx_dtype != SGV3D_DTYPE_BF16 && !bsm.sem && m.B * m.nchunks >= 4 * kNumSMs && make_context_map(m, context, &ctx_map)) {
    ctx_tma = 1;
    smem2 = std::max(smem2, (size_t)(kChunk / 32) * ((m.C + 7) / 8) * 1024 + 1024);
  }
  if (ctx_dtype == SGV3D_DTYPE_BF16) {
    if (int rc = set_smem(ls_lift_prep_kernel<__nv_bfloat16, false>, smem2)) return rc;
    ls_lift_prep_kernel<__nv_bfloat16, false><<<grid, kPrepThreads, smem2, s>>>(
        m, height, vec16, w.run_cnt, w.run_d, w.run_dst, w.w_pm, vm_out, static_cast<const __nv_bfloat16 *>(context),
        static_cast<__nv_bfloat16 *>(w.ctxT), row_perm(m), bsm, ctx_tma, ctx_map);
  } else {
    if (bsm.sem) {
      if (int rc = set_smem(ls_lift_prep_kernel<float, true>, smem2)) return rc;
      ls_lift_prep_kernel<float, true><<<grid, kPrepThreads, smem2, s>>>(m, height, vec16, w.run_cnt, w.run_d, w.run_dst, w.w_pm,
                                                                  vm_out, static_cast<const float *>(context),
                                                                  static_cast<float *>(w.ctxT), row_perm(m), bsm, ctx_tma, ctx_map);
    } else {
      if (int rc = set_smem(ls_lift_prep_kernel<float, false>, smem2)) return rc;
      ls_lift_prep_kernel<float, false><<<grid, kPrepThreads, smem2, s>>>(m, height, vec16, w.run_cnt, w.run_d, w.run_dst, w.w_pm,
                                                                   vm_out, static_cast<const float *>(context),
                                                                   static_cast<float *>(w.ctxT), row_perm(m), bsm, ctx_tma, ctx_map);
    }
  }
  SGV3D_CHECK_LAUNCH("ls_lift_prep_kernel");
  return SGV3D_OK;
}

int launch_backward_fused(const Dims &m, const Workspace &w, int ctx_dtype, const float *grad_bev,
                          const float *height, const void *context, float *grad_height, float *grad_context,
                          cudaStream_t s, BsmAssembly bsm = BsmAssembly{nullptr, 0, 0, 0.0f},
                          float *grad_semantic = nullptr) {
  // grad_bev -> one row per voxel, then everything else per pixel chunk in one kernel.  A channels-last gradient
  // (m.cl) already IS one row per voxel: the chunk kernel gathers from it directly, no copy.
  Workspace wg = w;
  if (m.cl) {
    SGV3D_REQUIRE(reinterpret_cast<uintptr_t>(grad_bev) % 16 == 0, "lift_splat_backward: channels-last grad_bev must be 16-byte aligned");
    wg.gT = const_cast<float *>(grad_bev);   // (read only by the chunk kernel)
  } else {
    CUtensorMap g_map;
    static const bool no_tma_load = getenv("SGV3D_NO_TMA_LOAD") != nullptr;
    if (!no_tma_load && make_bev_map(m, const_cast<float *>(grad_bev), &g_map)) {   // (same view as the forward's map)
      const size_t gsm = (size_t)((m.C + 7) / 8) * 2048;
      if (int rc = set_smem(ls_grad_rows_tma_kernel, gsm)) return rc;
      ls_grad_rows_tma_kernel<<<dim3(m.ntiles, m.B), 256, gsm, s>>>(m, w.tile_ptr, w.gT, row_perm(m), g_map);
      SGV3D_CHECK_LAUNCH("ls_grad_rows_tma_kernel");
    } else {
      const size_t gsm = sizeof(float) * (size_t)m.C * (kTileV + 1);
      if (int rc = set_smem(ls_grad_rows_kernel, gsm)) return rc;
      ls_grad_rows_kernel<<<dim3(m.ntiles, m.B), 256, gsm, s>>>(m, grad_bev, w.tile_ptr, w.gT, row_perm(m));
      SGV3D_CHECK_LAUNCH("ls_grad_rows_kernel");
    }
  }
  return ctx_dtype == SGV3D_DTYPE_BF16
             ? launch_backward_chunk<__nv_bfloat16>(m, wg, height, context, grad_height, grad_context, s)
             : launch_backward_chunk<float>(m, wg, height, context, grad_height, grad_context, s, bsm, grad_semantic);
}

// Which pipeline serves this descriptor: desc->reserved[0] = 0 (auto), 1 (voxel-tile pipeline of this file),
// 2 (pixel-block pipeline of lift_splat_block.cu, required).  Auto = voxel-tile: measured on B200 (profiles/README.md,
// round 2) it is the faster one for every reference shape -- DAIR-R50, 64 frames: plan + forward 452 vs 560 us,
// backward 470 vs 790 us -- although the pixel-block pipeline plans in one kernel (125 vs 194 us) and moves
// close to the algorithmic bytes; it stays selectable (tests run both) until its block kernels catch up.
bool use_block(const sgv3d_lift_splat_desc *desc, const Dims &m) {
  return desc->reserved[0] == 2 && block::supported(m);
}
// the pixel-block pipeline's workspace follows the voxel-tile pipeline's
void *block_ws(void *workspace, const Dims &m, int ctx_dtype) {
  return static_cast<char *>(workspace) + carve(nullptr, m, ctx_dtype).bytes;
}
int check_pipeline(const sgv3d_lift_splat_desc *desc, const Dims &m, const char *who) {
  SGV3D_REQUIRE(desc->reserved[0] >= 0 && desc->reserved[0] <= 2, "%s: bad pipeline selector %d", who, desc->reserved[0]);
  SGV3D_REQUIRE(desc->reserved[0] != 2 || block::supported(m), "%s: the pixel-block pipeline does not support this shape", who);
  SGV3D_REQUIRE((desc->reserved[1] & ~2) == 0, "%s: reserved[1] = %d", who, desc->reserved[1]);
  SGV3D_REQUIRE(!m.cl || (desc->reserved[0] != 2 && m.C % 16 == 0 && m.C <= 96),
                "%s: the channels-last BEV layout needs the voxel-tile pipeline and C in {16, 32, ..., 96}", who);
  return SGV3D_OK;
}
geom::Grid make_grid(const Dims &m, const float *lower3, const float *size3) {
  geom::Grid grid;
  for (int k = 0; k < 3; ++k) { grid.lower[k] = lower3[k]; grid.size[k] = size3[k]; }
  grid.X = m.X; grid.Y = m.Y; grid.Z = m.Z;
  geom::z_thresholds(grid.size[2], m.Z, &grid.zt_lo, &grid.zt_hi);
  grid.rcp_size[0] = 1.0f / grid.size[0];  // IEEE single division on the host: correctly rounded
  grid.rcp_size[1] = 1.0f / grid.size[1];
  return grid;
}

}  // namespace
}  // namespace sgv3d

using namespace sgv3d;

extern "C" int sgv3d_lift_splat_uses_block_pipeline(const sgv3d_lift_splat_desc *desc) {
  if (validate(desc, "lift_splat_uses_block_pipeline") != SGV3D_OK) return 0;
  return use_block(desc, make_dims(desc)) ? 1 : 0;
}

extern "C" size_t sgv3d_lift_splat_workspace_bytes(const sgv3d_lift_splat_desc *desc) {
  if (validate(desc, "lift_splat_workspace_bytes") != SGV3D_OK || desc->B == 0) return 0;
  const Dims m = make_dims(desc);
  return carve(nullptr, m, desc->ctx_dtype).bytes + (block::supported(m) ? block::workspace_bytes(m) : 0);
}

extern "C" int sgv3d_lift_splat_plan(const sgv3d_lift_splat_desc *desc, const float *u_tab,
                                     const float *v_tab, const float *z_tab, const float *ida_inv,
                                     const float *m_virtual, const float *m_ego, const float *bda,
                                     const float *ref_heights, const float *lower3,
                                     const float *size3, void *workspace, size_t workspace_bytes,
                                     sgv3d_stream_t stream) {
  if (int rc = validate(desc, "lift_splat_plan")) return rc;
  if (desc->B == 0) return SGV3D_OK;
  SGV3D_REQUIRE(u_tab && v_tab && z_tab && ida_inv && m_virtual && m_ego && ref_heights && lower3 && size3,
                "lift_splat_plan: null pointer");
  SGV3D_REQUIRE(size3[0] > 0.f && size3[1] > 0.f && size3[2] > 0.f, "lift_splat_plan: voxel size must be > 0");
  const Dims m = make_dims(desc);
  if (int rc = check_pipeline(desc, m, "lift_splat_plan")) return rc;
  Workspace w = carve(workspace, m, desc->ctx_dtype);
  if (block::supported(m)) w.bytes += block::workspace_bytes(m);
  if (int rc = check_ws(w, workspace, workspace_bytes, "lift_splat_plan")) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  prof_begin(s);
  const geom::Grid grid = make_grid(m, lower3, size3);
  if (use_block(desc, m))
    return block::plan(m, desc->arith, u_tab, v_tab, z_tab, ida_inv, m_virtual, m_ego, bda, ref_heights, grid,
                       block_ws(workspace, m, desc->ctx_dtype), s);

  dim3 gc(m.nchunks, m.B);
  const int gen_grid = std::min(m.nchunks * m.B, 5 * kNumSMs);  // 5 CTAs of the general kernel fit on an SM
  const size_t zsm = sizeof(float) * m.D + sizeof(int) * m.ntiles;
  const size_t zsm_fast = zsm + sizeof(float) * m.D;
#define SGV3D_PLAN_RUNS(A)                                                                                  \
  do {                                                                                                      \
    ls_plan_runs_fast_kernel<A><<<gc, kChunk, zsm_fast, s>>>(m, u_tab, v_tab, z_tab, ida_inv, m_virtual, m_ego,   \
                                                        bda, ref_heights, grid, w.run_cnt, w.run_vox,       \
                                                        w.run_d, w.hist, w.chunk_done);                     \
    SGV3D_CHECK_LAUNCH("ls_plan_runs_fast_kernel");                                                         \
    ls_plan_runs_kernel<A><<<gen_grid, kChunk, zsm, s>>>(m, u_tab, v_tab, z_tab, ida_inv, m_virtual, m_ego, bda,   \
                                                   ref_heights, grid, w.run_cnt, w.run_vox, w.run_d,        \
                                                   w.hist, w.chunk_done);                                   \
  } while (0)
  if (desc->arith == SGV3D_ARITH_PAIR) SGV3D_PLAN_RUNS(SGV3D_ARITH_PAIR);
  else if (desc->arith == SGV3D_ARITH_FMA) SGV3D_PLAN_RUNS(SGV3D_ARITH_FMA);
  else SGV3D_PLAN_RUNS(SGV3D_ARITH_SEQ);
#undef SGV3D_PLAN_RUNS
  SGV3D_CHECK_LAUNCH("ls_plan_runs_kernel");
  const size_t csm = sizeof(int) * ((kChunk / 32) * m.ntiles + kScatRows * kChunk + m.ntiles);
  if (m.ntiles <= kFusedScanMaxTiles && m.nchunks * m.ntiles <= kFusedScanMaxCells) {
    if (int rc = set_smem(ls_scatter_tiles_kernel<true>, csm)) return rc;
    ls_scatter_tiles_kernel<true><<<gc, kChunk, csm, s>>>(m, w.run_cnt, w.run_vox, w.hist, w.tile_ptr, w.bucket);
  } else {
    ls_scan_tiles_kernel<<<m.B, kScanThreads, 0, s>>>(m, w.hist, w.tile_ptr);
    SGV3D_CHECK_LAUNCH("ls_scan_tiles_kernel");
    if (int rc = set_smem(ls_scatter_tiles_kernel<false>, csm)) return rc;
    ls_scatter_tiles_kernel<false><<<gc, kChunk, csm, s>>>(m, w.run_cnt, w.run_vox, w.hist, w.tile_ptr, w.bucket);
  }
  SGV3D_CHECK_LAUNCH("ls_scatter_tiles_kernel");
  // heavy tiles (dense feature maps) get more warps: the kernel's tail is its most populated tile
  if ((long long)m.Nc * m.P * 25 / (m.ntiles * 2 / 3 + 1) > 1024)
    ls_finish_tiles_kernel<16><<<dim3(m.ntiles, m.B), 16 * 32, 0, s>>>(m, w.tile_ptr, w.bucket, w.vm_ent, w.run_dst);
  else
    ls_finish_tiles_kernel<4><<<dim3(m.ntiles, m.B), 4 * 32, 0, s>>>(m, w.tile_ptr, w.bucket, w.vm_ent, w.run_dst);
  SGV3D_CHECK_LAUNCH("ls_finish_tiles_kernel");
  return SGV3D_OK;
}

extern "C" int sgv3d_lift_splat_forward(const sgv3d_lift_splat_desc *desc, const float *height,
                                        const void *context, float *bev, void *workspace,
                                        size_t workspace_bytes, sgv3d_stream_t stream) {
  if (int rc = validate(desc, "lift_splat_forward")) return rc;
  if (desc->B == 0) return SGV3D_OK;
  SGV3D_REQUIRE(height && context && bev, "lift_splat_forward: null pointer");
  const Dims m = make_dims(desc);
  const Workspace w = carve(workspace, m, desc->ctx_dtype);
  if (int rc = check_ws(w, workspace, workspace_bytes, "lift_splat_forward")) return rc;
  if (int rc = check_pipeline(desc, m, "lift_splat_forward")) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  prof_begin(s);
  if (use_block(desc, m))
    return block::forward(m, desc->ctx_dtype, height, context, nullptr, 0, 0, 0.0f, bev,
                          block_ws(workspace, m, desc->ctx_dtype), s);
  if (int rc = launch_lift_prep(m, w, desc->ctx_dtype, height, context, true, s)) return rc;
  if (desc->ctx_dtype == SGV3D_DTYPE_BF16) return launch_reduce<__nv_bfloat16>(m, w, bev, s);
  return launch_reduce<float>(m, w, bev, s);
}

extern "C" int sgv3d_lift_splat_forward_bsm(const sgv3d_lift_splat_desc *desc, const float *height,
                                            const float *context, const float *semantic_logits,
                                            int semantic_channels, int64_t semantic_batch_stride,
                                            float background_threshold, float *bev, void *workspace,
                                            size_t workspace_bytes, sgv3d_stream_t stream) {
  if (int rc = validate(desc, "lift_splat_forward_bsm")) return rc;
  if (desc->B == 0) return SGV3D_OK;
  SGV3D_REQUIRE(height && context && semantic_logits && bev, "lift_splat_forward_bsm: null pointer");
  SGV3D_REQUIRE(desc->ctx_dtype == SGV3D_DTYPE_F32, "lift_splat_forward_bsm: fp32 context only");
  SGV3D_REQUIRE(semantic_channels > 0 && semantic_channels < desc->C,
                "lift_splat_forward_bsm: semantic_channels must be in (0, C)");
  SGV3D_REQUIRE(semantic_batch_stride >= 0, "lift_splat_forward_bsm: negative batch stride");
  Dims m = make_dims(desc);
  // `context` holds the C - Cs feature channels: its dense camera block is that much smaller
  if (!desc->ctx_batch_stride) m.cs = (long long)(m.C - semantic_channels) * m.P;
  const Workspace w = carve(workspace, m, desc->ctx_dtype);
  if (int rc = check_ws(w, workspace, workspace_bytes, "lift_splat_forward_bsm")) return rc;
  if (int rc = check_pipeline(desc, m, "lift_splat_forward_bsm")) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  prof_begin(s);
  if (use_block(desc, m))
    return block::forward(m, desc->ctx_dtype, height, context, semantic_logits, semantic_channels,
                          semantic_batch_stride ? semantic_batch_stride : (long long)semantic_channels * m.P,
                          background_threshold, bev, block_ws(workspace, m, desc->ctx_dtype), s);
  BsmAssembly bsm;
  bsm.sem = semantic_logits;
  bsm.sem_stride = semantic_batch_stride ? semantic_batch_stride : (long long)semantic_channels * m.P;
  bsm.Cs = semantic_channels;
  bsm.thr = background_threshold;
  if (int rc = launch_lift_prep(m, w, desc->ctx_dtype, height, context, true, s, bsm)) return rc;
  return launch_reduce<float>(m, w, bev, s, /*skip background pixels' rows*/ true);
}

extern "C" int sgv3d_lift_splat_backward(const sgv3d_lift_splat_desc *desc, const float *grad_bev,
                                         const float *height, const void *context,
                                         float *grad_height, float *grad_context, void *workspace,
                                         size_t workspace_bytes, sgv3d_stream_t stream) {
  if (int rc = validate(desc, "lift_splat_backward")) return rc;
  if (desc->B == 0) return SGV3D_OK;
  SGV3D_REQUIRE(grad_bev && height && context && grad_height && grad_context,
                "lift_splat_backward: null pointer");
  const Dims m = make_dims(desc);
  const Workspace w = carve(workspace, m, desc->ctx_dtype);
  if (int rc = check_ws(w, workspace, workspace_bytes, "lift_splat_backward")) return rc;
  if (int rc = check_pipeline(desc, m, "lift_splat_backward")) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  prof_begin(s);
  if (use_block(desc, m))
    return block::backward(m, desc->ctx_dtype, grad_bev, height, context, grad_height, grad_context,
                           block_ws(workspace, m, desc->ctx_dtype), s);
  const int gpad = m.Cpad;  // gradient rows share the context rows' (permuted) channel layout
  if (m.C <= 96) {
    // fused path: 4-lane gradient rows of its own (16 * ceil(C / 16) floats <= Cpad, so gT is large enough)
    Dims mb = m;
    mb.G = 4; mb.NV = ceil_div(m.C, 16); mb.Cpad = 16 * mb.NV;
    return launch_backward_fused(mb, w, desc->ctx_dtype, grad_bev, height, context, grad_height, grad_context, s);
  }
  if (int rc = launch_lift_prep(m, w, desc->ctx_dtype, height, context, false, s)) return rc;
  launch_transpose_pad<float, float, 1>(grad_bev, w.gT, m.B, m.C, m.V, m.V, (size_t)m.C * m.V, gpad,
                                        (size_t)m.V * gpad, s, row_perm(m));
  SGV3D_CHECK_LAUNCH("transpose_pad_kernel(grad_bev)");
  int rc = desc->ctx_dtype == SGV3D_DTYPE_BF16 ? launch_backward_gather<__nv_bfloat16>(m, w, gpad, s)
                                                : launch_backward_gather<float>(m, w, gpad, s);
  if (rc) return rc;
  dim3 gc(m.nchunks, m.B);
  if (m.logits) {
    const size_t smem = sizeof(float) * m.D * kChunk;
    if (int rc2 = set_smem(ls_expand_kernel<1>, smem)) return rc2;
    ls_expand_kernel<1><<<gc, kChunk, smem, s>>>(m, height, columns_vec16(height, m.hs, m.P) ? 1 : 0, w.run_cnt,
                                                 w.run_d, w.run_vox, w.w_pm, w.gw_pm, grad_height, nullptr);
  } else {
    ls_expand_kernel<0><<<gc, kChunk, 0, s>>>(m, nullptr, 0, w.run_cnt, w.run_d, w.run_vox, w.w_pm, w.gw_pm,
                                              grad_height, nullptr);
  }
  SGV3D_CHECK_LAUNCH("ls_expand_kernel");
  launch_transpose_pad<float, float, 2>(w.gctxT, grad_context, m.B * m.Nc, m.P, m.C, gpad,
                                        (size_t)m.P * gpad, m.P, (size_t)m.gcs, s, row_perm(m));
  SGV3D_CHECK_LAUNCH("transpose_pad_kernel(grad_context)");
  return SGV3D_OK;
}

extern "C" int sgv3d_lift_splat_backward_bsm(const sgv3d_lift_splat_desc *desc, const float *grad_bev,
                                             const float *height, const float *context,
                                             const float *semantic_logits, int semantic_channels,
                                             int64_t semantic_batch_stride, float background_threshold,
                                             float *grad_height, float *grad_context, float *grad_semantic,
                                             void *workspace, size_t workspace_bytes, sgv3d_stream_t stream) {
  if (int rc = validate(desc, "lift_splat_backward_bsm")) return rc;
  if (desc->B == 0) return SGV3D_OK;
  SGV3D_REQUIRE(grad_bev && height && context && semantic_logits && grad_height && grad_context && grad_semantic,
                "lift_splat_backward_bsm: null pointer");
  SGV3D_REQUIRE(desc->ctx_dtype == SGV3D_DTYPE_F32, "lift_splat_backward_bsm: fp32 context only");
  SGV3D_REQUIRE(semantic_channels > 0 && semantic_channels <= 8 && semantic_channels < desc->C,
                "lift_splat_backward_bsm: semantic_channels must be in [1, 8] and < C");
  SGV3D_REQUIRE(desc->C <= 96, "lift_splat_backward_bsm: C <= 96 (fused backward)");
  SGV3D_REQUIRE(semantic_batch_stride >= 0, "lift_splat_backward_bsm: negative batch stride");
  Dims m = make_dims(desc);
  SGV3D_REQUIRE(!use_block(desc, m), "lift_splat_backward_bsm: voxel-tile pipeline only");
  SGV3D_REQUIRE(desc->reserved[1] == 0, "lift_splat_backward_bsm: NCHW BEV gradient only");
  const int Cc = m.C - semantic_channels;
  // `context` / `grad_context` hold the C - Cs feature channels: their dense camera blocks are that much smaller
  if (!desc->ctx_batch_stride) m.cs = (long long)Cc * m.P;
  if (!desc->grad_ctx_batch_stride) m.gcs = (long long)Cc * m.P;
  const Workspace w = carve(workspace, m, desc->ctx_dtype);
  if (int rc = check_ws(w, workspace, workspace_bytes, "lift_splat_backward_bsm")) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  prof_begin(s);
  BsmAssembly bsm;
  bsm.sem = semantic_logits;
  bsm.sem_stride = semantic_batch_stride ? semantic_batch_stride : (long long)semantic_channels * m.P;
  bsm.Cs = semantic_channels;
  bsm.thr = background_threshold;
  Dims mb = m;
  mb.G = 4; mb.NV = ceil_div(m.C, 16); mb.Cpad = 16 * mb.NV;
  return launch_backward_fused(mb, w, desc->ctx_dtype, grad_bev, height, context, grad_height, grad_context, s, bsm,
                               grad_semantic);
}

extern "C" int sgv3d_lift_splat_plan_expand(const sgv3d_lift_splat_desc *desc, int32_t *vox_out,
                                            void *workspace, size_t workspace_bytes,
                                            sgv3d_stream_t stream) {
  if (int rc = validate(desc, "lift_splat_plan_expand")) return rc;
  if (desc->B == 0) return SGV3D_OK;
  SGV3D_REQUIRE(vox_out != nullptr, "lift_splat_plan_expand: null pointer");
  const Dims m = make_dims(desc);
  const Workspace w = carve(workspace, m, desc->ctx_dtype);
  if (int rc = check_ws(w, workspace, workspace_bytes, "lift_splat_plan_expand")) return rc;
  if (int rc = check_pipeline(desc, m, "lift_splat_plan_expand")) return rc;
  dim3 gc(m.nchunks, m.B);
  prof_begin(static_cast<cudaStream_t>(stream));
  if (use_block(desc, m))
    return block::plan_expand(m, vox_out, block_ws(workspace, m, desc->ctx_dtype), static_cast<cudaStream_t>(stream));
  ls_expand_kernel<2><<<gc, kChunk, 0, static_cast<cudaStream_t>(stream)>>>(
      m, nullptr, 0, w.run_cnt, w.run_d, w.run_vox, nullptr, nullptr, nullptr, vox_out);
  SGV3D_CHECK_LAUNCH("ls_expand_kernel(vox)");
  return SGV3D_OK;
}
