// Pixel-block pipeline of the fused lift-splat (B200, sm_100a).  Replaces, for rows of <= 96 channels and D <= 255,
// the voxel-tile pipeline of lift_splat.cu (global sort of the runs by voxel + L2 row gathers) with work that is
// local to 8 x 8 PIXEL BLOCKS, whose BEV footprint is small (~150 voxels at DAIR-R50, ~110 at SGV3D-BSM-R50:
// rays run along x, 64 neighbouring rays cover a narrow strip) and therefore fits in shared memory:
//
//   PLAN      bp_plan_kernel       CTA = pixel block, thread = pixel.  Bit-exact geometry (geometry.cuh; guarded linear
//                                  walk, ~14 instructions per height bin) -> voxel runs along D (ELL layout) and the
//                                  block's FOOTPRINT: one bit per voxel in a per-block bitmap over 32-voxel strips
//                                  (smem atomicOr: order independent), its popcount scan = dense "slot" numbers of the
//                                  touched voxels, the slot of every run, and a compact range of partial-sum rows.
//                                  No global sort, no histogram scan, no scatter.
//   FORWARD   bp_forward_kernel    CTA = pixel block, 8 lanes per pixel.  Height columns and context rows staged in
//                                  smem once (softmax over D fused); run weights accumulate into a slots x pixels
//                                  matrix W and a per-slot pixel bitmask; each 8-lane group then sums
//                                  sum_t W[s][t] * ctx_row[t] for its slots in pixel order (rows read from smem, FFMA2)
//                                  and writes one channels-last partial row per (block, slot).
//             bp_combine_kernel    CTA = strip of 32 voxels: adds the partial rows of the blocks whose footprint holds
//                                  the voxel, in block order, transposes through smem and writes NCHW lines once.
//   BACKWARD  bp_backward_kernel   CTA = pixel block.  grad_bev rows of the footprint staged in smem straight from the
//                                  NCHW gradient (no transposed copy in HBM), then per run: g_ctx += w G[slot],
//                                  gw = <ctx, G[slot]> from smem, softmax backward, coalesced g_height / g_ctx stores.
//
// Every sum has a fixed order (pixel order inside a slot, block order inside a voxel, run order inside a pixel):
// bitwise reproducible, no floating-point atomics on the data path.  (The only exceptions are degenerate geometries:
// a pixel whose ray re-enters a voxel adds its two weights with a shared-memory atomic -- two addends commute --
// and a frame whose footprints exceed the partial-row budget is completed by bp_fixup_kernel with global atomics.)
//
// Reference: layers/backbones/lss_fpn.py:462-495 (bsm_lss_fpn.py:523-559), ops/voxel_pooling/voxel_pooling.py:9-69,
// ops/voxel_pooling/src/voxel_pooling_forward_cuda.cu:9-36.
#include <cuda_bf16.h>

#include <algorithm>

#include "geometry.cuh"
#include "ls_block.cuh"
#include "ls_shared.cuh"
#include "transpose.cuh"

namespace sgv3d {
namespace block {
namespace {

constexpr int kBP = 64;                 // pixels per block (8 x 8)
constexpr int kFwdThreads = 8 * kBP;    // 8 lanes per pixel
constexpr int kSlotShift = 18, kDMask = 511;  // run descriptor: d0 | d1 << 9 | slot << 18

struct BDims {
  int B, Nc, D, fH, fW, C, X, Y, Z;
  int P, V;
  int nbw, nbh, nblk;   // pixel blocks per camera
  int NB;               // pixel blocks per frame = Nc * nblk
  int nstrips;          // ceil(V / 32)
  int NV, Cpad;         // row layout: 8 lanes x NV float4 (Cpad = 32 * NV floats)
  int esize;
  int logits;
  int rows_cap;         // partial-sum rows per frame
  long long hs, cs, ghs, gcs;
};

BDims make_bdims(const Dims &m) {
  BDims b;
  b.B = m.B; b.Nc = m.Nc; b.D = m.D; b.fH = m.fH; b.fW = m.fW; b.C = m.C; b.X = m.X; b.Y = m.Y; b.Z = m.Z;
  b.P = m.P; b.V = m.V;
  b.nbw = ceil_div(m.fW, 8); b.nbh = ceil_div(m.fH, 8); b.nblk = b.nbw * b.nbh;
  b.NB = m.Nc * b.nblk;
  b.nstrips = ceil_div(m.V, 32);
  b.NV = ceil_div(m.C, 32); b.Cpad = 32 * b.NV;
  b.esize = m.esize;
  b.logits = m.logits;
  // partial rows: measured 0.65 V (DAIR-R50) .. 2 V (SGV3D-BSM-R50); budget 3 V + 3 pixels per frame, capped by the
  // number of height-bin slots (every run its own voxel).  Overflow is handled (bp_fixup_kernel), never silent.
  const long long cap = std::min<long long>((long long)m.Nc * m.P * m.D, 3ll * m.V + 3ll * m.Nc * m.P);
  b.rows_cap = (int)std::max<long long>(cap, 64);
  b.hs = m.hs; b.cs = m.cs; b.ghs = m.ghs; b.gcs = m.gcs;
  return b;
}

struct BlockInfo {   // one per (frame, pixel block)
  int nslots;        // voxels in the footprint
  int row_base;      // first partial row of the block inside its frame's row pool, -1: pool exhausted
  int max_runs;      // longest run list of the block's pixels
  int pad;
};

struct BWorkspace {
  int *cnt;            // [B*NB][64]      runs per pixel
  int *vox;            // [B*NB][D][64]   voxel id of run r of pixel t
  int *rd;             // [B*NB][D][64]   d0 | d1 << 9 | slot << 18
  BlockInfo *info;     // [B*NB]
  unsigned *fmask;     // [B*NB][nstrips] footprint bitmap
  unsigned *fbase;     // [B*NB][nstrips] slot of the strip's first touched voxel
  int *alloc;          // [B] rows handed out per frame, then [B] overflow flags
  float *prow;         // [B][rows_cap][Cpad] partial sums
  float *gw;           // [B*NB][D][64] backward scratch: d BEV . ctx per run
  size_t bytes;
};

BWorkspace carve(void *ws, const BDims &m) {
  BWorkspace w;
  Carver c(ws);
  const size_t nb = (size_t)m.B * m.NB;
  w.cnt = c.take<int>(nb * kBP);
  w.vox = c.take<int>(nb * m.D * kBP);
  w.rd = c.take<int>(nb * m.D * kBP);
  w.info = reinterpret_cast<BlockInfo *>(c.take<int4>(nb));
  w.fmask = c.take<unsigned>(nb * m.nstrips);
  w.fbase = c.take<unsigned>(nb * m.nstrips);
  w.alloc = c.take<int>(2 * (size_t)m.B);
  w.prow = c.take<float>((size_t)m.B * m.rows_cap * m.Cpad);
  w.gw = c.take<float>(nb * m.D * kBP);
  w.bytes = c.used();
  return w;
}

// pixel t of block blk (inside its camera): image position, validity, linear pixel index
struct Pix {
  int h, w, p;
  bool valid;
};
__device__ __forceinline__ Pix pixel_of(const BDims &m, int blk, int t) {
  const int bi = blk / m.nbw, bj = blk - bi * m.nbw;
  Pix x;
  x.h = bi * 8 + (t >> 3);
  x.w = bj * 8 + (t & 7);
  x.valid = x.h < m.fH && x.w < m.fW;
  x.p = x.h * m.fW + x.w;
  return x;
}

// ---------------------------------------------------------------------------------------------
// PLAN.  grid (NB, B), 64 threads.
// ---------------------------------------------------------------------------------------------
template <int ARITH>
__global__ void __launch_bounds__(kBP, 12)
bp_plan_kernel(BDims m, const float *__restrict__ u_tab, const float *__restrict__ v_tab,
               const float *__restrict__ z_tab, const float *__restrict__ ida_inv, const float *__restrict__ mv,
               const float *__restrict__ me, const float *__restrict__ bda, const float *__restrict__ ref_h,
               geom::Grid grid, int *__restrict__ cnt_out, int *__restrict__ vox_out, int *__restrict__ rd_out,
               BlockInfo *__restrict__ info, unsigned *__restrict__ fmask, unsigned *__restrict__ fbase,
               int *__restrict__ alloc) {
  __shared__ geom::Camera cam;
  __shared__ int s_flags[4];   // 0: camera qualifies for the fast path, 1: z table finite, 2: hgt table valid
  __shared__ int s_red[2][2];
  extern __shared__ float psm[];
  float *zs = psm;                                             // [D] height-bin values
  float *hs = psm + m.D;                                       // [D] camera height above the bin's plane
  unsigned *fm = reinterpret_cast<unsigned *>(psm + 2 * m.D);  // [nstrips] footprint bitmap
  unsigned *fb = fm + m.nstrips;                               // [nstrips] slot bases
  const int b = blockIdx.y, kb = blockIdx.x;
  const int n = kb / m.nblk, blk = kb - n * m.nblk;
  const int bn = b * m.Nc + n;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const size_t fbk = (size_t)b * m.NB + kb;
  geom::load_camera(&cam, ida_inv, mv, me, bda, ref_h, bn, b);
  bool z_ok = true;
  float zmin = INFINITY, zmax = -INFINITY;
  for (int d = t; d < m.D; d += kBP) {
    const float z = z_tab[d];
    zs[d] = z;
    z_ok = z_ok && (fabsf(z) < INFINITY);
    zmin = fminf(zmin, z);
    zmax = fmaxf(zmax, z);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    zmin = fminf(zmin, __shfl_xor_sync(0xffffffffu, zmin, o));
    zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
  }
  if (lane == 0) { s_red[0][wid] = __float_as_int(zmin); s_red[1][wid] = __float_as_int(zmax); }
  for (int i = t; i < m.nstrips; i += kBP) fm[i] = 0u;
  const int all_z_ok = __syncthreads_and(z_ok);
  zmin = fminf(__int_as_float(s_red[0][0]), __int_as_float(s_red[0][1]));
  zmax = fmaxf(__int_as_float(s_red[1][0]), __int_as_float(s_red[1][1]));
  if (t == 0) {
    s_flags[0] = (all_z_ok && geom::camera_is_fast(cam)) ? 1 : 0;
    // row 2 of ida^-1 ignores (u, v): the bin heights are per camera, not per pixel
    s_flags[2] = (cam.A[8] == 0.0f && cam.A[9] == 0.0f) ? 1 : 0;
  }
  __syncthreads();
  const bool cam_fast = s_flags[0] != 0;
  const bool h_uniform = s_flags[2] != 0;
  if (cam_fast && h_uniform) {
    for (int d = t; d < m.D; d += kBP) {
      const float p0z = geom::dot2_tail<ARITH>(0.0f, cam.A + 8, zs[d], 1.0f);
      hs[d] = __fadd_rn(__fmul_rn(-1.0f, p0z), cam.ref_h);
    }
  }
  __syncthreads();

  const Pix px = pixel_of(m, blk, t);
  int *const rv = vox_out + fbk * m.D * kBP;
  int *const rdp = rd_out + fbk * m.D * kBP;
  int r = 0;
  // run emission: run r of this pixel sits at element r * 64 + t of the block's ELL tables
  int cur = -1, d0 = 0, eo = t;
  auto step = [&](int d, int vox) {
    if (vox != cur) {
      if (cur >= 0) {
        rv[eo] = cur;
        rdp[eo] = d0 | (d << 9);
        eo += kBP;
        atomicOr(&fm[cur >> 5], 1u << (cur & 31));
        ++r;
      }
      cur = vox;
      d0 = d;
    }
  };
  bool general = !cam_fast;
  for (int pass = 0; pass < 2; ++pass) {
    bool bad = false;
    r = 0; cur = -1; d0 = 0; eo = t;
    if (px.valid) {
      const float u = u_tab[px.w], v = v_tab[px.h];
      if (general) {
        geom::PixelRay<ARITH> ray;
        ray.init(cam, u, v);
        for (int d = 0; d < m.D; ++d) step(d, ray.voxel(cam, grid, zs[d]));
      } else {
        bad = !geom::walk_fast<ARITH>(cam, grid, zs, hs, h_uniform, m.D, u, v, zmin, zmax, step);
      }
      step(m.D, -2);  // sentinel closes the last run
    }
    if (general) break;
    if (!__syncthreads_or(bad)) break;
    // a point of this block left the domain in which the fast path is proven exact: redo the block with the general chain
    for (int i = t; i < m.nstrips; i += kBP) fm[i] = 0u;
    general = true;
    __syncthreads();
  }
  cnt_out[fbk * kBP + t] = r;
  int mr = r;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mr = max(mr, __shfl_xor_sync(0xffffffffu, mr, o));
  __syncthreads();  // footprint bitmap complete
  // slots: exclusive scan of the strips' popcounts (thread t owns a contiguous range of strips)
  const int spt = (m.nstrips + kBP - 1) / kBP;
  const int s_lo = min(t * spt, m.nstrips), s_hi = min(s_lo + spt, m.nstrips);
  int mine = 0;
  for (int i = s_lo; i < s_hi; ++i) mine += __popc(fm[i]);
  int x = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_red[0][wid] = x;
  if (lane == 0) s_red[1][wid] = mr;
  __syncthreads();
  int base = x - mine + (wid == 1 ? s_red[0][0] : 0);
  const int nslots = s_red[0][0] + s_red[0][1];
  for (int i = s_lo; i < s_hi; ++i) {
    fb[i] = (unsigned)base;
    base += __popc(fm[i]);
  }
  if (t == 0) {
    BlockInfo bi;
    bi.nslots = nslots;
    bi.max_runs = max(s_red[1][0], s_red[1][1]);
    bi.pad = 0;
    bi.row_base = -1;
    if (nslots > 0) {
      const int at = atomicAdd(&alloc[b], nslots);   // where the rows live does not affect any value
      if (at + nslots <= m.rows_cap) bi.row_base = at;
      else atomicOr(&alloc[m.B + b], 1);             // pool exhausted: this block is completed by bp_fixup_kernel
    } else {
      bi.row_base = 0;
    }
    info[fbk] = bi;
  }
  __syncthreads();
  unsigned *gm = fmask + fbk * m.nstrips, *gbs = fbase + fbk * m.nstrips;
  for (int i = t; i < m.nstrips; i += kBP) {
    gm[i] = fm[i];
    gbs[i] = fb[i];
  }
  // slot of every run (each thread re-reads what it wrote itself)
  for (int r0 = 0; r0 < r; r0 += 4) {
    int vv[4], dd[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      vv[k] = 0; dd[k] = 0;
      if (r0 + k < r) { vv[k] = rv[(r0 + k) * kBP + t]; dd[k] = rdp[(r0 + k) * kBP + t]; }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (r0 + k < r) {
        const unsigned w = fm[vv[k] >> 5];
        const int slot = (int)fb[vv[k] >> 5] + __popc(w & ((1u << (vv[k] & 31)) - 1u));
        rdp[(r0 + k) * kBP + t] = dd[k] | (slot << kSlotShift);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// shared staging helpers of the forward / backward block kernels (512 threads: pixel t = tid >> 3, lane l = tid & 7)
// ---------------------------------------------------------------------------------------------
// height bins (or logits) of the block: col[d * 64 + t]; pixels outside the image read as 0
__device__ __forceinline__ void stage_block_columns(float *col, const float *__restrict__ src /*camera base*/,
                                                    const BDims &m, int blk, bool vec16) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int bi = blk / m.nbw, bj = blk - bi * m.nbw;
  const int h0 = bi * 8, w0 = bj * 8;
  const bool full = h0 + 8 <= m.fH && w0 + 8 <= m.fW;
  if (vec16 && full) {
    // item = (d, row, half): one 16-byte piece
    const int items = m.D * 16;
    for (int i = tid; i < items; i += nthr) {
      const int d = i >> 4, row = (i >> 1) & 7, half = i & 1;
      cp_async_16(col + d * kBP + row * 8 + half * 4, src + (size_t)d * m.P + (size_t)(h0 + row) * m.fW + w0 + half * 4);
    }
  } else {
    const int items = m.D * kBP;
    for (int i = tid; i < items; i += nthr) {
      const int d = i >> 6, t = i & 63;
      const int h = h0 + (t >> 3), w = w0 + (t & 7);
      if (h < m.fH && w < m.fW) cp_async_4(col + i, src + (size_t)d * m.P + (size_t)h * m.fW + w);
      else col[i] = 0.0f;
    }
  }
}

// softmax over D of pixel t's column by its 8 lanes (lane l takes d = l mod 8): leaves exp(x - max) in col and
// returns 1 / sum.  Fixed butterfly order => deterministic, identical in forward and backward.
__device__ __forceinline__ float softmax_block_column(float *col, int D, int t, int l, unsigned gmask) {
  float mx = -INFINITY;
  for (int d = l; d < D; d += 8) mx = fmaxf(mx, col[d * kBP + t]);
  mx = fmaxf(mx, __shfl_xor_sync(gmask, mx, 1));
  mx = fmaxf(mx, __shfl_xor_sync(gmask, mx, 2));
  mx = fmaxf(mx, __shfl_xor_sync(gmask, mx, 4));
  float sm = 0.0f;
  for (int d = l; d < D; d += 8) {
    const float e = exp_ex2(__fsub_rn(col[d * kBP + t], mx));
    col[d * kBP + t] = e;
    sm = __fadd_rn(sm, e);
  }
  sm = __fadd_rn(sm, __shfl_xor_sync(gmask, sm, 1));
  sm = __fadd_rn(sm, __shfl_xor_sync(gmask, sm, 2));
  sm = __fadd_rn(sm, __shfl_xor_sync(gmask, sm, 4));
  return __fdiv_rn(1.0f, sm);
}

// Row layout (shared memory and partial rows): element 4 * (k * 8 + l) + e of a row holds channel l + 8 * (4k + e)
// (lane l of an 8-lane group owns NV float4: one 128-byte line per k and group).  In shared memory the 16-byte chunk
// j = k * 8 + l of row i is stored at chunk (k * 8) + ((l ^ i) & 7): rows are 32 * NV words apart, i.e. all rows
// start in bank 0, and the XOR spreads the same chunk of 8 consecutive rows over the 8 bank groups.
__device__ __forceinline__ int row_chunk(int i, int k, int l) { return k * 8 + ((l ^ i) & 7); }
__host__ __device__ __forceinline__ int chan_of(int k, int l, int e) { return l + 8 * (4 * k + e); }

template <typename CT>
__device__ __forceinline__ float ld_ctx(const CT *p) { return to_f32<CT>(*p); }

struct BsmArgs {
  const float *sem;      // [B*Nc, Cs, fH, fW] semantic logits, or nullptr
  long long sem_stride;
  int Cs;
  float thr;
};

// ---------------------------------------------------------------------------------------------
// FORWARD.  grid (NB, B), 512 threads, dynamic smem: col [D][64] | rows [64][Cpad] | W [cap][64] | pm [cap][2] | sem
// FIXUP: blocks whose partial rows did not fit the pool add their sums to the BEV map with atomics instead
// (launched after the combine kernel; every other block exits at once).
// ---------------------------------------------------------------------------------------------
template <typename CT, int NV, bool FIXUP>
__global__ void __launch_bounds__(kFwdThreads, 2)
bp_forward_kernel(BDims m, int cap, const float *__restrict__ height, int vec16, const CT *__restrict__ context,
                  BsmArgs bsm, const int *__restrict__ cnt_in, const int *__restrict__ rd_in,
                  const int *__restrict__ vox_in, const BlockInfo *__restrict__ info, const int *__restrict__ alloc,
                  float *__restrict__ prow, float *__restrict__ bev) {
  constexpr int kCpad = 32 * NV;
  extern __shared__ __align__(16) float fsm[];
  __shared__ float s_keep[kBP];
  const int b = blockIdx.y, kb = blockIdx.x;
  const size_t fbk = (size_t)b * m.NB + kb;
  const BlockInfo bi = info[fbk];
  if (FIXUP) {
    if (bi.row_base >= 0) return;
  } else {
    if (bi.nslots == 0 || bi.row_base < 0) return;
  }
  float *col = fsm;
  float *rows = col + m.D * kBP;
  float *Wm = rows + kBP * kCpad;
  unsigned *pm = reinterpret_cast<unsigned *>(Wm + (size_t)cap * kBP);
  float *sem_s = reinterpret_cast<float *>(pm + 2 * cap);   // [Cs][64] (BSM only)
  const int n = kb / m.nblk, blk = kb - n * m.nblk;
  const int bn = b * m.Nc + n;
  const int tid = threadIdx.x;
  const int t = tid >> 3, l = tid & 7;
  const unsigned gmask = 0xffu << ((tid & 31) & 24);

  stage_block_columns(col, height + (size_t)bn * m.hs, m, blk, vec16 != 0);
  // BSM context assembly (bsm_lss_fpn.py:524-529): per-pixel softmax over the Cs semantic channels in torch's
  // order (max, sum of exp(x - max) in channel order, exp / sum), background mask
  const int Cc = m.C - (bsm.sem ? bsm.Cs : 0);
  if (bsm.sem) {
    if (tid < kBP) {
      const Pix q = pixel_of(m, blk, tid);
      float keep = 1.0f;
      if (q.valid) {
        const float *ss = bsm.sem + (size_t)bn * bsm.sem_stride + q.p;
        float mx = ss[0];
        for (int k = 1; k < bsm.Cs; ++k) mx = fmaxf(mx, ss[(size_t)k * m.P]);
        float sum = 0.0f;
        for (int k = 0; k < bsm.Cs; ++k) sum = __fadd_rn(sum, expf(__fsub_rn(ss[(size_t)k * m.P], mx)));
        for (int k = 0; k < bsm.Cs; ++k)
          sem_s[k * kBP + tid] = __fdiv_rn(expf(__fsub_rn(ss[(size_t)k * m.P], mx)), sum);
        keep = sem_s[tid] > bsm.thr ? 0.0f : 1.0f;
      } else {
        for (int k = 0; k < bsm.Cs; ++k) sem_s[k * kBP + tid] = 0.0f;
      }
      s_keep[tid] = keep;
    }
    __syncthreads();
  }
  // context rows: item = (pixel, 16-byte chunk j = k * 8 + l'): four channels l' + 8 (4k + e) of one pixel
  {
    const CT *cb = context + (size_t)bn * m.cs;
    for (int i = tid; i < kBP * 8 * NV; i += kFwdThreads) {
      const int tp = i & 63, j = i >> 6;
      const int k = j >> 3, lp = j & 7;
      const Pix q = pixel_of(m, blk, tp);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q.valid) {
        float e[4];
#pragma unroll
        for (int ee = 0; ee < 4; ++ee) {
          const int c = chan_of(k, lp, ee);
          float x = 0.0f;
          if (c < Cc) x = ld_ctx<CT>(cb + (size_t)c * m.P + q.p);
          else if (c < m.C) x = sem_s[(c - Cc) * kBP + tp];
          if (bsm.sem) x = __fmul_rn(x, s_keep[tp]);
          e[ee] = x;
        }
        v = make_float4(e[0], e[1], e[2], e[3]);
      }
      *reinterpret_cast<float4 *>(rows + tp * kCpad + 4 * row_chunk(tp, k, lp)) = v;
    }
  }
  cp_async_wait_all();
  __syncthreads();

  const int cnt = cnt_in[fbk * kBP + t];
  float scale = 1.0f;
  if (m.logits) scale = softmax_block_column(col, m.D, t, l, gmask);
  const bool keep_px = !bsm.sem || s_keep[t] != 0.0f;   // masked pixels contribute exact zeros: skipped entirely
  const int *rdp = rd_in + fbk * m.D * kBP;
  const int *rvp = vox_in + fbk * m.D * kBP;
  float *pr = FIXUP ? nullptr : prow + ((size_t)b * m.rows_cap + bi.row_base) * kCpad;
  float *bevb = bev + (size_t)b * m.C * m.V;

  if constexpr (FIXUP) {
    // degenerate path (partial-row pool exhausted): w * ctx_row of every run goes to the BEV map with global
    // atomics, after the combine kernel has written it
    if (keep_px) {
      for (int r = l; r < cnt; r += 8) {
        const int rdv = rdp[r * kBP + t];
        const int vox = rvp[r * kBP + t];
        const int d0 = rdv & kDMask, d1 = (rdv >> 9) & kDMask;
        float w = 0.0f;
        for (int d = d0; d < d1; ++d) w = __fadd_rn(w, col[d * kBP + t]);
        if (m.logits) w = __fmul_rn(w, scale);
        for (int j = 0; j < 8 * NV; ++j) {
          const int k = j >> 3, lp = j & 7;
          const float4 x = *reinterpret_cast<const float4 *>(rows + t * kCpad + 4 * row_chunk(t, k, lp));
          const float xe[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = chan_of(k, lp, e);
            if (c < m.C) atomicAdd(bevb + (size_t)c * m.V + vox, __fmul_rn(w, xe[e]));
          }
        }
      }
    }
  } else {
  for (int lo = 0; lo < bi.nslots; lo += cap) {
    const int nr = min(cap, bi.nslots - lo);
    for (int i = tid; i < nr * kBP; i += kFwdThreads) Wm[i] = 0.0f;
    for (int i = tid; i < 2 * nr; i += kFwdThreads) pm[i] = 0u;
    __syncthreads();
    // run weights: lane l takes the runs r = l mod 8 of pixel t.  W[s][t] += w: a pixel meets a voxel once along its
    // ray, so this is a plain store in all but degenerate geometries (then two addends, which commute)
    if (keep_px) {
      for (int r = l; r < cnt; r += 8) {
        const int rdv = rdp[r * kBP + t];
        const int s = (int)((unsigned)rdv >> kSlotShift) - lo;
        if ((unsigned)s < (unsigned)nr) {
          const int d0 = rdv & kDMask, d1 = (rdv >> 9) & kDMask;
          float w = 0.0f;
          for (int d = d0; d < d1; ++d) w = __fadd_rn(w, col[d * kBP + t]);
          if (m.logits) w = __fmul_rn(w, scale);
          atomicAdd(&Wm[s * kBP + t], w);
          atomicOr(&pm[2 * s + (t >> 5)], 1u << (t & 31));
        }
      }
    }
    __syncthreads();
    // accumulate: group t sums its slots in pixel order, rows from shared memory
    for (int s = t; s < nr; s += kBP) {
      float acc[NV][4];
#pragma unroll
      for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[k][e] = 0.0f;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        unsigned bits = pm[2 * s + half];
        while (bits) {
          const int tp = half * 32 + __ffs(bits) - 1;
          bits &= bits - 1;
          const float w = Wm[s * kBP + tp];
          const float *row = rows + tp * kCpad;
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            const float4 x = *reinterpret_cast<const float4 *>(row + 4 * row_chunk(tp, k, l));
            fma2(acc[k][0], acc[k][1], w, x.x, x.y);
            fma2(acc[k][2], acc[k][3], w, x.z, x.w);
          }
        }
      }
      float *dst = pr + (size_t)(lo + s) * kCpad + 4 * l;
#pragma unroll
      for (int k = 0; k < NV; ++k)
        *reinterpret_cast<float4 *>(dst + 32 * k) = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
    }
    __syncthreads();
  }
  }
}

// ---------------------------------------------------------------------------------------------
// COMBINE.  grid (nstrips, B), 256 threads: 8-lane group g <-> voxel g of the strip.
// ---------------------------------------------------------------------------------------------
constexpr int kCombThreads = 256;
constexpr int kCombList = 1024;   // contributing blocks kept in shared memory per pass

template <int NV>
__global__ void __launch_bounds__(kCombThreads)
bp_combine_kernel(BDims m, const BlockInfo *__restrict__ info, const unsigned *__restrict__ fmask,
                  const unsigned *__restrict__ fbase, const float *__restrict__ prow, float *__restrict__ bev) {
  constexpr int kCpad = 32 * NV;
  __shared__ unsigned l_mask[kCombList];
  __shared__ int l_row[kCombList];
  __shared__ int s_wcnt[kCombThreads / 32];
  __shared__ int s_n;
  __shared__ float tile[kCpad * 33];
  const int b = blockIdx.y, strip = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int g = tid >> 3, l = tid & 7;   // voxel of the strip, lane of its group
  const unsigned below = (1u << g) - 1u;
  const float *pr = prow + (size_t)b * m.rows_cap * kCpad;
  float acc[NV][4];
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[k][e] = 0.0f;
  // blocks are visited in ascending order (fixed summation order); their partial rows are added whenever the
  // contributor list fills up, and after the last block
  if (tid == 0) s_n = 0;
  __syncthreads();
  for (int kb0 = 0; kb0 < m.NB; kb0 += kCombThreads) {
    const int kb = kb0 + tid;
    unsigned mk = 0u;
    int row = 0;
    if (kb < m.NB) {
      const size_t fbk = (size_t)b * m.NB + kb;
      mk = fmask[fbk * m.nstrips + strip];
      if (mk) {
        const int rb = info[fbk].row_base;
        if (rb < 0) mk = 0u;   // completed by the fix-up launch
        else row = rb + (int)fbase[fbk * m.nstrips + strip];
      }
    }
    // ordered compaction of the contributing blocks among kb0 .. kb0 + 255
    const unsigned bal = __ballot_sync(0xffffffffu, mk != 0u);
    if (lane == 0) s_wcnt[wid] = __popc(bal);
    __syncthreads();
    int off = s_n, total = s_n;
    for (int w = 0; w < kCombThreads / 32; ++w) {
      if (w < wid) off += s_wcnt[w];
      total += s_wcnt[w];
    }
    if (mk) {
      const int at = off + __popc(bal & ((1u << lane) - 1u));
      l_mask[at] = mk;
      l_row[at] = row;
    }
    __syncthreads();
    if (total > kCombList - kCombThreads || kb0 + kCombThreads >= m.NB) {
      for (int i = 0; i < total; ++i) {
        const unsigned mk_i = l_mask[i];
        if ((mk_i >> g) & 1u) {
          const float *src = pr + (size_t)(l_row[i] + __popc(mk_i & below)) * kCpad + 4 * l;
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            const float4 x = __ldg(reinterpret_cast<const float4 *>(src + 32 * k));
            acc[k][0] = __fadd_rn(acc[k][0], x.x); acc[k][1] = __fadd_rn(acc[k][1], x.y);
            acc[k][2] = __fadd_rn(acc[k][2], x.z); acc[k][3] = __fadd_rn(acc[k][3], x.w);
          }
        }
      }
      total = 0;
    }
    __syncthreads();
    if (tid == 0) s_n = total;
  }
  // transpose through shared memory: tile[c][voxel]
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int e = 0; e < 4; ++e) tile[chan_of(k, l, e) * 33 + g] = acc[k][e];
  __syncthreads();
  const int v = strip * 32 + lane;
  if (v < m.V) {
    float *out = bev + (size_t)b * m.C * m.V + v;
    for (int c = wid; c < m.C; c += kCombThreads / 32) stg_stream_f1(out + (size_t)c * m.V, tile[c * 33 + lane]);
  }
}

// ---------------------------------------------------------------------------------------------
// BACKWARD.  grid (NB, B), 512 threads, dynamic smem: col [D][64] | tile [Cpad][65] | G [cap][Cpad] | strip list
// ---------------------------------------------------------------------------------------------
template <typename CT, int NV>
__global__ void __launch_bounds__(kFwdThreads, 2)
bp_backward_kernel(BDims m, int cap, const float *__restrict__ height, int vec16, const CT *__restrict__ context,
                   const float *__restrict__ grad_bev, const int *__restrict__ cnt_in,
                   const int *__restrict__ rd_in, const BlockInfo *__restrict__ info,
                   const unsigned *__restrict__ fmask, const unsigned *__restrict__ fbase,
                   float *__restrict__ gw_ws, float *__restrict__ g_height, float *__restrict__ g_context) {
  constexpr int kCpad = 32 * NV;
  constexpr int kLd = kBP + 1;
  extern __shared__ __align__(16) float bsm_[];
  __shared__ int s_nlist;
  const int b = blockIdx.y, kb = blockIdx.x;
  const size_t fbk = (size_t)b * m.NB + kb;
  const BlockInfo bi = info[fbk];
  float *col = bsm_;
  float *tile = col + m.D * kBP;                 // [Cpad][65]: context in, g_ctx out
  float *G = tile + kCpad * kLd;                 // [cap][Cpad], chunks XOR-swizzled by the slot
  int *list = reinterpret_cast<int *>(G + (size_t)cap * kCpad);   // [cap + 1][3]: strip, mask, slot base
  const int n = kb / m.nblk, blk = kb - n * m.nblk;
  const int bn = b * m.Nc + n;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int t = tid >> 3, l = tid & 7;
  const unsigned gmask = 0xffu << (lane & 24);
  const int glane0 = lane & 24;

  // ---- stage ----------------------------------------------------------------------------------------
  stage_block_columns(col, height + (size_t)bn * m.hs, m, blk, vec16 != 0);
  {
    const CT *cb = context + (size_t)bn * m.cs;
    for (int i = tid; i < m.C * kBP; i += kFwdThreads) {
      const int c = i >> 6, tp = i & 63;
      const Pix q = pixel_of(m, blk, tp);
      if (q.valid) {
        if (sizeof(CT) == 4) cp_async_4(tile + c * kLd + tp, reinterpret_cast<const float *>(cb) + (size_t)c * m.P + q.p);
        else tile[c * kLd + tp] = ld_ctx<CT>(cb + (size_t)c * m.P + q.p);
      } else {
        tile[c * kLd + tp] = 0.0f;
      }
    }
  }
  cp_async_wait_all();
  __syncthreads();

  const int cnt = cnt_in[fbk * kBP + t];
  float scale = 1.0f;
  if (m.logits) scale = softmax_block_column(col, m.D, t, l, gmask);
  float cx[NV][4], acc[NV][4];
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = chan_of(k, l, e);
      cx[k][e] = c < m.C ? tile[c * kLd + t] : 0.0f;
      acc[k][e] = 0.0f;
    }
  const int *rdp = rd_in + fbk * m.D * kBP;
  float *gwp = gw_ws + fbk * m.D * kBP;
  const unsigned *gm = fmask + fbk * m.nstrips, *gbs = fbase + fbk * m.nstrips;
  const float *gb = grad_bev + (size_t)b * m.C * m.V;
  float S = 0.0f;

  for (int lo = 0; lo < bi.nslots; lo += cap) {
    const int nr = min(cap, bi.nslots - lo);
    if (tid == 0) s_nlist = 0;
    __syncthreads();   // (also: the previous round's reads of G are done)
    // strips of the footprint that hold a slot of this round (order irrelevant: each row is written once)
    for (int i = tid; i < m.nstrips; i += kFwdThreads) {
      const unsigned mk = gm[i];
      if (mk) {
        const int sb = (int)gbs[i];
        if (sb < lo + nr && sb + __popc(mk) > lo) {
          const int at = atomicAdd(&s_nlist, 1);
          list[3 * at] = i; list[3 * at + 1] = (int)mk; list[3 * at + 2] = sb;
        }
      }
    }
    __syncthreads();
    const int nlist = s_nlist;
    // gradient rows of those voxels: lane <-> voxel of the strip, 128-byte line per channel and strip
    for (int i = wid; i < nlist; i += kFwdThreads / 32) {
      const int strip = list[3 * i];
      const unsigned mk = (unsigned)list[3 * i + 1];
      const int s = list[3 * i + 2] + __popc(mk & ((1u << lane) - 1u)) - lo;
      const int v = strip * 32 + lane;
      const bool on = ((mk >> lane) & 1u) && (unsigned)s < (unsigned)nr;
      const float *src = gb + v;
      float *dst = G + (size_t)s * kCpad;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
#pragma unroll 2
        for (int lp = 0; lp < 8; ++lp) {
          float e[4];
#pragma unroll
          for (int ee = 0; ee < 4; ++ee) {
            const int c = chan_of(k, lp, ee);
            e[ee] = (on && c < m.C) ? __ldg(src + (size_t)c * m.V) : 0.0f;
          }
          if (on) *reinterpret_cast<float4 *>(dst + 4 * row_chunk(s, k, lp)) = make_float4(e[0], e[1], e[2], e[3]);
        }
      }
    }
    __syncthreads();
    // runs: lane l fetches descriptor and weight of run r0 + l; the group then walks the 8 runs together
    for (int r0 = 0; r0 < cnt; r0 += 8) {
      int my_rd = 0;
      float my_w = 0.0f, my_gw = 0.0f;
      const bool mine = r0 + l < cnt;
      if (mine) {
        my_rd = rdp[(r0 + l) * kBP + t];
        const int d0 = my_rd & kDMask, d1 = (my_rd >> 9) & kDMask;
        float w = 0.0f;
        for (int d = d0; d < d1; ++d) w = __fadd_rn(w, col[d * kBP + t]);
        my_w = m.logits ? __fmul_rn(w, scale) : w;
      }
      const int nj = min(8, cnt - r0);
      for (int j = 0; j < nj; ++j) {
        const int rdj = __shfl_sync(gmask, my_rd, glane0 + j);
        const float wj = __shfl_sync(gmask, my_w, glane0 + j);
        const int s = (int)((unsigned)rdj >> kSlotShift) - lo;
        if ((unsigned)s < (unsigned)nr) {
          const float *grow = G + (size_t)s * kCpad;
          float da = 0.0f, db = 0.0f;
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            const float4 x = *reinterpret_cast<const float4 *>(grow + 4 * row_chunk(s, k, l));
            fma2(acc[k][0], acc[k][1], wj, x.x, x.y);
            fma2(acc[k][2], acc[k][3], wj, x.z, x.w);
            da = __fmaf_rn(cx[k][0], x.x, da); db = __fmaf_rn(cx[k][1], x.y, db);
            da = __fmaf_rn(cx[k][2], x.z, da); db = __fmaf_rn(cx[k][3], x.w, db);
          }
          float dot = __fadd_rn(da, db);
          dot = __fadd_rn(dot, __shfl_xor_sync(gmask, dot, 1));
          dot = __fadd_rn(dot, __shfl_xor_sync(gmask, dot, 2));
          dot = __fadd_rn(dot, __shfl_xor_sync(gmask, dot, 4));
          if (l == j) my_gw = dot;
          S = __fmaf_rn(wj, dot, S);
        }
      }
      if (mine) {
        const int s = (int)((unsigned)my_rd >> kSlotShift) - lo;
        if ((unsigned)s < (unsigned)nr) gwp[(r0 + l) * kBP + t] = my_gw;
      }
    }
  }
  __syncthreads();   // every thread holds its context values; gw of every run is visible to the block

  // ---- g_ctx row -> tile column; g_height in place of the staged column ------------------------------------
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = chan_of(k, l, e);
      if (c < m.C) tile[c * kLd + t] = acc[k][e];
    }
  {
    // lane l owns the bins d = l mod 8; every lane walks the pixel's runs in order
    int dc = l;
    auto put = [&](int d, float gv) {
      float v = gv;
      if (m.logits) v = __fmul_rn(__fmul_rn(col[d * kBP + t], scale), __fsub_rn(gv, S));
      col[d * kBP + t] = v;
    };
    for (int r = 0; r < cnt; ++r) {
      const int rdv = rdp[r * kBP + t];
      const float gv = gwp[r * kBP + t];
      const int d0 = rdv & kDMask, d1 = (rdv >> 9) & kDMask;
      for (; dc < d0; dc += 8) put(dc, 0.0f);
      for (; dc < d1; dc += 8) put(dc, gv);
    }
    for (; dc < m.D; dc += 8) put(dc, 0.0f);
  }
  __syncthreads();
  // ---- coalesced stores --------------------------------------------------------------------------------
  {
    const int bi_ = blk / m.nbw, bj_ = blk - bi_ * m.nbw;
    const int h0 = bi_ * 8, w0 = bj_ * 8;
    float *gh = g_height + (size_t)bn * m.ghs;
    float *gc = g_context + (size_t)bn * m.gcs;
    const bool full = h0 + 8 <= m.fH && w0 + 8 <= m.fW;
    const bool v16h = full && (m.fW % 4 == 0) && (m.ghs % 4 == 0) && (reinterpret_cast<uintptr_t>(g_height) % 16 == 0);
    if (v16h) {
      for (int i = tid; i < m.D * 16; i += kFwdThreads) {
        const int d = i >> 4, row = (i >> 1) & 7, half = i & 1;
        const float4 v = *reinterpret_cast<const float4 *>(col + d * kBP + row * 8 + half * 4);
        stg_stream_f4(reinterpret_cast<float4 *>(gh + (size_t)d * m.P + (size_t)(h0 + row) * m.fW + w0 + half * 4), v);
      }
    } else {
      for (int i = tid; i < m.D * kBP; i += kFwdThreads) {
        const int d = i >> 6, tp = i & 63;
        const int h = h0 + (tp >> 3), w = w0 + (tp & 7);
        if (h < m.fH && w < m.fW) stg_stream_f1(gh + (size_t)d * m.P + (size_t)h * m.fW + w, col[i]);
      }
    }
    for (int i = tid; i < m.C * kBP; i += kFwdThreads) {
      const int c = i >> 6, tp = i & 63;
      const int h = h0 + (tp >> 3), w = w0 + (tp & 7);
      if (h < m.fH && w < m.fW) stg_stream_f1(gc + (size_t)c * m.P + (size_t)h * m.fW + w, tile[c * kLd + tp]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Debug / parity: voxel id per point from the block plan.  grid (NB, B), 64 threads.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBP)
bp_expand_kernel(BDims m, const int *__restrict__ cnt_in, const int *__restrict__ rd_in,
                 const int *__restrict__ vox_in, int *__restrict__ vox_out) {
  const int b = blockIdx.y, kb = blockIdx.x;
  const int n = kb / m.nblk, blk = kb - n * m.nblk;
  const size_t fbk = (size_t)b * m.NB + kb;
  const int t = threadIdx.x;
  const Pix px = pixel_of(m, blk, t);
  if (!px.valid) return;
  const int cnt = cnt_in[fbk * kBP + t];
  int *out = vox_out + (size_t)(b * m.Nc + n) * m.D * m.P + px.p;
  int dc = 0;
  for (int r = 0; r < cnt; ++r) {
    const int rdv = rd_in[(fbk * m.D + r) * kBP + t];
    const int vv = vox_in[(fbk * m.D + r) * kBP + t];
    const int d0 = rdv & kDMask, d1 = (rdv >> 9) & kDMask;
    for (; dc < d0; ++dc) out[(size_t)dc * m.P] = -1;
    for (; dc < d1; ++dc) out[(size_t)dc * m.P] = vv;
  }
  for (; dc < m.D; ++dc) out[(size_t)dc * m.P] = -1;
}

// two CTAs per SM: 228 KB per SM, 1 KB of it reserved per CTA, plus the kernels' few static words
size_t smem_budget() { return 111 * 1024; }

int forward_cap(const BDims &m, int Cs) {
  const size_t fixed = sizeof(float) * ((size_t)m.D * kBP + (size_t)kBP * m.Cpad + (size_t)Cs * kBP);
  const long long left = (long long)smem_budget() - (long long)fixed;
  return (int)std::max<long long>(left / (kBP * 4 + 8), 0);
}
int backward_cap(const BDims &m) {
  const size_t fixed = sizeof(float) * ((size_t)m.D * kBP + (size_t)m.Cpad * (kBP + 1)) + 16;
  const long long left = (long long)smem_budget() - (long long)fixed;
  return (int)std::max<long long>(left / (m.Cpad * 4 + 12) - 1, 0);
}

template <typename CT, int NV>
int launch_forward(const BDims &m, const BWorkspace &w, const float *height, const void *context, BsmArgs bsm,
                   float *bev, cudaStream_t s) {
  const int Cs = bsm.sem ? bsm.Cs : 0;
  const int cap = std::min(forward_cap(m, Cs), 4096);
  const size_t smem = sizeof(float) * ((size_t)m.D * kBP + (size_t)kBP * m.Cpad + (size_t)cap * kBP + (size_t)Cs * kBP) +
                      8 * (size_t)cap;
  const int vec16 = columns_vec16(height, m.hs, m.P) && (m.fW % 4 == 0);
  dim3 grid(m.NB, m.B);
  if (int rc = set_smem(bp_forward_kernel<CT, NV, false>, smem)) return rc;
  bp_forward_kernel<CT, NV, false><<<grid, kFwdThreads, smem, s>>>(
      m, cap, height, vec16, static_cast<const CT *>(context), bsm, w.cnt, w.rd, w.vox, w.info, w.alloc, w.prow, bev);
  SGV3D_CHECK_LAUNCH("bp_forward_kernel");
  bp_combine_kernel<NV><<<dim3(m.nstrips, m.B), kCombThreads, 0, s>>>(m, w.info, w.fmask, w.fbase, w.prow, bev);
  SGV3D_CHECK_LAUNCH("bp_combine_kernel");
  if (int rc = set_smem(bp_forward_kernel<CT, NV, true>, smem)) return rc;
  bp_forward_kernel<CT, NV, true><<<grid, kFwdThreads, smem, s>>>(
      m, cap, height, vec16, static_cast<const CT *>(context), bsm, w.cnt, w.rd, w.vox, w.info, w.alloc, w.prow, bev);
  SGV3D_CHECK_LAUNCH("bp_fixup_kernel");
  return SGV3D_OK;
}

template <typename CT>
int launch_forward_nv(const BDims &m, const BWorkspace &w, const float *height, const void *context, BsmArgs bsm,
                      float *bev, cudaStream_t s) {
  switch (m.NV) {
    case 1: return launch_forward<CT, 1>(m, w, height, context, bsm, bev, s);
    case 2: return launch_forward<CT, 2>(m, w, height, context, bsm, bev, s);
    default: return launch_forward<CT, 3>(m, w, height, context, bsm, bev, s);
  }
}

template <typename CT, int NV>
int launch_backward(const BDims &m, const BWorkspace &w, const float *grad_bev, const float *height,
                    const void *context, float *g_height, float *g_context, cudaStream_t s) {
  const int cap = std::min(backward_cap(m), 4096);
  const size_t smem = sizeof(float) * ((size_t)m.D * kBP + (size_t)m.Cpad * (kBP + 1) + (size_t)cap * m.Cpad) +
                      12 * (size_t)(cap + 1);
  const int vec16 = columns_vec16(height, m.hs, m.P) && (m.fW % 4 == 0);
  if (int rc = set_smem(bp_backward_kernel<CT, NV>, smem)) return rc;
  bp_backward_kernel<CT, NV><<<dim3(m.NB, m.B), kFwdThreads, smem, s>>>(
      m, cap, height, vec16, static_cast<const CT *>(context), grad_bev, w.cnt, w.rd, w.info, w.fmask, w.fbase, w.gw,
      g_height, g_context);
  SGV3D_CHECK_LAUNCH("bp_backward_kernel");
  return SGV3D_OK;
}

template <typename CT>
int launch_backward_nv(const BDims &m, const BWorkspace &w, const float *grad_bev, const float *height,
                       const void *context, float *g_height, float *g_context, cudaStream_t s) {
  switch (m.NV) {
    case 1: return launch_backward<CT, 1>(m, w, grad_bev, height, context, g_height, g_context, s);
    case 2: return launch_backward<CT, 2>(m, w, grad_bev, height, context, g_height, g_context, s);
    default: return launch_backward<CT, 3>(m, w, grad_bev, height, context, g_height, g_context, s);
  }
}

}  // namespace

bool supported(const Dims &d) {
  if (d.C > 96 || d.D > 255) return false;
  const BDims m = make_bdims(d);
  // the plan kernel keeps two words per strip in shared memory; the block kernels need room for >= 32 slots
  if ((size_t)8 * m.nstrips + 8 * m.D > 200 * 1024) return false;
  return forward_cap(m, 16) >= 32 && backward_cap(m) >= 32;
}

size_t workspace_bytes(const Dims &d) { return carve(nullptr, make_bdims(d)).bytes; }

int plan(const Dims &d, int arith, const float *u_tab, const float *v_tab, const float *z_tab, const float *ida_inv,
         const float *m_virtual, const float *m_ego, const float *bda, const float *ref_heights,
         const geom::Grid &grid, void *ws, cudaStream_t s) {
  const BDims m = make_bdims(d);
  const BWorkspace w = carve(ws, m);
  SGV3D_CUDA(cudaMemsetAsync(w.alloc, 0, sizeof(int) * 2 * (size_t)m.B, s));
  const size_t smem = sizeof(float) * 2 * (size_t)m.D + sizeof(unsigned) * 2 * (size_t)m.nstrips;
  dim3 g(m.NB, m.B);
#define SGV3D_BP_PLAN(A)                                                                                         \
  do {                                                                                                           \
    if (int rc = set_smem(bp_plan_kernel<A>, smem)) return rc;                                                   \
    bp_plan_kernel<A><<<g, kBP, smem, s>>>(m, u_tab, v_tab, z_tab, ida_inv, m_virtual, m_ego, bda, ref_heights,  \
                                           grid, w.cnt, w.vox, w.rd, w.info, w.fmask, w.fbase, w.alloc);         \
  } while (0)
  if (arith == SGV3D_ARITH_PAIR) SGV3D_BP_PLAN(SGV3D_ARITH_PAIR);
  else if (arith == SGV3D_ARITH_FMA) SGV3D_BP_PLAN(SGV3D_ARITH_FMA);
  else SGV3D_BP_PLAN(SGV3D_ARITH_SEQ);
#undef SGV3D_BP_PLAN
  SGV3D_CHECK_LAUNCH("bp_plan_kernel");
  return SGV3D_OK;
}

int forward(const Dims &d, int ctx_dtype, const float *height, const void *context, const float *sem, int Cs,
            long long sem_stride, float thr, float *bev, void *ws, cudaStream_t s) {
  const BDims m = make_bdims(d);
  const BWorkspace w = carve(ws, m);
  BsmArgs bsm;
  bsm.sem = sem; bsm.Cs = Cs; bsm.sem_stride = sem_stride; bsm.thr = thr;
  if (ctx_dtype == SGV3D_DTYPE_BF16) return launch_forward_nv<__nv_bfloat16>(m, w, height, context, bsm, bev, s);
  return launch_forward_nv<float>(m, w, height, context, bsm, bev, s);
}

int backward(const Dims &d, int ctx_dtype, const float *grad_bev, const float *height, const void *context,
             float *g_height, float *g_context, void *ws, cudaStream_t s) {
  const BDims m = make_bdims(d);
  const BWorkspace w = carve(ws, m);
  if (ctx_dtype == SGV3D_DTYPE_BF16)
    return launch_backward_nv<__nv_bfloat16>(m, w, grad_bev, height, context, g_height, g_context, s);
  return launch_backward_nv<float>(m, w, grad_bev, height, context, g_height, g_context, s);
}

int plan_expand(const Dims &d, int *vox_out, void *ws, cudaStream_t s) {
  const BDims m = make_bdims(d);
  const BWorkspace w = carve(ws, m);
  bp_expand_kernel<<<dim3(m.NB, m.B), kBP, 0, s>>>(m, w.cnt, w.rd, w.vox, vox_out);
  SGV3D_CHECK_LAUNCH("bp_expand_kernel");
  return SGV3D_OK;
}

}  // namespace block
}  // namespace sgv3d
