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
  uint2 *ftab;         // [B*NB][nstrips] footprint: {bitmap of the strip, slot of its first touched voxel}
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
  w.ftab = c.take<uint2>(nb * m.nstrips);
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
               BlockInfo *__restrict__ info, uint2 *__restrict__ ftab, int *__restrict__ alloc) {
  __shared__ geom::Camera cam;
  __shared__ int s_flags[4];   // 0: camera qualifies for the fast path, 1: z table finite, 2: hgt table valid
  __shared__ __align__(16) int s_red[2][2];   // (own 16 bytes: the compiler reads it with one 128-bit load)
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
  uint2 *gt = ftab + fbk * m.nstrips;
  for (int i = t; i < m.nstrips; i += kBP) gt[i] = make_uint2(fm[i], fb[i]);
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
// shared helpers of the forward / backward block kernels (512 threads: pixel t = tid >> 3, lane l = tid & 7)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_4_zfill(float *smem_dst, const float *gsrc, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 4 : 0;   // src-size 0: four zero bytes are written, the source is not read
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(s), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct Org {   // image position of a pixel block
  int h0, w0;
  bool full;   // all 64 pixels inside the image
};
__device__ __forceinline__ Org block_origin(const BDims &m, int blk) {
  const int bi = blk / m.nbw, bj = blk - bi * m.nbw;
  Org o;
  o.h0 = bi * 8; o.w0 = bj * 8;
  o.full = o.h0 + 8 <= m.fH && o.w0 + 8 <= m.fW;
  return o;
}

// height bins (or logits) of the block: col[d * 64 + t]; pixels outside the image read as 0.  cp.async, not waited for.
__device__ __forceinline__ void stage_block_columns(float *col, const float *__restrict__ src /*camera base*/,
                                                    const BDims &m, const Org &o, bool vec16) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const float *base = src + o.h0 * m.fW + o.w0;
  if (vec16 && o.full) {
    // item = (d, row, half): one 16-byte piece
    const int row = (tid >> 1) & 7, half = tid & 1;
    const int off = row * m.fW + half * 4, soff = row * 8 + half * 4;
    for (int d = tid >> 4; d < m.D; d += nthr >> 4) cp_async_16(col + d * kBP + soff, base + (size_t)d * m.P + off);
  } else {
    const int t = tid & 63;
    const int r = t >> 3, c = t & 7;
    const bool ok = o.h0 + r < m.fH && o.w0 + c < m.fW;
    const int off = r * m.fW + c;
    for (int d = tid >> 6; d < m.D; d += nthr >> 6) cp_async_4_zfill(col + d * kBP + t, ok ? base + (size_t)d * m.P + off : src, ok);
  }
}

// softmax over D of pixel t's column by its 8 lanes (lane l takes d = l mod 8): leaves exp(x - max) in col and
// returns 1 / sum.  Fixed butterfly order => deterministic, identical in forward and backward.
__device__ __forceinline__ float softmax_block_column(float *col, int D, int t, int l, unsigned gmask) {
  float *c0 = col + l * kBP + t;
  float mx = -INFINITY;
  for (int d = l; d < D; d += 8) mx = fmaxf(mx, c0[(d - l) * kBP]);
  mx = fmaxf(mx, __shfl_xor_sync(gmask, mx, 1));
  mx = fmaxf(mx, __shfl_xor_sync(gmask, mx, 2));
  mx = fmaxf(mx, __shfl_xor_sync(gmask, mx, 4));
  float sm = 0.0f;
  for (int d = l; d < D; d += 8) {
    const float e = exp_ex2(__fsub_rn(c0[(d - l) * kBP], mx));
    c0[(d - l) * kBP] = e;
    sm = __fadd_rn(sm, e);
  }
  sm = __fadd_rn(sm, __shfl_xor_sync(gmask, sm, 1));
  sm = __fadd_rn(sm, __shfl_xor_sync(gmask, sm, 2));
  sm = __fadd_rn(sm, __shfl_xor_sync(gmask, sm, 4));
  return __fdiv_rn(1.0f, sm);
}

// Rows (context rows, gradient rows, partial sums) hold the channels in natural order, 32 * NV floats per row; lane l
// of an 8-lane group owns the 16-byte chunks l, 8 + l, 16 + l (channels 32k + 4l .. + 3).  In SHARED memory the chunk j
// of row i sits at chunk position (j & ~7) | ((j ^ i) & 7): rows are a multiple of 32 words apart, so without the XOR
// the same chunk of different rows would share its banks.
__device__ __forceinline__ int swz(int j, int i) { return (j & ~7) | ((j ^ i) & 7); }

struct BsmArgs {
  const float *sem;      // [B*Nc, Cs, fH, fW] semantic logits, or nullptr
  long long sem_stride;
  int Cs;
  float thr;
};

// run weight: sum of the (exponentiated) bins d0 .. d1 - 1 of pixel t, ascending
__device__ __forceinline__ float run_weight(const float *col, int t, int rdv) {
  const int d0 = rdv & kDMask, d1 = (rdv >> 9) & kDMask;
  const float *c = col + d0 * kBP + t;
  float w = 0.0f;
  for (int d = d0; d < d1; ++d, c += kBP) w = __fadd_rn(w, *c);
  return w;
}

// ---------------------------------------------------------------------------------------------
// FORWARD.  grid (NB, B), 512 threads, dynamic smem: col [D][64] | rows [64][Cpad] | W [cap][64] | pm [cap][2] | sem
// ---------------------------------------------------------------------------------------------
template <typename CT, int NV>
__global__ void __launch_bounds__(kFwdThreads, 2)
bp_forward_kernel(BDims m, int cap, const float *__restrict__ height, int vec16, const CT *__restrict__ context,
                  BsmArgs bsm, const int *__restrict__ cnt_in, const int *__restrict__ rd_in,
                  const BlockInfo *__restrict__ info, float *__restrict__ prow) {
  constexpr int kCpad = 32 * NV;
  extern __shared__ __align__(16) float fsm[];
  __shared__ float s_keep[kBP];
  const int b = blockIdx.y, kb = blockIdx.x;
  const size_t fbk = (size_t)b * m.NB + kb;
  const BlockInfo bi = info[fbk];
  if (bi.nslots == 0 || bi.row_base < 0) return;
  float *col = fsm;
  float *rows = col + m.D * kBP;
  float *Wm = rows + kBP * kCpad;
  unsigned *pm = reinterpret_cast<unsigned *>(Wm + (size_t)cap * kBP);
  float *sem_s = reinterpret_cast<float *>(pm + 2 * cap);   // [Cs][64] (BSM only)
  const int n = kb / m.nblk, blk = kb - n * m.nblk;
  const int bn = b * m.Nc + n;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int t = tid >> 3, l = tid & 7;
  const unsigned gmask = 0xffu << (lane & 24);
  const Org o = block_origin(m, blk);

  stage_block_columns(col, height + (size_t)bn * m.hs, m, o, vec16 != 0);
  cp_async_commit();
  // BSM context assembly (bsm_lss_fpn.py:524-529): per-pixel softmax over the Cs semantic channels in torch's
  // order (max, sum of exp(x - max) in channel order, exp / sum), background mask
  const int Cc = m.C - (bsm.sem ? bsm.Cs : 0);
  if (bsm.sem) {
    if (tid < kBP) {
      const int h = o.h0 + (tid >> 3), w = o.w0 + (tid & 7);
      float keep = 1.0f;
      if (h < m.fH && w < m.fW) {
        const float *ss = bsm.sem + (size_t)bn * bsm.sem_stride + h * m.fW + w;
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
  // context rows: warp item = (image row r of the block, chunk j); lane = (pixel of the row, channel of the chunk):
  // one request reads 4 channel planes x 32 bytes and writes one chunk of 8 rows (8 different bank groups)
  {
    const CT *cb = context + (size_t)bn * m.cs;
    const int e = lane & 3, tpl = lane >> 2;
    const bool wok = o.w0 + tpl < m.fW;
    for (int it = wid; it < 8 * 8 * NV; it += kFwdThreads / 32) {
      const int r = it & 7, j = it >> 3;
      const int c = 4 * j + e, tp = r * 8 + tpl;
      const bool ok = wok && o.h0 + r < m.fH;
      float *dst = rows + tp * kCpad + 4 * swz(j, tp) + e;
      const int off = c * m.P + (o.h0 + r) * m.fW + o.w0 + tpl;
      if (sizeof(CT) == 4 && !bsm.sem) {
        cp_async_4_zfill(dst, reinterpret_cast<const float *>(cb) + (ok && c < m.C ? off : 0), ok && c < m.C);
      } else {
        float x = 0.0f;
        if (ok) {
          if (c < Cc) x = to_f32<CT>(cb[off]);
          else if (c < m.C) x = sem_s[(c - Cc) * kBP + tp];
          if (bsm.sem) x = __fmul_rn(x, s_keep[tp]);
        }
        *dst = x;
      }
    }
  }
  cp_async_commit();
  const int cnt = cnt_in[fbk * kBP + t];
  const int *rdp = rd_in + fbk * m.D * kBP + t;
  // this lane's first run descriptors travel while the columns land
  int rd0 = 0, rd1 = 0;
  if (l < cnt) rd0 = rdp[l * kBP];
  if (l + 8 < cnt) rd1 = rdp[(l + 8) * kBP];
  cp_async_wait_group<1>();   // columns
  __syncthreads();
  float scale = 1.0f;
  if (m.logits) scale = softmax_block_column(col, m.D, t, l, gmask);
  const bool keep_px = !bsm.sem || s_keep[t] != 0.0f;   // masked pixels contribute exact zeros: skipped entirely
  float *pr = prow + ((size_t)b * m.rows_cap + bi.row_base) * kCpad;
  cp_async_wait_group<0>();   // context rows

  for (int lo = 0; lo < bi.nslots; lo += cap) {
    const int nr = min(cap, bi.nslots - lo);
    {
      float4 *w4 = reinterpret_cast<float4 *>(Wm);
      for (int i = tid; i < nr * (kBP / 4); i += kFwdThreads) w4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int i = tid; i < 2 * nr; i += kFwdThreads) pm[i] = 0u;
    }
    __syncthreads();
    // run weights: lane l takes the runs r = l mod 8 of pixel t.  W[s][t] += w: a pixel meets a voxel once along its
    // ray, so this is a plain store in all but degenerate geometries (then two addends, which commute)
    if (keep_px) {
      for (int r = l; r < cnt; r += 8) {
        const int rdv = r == l ? rd0 : (r == l + 8 ? rd1 : rdp[r * kBP]);
        const int s = (int)((unsigned)rdv >> kSlotShift) - lo;
        if ((unsigned)s < (unsigned)nr) {
          float w = run_weight(col, t, rdv);
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
      const float *wrow = Wm + s * kBP;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        unsigned bits = pm[2 * s + half];
        while (bits) {
          const int tp = half * 32 + __ffs(bits) - 1;
          bits &= bits - 1;
          const float w = wrow[tp];
          const float *row = rows + tp * kCpad + 4 * ((l ^ tp) & 7);
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            const float4 x = *reinterpret_cast<const float4 *>(row + 32 * k);
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

// ---------------------------------------------------------------------------------------------
// FIX-UP (degenerate path): blocks whose partial rows did not fit the frame's row pool add w * ctx of every run to the
// BEV map with global atomics, after the combine kernel has written it.  A small persistent grid: it first looks at
// the frames' overflow flags and normally exits at once.
// ---------------------------------------------------------------------------------------------
template <typename CT>
__global__ void __launch_bounds__(256)
bp_fixup_kernel(BDims m, const float *__restrict__ height, const CT *__restrict__ context, BsmArgs bsm,
                const int *__restrict__ cnt_in, const int *__restrict__ rd_in, const int *__restrict__ vox_in,
                const BlockInfo *__restrict__ info, const int *__restrict__ alloc, float *__restrict__ bev) {
  int any = 0;
  for (int i = threadIdx.x; i < m.B; i += blockDim.x) any |= alloc[m.B + i];
  if (!__syncthreads_or(any)) return;
  const int Cc = m.C - (bsm.sem ? bsm.Cs : 0);
  const int total = m.B * m.NB;
  // warp per pixel-block pixel: plain, slow, correct
  for (int fb = blockIdx.x; fb < total; fb += gridDim.x) {
    if (info[fb].row_base >= 0) continue;
    const int b = fb / m.NB, kb = fb - b * m.NB;
    const int n = kb / m.nblk, blk = kb - n * m.nblk;
    const int bn = b * m.Nc + n;
    const Org o = block_origin(m, blk);
    const float *hb = height + (size_t)bn * m.hs;
    const CT *cb = context + (size_t)bn * m.cs;
    float *bevb = bev + (size_t)b * m.C * m.V;
    for (int t = threadIdx.x >> 5; t < kBP; t += blockDim.x >> 5) {
      const int lane = threadIdx.x & 31;
      const int h = o.h0 + (t >> 3), w = o.w0 + (t & 7);
      if (h >= m.fH || w >= m.fW) continue;
      const int p = h * m.fW + w;
      const int cnt = cnt_in[(size_t)fb * kBP + t];
      float keep = 1.0f, semp[16];
      if (bsm.sem) {
        const float *ss = bsm.sem + (size_t)bn * bsm.sem_stride + p;
        float mx = ss[0];
        for (int k = 1; k < bsm.Cs; ++k) mx = fmaxf(mx, ss[(size_t)k * m.P]);
        float sum = 0.0f;
        for (int k = 0; k < bsm.Cs; ++k) sum = __fadd_rn(sum, expf(__fsub_rn(ss[(size_t)k * m.P], mx)));
        for (int k = 0; k < bsm.Cs && k < 16; ++k) semp[k] = __fdiv_rn(expf(__fsub_rn(ss[(size_t)k * m.P], mx)), sum);
        keep = semp[0] > bsm.thr ? 0.0f : 1.0f;
      }
      if (keep == 0.0f) continue;
      float mx = -INFINITY, sm = 1.0f;
      if (m.logits) {
        for (int d = 0; d < m.D; ++d) mx = fmaxf(mx, hb[(size_t)d * m.P + p]);
        sm = 0.0f;
        for (int d = 0; d < m.D; ++d) sm = __fadd_rn(sm, exp_ex2(__fsub_rn(hb[(size_t)d * m.P + p], mx)));
      }
      for (int r = 0; r < cnt; ++r) {
        const int rdv = rd_in[((size_t)fb * m.D + r) * kBP + t];
        const int vox = vox_in[((size_t)fb * m.D + r) * kBP + t];
        const int d0 = rdv & kDMask, d1 = (rdv >> 9) & kDMask;
        float wgt = 0.0f;
        for (int d = d0; d < d1; ++d) {
          const float x = hb[(size_t)d * m.P + p];
          wgt = __fadd_rn(wgt, m.logits ? exp_ex2(__fsub_rn(x, mx)) : x);
        }
        if (m.logits) wgt = __fmul_rn(wgt, __fdiv_rn(1.0f, sm));
        for (int c = lane; c < m.C; c += 32) {
          float x = c < Cc ? to_f32<CT>(cb[(size_t)c * m.P + p]) : semp[min(c - Cc, 15)];
          atomicAdd(bevb + (size_t)c * m.V + vox, __fmul_rn(wgt, x));
        }
      }
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
bp_combine_kernel(BDims m, const BlockInfo *__restrict__ info, const uint2 *__restrict__ ftab,
                  const float *__restrict__ prow, float *__restrict__ bev) {
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
  const float *pr = prow + (size_t)b * m.rows_cap * kCpad + 4 * l;
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
      const uint2 e = __ldg(ftab + fbk * m.nstrips + strip);   // {mask, slot of the strip's first touched voxel}
      const int rb = __ldg(&info[fbk].row_base);
      mk = rb < 0 ? 0u : e.x;   // rb < 0: completed by the fix-up launch
      row = rb + (int)e.y;
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
          const float *src = pr + (size_t)(l_row[i] + __popc(mk_i & below)) * kCpad;
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
    for (int e = 0; e < 4; ++e) tile[(32 * k + 4 * l + e) * 33 + g] = acc[k][e];
  __syncthreads();
  const int v = strip * 32 + lane;
  if (v < m.V) {
    float *out = bev + (size_t)b * m.C * m.V + v;
    for (int c = wid; c < m.C; c += kCombThreads / 32) stg_stream_f1(out + (size_t)c * m.V, tile[c * 33 + lane]);
  }
}

// ---------------------------------------------------------------------------------------------
// BACKWARD.  grid (NB, B), 512 threads, dynamic smem:
//   col [D][64] | tile [Cpad][65] | G [cap][Cpad] | svox [cap] | rd_s [kRunsStaged][64] | gw_s [kRunsStaged][64]
// ---------------------------------------------------------------------------------------------
constexpr int kRunsStaged = 32;   // run descriptors / run gradients of a pixel kept in shared memory; the rest in HBM

template <typename CT, int NV>
__global__ void __launch_bounds__(kFwdThreads, 2)
bp_backward_kernel(BDims m, int cap, const float *__restrict__ height, int vec16, const CT *__restrict__ context,
                   const float *__restrict__ grad_bev, const int *__restrict__ cnt_in,
                   const int *__restrict__ rd_in, const BlockInfo *__restrict__ info,
                   const uint2 *__restrict__ ftab, float *__restrict__ gw_ws, float *__restrict__ g_height,
                   float *__restrict__ g_context) {
  constexpr int kCpad = 32 * NV;
  constexpr int kLd = kBP + 1;
  extern __shared__ __align__(16) float bsm_[];
  __shared__ __align__(8) unsigned long long s_bar;
  const int b = blockIdx.y, kb = blockIdx.x;
  const size_t fbk = (size_t)b * m.NB + kb;
  const BlockInfo bi = info[fbk];
  float *col = bsm_;
  float *tile = col + m.D * kBP;                 // [Cpad][65]: context in, g_ctx out
  float *G = tile + kCpad * kLd;                 // [cap][Cpad], chunks XOR-swizzled by the slot
  int *svox = reinterpret_cast<int *>(G + (size_t)cap * kCpad);   // [cap] voxel of the round's slots
  int *rd_s = svox + cap;                        // [kRunsStaged][64]
  float *gw_s = reinterpret_cast<float *>(rd_s + kRunsStaged * kBP);   // [kRunsStaged][64]
  const int n = kb / m.nblk, blk = kb - n * m.nblk;
  const int bn = b * m.Nc + n;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int t = tid >> 3, l = tid & 7;
  const unsigned gmask = 0xffu << (lane & 24);
  const int glane0 = lane & 24;
  const Org o = block_origin(m, blk);

  // ---- stage (all asynchronous) ------------------------------------------------------------------------
  stage_block_columns(col, height + (size_t)bn * m.hs, m, o, vec16 != 0);
  const int nst = min(bi.max_runs, kRunsStaged);
  if (nst > 0) {
    // run descriptors of the first kRunsStaged runs: one contiguous block, one bulk-async copy (TMA engine)
    if (tid == 0) {
      mbar_init(&s_bar, 1);
      mbar_arrive_expect_tx(&s_bar, (unsigned)nst * kBP * 4);
      bulk_copy_g2s(rd_s, rd_in + fbk * m.D * kBP, (unsigned)nst * kBP * 4, &s_bar);
    }
  }
  cp_async_commit();
  {
    const CT *cb = context + (size_t)bn * m.cs;
    const int tp = tid & 63;
    const int r = tp >> 3, cc = tp & 7;
    const bool ok = o.h0 + r < m.fH && o.w0 + cc < m.fW;
    const int off = (o.h0 + r) * m.fW + o.w0 + cc;
    for (int c = tid >> 6; c < m.C; c += kFwdThreads >> 6) {
      if (sizeof(CT) == 4) cp_async_4_zfill(tile + c * kLd + tp, reinterpret_cast<const float *>(cb) + (ok ? c * m.P + off : 0), ok);
      else tile[c * kLd + tp] = ok ? to_f32<CT>(cb[c * m.P + off]) : 0.0f;
    }
  }
  cp_async_commit();
  const int cnt = cnt_in[fbk * kBP + t];
  const int *rdp = rd_in + fbk * m.D * kBP + t;
  float *gwp = gw_ws + fbk * m.D * kBP + t;
  const uint2 *ft = ftab + fbk * m.nstrips;
  const float *gb = grad_bev + (size_t)b * m.C * m.V;
  cp_async_wait_group<1>();   // columns
  __syncthreads();            // (also: the mbarrier initialised by thread 0 is visible to every waiter below)
  float scale = 1.0f;
  if (m.logits) scale = softmax_block_column(col, m.D, t, l, gmask);
  cp_async_wait_group<0>();   // context
  if (nst > 0) mbar_wait(&s_bar, 0);   // run descriptors
  __syncthreads();
  float cx[NV][4], acc[NV][4];
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = 32 * k + 4 * l + e;
      cx[k][e] = c < m.C ? tile[c * kLd + t] : 0.0f;
      acc[k][e] = 0.0f;
    }
  float S = 0.0f;
  auto load_rd = [&](int r) -> int { return r < kRunsStaged ? rd_s[r * kBP + t] : rdp[r * kBP]; };

  for (int lo = 0; lo < bi.nslots; lo += cap) {
    const int nr = min(cap, bi.nslots - lo);
    __syncthreads();   // the previous round's reads of G / svox are done
    // voxel of every slot of this round: expand the footprint bitmap (strips whose slots overlap [lo, lo + nr))
    for (int i = tid; i < m.nstrips; i += kFwdThreads) {
      const uint2 e = __ldg(ft + i);
      if (e.x) {
        int s = (int)e.y - lo;
        if (s < nr && s + __popc(e.x) > 0) {
          unsigned bits = e.x;
          while (bits) {
            const int vt = __ffs(bits) - 1;
            bits &= bits - 1;
            if ((unsigned)s < (unsigned)nr) svox[s] = i * 32 + vt;
            ++s;
          }
        }
      }
    }
    __syncthreads();
    // gradient rows of those voxels, straight from the NCHW gradient: lane <-> slot (consecutive slots are mostly
    // consecutive voxels: x-runs), four channel planes per 16-byte chunk, 16 loads in flight
    for (int s0 = wid * 32; s0 < nr; s0 += kFwdThreads) {
      const int s = s0 + lane;
      const bool on = s < nr;
      const float *src = gb + (on ? svox[s] : 0);
      float *dst = G + (size_t)s * kCpad;
      for (int j0 = 0; j0 < kCpad; j0 += 16) {   // (pad channels are written as zeros: they meet cx = 0 in the dot product)
        float x[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) x[u] = (on && j0 + u < m.C) ? __ldg(src + (size_t)(j0 + u) * m.V) : 0.0f;
        if (on) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4 *>(dst + 4 * swz((j0 >> 2) + q, s)) = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
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
        my_rd = load_rd(r0 + l);
        const float w = run_weight(col, t, my_rd);
        my_w = m.logits ? __fmul_rn(w, scale) : w;
      }
      const int nj = min(8, cnt - r0);
      for (int j = 0; j < nj; ++j) {
        const int rdj = __shfl_sync(gmask, my_rd, glane0 + j);
        const float wj = __shfl_sync(gmask, my_w, glane0 + j);
        const int s = (int)((unsigned)rdj >> kSlotShift) - lo;
        if ((unsigned)s < (unsigned)nr) {
          const float *grow = G + (size_t)s * kCpad + 4 * ((l ^ s) & 7);
          float da = 0.0f, db = 0.0f;
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            const float4 x = *reinterpret_cast<const float4 *>(grow + 32 * k);
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
        if ((unsigned)s < (unsigned)nr) {
          if (r0 + l < kRunsStaged) gw_s[(r0 + l) * kBP + t] = my_gw;
          else gwp[(r0 + l) * kBP] = my_gw;
        }
      }
    }
  }
  __syncthreads();   // every thread holds its context values; gw of every run is visible to the block

  // ---- g_ctx row -> tile column; g_height in place of the staged column ------------------------------------
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = 32 * k + 4 * l + e;
      if (c < m.C) tile[c * kLd + t] = acc[k][e];
    }
  {
    // lane l owns the bins d = l mod 8; every lane walks the pixel's runs in order
    int dc = l;
    float *cp = col + t;
    auto put = [&](int d, float gv) {
      float v = gv;
      if (m.logits) v = __fmul_rn(__fmul_rn(cp[d * kBP], scale), __fsub_rn(gv, S));
      cp[d * kBP] = v;
    };
    for (int r = 0; r < cnt; ++r) {
      const int rdv = load_rd(r);
      const float gv = r < kRunsStaged ? gw_s[r * kBP + t] : gwp[r * kBP];
      const int d0 = rdv & kDMask, d1 = (rdv >> 9) & kDMask;
      for (; dc < d0; dc += 8) put(dc, 0.0f);
      for (; dc < d1; dc += 8) put(dc, gv);
    }
    for (; dc < m.D; dc += 8) put(dc, 0.0f);
  }
  __syncthreads();
  // ---- coalesced stores --------------------------------------------------------------------------------
  {
    float *gh = g_height + (size_t)bn * m.ghs + o.h0 * m.fW + o.w0;
    float *gc = g_context + (size_t)bn * m.gcs + o.h0 * m.fW + o.w0;
    const bool v16h = o.full && (m.fW % 4 == 0) && (m.ghs % 4 == 0) && (reinterpret_cast<uintptr_t>(g_height) % 16 == 0);
    if (v16h) {
      const int row = (tid >> 1) & 7, half = tid & 1;
      const int off = row * m.fW + half * 4, soff = row * 8 + half * 4;
      for (int d = tid >> 4; d < m.D; d += kFwdThreads >> 4)
        stg_stream_f4(reinterpret_cast<float4 *>(gh + (size_t)d * m.P + off), *reinterpret_cast<const float4 *>(col + d * kBP + soff));
    } else {
      const int tp = tid & 63;
      const int r = tp >> 3, cc = tp & 7;
      if (o.h0 + r < m.fH && o.w0 + cc < m.fW)
        for (int d = tid >> 6; d < m.D; d += kFwdThreads >> 6) stg_stream_f1(gh + (size_t)d * m.P + r * m.fW + cc, col[d * kBP + tp]);
    }
    {
      const int tp = tid & 63;
      const int r = tp >> 3, cc = tp & 7;
      if (o.h0 + r < m.fH && o.w0 + cc < m.fW)
        for (int c = tid >> 6; c < m.C; c += kFwdThreads >> 6) stg_stream_f1(gc + (size_t)c * m.P + r * m.fW + cc, tile[c * kLd + tp]);
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
  return (int)std::max<long long>(left / (kBP * 4 + 8), 0) & ~3;
}
size_t backward_fixed_smem(const BDims &m) {
  return sizeof(float) * ((size_t)m.D * kBP + (size_t)m.Cpad * (kBP + 1)) + 2 * sizeof(float) * kRunsStaged * kBP + 16;
}
int backward_cap(const BDims &m) {
  const long long left = (long long)smem_budget() - (long long)backward_fixed_smem(m);
  return (int)std::max<long long>(left / (m.Cpad * 4 + 4), 0) & ~3;
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
  if (int rc = set_smem(bp_forward_kernel<CT, NV>, smem)) return rc;
  bp_forward_kernel<CT, NV><<<grid, kFwdThreads, smem, s>>>(m, cap, height, vec16, static_cast<const CT *>(context), bsm,
                                                           w.cnt, w.rd, w.info, w.prow);
  SGV3D_CHECK_LAUNCH("bp_forward_kernel");
  bp_combine_kernel<NV><<<dim3(m.nstrips, m.B), kCombThreads, 0, s>>>(m, w.info, w.ftab, w.prow, bev);
  SGV3D_CHECK_LAUNCH("bp_combine_kernel");
  bp_fixup_kernel<CT><<<std::min(m.B * m.NB, 2 * kNumSMs), 256, 0, s>>>(m, height, static_cast<const CT *>(context), bsm, w.cnt,
                                                                     w.rd, w.vox, w.info, w.alloc, bev);
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
  const size_t smem = backward_fixed_smem(m) + (size_t)cap * (m.Cpad * 4 + 4);
  const int vec16 = columns_vec16(height, m.hs, m.P) && (m.fW % 4 == 0);
  if (int rc = set_smem(bp_backward_kernel<CT, NV>, smem)) return rc;
  bp_backward_kernel<CT, NV><<<dim3(m.NB, m.B), kFwdThreads, smem, s>>>(
      m, cap, height, vec16, static_cast<const CT *>(context), grad_bev, w.cnt, w.rd, w.info, w.ftab, w.gw, g_height,
      g_context);
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
                                           grid, w.cnt, w.vox, w.rd, w.info, w.ftab, w.alloc);                 \
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
