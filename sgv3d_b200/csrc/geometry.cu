// sgv3d_geometry_quantize: frustum geometry + voxel-index quantisation as a stand-alone kernel.
// Produces the int32 (x, y, z) index tensor the reference hands to voxel_pooling
// (layers/backbones/lss_fpn.py:478-488), so the op-level drop-in path and the parity tests can
// consume exactly what `get_geometry(...)` + `.int()` would have produced.
#include "geometry.cuh"

namespace sgv3d {
namespace {

constexpr int kThreads = 128;

// One thread per pixel of one camera; the thread walks all D height bins.
// Pure compute: the only global reads are three tiny tables; writes are 12 B/point.
template <int ARITH>
__global__ void __launch_bounds__(kThreads)
geometry_quantize_kernel(int Nc, int D, int fH, int fW, const float *__restrict__ u_tab,
                         const float *__restrict__ v_tab, const float *__restrict__ z_tab,
                         const float *__restrict__ ida_inv, const float *__restrict__ mv,
                         const float *__restrict__ me, const float *__restrict__ bda,
                         const float *__restrict__ ref_h, geom::Grid grid,
                         int32_t *__restrict__ idx_out, float *__restrict__ xyz_out) {
  __shared__ geom::Camera cam;
  extern __shared__ float z_s[];
  const int bn = blockIdx.y, b = bn / Nc;
  geom::load_camera(&cam, ida_inv, mv, me, bda, ref_h, bn, b);
  for (int d = threadIdx.x; d < D; d += kThreads) z_s[d] = z_tab[d];
  __syncthreads();
  const int P = fH * fW;
  const int p = blockIdx.x * kThreads + threadIdx.x;
  if (p >= P) return;
  const int h = p / fW, w = p - h * fW;
  geom::PixelRay<ARITH> ray;
  ray.init(cam, u_tab[w], v_tab[h]);
  size_t o = ((size_t)bn * D * P + p) * 3;
  for (int d = 0; d < D; ++d, o += (size_t)P * 3) {
    float gx, gy, gz;
    ray.point(cam, z_s[d], gx, gy, gz);
    if (xyz_out) {
      xyz_out[o] = gx; xyz_out[o + 1] = gy; xyz_out[o + 2] = gz;
    }
    if (idx_out) {
      idx_out[o] = geom::quantize1(gx, grid.lower[0], grid.size[0]);
      idx_out[o + 1] = geom::quantize1(gy, grid.lower[1], grid.size[1]);
      idx_out[o + 2] = geom::quantize1(gz, grid.lower[2], grid.size[2]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Batched 4x4 inverse with the arithmetic of torch.inverse / Tensor.inverse() on a CUDA tensor
// (lss_fpn.py:361,367,392), i.e. torch.linalg.inv_ex -> cuBLAS getrfBatched + getrsBatched against the
// identity.  Its rounding sequence was established offline from (A, LU, pivots, inverse) dumps of
// tools/probe_inverse.py (12288 calibration-like and random matrices reproduced bit for bit):
//   LU, partial pivoting (first row of maximal |a_ik|):  r = 1 / a_kk;  l_ik = a_ik * r;
//                                                         a_ij = fma(-l_ik, a_kj, a_ij)
//   L y = P e_c  (unit lower):  y_i = fma(-l_ij, y_j, y_i), j ascending
//   U x = y:                    x_i = fma(-u_ij, x_j, x_i), j descending;  x_i = x_i / u_ii
// One thread per matrix (a batch holds a few hundred of them); everything in registers.
// The host side checks the kernel against torch's own routine once per process and device before using it.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load4x4(const float *p, float (&a)[4][4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float4 v = *reinterpret_cast<const float4 *>(p + 4 * r);
    a[r][0] = v.x; a[r][1] = v.y; a[r][2] = v.z; a[r][3] = v.w;
  }
}
__device__ __forceinline__ void store4x4(float *p, const float (&x)[4][4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) *reinterpret_cast<float4 *>(p + 4 * r) = make_float4(x[r][0], x[r][1], x[r][2], x[r][3]);
}

// x = inverse(a); a is overwritten with its LU factors
__device__ __forceinline__ void invert4x4(float (&a)[4][4], float (&x)[4][4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) x[r][c] = r == c ? 1.0f : 0.0f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    // pivot: first row i >= k with the largest |a_ik|; the row swap is applied to A and to the right-hand side
    int p = k;
    float best = fabsf(a[k][k]);
#pragma unroll
    for (int r = k + 1; r < 4; ++r)
      if (fabsf(a[r][k]) > best) { best = fabsf(a[r][k]); p = r; }
#pragma unroll
    for (int r = k + 1; r < 4; ++r)
      if (p == r) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float t = a[k][c]; a[k][c] = a[r][c]; a[r][c] = t;
          t = x[k][c]; x[k][c] = x[r][c]; x[r][c] = t;
        }
      }
    if (k < 3) {
      const float rcp = __fdiv_rn(1.0f, a[k][k]);
#pragma unroll
      for (int r = k + 1; r < 4; ++r) {
        a[r][k] = __fmul_rn(a[r][k], rcp);
#pragma unroll
        for (int c = k + 1; c < 4; ++c) a[r][c] = __fmaf_rn(-a[r][k], a[k][c], a[r][c]);
      }
    }
  }
  // forward substitution with the unit lower factor, all four columns of the right-hand side
#pragma unroll
  for (int r = 1; r < 4; ++r)
#pragma unroll
    for (int j = 0; j < r; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) x[r][c] = __fmaf_rn(-a[r][j], x[j][c], x[r][c]);
  // back substitution with the upper factor
#pragma unroll
  for (int r = 3; r >= 0; --r) {
#pragma unroll
    for (int j = 3; j > r; --j)
#pragma unroll
      for (int c = 0; c < 4; ++c) x[r][c] = __fmaf_rn(-a[r][j], x[j][c], x[r][c]);
#pragma unroll
    for (int c = 0; c < 4; ++c) x[r][c] = __fdiv_rn(x[r][c], a[r][r]);
  }
}

__global__ void __launch_bounds__(128)
inverse4x4_kernel(const float *__restrict__ in0, const float *__restrict__ in1, const float *__restrict__ in2,
                  float *__restrict__ out0, float *__restrict__ out1, float *__restrict__ out2, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float *in = blockIdx.y == 0 ? in0 : (blockIdx.y == 1 ? in1 : in2);
  float *out = blockIdx.y == 0 ? out0 : (blockIdx.y == 1 ? out1 : out2);
  float a[4][4], x[4][4];
  load4x4(in + (size_t)i * 16, a);
  invert4x4(a, x);
  store4x4(out + (size_t)i * 16, x);
}

// c = a @ b with the k-ascending rounding order torch's CUDA matmul uses for these 4x4 batches
// (tools/probe_matmul.py): one matrix -> plain products and sums (SEQ); two or more -> FMA chain.
template <int ARITH>
__device__ __forceinline__ void matmul4x4(const float (&a)[4][4], const float (&b)[4][4], float (&c)[4][4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[r][j] = geom::dot4<ARITH>(a[r], b[0][j], b[1][j], b[2][j], b[3][j]);
}

// The whole per-camera prep of lss_fpn.py:361,367,392 in one launch; blockIdx.y selects one of the three
// independent tasks of a camera (so that their latency chains run side by side):
//   0: ida_inv = inverse(ida)   1: m_virtual = sensor2virtual @ inverse(intrin)   2: m_ego = sensor2ego @ inverse(sensor2virtual)
template <int ARITH>
__global__ void __launch_bounds__(64)
camera_prep_kernel(int n, const float *__restrict__ ida, const float *__restrict__ intrin,
                   const float *__restrict__ s2v, const float *__restrict__ s2e, float *__restrict__ ida_inv,
                   float *__restrict__ m_virtual, float *__restrict__ m_ego) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t o = (size_t)i * 16;
  const int task = blockIdx.y;
  float a[4][4], x[4][4], l[4][4], c[4][4];
  load4x4((task == 0 ? ida : (task == 1 ? intrin : s2v)) + o, a);
  if (task != 0) load4x4((task == 1 ? s2v : s2e) + o, l);   // left factor of the product, in flight during the LU
  invert4x4(a, x);
  if (task == 0) {
    store4x4(ida_inv + o, x);
  } else {
    matmul4x4<ARITH>(l, x, c);
    store4x4((task == 1 ? m_virtual : m_ego) + o, c);
  }
}

}  // namespace
}  // namespace sgv3d

extern "C" int sgv3d_inverse4x4(int n, const float *a0, const float *a1, const float *a2, float *inv0,
                                float *inv1, float *inv2, sgv3d_stream_t stream) {
  using namespace sgv3d;
  SGV3D_REQUIRE(n >= 0, "inverse4x4: bad count");
  if (n == 0) return SGV3D_OK;
  SGV3D_REQUIRE(a0 && inv0, "inverse4x4: null pointer");
  SGV3D_REQUIRE((a1 == nullptr) == (inv1 == nullptr) && (a2 == nullptr) == (inv2 == nullptr) && (a1 || !a2),
                "inverse4x4: input / output sets must pair up");
  for (const void *q : {(const void *)a0, (const void *)a1, (const void *)a2, (const void *)inv0, (const void *)inv1,
                        (const void *)inv2})
    SGV3D_REQUIRE(reinterpret_cast<uintptr_t>(q) % 16 == 0, "inverse4x4: matrices must be 16-byte aligned");
  const int sets = a2 ? 3 : (a1 ? 2 : 1);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  prof_begin(s);
  inverse4x4_kernel<<<dim3(ceil_div(n, 128), sets), 128, 0, s>>>(a0, a1, a2, inv0, inv1, inv2, n);
  SGV3D_CHECK_LAUNCH("inverse4x4_kernel");
  return SGV3D_OK;
}

extern "C" int sgv3d_camera_prep(int n, int product_arith, const float *ida, const float *intrin,
                                 const float *sensor2virtual, const float *sensor2ego, float *ida_inv,
                                 float *m_virtual, float *m_ego, sgv3d_stream_t stream) {
  using namespace sgv3d;
  SGV3D_REQUIRE(n >= 0, "camera_prep: bad count");
  if (n == 0) return SGV3D_OK;
  SGV3D_REQUIRE(product_arith == SGV3D_ARITH_SEQ || product_arith == SGV3D_ARITH_FMA,
                "camera_prep: product_arith must be SEQ or FMA");
  SGV3D_REQUIRE(ida && intrin && sensor2virtual && sensor2ego && ida_inv && m_virtual && m_ego,
                "camera_prep: null pointer");
  for (const void *q : {(const void *)ida, (const void *)intrin, (const void *)sensor2virtual,
                        (const void *)sensor2ego, (const void *)ida_inv, (const void *)m_virtual, (const void *)m_ego})
    SGV3D_REQUIRE(reinterpret_cast<uintptr_t>(q) % 16 == 0, "camera_prep: matrices must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  prof_begin(s);
  if (product_arith == SGV3D_ARITH_FMA)
    camera_prep_kernel<SGV3D_ARITH_FMA><<<dim3(ceil_div(n, 64), 3), 64, 0, s>>>(n, ida, intrin, sensor2virtual, sensor2ego,
                                                                               ida_inv, m_virtual, m_ego);
  else
    camera_prep_kernel<SGV3D_ARITH_SEQ><<<dim3(ceil_div(n, 64), 3), 64, 0, s>>>(n, ida, intrin, sensor2virtual, sensor2ego,
                                                                               ida_inv, m_virtual, m_ego);
  SGV3D_CHECK_LAUNCH("camera_prep_kernel");
  return SGV3D_OK;
}

extern "C" int sgv3d_geometry_quantize(int arith, int B, int Nc, int D, int fH, int fW,
                                       const float *u_tab, const float *v_tab, const float *z_tab,
                                       const float *ida_inv, const float *m_virtual,
                                       const float *m_ego, const float *bda,
                                       const float *ref_heights, const float *lower3,
                                       const float *size3, int32_t *idx_out, float *xyz_out,
                                       sgv3d_stream_t stream) {
  using namespace sgv3d;
  SGV3D_REQUIRE(B >= 0 && Nc > 0 && D > 0 && fH > 0 && fW > 0, "geometry_quantize: bad sizes");
  SGV3D_REQUIRE(arith >= SGV3D_ARITH_SEQ && arith <= SGV3D_ARITH_PAIR, "geometry_quantize: bad arith %d", arith);
  SGV3D_REQUIRE(u_tab && v_tab && z_tab && ida_inv && m_virtual && m_ego && ref_heights && lower3 && size3,
                "geometry_quantize: null pointer");
  SGV3D_REQUIRE((long long)B * Nc <= 65535, "geometry_quantize: B*Nc > 65535");
  if (B == 0 || (!idx_out && !xyz_out)) return SGV3D_OK;
  geom::Grid grid;
  for (int k = 0; k < 3; ++k) { grid.lower[k] = lower3[k]; grid.size[k] = size3[k]; }
  grid.X = grid.Y = grid.Z = 0;
  dim3 g(ceil_div(fH * fW, kThreads), B * Nc);
  const size_t smem = sizeof(float) * D;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  prof_begin(s);
  if (arith == SGV3D_ARITH_PAIR)
    geometry_quantize_kernel<SGV3D_ARITH_PAIR><<<g, kThreads, smem, s>>>(
        Nc, D, fH, fW, u_tab, v_tab, z_tab, ida_inv, m_virtual, m_ego, bda, ref_heights, grid, idx_out, xyz_out);
  else if (arith == SGV3D_ARITH_FMA)
    geometry_quantize_kernel<SGV3D_ARITH_FMA><<<g, kThreads, smem, s>>>(
        Nc, D, fH, fW, u_tab, v_tab, z_tab, ida_inv, m_virtual, m_ego, bda, ref_heights, grid, idx_out, xyz_out);
  else
    geometry_quantize_kernel<SGV3D_ARITH_SEQ><<<g, kThreads, smem, s>>>(
        Nc, D, fH, fW, u_tab, v_tab, z_tab, ida_inv, m_virtual, m_ego, bda, ref_heights, grid, idx_out, xyz_out);
  SGV3D_CHECK_LAUNCH("geometry_quantize_kernel");
  return SGV3D_OK;
}
