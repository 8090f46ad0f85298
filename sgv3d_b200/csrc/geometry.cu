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

}  // namespace
}  // namespace sgv3d

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
