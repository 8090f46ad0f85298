// Interface of the pixel-block pipeline (lift_splat_block.cu) towards the C-ABI entry points in lift_splat.cu.
#pragma once

#include "geometry.cuh"
#include "ls_shared.cuh"

namespace sgv3d {
namespace block {

// rows of <= 96 channels, D <= 255, block kernels fit the shared-memory budget
bool supported(const Dims &d);
size_t workspace_bytes(const Dims &d);
int plan(const Dims &d, int arith, const float *u_tab, const float *v_tab, const float *z_tab, const float *ida_inv,
         const float *m_virtual, const float *m_ego, const float *bda, const float *ref_heights,
         const geom::Grid &grid, void *ws, cudaStream_t s);
// sem != nullptr: BSM context assembly fused (Cs semantic logits behind the d.C - Cs context channels)
int forward(const Dims &d, int ctx_dtype, const float *height, const void *context, const float *sem, int Cs,
            long long sem_stride, float thr, float *bev, void *ws, cudaStream_t s);
int backward(const Dims &d, int ctx_dtype, const float *grad_bev, const float *height, const void *context,
             float *g_height, float *g_context, void *ws, cudaStream_t s);
int plan_expand(const Dims &d, int *vox_out, void *ws, cudaStream_t s);

}  // namespace block
}  // namespace sgv3d
