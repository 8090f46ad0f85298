// Shared host/device helpers for libsgv3d_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/sgv3d_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libsgv3d_b200 is written for sm_100a (B200) only"
#endif

namespace sgv3d {

constexpr int kWarp = 32;
constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ---- host-side status plumbing --------------------------------------------------------------
void set_error(const char *fmt, ...);
int64_t &launch_counter();
// optional per-kernel CUDA-event timing (sgv3d_profile_enable / sgv3d_profile_report):
// prof_begin() marks the start of an API call on its stream, prof_mark() is recorded right after
// every launch; a kernel's time is the span between its mark and the previous one.
void prof_begin(cudaStream_t stream);
void prof_mark(const char *name);

#define SGV3D_REQUIRE(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      ::sgv3d::set_error(__VA_ARGS__);      \
      return SGV3D_ERR_INVALID_ARGUMENT;    \
    }                                       \
  } while (0)

// Check the launch just issued (no host sync: only launch-configuration errors surface here).
#define SGV3D_CHECK_LAUNCH(name)                                                      \
  do {                                                                                \
    ++::sgv3d::launch_counter();                                                      \
    ::sgv3d::prof_mark(name);                                                         \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess) {                                                         \
      ::sgv3d::set_error("kernel %s failed to launch: %s", name, cudaGetErrorString(e__)); \
      return SGV3D_ERR_CUDA;                                                          \
    }                                                                                 \
  } while (0)

#define SGV3D_CUDA(call)                                                              \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      ::sgv3d::set_error("%s: %s", #call, cudaGetErrorString(e__));                   \
      return SGV3D_ERR_CUDA;                                                          \
    }                                                                                 \
  } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Carves a caller-provided workspace into 256-byte aligned arrays.
struct Carver {
  char *base;
  size_t off = 0;
  explicit Carver(void *p) : base(static_cast<char *>(p)) {}
  template <typename T>
  T *take(size_t count) {
    off = align_up(off, 256);
    T *r = reinterpret_cast<T *>(base + off);
    off += count * sizeof(T);
    return r;
  }
  size_t used() const { return align_up(off, 256); }
};

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// streaming (read-once) 128-bit / 32-bit loads that do not pollute L1
__device__ __forceinline__ float4 ldg_stream_f4(const float4 *p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_stream_f1(const float *p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
// streaming stores (write-once outputs)
__device__ __forceinline__ void stg_stream_f4(float4 *p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void stg_stream_f1(float *p, float v) {
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

}  // namespace sgv3d
