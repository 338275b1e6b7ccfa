// Shared helpers for libstreamyolo_sm100 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/streamyolo_sm100.h"

namespace sy {

// thread-local last error text (sy_last_error_string)
void set_error(const char* fmt, ...);

#define SY_REQUIRE(cond, code, ...)      \
  do {                                   \
    if (!(cond)) {                       \
      ::sy::set_error(__VA_ARGS__);      \
      return (code);                     \
    }                                    \
  } while (0)

#define SY_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      ::sy::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                      __LINE__);                                                        \
      return SY_ELAUNCH;                                                                \
    }                                                                                   \
  } while (0)

inline int launch_status(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return SY_ELAUNCH;
  }
  return SY_OK;
}

inline bool view_ok(const SyTensor& t) {
  return t.ptr != nullptr && t.n > 0 && t.h > 0 && t.w > 0 && t.c > 0 && t.pitch >= t.c &&
         (t.c % 8) == 0 && (t.pitch % 8) == 0 && ((uintptr_t)t.ptr % 16) == 0;
}

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch (PDL) -------------------------------------------------
// Kernels launched through launch_pdl() may be scheduled while their predecessor on the stream is still
// draining: they call pdl_launch_dependents() first thing (lets the *next* kernel do the same) and pdl_wait()
// before their first access to global memory (returns once every prerequisite grid has completed and its
// writes are visible).  Both are no-ops for a normally launched grid.  SY_PDL=0 turns the attribute off.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// the same, as thread-block clusters of `cluster_x` consecutive CTAs (grid.x must be a multiple of it)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), int cluster_x, dim3 grid, dim3 block, size_t smem,
                                      cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = (unsigned)cluster_x;
  at[1].val.clusterDim.y = 1;
  at[1].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- bf16 pack helpers -------------------------------------------------------
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);  // .x = a (low half), .y = b
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float round_bf16(float a) { return __bfloat162float(__float2bfloat16_rn(a)); }

// SiLU = v * rcp(1 + 2^(-v * log2 e)) on the two approximate SFU ops (ex2.approx, rcp.approx: ~2 ulp fp32, far below
// the bf16 output rounding) = 5 instructions per value; the normalise pass is issue-bound otherwise (a correctly
// rounded divide costs ~10 instructions, div.approx adds range fix-ups).  v << 0: 2^x overflows to +inf, rcp(inf) = 0,
// the product is -0, the correct limit.
__device__ __forceinline__ float silu_f(float v) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return v * r;
}

}  // namespace sy
