// Shared device helpers for the ddgan_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define DDG_OK 0
#define DDG_ERR_ARG -1
#define DDG_ERR_UNSUPPORTED -2
#define DDG_ERR_LAUNCH -3

#define DDG_CHECK_LAUNCH()                                   \
  do {                                                       \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) { ddg_set_last_error(cudaGetErrorString(e__)); return DDG_ERR_LAUNCH; } \
  } while (0)

void ddg_set_last_error(const char* msg);
int ddg_pdl_enabled(void);   // api_common.cu: programmatic dependent launch switch (ddg_set_pdl / DDG_PDL)

namespace ddg {

constexpr float kRsqrt2 = 0.70710678118654752440f;

enum Act : int { ACT_NONE = 0, ACT_SILU = 1, ACT_LEAKY = 2, ACT_TANH = 3 };

__device__ __forceinline__ float silu_f(float y) { return __fdividef(y, 1.0f + __expf(-y)); }
__device__ __forceinline__ float leaky_f(float y) { return y > 0.f ? y : 0.2f * y; }

__device__ __forceinline__ float apply_act(float y, int act) {
  switch (act) {
    case ACT_SILU: return silu_f(y);
    case ACT_LEAKY: return leaky_f(y);
    case ACT_TANH: return tanhf(y);
    default: return y;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 128-bit global access (no L1 allocation)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}

// split two fp32 into bf16 hi pair and bf16 lo (residual) pair, packed little-endian (a in low half)
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  float ra = a - __low2float(h);
  float rb = b - __high2float(h);
  __nv_bfloat162 l = __floats2bfloat162_rn(ra, rb);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------------------
// A kernel launched through launch_pdl() may start while the kernel in front of it on the stream is still draining: its CTAs are
// scheduled, set up their shared memory / barriers / TMEM, and block in pdl_wait() until the predecessor has completed and its writes
// are visible.  Rules every such kernel follows here: no global memory access before pdl_wait(), and pdl_trigger() only after
// pdl_wait() -- so a dependent kernel can overlap its direct predecessor only, never anything older.  Without the launch attribute both
// instructions are no-ops.  Works inside stream capture (the edge becomes a programmatic dependency of the CUDA graph).
// Compiled in only with -DDDG_ENABLE_PDL: inside the captured loops PDL measured no gain, so the default build carries neither
// instruction and launches with the plain <<< >>> syntax.
#ifdef DDG_ENABLE_PDL
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else
__device__ __forceinline__ void pdl_wait() {}
__device__ __forceinline__ void pdl_trigger() {}
#endif

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
#ifndef DDG_ENABLE_PDL
  kern<<<grid, block, smem, stream>>>(KArgs(args)...);
  return cudaSuccess;
#endif
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
#ifdef DDG_ENABLE_PDL
  cfg.numAttrs = ddg_pdl_enabled() ? 1 : 0;
#else
  cfg.numAttrs = 0;
#endif
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

}  // namespace ddg
