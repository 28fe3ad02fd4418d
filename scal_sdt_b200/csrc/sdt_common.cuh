// Shared helpers for libsdt_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/sdt_b200.h"

namespace sdt {

void set_error(const char* fmt, ...);   // defined in api.cu (thread-local message)
void count_launch();                    // api.cu: process-wide count of kernels this library launched

#define SDT_REQUIRE(cond, code, ...)                      \
  do {                                                    \
    if (!(cond)) {                                        \
      ::sdt::set_error(__VA_ARGS__);                      \
      return (code);                                      \
    }                                                     \
  } while (0)

#define SDT_CUDA_OK(expr)                                                                  \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::sdt::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,   \
                       __LINE__);                                                          \
      return SDT_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

// launch-error check that is legal during stream capture (no sync)
#define SDT_LAUNCH_OK(what)                                                                 \
  do {                                                                                      \
    cudaError_t _e = cudaPeekAtLastError();                                                 \
    if (_e != cudaSuccess) {                                                                \
      (void)cudaGetLastError();                                                             \
      ::sdt::set_error("launch of %s failed: %s", what, cudaGetErrorString(_e));            \
      return SDT_ERR_CUDA;                                                                  \
    }                                                                                       \
    ::sdt::count_launch();                                                                  \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int num_sms();   // cached multiProcessorCount of the current device (api.cu)

// ---- programmatic dependent launch (PDL) ----------------------------------------------------
// Kernels that call pdl_wait() before their first global-memory access may be launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: the grid is then scheduled while its predecessor in the stream is
// still draining, runs its prologue (barrier init, TMEM allocation, tensor-map prefetch, constant operands) and blocks in
// `griddepcontrol.wait` until the predecessor has completed and flushed.  Works inside stream capture (the edge becomes a
// programmatic dependency of the graph).  SDT_PDL=0 in the environment or sdt_debug_set(24, 1) turns it off (A/B).
bool pdl_enabled();   // api.cu

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- device helpers -----------------------------------------------------------------------
// PDL: block until every prerequisite grid has completed and its memory is visible (no-op without the launch attribute)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// PDL: this CTA no longer holds back the launch of the dependent grid (it still waits for our completion in ITS pdl_wait)
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 128-bit accesses: data touched once, keep it out of L1
__device__ __forceinline__ uint4 ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// read-write arenas (EMA shadow): plain ld (not .nc) because the same kernel writes it
__device__ __forceinline__ uint4 ld_rw(const void* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ float bf16_bits_to_f32(uint32_t lo16) { return __uint_as_float(lo16 << 16); }
__device__ __forceinline__ uint32_t f32_to_bf16_bits(float f) {
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(f));
}
// two f32 -> one packed bf16 pair (round to nearest even), lo in bits [0,16): a single cvt.rn.bf16x2.f32
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
// round an f32 to the nearest bf16 value, result as f32 (what torch does after every bf16 op)
__device__ __forceinline__ float round_bf16(float f) { return __bfloat162float(__float2bfloat16_rn(f)); }

// Standard normal CDF Phi(x) = 0.5 (1 + erf(x / sqrt 2)) to 1e-7 absolute and e = exp(-x^2 / 2), from Abramowitz & Stegun 7.1.26
// (0.5 erfc(a) = 0.5 poly(t) exp(-a^2), t = 1 / (1 + p a), a = |x| / sqrt 2): two MUFU (rcp.approx, ex2.approx) and ten FP32
// instructions, far below bf16 / fp16 resolution.  erff() is ~3x that; an IEEE-rounded reciprocal (__frcp_rn) alone added a
// Newton step and a slow-path call per element, and the GEGLU kernels were bound by instruction issue (48 instructions per
// element, 4.2 TB/s), not by HBM.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void normal_cdf_pdf(float x, float& cdf, float& e) {
  const float a = fabsf(x) * 0.70710678118654752f;
  const float t = rcp_approx(fmaf(0.3275911f, a, 1.0f));
  const float poly = fmaf(fmaf(fmaf(fmaf(1.061405429f, t, -1.453152027f), t, 1.421413741f), t, -0.284496736f), t, 0.254829592f) * t;
  e = __expf(-a * a);
  const float tail = 0.5f * poly * e;                  // Phi(-|x|)
  cdf = x < 0.0f ? tail : 1.0f - tail;
}
// erf GELU (what torch.nn.functional.gelu evaluates, approximate="none") and its derivative Phi(x) + x phi(x)
__device__ __forceinline__ float gelu_erf(float x) {
  float cdf, e;
  normal_cdf_pdf(x, cdf, e);
  return x * cdf;
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float cdf, e;
  normal_cdf_pdf(x, cdf, e);
  return fmaf(x * 0.39894228040143268f, e, cdf);
}

// ---- 16-bit activation format of the tensor-core path: bf16 (default) or IEEE fp16 (the reference's stock
// `trainer.precision: 16`).  A launch-uniform flag selects it; the tensor core takes the format from the instruction descriptor.
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  const __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack_act2(float lo, float hi, bool f16) { return f16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi); }
__device__ __forceinline__ float round_act(float f, bool f16) { return f16 ? __half2float(__float2half_rn(f)) : round_bf16(f); }
// 32 f32 accumulator values (as raw bits) -> 16 packed pairs; the branch is outside the unrolled loops
__device__ __forceinline__ void pack_acc32(const uint32_t (&v)[32], uint32_t (&pk)[16], bool f16) {
  if (f16) {
#pragma unroll
    for (int j = 0; j < 16; ++j) pk[j] = pack_f16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
  }
}

}  // namespace sdt
