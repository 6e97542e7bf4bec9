// Shared helpers for the hifihr_b200 kernels (sm_100a).
//
// Math that has to be reproducible bit for bit against the CPU oracle goes through the
// X* wrappers: on the device they are the round-to-nearest intrinsics, which nvcc never
// contracts into FMAs; on the host (tests/host_emul builds the same headers with g++
// -ffp-contract=off) they are the plain operators.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "hifihr_b200.h"

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define HFR_HD __host__ __device__ __forceinline__
#else
#define HFR_HD static inline
struct alignas(16) float4 { float x, y, z, w; };   // host build of the math headers (tests/host_emul)
#endif

#if defined(__CUDA_ARCH__)
#define HFR_LDG(p) __ldg(p)
#else
#define HFR_LDG(p) (*(p))
#endif

#if defined(__CUDA_ARCH__)
#define XMUL(a, b) __fmul_rn((a), (b))
#define XADD(a, b) __fadd_rn((a), (b))
#define XSUB(a, b) __fsub_rn((a), (b))
#define XDIV(a, b) __fdiv_rn((a), (b))
#else
#define XMUL(a, b) ((a) * (b))
#define XADD(a, b) ((a) + (b))
#define XSUB(a, b) ((a) - (b))
#define XDIV(a, b) ((a) / (b))
#endif

// Fast (few-ulp) variants for math that does NOT have to be bit-exact against the oracle: shading,
// blending and every backward formula (tolerances in DESIGN.md §5).  The host build keeps the
// plain operators.
#if defined(__CUDA_ARCH__)
#define HFR_FDIV(a, b) __fdividef((a), (b))
#define HFR_RCP(x) __frcp_rn(x)
#define HFR_EXP(x) __expf(x)
#define HFR_POW(x, y) __powf((x), (y))
#else
#define HFR_FDIV(a, b) ((a) / (b))
#define HFR_RCP(x) (1.0f / (x))
#define HFR_EXP(x) expf(x)
#define HFR_POW(x, y) powf((x), (y))
#endif

HFR_HD float hfr_min3(float a, float b, float c) { return fminf(fminf(a, b), c); }
HFR_HD float hfr_max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
HFR_HD float hfr_clamp01(float t) { return fminf(fmaxf(t, 0.0f), 1.0f); }

// ---------------------------------------------------------------- host-side error plumbing
void hfr_set_error(const char* fmt, ...);

#ifdef __CUDACC__
#define HFR_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      hfr_set_error(__VA_ARGS__);           \
      return HFR_EINVAL;                    \
    }                                       \
  } while (0)

#define HFR_CHECK_LAUNCH(name)                                                     \
  do {                                                                             \
    cudaError_t e_ = cudaGetLastError();                                           \
    if (e_ != cudaSuccess) {                                                       \
      hfr_set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e_));   \
      return HFR_ECUDA;                                                            \
    }                                                                              \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming (evict-first) 128-bit store: Fragments are written once and not re-read soon
__device__ __forceinline__ void st_cs_f4(float* p, float a, float b, float c, float d) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void st_cs_i64x2(int64_t* p, int64_t a, int64_t b) {
  asm volatile("st.global.cs.v2.s64 [%0], {%1,%2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
#endif
