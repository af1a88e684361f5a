// Shared device/host helpers for libisdf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace isdf {

typedef double2 cplx;  // complex128, interleaved (re, im) -- numpy/torch layout

// ---- error plumbing (never throws across the C ABI) -------------------------------
struct Handle {
  int device;
  int sm_count;
  int max_smem_optin;
  void* scratch; size_t scratch_bytes;   // grow-only device scratch owned by the handle (split-K partials)
  char err[512];
};

#define ISDF_OK 0
#define ISDF_EARG (-1)
#define ISDF_ESIZE (-2)

#define ISDF_CHECK_ARG(h, cond, msg)                                              \
  do {                                                                            \
    if (!(cond)) {                                                                \
      if (h) snprintf((h)->err, sizeof((h)->err), "%s:%d: bad argument: %s", __FILE__, __LINE__, msg); \
      return ISDF_EARG;                                                           \
    }                                                                             \
  } while (0)

#define ISDF_CUDA(h, call)                                                        \
  do {                                                                            \
    cudaError_t e_ = (call);                                                      \
    if (e_ != cudaSuccess) {                                                      \
      if (h) snprintf((h)->err, sizeof((h)->err), "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return (int)e_;                                                             \
    }                                                                             \
  } while (0)

#define ISDF_LAUNCH_CHECK(h) ISDF_CUDA(h, cudaGetLastError())

// ---- PTX wrappers -------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// 16-byte async copy global->shared, zero-filled when !valid (src must still be a legal address)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// FP64 tensor-core MMA: D(8x8) += A(8x4) * B(4x8).  Lane (g = lane>>2, t = lane&3) holds
// a = A[g][t], b = B[t][g], c0/c1 = C[g][2t], C[g][2t+1].  Lowers to DMMA.8x8x4 on sm_100a.
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) {  // a * conj(b)
  return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ void cfma(cplx& acc, cplx a, cplx b) {  // acc += a*b
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}

// atomic max for non-negative doubles (bit pattern order == numeric order)
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
  atomicMax((unsigned long long*)addr, (unsigned long long)__double_as_longlong(v));
}

}  // namespace isdf
