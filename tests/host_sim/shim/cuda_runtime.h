// tests/host_sim/shim/cuda_runtime.h -- TEST ONLY.  A stand-in for <cuda_runtime.h> that lets g++ compile
// fp8_quantization_b200/csrc/fp8fq_kernels.cu unchanged (built with -DFP8FQ_HOST_SIM) and run the kernels' own code on
// the CPU: one fiber per CUDA thread, CTAs executed one after another, __syncthreads / __shfl_xor_sync as cooperative
// barriers.  It exists so that the -m "not gpu" tests can check the kernels' index arithmetic, tails, layout variants and
// reductions against the CPU oracle without a GPU.  It is never linked into libfp8fq.so and the product package never
// loads it; it says nothing about performance.
//
// What is modelled: grid / block indices, shared memory (static storage: one CTA is live at a time), dynamic shared
// memory, barriers, warp shuffles over the full mask, atomics (trivially, one OS thread), the runtime calls the C ABI
// makes (memset / memcpy / malloc / streams as no-ops).  What is not: memory-model effects, warp divergence rules,
// asynchrony, anything about timing.  libm stands in for libdevice (log2f / powf / rsqrtf may differ by an ulp, which is
// why the simulation is compared with the C oracle built on the same libm, not with GPU outputs).
#pragma once

#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#ifndef FP8FQ_HOST_SIM
#error "the CUDA shim is only for the FP8FQ_HOST_SIM build"
#endif

// ---- qualifiers --------------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static
#define __align__(n) alignas(n)

// ---- vector types ------------------------------------------------------------------------------------------------
struct uint3 {
  unsigned x, y, z;
};
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 {
  float x, y;
};
struct alignas(16) float4 {
  float x, y, z, w;
};
struct alignas(16) int4 {
  int x, y, z, w;
};
struct alignas(16) uint4 {
  unsigned int x, y, z, w;
};
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }

// ---- built-in variables (one OS thread: plain globals, set by the scheduler at every switch) ------------------------
inline uint3 threadIdx{0, 0, 0};
inline uint3 blockIdx{0, 0, 0};
inline dim3 blockDim;
inline dim3 gridDim;

// ---- runtime API subset --------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1 };
typedef void* cudaStream_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 };
enum { cudaDevAttrMultiProcessorCount = 16 };
enum { cudaStreamNonBlocking = 1 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes {
  cudaMemoryType type;
};

namespace fp8fq_sim {
inline int g_sm_count = 148;          // B200; tests may change it to drive other grid sizes
inline bool g_report_pinned = false;  // what cudaPointerGetAttributes says about host pointers
}  // namespace fp8fq_sim

inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = fp8fq_sim::g_sm_count; return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
  memmove(d, s, n);
  return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
template <typename T>
inline cudaError_t cudaMalloc(T** p, size_t n) {
  *p = static_cast<T*>(aligned_alloc(256, (n + 255) & ~size_t(255)));
  return *p ? cudaSuccess : cudaErrorInvalidValue;
}
template <typename T>
inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) {
  a->type = fp8fq_sim::g_report_pinned ? cudaMemoryTypeHost : cudaMemoryTypeUnregistered;
  return cudaSuccess;
}

// ---- device intrinsics ---------------------------------------------------------------------------------------------
template <typename T>
inline T __ldg(const T* p) { return *p; }
template <typename T>
inline T __ldcs(const T* p) { return *p; }
template <typename T>
inline T __ldcg(const T* p) { return *p; }
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline void __threadfence() {}
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline double atomicAdd(double* p, double v) { double o = *p; *p = o + v; return o; }
inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { unsigned int o = *p; *p = o + v; return o; }

// ---- the grid runner -------------------------------------------------------------------------------------------------
namespace fp8fq_sim {

float* dynamic_smem();
void cta_barrier();                       // __syncthreads
float warp_exchange(float v, int lane_xor);  // __shfl_xor_sync over the full mask
// cooperative = false: the kernel is declared barrier-free (no __syncthreads, no shuffles) and its threads run as plain
// calls, one after another, without fibers (10x faster); reaching a barrier in that mode aborts with a message.
void run_grid(dim3 grid, dim3 block, size_t smem, bool cooperative, const std::function<void()>& thread_body);
int64_t launches();
int64_t ctas_run();

template <typename... KArgs, typename... Args>
void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, bool cooperative, Args&&... args) {
  // kernel parameters are passed by value, converted to the parameter types like a real launch does
  std::tuple<std::decay_t<KArgs>...> params(static_cast<std::decay_t<KArgs>>(std::forward<Args>(args))...);
  run_grid(grid, block, smem, cooperative, [&]() { std::apply(kernel, params); });
}

}  // namespace fp8fq_sim

inline void __syncthreads() { fp8fq_sim::cta_barrier(); }
inline float __shfl_xor_sync(unsigned mask, float v, int lane_xor) {
  if (mask != 0xffffffffu) {
    fprintf(stderr, "fp8fq_sim: only full-mask shuffles are modelled\n");
    abort();
  }
  return fp8fq_sim::warp_exchange(v, lane_xor);
}
