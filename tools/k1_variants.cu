// tools/k1_variants.cu -- micro-benchmark of K1 launch-shape / cache-policy variants with the real quantiser math.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/k1_variants tools/k1_variants.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../fp8_quantization_b200/csrc/fp8fq_core.h"
using namespace fp8fq;

struct Ctx { float hi, lo, guard, t2, t3, s1, s2, s3, r1, r2, r3; };

template <int LD> __device__ __forceinline__ float4 ld4(const float4* p) {
  if (LD == 0) return *p;
  if (LD == 1) return __ldcs(p);
  if (LD == 2) return __ldg(p);
  float4 r;
  if (LD == 3) asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  if (LD == 4) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
template <int ST> __device__ __forceinline__ void st4(float4* p, float4 v) {
  if (ST == 0) *p = v;
  if (ST == 1) __stcs(p, v);
  if (ST == 2) __stwt(p, v);
  if (ST == 3) __stcg(p, v);
}

__device__ __forceinline__ void q4(float4& v, const Ctx& c) {
  float x[4] = {v.x, v.y, v.z, v.w}, s[4], rs[4], q[4], xc[4];
  bool slow = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    xc[k] = min_nan(max_nan(x[k], c.lo), c.hi);
    float a = fabsf(xc[k]);
    bool p2 = a >= c.t2, p3 = a >= c.t3;
    s[k] = p3 ? c.s3 : (p2 ? c.s2 : c.s1);
    rs[k] = p3 ? c.r3 : (p2 ? c.r2 : c.r1);
    float r = mul_rn(xc[k], rs[k]);
    q[k] = nearbyintf(r);
    slow |= !(fabsf(r - q[k]) < c.guard);
  }
  if (slow) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float r = mul_rn(xc[k], rs[k]);
      if (!(fabsf(r - q[k]) < c.guard)) q[k] = nearbyintf(div_rn(xc[k], s[k]));
    }
  }
  v.x = mul_rn(q[0], s[0]); v.y = mul_rn(q[1], s[1]); v.z = mul_rn(q[2], s[2]); v.w = mul_rn(q[3], s[3]);
}

template <int THREADS, int UNROLL, int PERSIST, int LD, int ST, int MINB, int MATH>
__global__ void __launch_bounds__(THREADS, MINB) kern(const float4* __restrict__ x, float4* __restrict__ y, int64_t nvec, Ctx c) {
  const int64_t tile = (int64_t)THREADS * UNROLL;
  const int64_t ntiles = (nvec + tile - 1) / tile;
  for (int64_t t = blockIdx.x; t < ntiles; t += (PERSIST ? gridDim.x : ntiles)) {
    const int64_t base = t * tile + threadIdx.x;
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { int64_t i = base + (int64_t)u * THREADS; if (i < nvec) v[u] = ld4<LD>(x + i); }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { int64_t i = base + (int64_t)u * THREADS; if (i < nvec) { if (MATH) q4(v[u], c); st4<ST>(y + i, v[u]); } }
  }
}

struct f8 { float4 a, b; };
template <int LD> __device__ __forceinline__ f8 ld8(const float4* p) {
  f8 r;
  if (LD == 0) asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w) : "l"(p));
  if (LD == 1) asm volatile("ld.global.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w) : "l"(p));
  if (LD == 2) asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w) : "l"(p));
  return r;
}
template <int ST> __device__ __forceinline__ void st8(float4* p, const f8& v) {
  if (ST == 0) asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p), "f"(v.a.x), "f"(v.a.y), "f"(v.a.z), "f"(v.a.w), "f"(v.b.x), "f"(v.b.y), "f"(v.b.z), "f"(v.b.w) : "memory");
  if (ST == 1) asm volatile("st.global.L2::evict_first.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p), "f"(v.a.x), "f"(v.a.y), "f"(v.a.z), "f"(v.a.w), "f"(v.b.x), "f"(v.b.y), "f"(v.b.z), "f"(v.b.w) : "memory");
}
template <int THREADS, int UNROLL, int PERSIST, int LD, int ST, int MINB, int MATH>
__global__ void __launch_bounds__(THREADS, MINB) kern8(const float4* __restrict__ x, float4* __restrict__ y, int64_t nvec, Ctx c) {
  const int64_t nv8 = nvec / 2;
  const int64_t tile = (int64_t)THREADS * UNROLL;
  const int64_t ntiles = (nv8 + tile - 1) / tile;
  for (int64_t t = blockIdx.x; t < ntiles; t += (PERSIST ? gridDim.x : ntiles)) {
    const int64_t base = t * tile + threadIdx.x;
    f8 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { int64_t i = base + (int64_t)u * THREADS; if (i < nv8) v[u] = ld8<LD>(x + 2 * i); }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { int64_t i = base + (int64_t)u * THREADS; if (i < nv8) { if (MATH) { q4(v[u].a, c); q4(v[u].b, c); } st8<ST>(y + 2 * i, v[u]); } }
  }
}

static int g_sms = 148;
template <int THREADS, int UNROLL, int PERSIST, int LD, int ST, int MINB, int MATH>
void run8(const char* name, const float4* x, float4* y, int64_t nvec, Ctx c) {
  auto k = kern8<THREADS, UNROLL, PERSIST, LD, ST, MINB, MATH>;
  int occ = 1; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, THREADS, 0);
  const int64_t tile = (int64_t)THREADS * UNROLL;
  int64_t ntiles = (nvec / 2 + tile - 1) / tile;
  int64_t grid = PERSIST ? std::min<int64_t>(ntiles, (int64_t)g_sms * occ * PERSIST) : ntiles;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<float> ts;
  for (int it = 0; it < 13; ++it) {
    cudaEventRecord(e0);
    k<<<(unsigned)grid, THREADS>>>(x, y, nvec, c);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (it >= 3) ts.push_back(ms);
  }
  std::sort(ts.begin(), ts.end());
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k);
  double gbs = 32.0 * nvec / (ts[ts.size() / 2] * 1e-3) / 1e9;
  printf("256b %-41s thr=%3d unr=%d persist=%d ld=%d st=%d regs=%3d occ=%2d grid=%8lld  med=%.4f ms  %.0f GB/s  (min %.0f)\n", name, THREADS, UNROLL,
         PERSIST, LD, ST, fa.numRegs, occ, (long long)grid, ts[ts.size() / 2], gbs, 32.0 * nvec / (ts[0] * 1e-3) / 1e9);
}

template <int THREADS, int UNROLL, int PERSIST, int LD, int ST, int MINB, int MATH>
void run(const char* name, const float4* x, float4* y, int64_t nvec, Ctx c) {
  auto k = kern<THREADS, UNROLL, PERSIST, LD, ST, MINB, MATH>;
  int occ = 1; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, THREADS, 0);
  const int64_t tile = (int64_t)THREADS * UNROLL;
  int64_t ntiles = (nvec + tile - 1) / tile;
  int64_t grid = PERSIST ? std::min<int64_t>(ntiles, (int64_t)g_sms * occ * PERSIST) : ntiles;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<float> ts;
  for (int it = 0; it < 13; ++it) {
    cudaEventRecord(e0);
    k<<<(unsigned)grid, THREADS>>>(x, y, nvec, c);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (it >= 3) ts.push_back(ms);
  }
  std::sort(ts.begin(), ts.end());
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k);
  double gbs = 32.0 * nvec / (ts[ts.size() / 2] * 1e-3) / 1e9;
  printf("%-46s thr=%3d unr=%d persist=%d ld=%d st=%d regs=%3d occ=%2d grid=%8lld  med=%.4f ms  %.0f GB/s  (min %.0f)\n", name, THREADS, UNROLL,
         PERSIST, LD, ST, fa.numRegs, occ, (long long)grid, ts[ts.size() / 2], gbs, 32.0 * nvec / (ts[0] * 1e-3) / 1e9);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); g_sms = p.multiProcessorCount;
  const int64_t n = 1ll << 28, nvec = n / 4;
  float *x, *y; cudaMalloc(&x, n * 4); cudaMalloc(&y, n * 4);
  std::vector<float> h(1 << 20); srand(1);
  for (auto& v : h) v = (float)((rand() / (double)RAND_MAX - 0.5) * 6.0);
  for (int64_t o = 0; o < n; o += (1 << 20)) cudaMemcpy(x + o, h.data(), 4 << 20, cudaMemcpyHostToDevice);
  Ctx c{3.0f, -3.0f, 0.5f - ldexpf(1.0f, 5 - 20), 0.75f, 1.5f, 0.0117f, 0.0234f, 0.0469f, 1 / 0.0117f, 1 / 0.0234f, 1 / 0.0469f};
  const float4* xv = (const float4*)x; float4* yv = (float4*)y;
  cudaMemcpy(y, x, n * 4, cudaMemcpyDeviceToDevice);
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); std::vector<float> ts;
    for (int it = 0; it < 13; ++it) { cudaEventRecord(e0); cudaMemcpyAsync(y, x, n * 4, cudaMemcpyDeviceToDevice); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (it >= 3) ts.push_back(ms); }
    std::sort(ts.begin(), ts.end()); printf("cudaMemcpy D2D: %.4f ms %.0f GB/s\n", ts[5], 8.0 * n / (ts[5] * 1e-3) / 1e9);
  }
  //            THREADS UNR PERS LD ST MINB MATH
  run<256, 4, 1, 1, 0, 1, 1>("current (persist, ldcs, st)", xv, yv, nvec, c);
  run<256, 4, 1, 1, 0, 1, 0>("copy only, same shape", xv, yv, nvec, c);
  run<256, 4, 0, 1, 0, 1, 1>("one-shot grid", xv, yv, nvec, c);
  run<256, 4, 0, 1, 0, 1, 0>("one-shot grid copy only", xv, yv, nvec, c);
  run<256, 4, 1, 0, 0, 1, 1>("plain ld", xv, yv, nvec, c);
  run<256, 4, 1, 2, 0, 1, 1>("ldg", xv, yv, nvec, c);
  run<256, 4, 1, 3, 0, 1, 1>("ld no_allocate", xv, yv, nvec, c);
  run<256, 4, 1, 4, 0, 1, 1>("ld nc no_alloc evict_first", xv, yv, nvec, c);
  run<256, 4, 1, 1, 1, 1, 1>("ldcs + stcs", xv, yv, nvec, c);
  run<256, 4, 1, 1, 2, 1, 1>("ldcs + stwt", xv, yv, nvec, c);
  run<256, 4, 1, 1, 3, 1, 1>("ldcs + stcg", xv, yv, nvec, c);
  run<256, 4, 0, 1, 1, 1, 1>("one-shot ldcs + stcs", xv, yv, nvec, c);
  run<256, 8, 1, 1, 0, 1, 1>("unroll 8", xv, yv, nvec, c);
  run<256, 8, 1, 1, 1, 1, 1>("unroll 8 stcs", xv, yv, nvec, c);
  run<256, 2, 1, 1, 0, 1, 1>("unroll 2", xv, yv, nvec, c);
  run<128, 4, 1, 1, 0, 1, 1>("128 thr", xv, yv, nvec, c);
  run<128, 8, 1, 1, 0, 1, 1>("128 thr unroll 8", xv, yv, nvec, c);
  run<512, 4, 1, 1, 0, 1, 1>("512 thr", xv, yv, nvec, c);
  run<512, 2, 1, 1, 0, 1, 1>("512 thr unroll 2", xv, yv, nvec, c);
  run<256, 4, 1, 1, 0, 8, 1>("minblocks 8 (<=32 regs)", xv, yv, nvec, c);
  run<256, 4, 1, 1, 0, 6, 1>("minblocks 6", xv, yv, nvec, c);
  run<256, 4, 2, 1, 0, 1, 1>("persist x2 waves", xv, yv, nvec, c);
  run<256, 4, 4, 1, 0, 1, 1>("persist x4 waves", xv, yv, nvec, c);
  run<1024, 1, 0, 1, 0, 1, 1>("1024 thr unroll 1 one-shot", xv, yv, nvec, c);
  run<128, 4, 0, 1, 0, 1, 1>("128 thr one-shot", xv, yv, nvec, c);
  run<128, 4, 0, 1, 1, 1, 1>("128 thr one-shot stcs", xv, yv, nvec, c);
  run<128, 4, 0, 0, 0, 1, 0>("128 thr one-shot plain copy (torch-like)", xv, yv, nvec, c);
  run8<256, 2, 1, 0, 0, 1, 1>("v8 plain", xv, yv, nvec, c);
  run8<256, 2, 1, 1, 0, 1, 1>("v8 ld evict_first", xv, yv, nvec, c);
  run8<256, 2, 1, 2, 0, 1, 1>("v8 ld nc no_allocate", xv, yv, nvec, c);
  run8<256, 2, 1, 1, 1, 1, 1>("v8 ld+st evict_first", xv, yv, nvec, c);
  run8<256, 4, 1, 1, 0, 1, 1>("v8 unroll 4", xv, yv, nvec, c);
  run8<256, 2, 0, 1, 0, 1, 1>("v8 one-shot", xv, yv, nvec, c);
  run8<128, 2, 0, 1, 0, 1, 1>("v8 128thr one-shot", xv, yv, nvec, c);
  run8<128, 4, 1, 1, 0, 1, 1>("v8 128thr unroll 4", xv, yv, nvec, c);
  run8<256, 2, 1, 0, 0, 1, 0>("v8 copy only", xv, yv, nvec, c);
  run8<256, 1, 1, 1, 0, 1, 1>("v8 unroll 1", xv, yv, nvec, c);
  cudaFree(x); cudaFree(y);
  return 0;
}
