// fp8fq_kernels.cu -- sm_100a kernels + C ABI of libfp8fq.so (see include/fp8fq.h).
//
// Kernel families (DESIGN.md has the roofline of each):
//   prepare_kernel        per-channel prologue of quantize_to_fp8_ste_MM (fp8_quantizer.py:105-113,128,130)
//   fq_stream_kernel      K1: streaming fake-quant, optionally fused with BN+act or residual-add+act
//                         (fp8_quantizer.py:113-132; quantized_folded_bn.py:39-55; resnet_quantized.py:43-46)
//   fq_rows_kernel        K1 per-channel (weights, channel = dim 0)
//   minmax_*_kernel       K2a: NaN-propagating min/max + estimator update (+ fused prologue)
//                         (range_estimators.py:61-125)
//   mse_grid_kernel       K2b: FP_MSE_Estimator's candidate loop (range_estimators.py:337-347)
//
//   fq_backward_kernel    STE backward (autograd through fp8_quantizer.py:112-132)
//
// All HBM-bound kernels use 128-bit coalesced global accesses with several independent loads in flight per thread,
// quantiser tables in registers (read with uniform loads), and short-lived one-tile CTAs (see fq_stream_kernel).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <type_traits>
#include <utility>

#include "../../include/fp8fq.h"
#include "fp8fq_core.h"

using namespace fp8fq;

namespace {

std::atomic<int64_t> g_launches{0};

// SM count per device, read once (0 = not read yet; concurrent first callers read the same attribute value).
std::atomic<int> g_sms[64];

int sm_count() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = g_sms[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    g_sms[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

// Programmatic dependent launch (PDL): kernels of this library that follow each other on a stream (23 per ResNet-18
// forward) overlap the next kernel's launch latency, CTA rasterisation and prologue with the previous kernel's
// last wave.  Every participating kernel calls pdl_prologue() before its first global-memory access:
//   griddepcontrol.launch_dependents  -- the next kernel in the stream may be scheduled once all CTAs of this grid
//                                        have started (i.e. during this grid's last wave);
//   griddepcontrol.wait               -- block until the previous grid has completed and its writes are visible.
// Launched without the attribute (FP8FQ_PDL=0) both instructions are no-ops.
__device__ __forceinline__ void pdl_prologue() {
#ifndef FP8FQ_HOST_SIM
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("FP8FQ_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

// Every kernel launch of the library goes through launch_impl.  `pdl`: with the programmatic-stream-serialization
// attribute (kernels that call pdl_prologue()); without it, it is the launch a triple-chevron call issues.
// FP8FQ_HOST_SIM (tests/host_sim: a g++ build of this file that runs the kernels' code on the CPU for the
// -m "not gpu" tests, never part of libfp8fq.so) substitutes its own grid runner here.
template <typename... KArgs, typename... Args>
cudaError_t launch_impl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                        Args&&... args) {
#ifdef FP8FQ_HOST_SIM
  (void)st;  // the PDL kernels are exactly the barrier-free streaming kernels: the simulation runs those without fibers
  fp8fq_sim::launch(kernel, grid, block, smem, /*cooperative=*/!pdl, std::forward<Args>(args)...);
  return cudaSuccess;
#else
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
#endif
}
template <typename... KArgs, typename... Args>
cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  return launch_impl(true, kernel, grid, block, smem, st, std::forward<Args>(args)...);
}
template <typename... KArgs, typename... Args>
cudaError_t launch_plain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  return launch_impl(false, kernel, grid, block, smem, st, std::forward<Args>(args)...);
}

inline int launch_status() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? FP8FQ_OK : (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// vector access helpers
// ------------------------------------------------------------------------------------------------
template <int VEC>
struct Pack;
template <>
struct Pack<4> {
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    float4 t = __ldcs(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct Pack<1> {
  float v[1];
  __device__ __forceinline__ void load(const float* p) { v[0] = __ldcs(p); }
  __device__ __forceinline__ void store(float* p) const { *p = v[0]; }
};
template <int VEC>
struct IPack;
template <>
struct IPack<4> {
  int32_t v[4];
  __device__ __forceinline__ void store(int32_t* p) const {
    *reinterpret_cast<int4*>(p) = make_int4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct IPack<1> {
  int32_t v[1];
  __device__ __forceinline__ void store(int32_t* p) const { *p = v[0]; }
};

// exact unsigned division by a runtime-constant divisor (n < 2^32): q = (umulhi(n, m) + n) >> s
struct FastDiv {
  uint32_t m, s, d;
};
FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  if (d == 1) { f.m = 0; f.s = 0; return f; }
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;  // ceil(log2 d)
  uint64_t m = ((1ull << 32) * ((1ull << l) - d)) / d + 1;
  f.m = (uint32_t)m;
  f.s = l;
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
  uint64_t t = (uint64_t)__umulhi(n, f.m) + n;
  return (uint32_t)(t >> f.s);
}

// ------------------------------------------------------------------------------------------------
// prologue: one CTA builds the table of one channel
// ------------------------------------------------------------------------------------------------
__device__ void prepare_channel(float mv, int M, int E, int K, int sign_bits, float* tab) {
  __shared__ float s_bias;
  const int tid = threadIdx.x;
  if (tid == 0) s_bias = prep_header(tab, mv, M, E, K, sign_bits);
  __syncthreads();
  const float bias = s_bias;
  for (int k = tid + 1; k <= K; k += blockDim.x) prep_entry(tab, k, M, K, bias);
  __syncthreads();
  if (tid == 0) prep_finish(tab, M, K, mv);
  __syncthreads();
}

// set_quant_range (fp8_quantizer.py:236-237): maxval = | max(|x_min|, x_max) |
__device__ __forceinline__ float range_to_maxval(float xmin, float xmax) {
  return fabsf(max_nan(fabsf(xmin), xmax));
}

__global__ void prepare_kernel(const float* __restrict__ maxval, const float* __restrict__ xmin,
                               const float* __restrict__ xmax, float* __restrict__ maxval_out, int64_t C,
                               int M, int E, int K, int sign_bits, float* __restrict__ table) {
  const int stride = table_stride(K);
  for (int64_t c = blockIdx.x; c < C; c += gridDim.x) {
    float mv;
    if (xmin != nullptr) {
      mv = range_to_maxval(xmin[c], xmax[c]);
      if (maxval_out != nullptr && threadIdx.x == 0) maxval_out[c] = mv;
    } else {
      mv = maxval[c];
    }
    prepare_channel(mv, M, E, K, sign_bits, table + c * stride);
  }
}

// ------------------------------------------------------------------------------------------------
// K1: streaming fake-quant
// ------------------------------------------------------------------------------------------------
// PRE_AFFINE  : y = Q(act(bn(x)))                 tile-local row arithmetic, bn of the tile's rows in smem
// PRE_AFFINE_G: same, any shape                   channel from a flat 64-bit-safe division, bn from global
// PRE_ADD     : y = Q(act(a + b))
// PRE_BNQ_ADD : y = Q2(act(Q1(bn(x)) + b))        the whole residual-block tail in one pass (12 B/elem)
// *_PL        : the same when H*W is not a multiple of the vector width (per-lane rows)
//
// Launch shape: ONE tile of 256 x VEC x 4 elements per CTA (grid = number of tiles).  Measured on B200
// (tools/k1_variants.cu, profiles/k1_variants_r01.txt): a persistent grid of SMs x occupancy CTAs reaches
// 5.85 TB/s on this kernel, one tile per CTA 6.80 TB/s -- resident CTAs in different phases (load / math / store)
// keep HBM reads and writes overlapped, CTAs marching in lock step do not.  The per-CTA prologue is therefore kept
// to a 20-float table copy, and batch norm is folded only for the <= 128 rows a tile touches, after the tile's
// loads have been issued.
// PRE_AFFINE_PL: PRE_AFFINE when H*W is not a multiple of the vector width (per-lane rows)
// PRE_AFFINE_CL / PRE_BNQ_ADD_CL: the same two fusions for channel-innermost memory ([pixels, C]: channels_last
//               activations -- the layout cuDNN's tensor-core convolutions produce natively -- and Linear outputs):
//               channel = flat index % C, a 128-bit vector spans 4 consecutive channels, and when C divides the
//               per-pass stride (1024 elements) every vector of a thread has the SAME 4 channels, so the
//               batch-norm parameters are loaded once per tile.
enum { PRE_PLAIN = 0, PRE_AFFINE = 1, PRE_ADD = 2, PRE_AFFINE_G = 3, PRE_BNQ_ADD = 4, PRE_AFFINE_PL = 5, PRE_BNQ_ADD_PL = 6,
       PRE_AFFINE_CL = 7, PRE_BNQ_ADD_CL = 8 };

template <int PRE>
struct PreTraits {
  static constexpr bool kCL = (PRE == PRE_AFFINE_CL || PRE == PRE_BNQ_ADD_CL);
  static constexpr bool kTail = (PRE == PRE_BNQ_ADD || PRE == PRE_BNQ_ADD_PL || PRE == PRE_BNQ_ADD_CL);
  static constexpr bool kPerLane = (PRE == PRE_AFFINE_PL || PRE == PRE_BNQ_ADD_PL);
  static constexpr bool kLocalRows = (PRE == PRE_AFFINE || PRE == PRE_AFFINE_PL || PRE == PRE_BNQ_ADD || PRE == PRE_BNQ_ADD_PL);
  static constexpr bool kTwoIn = (PRE == PRE_ADD || kTail);
  static constexpr bool kHasBn = (kLocalRows || kCL || PRE == PRE_AFFINE_G);
  static constexpr bool kBnAct = kHasBn && !kTail;   // y = Q(act(bn(x)))
};

struct StreamArgs {
  const float* x;
  const float* x2;      // PRE_ADD / PRE_BNQ_ADD: second addend
  float* y;
  int32_t* codes;
  const float* table;   // per-tensor table (PRE_BNQ_ADD: of the inner quantiser Q1)
  const float* table2;  // PRE_BNQ_ADD: table of the outer quantiser Q2
  int64_t n;
  int K, K2;
  int act;              // activation before the (outer) quantiser
  const float* bn_p[2]; // bn_mode 0: (scale, shift), [Cbn] each; bn_mode 1: bn_p[0] = packed [4 * Cbn]
  int bn_mode;          // 0: y = fma(x, scale, shift); 1: ATen-CUDA's eval batch norm, bit for bit (see BnParams)
  uint32_t hw, Cbn;
  uint32_t hw_rcp;      // ceil(2^32 / hw): umulhi(p, hw_rcp) == p / hw for p * hw < 2^32
  FastDiv hw_div, c_div;
  int cl_same;          // *_CL: Cbn divides threads * VEC, i.e. a thread sees the same channels in every vector
  int threads;          // *_CL: threads per CTA (<= kMaxDynThreads), chosen so that cl_same holds; 0 = kThreads
  int64_t ntiles;       // tiles of the launch (filled by launch_stream_t: the kernel need not divide by its tile size)
};

struct RegTab {  // K <= 3: everything in registers
  float t2, t3, s1, s2, s3, r1, r2, r3;
};

template <int KMODE>
struct ElemCtx {
  float hi, lo, guard;
  RegTab rt;            // KMODE 0
  const float* stab;    // KMODE 1: the channel table (global memory, read through L1; or a shared-memory copy)
  int K;
  uint32_t base;
  bool irregular;
  uint32_t ref, band;   // KMODE 1: exponent-arithmetic fast path (lookup_code_fast)
  bool dbl;             // KMODE 1: FLAG_SDOUBLE table -- (s, 1/s) by integer arithmetic (lookup_scale_fast), no load
  uint32_t s1b, r1b, tmax;
  bool magic;           // KMODE 0 / 1: FLAG_MAGIC table -- the element path in the scaled domain (quant_magic), no look-up
  MagicConsts mc;
  int folded_act;       // activation folded into (lo, hi) by fold_act (quant_vec_cold rebuilds the context from the table)
};

__device__ __forceinline__ float apply_act(float v, int act) {
  // torch.relu / relu6 propagate NaN (clamp semantics)
  if (act == FP8FQ_ACT_RELU) return max_nan(v, 0.0f);
  if (act == FP8FQ_ACT_RELU6) return min_nan(max_nan(v, 0.0f), 6.0f);
  return v;
}

// FP8FQ_FOLD_ACT (build option, default ON since round 2: 384 fused calls hash-identical to the unfolded build on the
// B200, step 0.573 -> 0.551 ms, profiles/ab_build_options_r02a.json): ReLU / ReLU6 in front of an FP quantiser are
// folded into the quantiser's clamp,
//   min(max(act(v), lo), hi) == min(max(v, max(lo, 0)), min(hi, 6 for ReLU6)),
// which holds for every input including NaN (max.NaN / min.NaN propagate), +-inf and signed zeros: with lo < 0 or
// lo == +0.0 the folded clamp executes the very instruction the activation did, max.NaN(v, +0.0), on the same operands
// (so the hardware's ordering of -0.0 / +0.0 does not enter), and min(min(a, 6), hi) == min(a, min(hi, 6)) for
// a >= 0, hi > 0; a zero range (lo == -0.0) yields NaN outputs either way.  The two bounds are adjusted once per thread
// and the 1-2 min/max per element of the activation disappear from the issue-bound BN variants.  Not applied to the INT quantisers (KMODE 2), whose
// clamp comes after the rounding.
#ifndef FP8FQ_FOLD_ACT
#define FP8FQ_FOLD_ACT 1
#endif
// FP8FQ_FULL_TILE (build option, default ON since round 2, same evidence; with FOLD_ACT: step 0.573 -> 0.546 ms,
// in-step roofline 0.808 -> 0.849): fq_stream_kernel instantiates its tile body a second time
// without the per-vector bounds predicates for the tiles that are full (all but the last one of a launch), so that the
// four loads / stores of a thread share one base address and the 64-bit bounds tests disappear from the common path.
#ifndef FP8FQ_FULL_TILE
#define FP8FQ_FULL_TILE 1
#endif
// FP8FQ_PIN_SEL (build option): the K <= 3 code select is  s = p3 ? s3 : (p2 ? s2 : s1)  (and the same for 1/s).  ptxas
// keeps the six table values in UNIFORM registers, and an FSEL takes at most one uniform operand, so every element pays
// two extra moves (uniform -> vector register) for the inner selects.  The option keeps s2 and 1/s2 in vector registers:
// they are re-loaded through an address ptxas cannot prove warp-uniform (threadIdx.y is 0 in every launch of this
// library) -- a value it knows to be uniform goes back to a uniform register whatever the source says.
// Measured (round 2 A/B): -24 executed instructions per tile, +-0.5 % in time -- off.
#ifndef FP8FQ_PIN_SEL
#define FP8FQ_PIN_SEL 0
#endif
__device__ __forceinline__ void pin_vreg(float& v, const float* src) {
#if FP8FQ_PIN_SEL && defined(__CUDA_ARCH__) && !defined(FP8FQ_HOST_SIM)
  v = __ldg(src + threadIdx.y);
#else
  (void)v; (void)src;
#endif
}
// FP8FQ_SDOUBLE (build option, default off): for tables whose scales are exact doublings of each other (FLAG_SDOUBLE,
// the usual case) the K > 3 element path derives (s, 1/s) from the exponent code by integer arithmetic
// (lookup_scale_fast) instead of the per-element 64-bit table load.  Bit-identical (host simulation + GPU tests of
// round 2), but measured SLOWER on the B200: the K > 3 kernels are issue / ALU bound, not L1 bound -- MobileNetV2's
// channels_last BN+ReLU6+E3M4 sites 0.797 -> 0.764 of the HBM peak at [128,96,112,112] (profiles/kernels_ab_r02c.json).
#ifndef FP8FQ_SDOUBLE
#define FP8FQ_SDOUBLE 0
#endif
// FP8FQ_MAGIC (default ON, round 2): K > 3 formats whose table carries FLAG_MAGIC (exact-doubling scales and switching
// points within kMagicMaxDev ulps of their ideal positions -- the usual case, see prep_finish in fp8fq_core.h) run the
// element path in the scaled domain u = |xc| / s_1: y = s_1 * (u rounded to M + 1 significant bits) by adding and
// subtracting a magic constant built from u's exponent field (quant_magic).  11 instructions per element after the
// clamp instead of 15+ (no code look-up, no (s, 1/s) gather, no FRND), same bits: the tie guard of the reciprocal
// multiply is kept, and a vector with a lane inside it (or a NaN) re-runs the look-up path.  The code-plane variant
// (CODES) keeps the look-up path: its exponent codes follow the reference's switching points exactly.
// Two-group tables (prep_finish) pay one compare and one select more for the scale of the element's group.
// FP8FQ_MAGIC_K0: also for the formats with <= 3 exponent codes (M >= 5), in place of their register select.
#ifndef FP8FQ_MAGIC
#define FP8FQ_MAGIC 1
#endif
#ifndef FP8FQ_MAGIC_K0
#define FP8FQ_MAGIC_K0 0
#endif
// FP8FQ_MAGIC_ONEPATH: one instantiation of the scaled-domain loop for one- and two-group tables (always the select).
// FP8FQ_MAGIC_SELFSLOW: lanes inside the tie guard are finished by an IEEE division inside the scaled-domain path
// instead of re-running the whole vector through the look-up path.
#ifndef FP8FQ_MAGIC_HOIST
#define FP8FQ_MAGIC_HOIST 0
#endif
// FP8FQ_COLD_CALL: in the K > 3 stream / row kernels everything that is not the scaled-domain loop -- the look-up path of
// the few tables without FLAG_MAGIC, and the vectors with a lane inside the tie guard -- is ONE out-of-line function
// (quant_vec_cold) instead of being inlined into each of the unrolled vector bodies: the look-up, the IEEE-division
// fallback and their slow paths are ~4/5 of those kernels' 70-80 KB of SASS, far beyond the instruction caches, and every
// variant added next to the hot loop had cost it 1-3 % (round 2 calls p -> q -> s).
// MEASURED (round 2, call t): 0.85-0.89 -> 0.59 of the HBM peak -- a call in the kernel takes the address of the vector
// arrays, which then live in local memory on the hot path too.  Off; kept as a negative result.
#ifndef FP8FQ_COLD_CALL
#define FP8FQ_COLD_CALL 0
#endif
// FP8FQ_MAGIC_TWO (default ON): two-group tables on the scaled-domain path too -- a second instantiation of its loop in
// every vector body, with the compare + select of the element's scale group, chosen by a uniform branch on a bit of the
// table's flags word.  Measured (round 2, E3M4 channel-innermost sites, mean fraction of the HBM peak): one-group tables
// 0.920 -> 0.904, two-group ones 0.779 (look-up path) -> 0.869; BASELINE config 3 with random-init weights (62 of its 63
// calibrated tables are one-group) 0.865 -> 0.854.  About half of the ranges in [2, 8) -- a ReLU6 network saturating at
// maxval = 6 -- and a sixth of all ranges are two-group, so ON is the better or equal choice everywhere but on that
// synthetic workload.  (The first version branched on a bool derived from a float compare, which ptxas re-evaluated
// per vector in vector registers: +14 instructions per vector, one-group tables at 0.87 -- calls u, v.)
#ifndef FP8FQ_MAGIC_TWO
#define FP8FQ_MAGIC_TWO 1
#endif
#if defined(FP8FQ_HOST_SIM)
#define FQ_NOINLINE __attribute__((noinline))
#else
#define FQ_NOINLINE __noinline__
#endif
#ifndef FP8FQ_MAGIC_ONEPATH
#define FP8FQ_MAGIC_ONEPATH 0
#endif
#ifndef FP8FQ_MAGIC_SELFSLOW
#define FP8FQ_MAGIC_SELFSLOW 0
#endif
// FP8FQ_PACK2 (build option): the independent fp32 multiplies / adds / FMAs of neighbouring elements are issued as
// sm_100's two-wide instructions (FMUL2 / FADD2 / FFMA2: same IEEE round-to-nearest results, half the issue slots).
// Measured (round 2 A/B): K <= 3 stream kernels 0 %, K > 3 +1..2 %, MSE-grid kernel -2..4 % -- off.
#ifndef FP8FQ_PACK2
#define FP8FQ_PACK2 0
#endif
__device__ __forceinline__ void mul2_rn(float a0, float a1, float b0, float b1, float& r0, float& r1) {
#if FP8FQ_PACK2 && defined(__CUDA_ARCH__)
  const float2 r = __fmul2_rn(make_float2(a0, a1), make_float2(b0, b1));
  r0 = r.x; r1 = r.y;
#else
  r0 = mul_rn(a0, b0); r1 = mul_rn(a1, b1);
#endif
}
__device__ __forceinline__ void sub2_rn(float a0, float a1, float b0, float b1, float& r0, float& r1) {
#if FP8FQ_PACK2 && defined(__CUDA_ARCH__)
  const float2 r = __fadd2_rn(make_float2(a0, a1), make_float2(-b0, -b1));
  r0 = r.x; r1 = r.y;
#else
  r0 = sub_rn(a0, b0); r1 = sub_rn(a1, b1);
#endif
}
__device__ __forceinline__ void fma2_rn(float a0, float a1, float b0, float b1, float c0, float c1, float& r0, float& r1) {
#if FP8FQ_PACK2 && defined(__CUDA_ARCH__)
  const float2 r = __ffma2_rn(make_float2(a0, a1), make_float2(b0, b1), make_float2(c0, c1));
  r0 = r.x; r1 = r.y;
#else
  r0 = fmaf(a0, b0, c0); r1 = fmaf(a1, b1, c1);
#endif
}

template <int KMODE>
__device__ __forceinline__ void fold_act(ElemCtx<KMODE>& c, int act) {
  if (act == FP8FQ_ACT_RELU || act == FP8FQ_ACT_RELU6) c.lo = max_nan(c.lo, 0.0f);
  if (act == FP8FQ_ACT_RELU6) c.hi = min_nan(c.hi, 6.0f);
  c.folded_act = act;
}

// Quantises N values.  One exactness check per vector: the IEEE-division fallback is entered by the whole
// vector when any lane is within the guard band of a rounding tie (probability ~ N * 2^(M-19)).
// GUARD = false drops the exactness check of the reciprocal multiply (only for consumers that tolerate q off by one
// on an exact rounding tie -- the MSE kernel, where either neighbour is equally far from x -- and only for tables
// without FLAG_RSNAN).
// SIGNED_OUT = false: the caller only needs |y| (the MSE kernel, which forms |x| - |y|); the magic path then skips
// restoring the sign.
template <int KMODE, bool CODES, int N>
__device__ FQ_NOINLINE void quant_vec_cold(const float* v, float* y, int32_t* code, const float* table, int K, int folded_act);

// MM: what the caller already knows about the table (fq_stream_kernel decides once per launch and instantiates its tile
// loop per case, FP8FQ_MAGIC_HOIST): 0 = nothing, look at c.magic / c.mc.two here; 1 = a one-group FLAG_MAGIC table;
// 2 = a two-group one; 3 = not a FLAG_MAGIC table.  With MM = 1 / 2 the lanes inside the tie guard are finished inside
// the scaled-domain path, so those instantiations carry no look-up code at all.
// NONNEG: the caller passes |x| and the format is signed (clamp bounds -maxval / +maxval): |clamp(x)| = min(|x|, maxval),
// so the lower clamp is skipped (the MSE kernel, which takes |x| once and sweeps hundreds of candidate tables over it).
template <int KMODE, bool CODES, int N, bool STAB_SHARED = false, bool GUARD = true, bool SIGNED_OUT = true, int MM = 0,
          bool NONNEG = false>
__device__ __forceinline__ void quant_vec(const float (&v)[N], const ElemCtx<KMODE>& c, float (&y)[N], int32_t (&code)[N],
                                          float* s_out = nullptr) {
  if (KMODE == 2) {  // INT uniform quantiser: c.rt = {zp, sat, scale, -, -, 1/scale, -, -}, c.lo/hi = int_min/int_max
    float q[N];
    bool slow = false;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const float r = mul_rn(v[k], c.rt.r1);
      q[k] = nearbyintf(r);
      slow |= !(fabsf(r - q[k]) < c.guard) && !(fabsf(r) >= c.rt.t3);
    }
    if (slow) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const float r = mul_rn(v[k], c.rt.r1);
        if (!(fabsf(r - q[k]) < c.guard) && !(fabsf(r) >= c.rt.t3)) q[k] = nearbyintf(div_rn(v[k], c.rt.s1));
      }
    }
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const float xi = min_nan(max_nan(add_rn(q[k], c.rt.t2), c.lo), c.hi);
      y[k] = mul_rn(c.rt.s1, sub_rn(xi, c.rt.t2));
      if (CODES) code[k] = (y[k] != y[k]) ? 0x7fffffff : (int32_t)xi;
    }
    return;
  }
  float xc[N], s[N], rs[N], q[N];
  int e[N];
  bool slow = false;
#pragma unroll
  for (int k = 0; k < N; ++k) xc[k] = NONNEG ? min_nan(v[k], c.hi) : min_nan(max_nan(v[k], c.lo), c.hi);
  constexpr bool kCanMagic = (KMODE == 1 || (KMODE == 0 && FP8FQ_MAGIC_K0)) && FP8FQ_MAGIC && !CODES;
  if (kCanMagic && s_out == nullptr && MM != 3 && (MM != 0 || c.magic)) {
    // scaled-domain path (FP8FQ_MAGIC above); one exactness check per vector, like the look-up path
    bool all_ok = true;
    auto body = [&](auto two_tag) {
      constexpr bool kTwo = decltype(two_tag)::value;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const float a = fabsf(xc[k]);
        const float u = mul_rn(a, c.mc.r1);
        uint32_t cb = f2u(u) & 0x7f800000u;
        cb = (cb < c.mc.lo ? c.mc.lo : cb) + c.mc.add;
        const float C = u2f(cb);
        const float qu = sub_rn(add_rn(u, C), C);
        if (GUARD) all_ok &= fabsf(sub_rn(u, qu)) < mul_rn(C, c.mc.kap);
        const float ya = mul_rn(qu, kTwo ? (a >= c.mc.tb ? c.mc.sb : c.mc.s1) : c.mc.s1);
        y[k] = SIGNED_OUT ? u2f(f2u(ya) | (f2u(xc[k]) & 0x80000000u)) : ya;
      }
    };
    if (FP8FQ_MAGIC_TWO && (MM == 2 || (MM == 0 && (FP8FQ_MAGIC_ONEPATH || c.mc.two)))) body(std::true_type{});   // (one group: tb is NaN, the select keeps s1)
    else body(std::false_type{});
    if (!GUARD || all_ok) return;
    if (FP8FQ_MAGIC_SELFSLOW || MM != 0) {
    // a lane within the guard band of a rounding tie (or NaN): the reference's own arithmetic for that lane -- IEEE
    // division by the scale of its code, s' * 2^p (a tie is half a step away from every code boundary, so p is the
    // reference's code here) -- without leaving the scaled-domain path
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const float a = fabsf(xc[k]);
      const float u = mul_rn(a, c.mc.r1);
      uint32_t t = f2u(u) & 0x7f800000u;
      t = t < c.mc.lo ? c.mc.lo : t;
      const float C = u2f(t + c.mc.add);
      const float qu = sub_rn(add_rn(u, C), C);
      if (!(fabsf(sub_rn(u, qu)) < mul_rn(C, c.mc.kap))) {
        const float se = u2f(f2u(a >= c.mc.tb ? c.mc.sb : c.mc.s1) + (t - c.mc.lo));
        const float ya = mul_rn(nearbyintf(div_rn(a, se)), se);
        y[k] = SIGNED_OUT ? u2f(f2u(ya) | (f2u(xc[k]) & 0x80000000u)) : ya;
      }
    }
    return;
    }
  }
  if (FP8FQ_COLD_CALL && kCanMagic && KMODE == 1 && MM == 0 && !STAB_SHARED && GUARD && SIGNED_OUT && !NONNEG && s_out == nullptr) {
    quant_vec_cold<KMODE, CODES, N>(v, y, code, c.stab, c.K, c.folded_act);
    return;
  }
  if (KMODE == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const float a = fabsf(xc[k]);
      const bool p2 = a >= c.rt.t2, p3 = a >= c.rt.t3;
      s[k] = p3 ? c.rt.s3 : (p2 ? c.rt.s2 : c.rt.s1);
      rs[k] = p3 ? c.rt.r3 : (p2 ? c.rt.r2 : c.rt.r1);
      e[k] = 1 + (p2 ? 1 : 0) + (p3 ? 1 : 0);
    }
  } else {
    // exponent-arithmetic code; the table lookup runs for the whole vector only when a lane's mantissa lies in the
    // (band + 1) / 2^23 ambiguous band.  c.stab is the global table (stream/row kernels, L1-resident) or a
    // shared-memory copy (MSE kernel), hence generic-address loads.
    bool amb = false;
    if (c.dbl) {
      // exact-doubling scale table (the usual case): the scale pair comes out of the same integer arithmetic as the
      // code -- no per-element load at all (ncu, round 2: the per-element 64-bit gathers put the L1 data pipe of the
      // K > 3 kernels at 64-69 % next to 74 % issue utilisation)
#pragma unroll
      for (int k = 0; k < N; ++k) {
        bool ak;
        const uint32_t t = lookup_scale_fast(fabsf(xc[k]), c.ref, c.band, c.tmax, c.s1b, c.r1b, &s[k], &rs[k], &ak);
        e[k] = (int)(t >> 23) + 1;
        amb |= ak;
      }
    } else {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        bool ak;
        e[k] = lookup_code_fast(fabsf(xc[k]), c.ref, c.band, c.K, &ak);
        amb |= ak;
      }
    }
    if (amb) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const int ee = lookup_code(fabsf(xc[k]), c.stab, c.K, c.base, c.irregular, [](const float* p) { return *p; });
        e[k] = ee < 1 ? 1 : ee;
      }
    }
    const float2* sr = reinterpret_cast<const float2*>(c.stab + off_sr(c.K));
    if (c.dbl && !amb) {
      // (s, 1/s) already in registers
    } else
#ifndef FP8FQ_HOST_SIM
    if (STAB_SHARED) {  // the table is a shared-memory copy: ld.shared instead of a generic-address load
      const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sr);
#pragma unroll
      for (int k = 0; k < N; ++k)
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(s[k]), "=f"(rs[k]) : "r"(sbase + 8u * (uint32_t)e[k]));
    } else
#endif
    {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const float2 p = sr[e[k]];
        s[k] = p.x;
        rs[k] = p.y;
      }
    }
  }
  if (FP8FQ_PACK2 && N % 2 == 0) {
#pragma unroll
    for (int k = 0; k + 1 < N; k += 2) {
      float r0, r1, d0, d1;
      mul2_rn(xc[k], xc[k + 1], rs[k], rs[k + 1], r0, r1);
      q[k] = nearbyintf(r0);
      q[k + 1] = nearbyintf(r1);
      if (GUARD) {
        sub2_rn(r0, r1, q[k], q[k + 1], d0, d1);
        slow |= !(fabsf(d0) < c.guard) | !(fabsf(d1) < c.guard);
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const float r = mul_rn(xc[k], rs[k]);
      q[k] = nearbyintf(r);
      if (GUARD) slow |= !(fabsf(r - q[k]) < c.guard);
    }
  }
  if (GUARD && slow) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const float r = mul_rn(xc[k], rs[k]);
      if (!(fabsf(r - q[k]) < c.guard)) q[k] = nearbyintf(div_rn(xc[k], s[k]));  // near a tie, or rs unusable
    }
  }
  if (FP8FQ_PACK2 && N % 2 == 0) {
#pragma unroll
    for (int k = 0; k + 1 < N; k += 2) mul2_rn(q[k], q[k + 1], s[k], s[k + 1], y[k], y[k + 1]);
  } else {
#pragma unroll
    for (int k = 0; k < N; ++k) y[k] = mul_rn(q[k], s[k]);
  }
#pragma unroll
  for (int k = 0; k < N; ++k) {
    if (s_out) s_out[k] = s[k];
    if (CODES) {
      if (y[k] != y[k]) code[k] = 0x7fffffff;
      else code[k] = (int32_t)((f2u(y[k]) & 0x80000000u) | ((uint32_t)e[k] << 16) | (uint32_t)fabsf(q[k]));
    }
  }
}

template <int KMODE, bool CODES>
__device__ __forceinline__ float quant_elem(float v, const ElemCtx<KMODE>& c, int32_t* code) {
  float vi[1] = {v}, yo[1];
  int32_t cd[1];
  quant_vec<KMODE, CODES, 1>(vi, c, yo, cd);
  if (CODES) *code = cd[0];
  return yo[0];
}

// Fills the element context from a channel table through `ld` (global: uniform __ldg loads -- every thread reads the
// tiny, L1/L2-resident table itself, no shared memory, no barrier, so a CTA is a fully independent streaming unit;
// the MSE kernel passes a shared-memory copy).
template <int KMODE, typename Ld>
__device__ __forceinline__ void load_ctx(ElemCtx<KMODE>& c, const float* tab, int K, Ld ld) {
  c.folded_act = FP8FQ_ACT_NONE;
  if (KMODE == 2) {  // uniform table
    c.hi = ld(tab + U_IMAX);
    c.lo = ld(tab + U_IMIN);
    c.guard = ld(tab + U_GUARD);
    c.rt.s1 = ld(tab + U_SCALE);
    c.rt.r1 = ld(tab + U_RS);
    c.rt.t2 = ld(tab + U_ZP);
    c.rt.t3 = ld(tab + U_SAT);
    c.rt.s2 = c.rt.s3 = c.rt.r2 = c.rt.r3 = 0.0f;
    c.K = 1; c.stab = tab; c.base = 0; c.irregular = false; c.ref = 0; c.band = 0;
    c.dbl = false; c.s1b = c.r1b = c.tmax = 0;
    c.magic = false;
    return;
  }
  c.hi = ld(tab + H_HI);
  c.lo = ld(tab + H_LO);
  c.guard = ld(tab + H_GUARD);
  c.K = K;
  c.stab = tab;
  if (KMODE == 0) {
    const float never = __int_as_float(0x7fc00000);  // NaN: "a >= never" is false
    const float* thr = tab + kHdr;
    const float* sr = tab + off_sr(K);
    c.rt.t2 = K >= 2 ? ld(thr + 1) : never;
    c.rt.t3 = K >= 3 ? ld(thr + 2) : never;
    c.rt.s1 = ld(sr + 2); c.rt.r1 = ld(sr + 3);
    c.rt.s2 = K >= 2 ? ld(sr + 4) : c.rt.s1; c.rt.r2 = K >= 2 ? ld(sr + 5) : c.rt.r1;
    c.rt.s3 = K >= 3 ? ld(sr + 6) : c.rt.s2; c.rt.r3 = K >= 3 ? ld(sr + 7) : c.rt.r2;
    c.base = 0;
    c.irregular = false;
    c.ref = 0;
    c.band = 0;
    c.dbl = false; c.s1b = c.r1b = c.tmax = 0;
    c.magic = false;
    if (FP8FQ_MAGIC && FP8FQ_MAGIC_K0) {
      const uint32_t fl = f2u(ld(tab + H_FLAGS));
      c.mc = magic_consts(tab, K, fl, ld);
      c.magic = (fl & FLAG_MAGIC) != 0 && (FP8FQ_MAGIC_TWO || !c.mc.two);
    }
  } else {
    const uint32_t fl = f2u(ld(tab + H_FLAGS));
    c.base = f2u(ld(tab + H_BASE));
    c.irregular = (fl & FLAG_IRREGULAR) != 0;
    c.ref = f2u(ld(tab + H_REF));
    c.band = flags_band(fl);
    c.dbl = FP8FQ_SDOUBLE && (fl & FLAG_SDOUBLE) != 0;
    c.s1b = f2u(ld(tab + off_sr(K) + 2));
    c.r1b = f2u(ld(tab + off_sr(K) + 3));
    c.tmax = (uint32_t)(K - 1) << 23;
    c.mc = magic_consts(tab, K, fl, ld);
    c.magic = FP8FQ_MAGIC && (fl & FLAG_MAGIC) != 0 && (FP8FQ_MAGIC_TWO || !c.mc.two);
  }
}
template <int KMODE>
__device__ __forceinline__ void load_ctx_direct(ElemCtx<KMODE>& c, const float* __restrict__ gtab, int K) {
  load_ctx<KMODE>(c, gtab, K, [](const float* p) { return __ldg(p); });
}

// The out-of-line rest of quant_vec (FP8FQ_COLD_CALL): rebuilds the element context from the (global-memory) table and
// runs the vector through the look-up path (MM = 3) -- the IEEE-division fallback included.
template <int KMODE, bool CODES, int N>
__device__ FQ_NOINLINE void quant_vec_cold(const float* v, float* y, int32_t* code, const float* table, int K, int folded_act) {
  ElemCtx<KMODE> c;
  load_ctx_direct<KMODE>(c, table, K);
  if (folded_act != FP8FQ_ACT_NONE) fold_act(c, folded_act);
  float vv[N], yy[N];
  int32_t cc[N];
#pragma unroll
  for (int k = 0; k < N; ++k) vv[k] = v[k];
  quant_vec<KMODE, CODES, N, false, true, true, 3>(vv, c, yy, cc);
#pragma unroll
  for (int k = 0; k < N; ++k) {
    y[k] = yy[k];
    if (CODES) code[k] = cc[k];
  }
}

// Per-channel batch-norm parameters of one vector.
//   mode 0 (affine):  y = fma(x, scale, shift), scale = gamma / sqrt(var + eps), shift = beta - mean * scale
//   mode 1 (exact) :  y = fma(gamma * (x - mean), rsqrtf(var + eps), beta) -- the arithmetic of ATen's eval-mode
//                     batch norm on CUDA (native batch_norm_transform_input_kernel), measured bit-identical to
//                     F.batch_norm on B200 for every element (tools/bn_formula.py, profiles/bn_formula_r01.json)
struct BnParams {
  float a, b, c, d;
};
template <int BNM>
__device__ __forceinline__ BnParams bn_load(const StreamArgs& a, uint32_t ch) {
  BnParams p;
  if (BNM == 0) {
    p.a = __ldg(a.bn_p[0] + ch);
    p.b = __ldg(a.bn_p[1] + ch);
    p.c = 0.0f;
    p.d = 0.0f;
  } else {
    // one 128-bit load: {mean, gamma, rsqrtf(var + eps), beta} of channel ch
    const float4 q = __ldg(reinterpret_cast<const float4*>(a.bn_p[0]) + ch);
    p.a = q.x; p.b = q.y; p.c = q.z; p.d = q.w;
  }
  return p;
}
template <int BNM>
__device__ __forceinline__ float bn_apply(float v, const BnParams& p) {
  if (BNM == 0) return fmaf(v, p.a, p.b);
  return fmaf(mul_rn(p.b, sub_rn(v, p.a)), p.c, p.d);
}

// batch norm of the VEC lanes of one vector that share a channel (NCHW rows)
template <int BNM, int VEC>
__device__ __forceinline__ void bn_apply_vec(const float (&in)[VEC], const BnParams& p, float (&v)[VEC]) {
  if (FP8FQ_PACK2 && VEC % 2 == 0) {
#pragma unroll
    for (int k = 0; k + 1 < VEC; k += 2) {
      if (BNM == 0) {
        fma2_rn(in[k], in[k + 1], p.a, p.a, p.b, p.b, v[k], v[k + 1]);
      } else {
        float t0, t1;
        sub2_rn(in[k], in[k + 1], p.a, p.a, t0, t1);
        mul2_rn(p.b, p.b, t0, t1, t0, t1);
        fma2_rn(t0, t1, p.c, p.c, p.d, p.d, v[k], v[k + 1]);
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < VEC; ++k) v[k] = bn_apply<BNM>(in[k], p);
  }
}

constexpr int kThreads = 256;
// independent 128-bit loads in flight per thread: 4 vectors, or 2 vectors x 2 inputs for the residual variants
// (same bytes in flight, 16 fewer live registers -> no spills at 5-6 resident CTAs per SM)
template <int PRE>
struct StreamUnroll {
  static constexpr int value = PreTraits<PRE>::kTwoIn ? 2 : 4;
};

// Resident CTAs per SM the kernel is compiled for.  Measured (tools/bench_kernels.py, B200, [128,64,112,112]):
// one-tile CTAs are short lived, so occupancy is what keeps HBM busy -- plain E2M5 5.9 -> 6.5 TB/s going from 4 to
// 6 CTAs/SM; the two-input variants spill at 6 and are best at 5.  The channel-innermost variants fit 40 registers since
// the lane-major batch norm and FOLD_ACT: 6 CTAs/SM for the constant-CTA-size instantiations (round 2 A/B,
// profiles/ab_build_options_r02i.json: BN+ReLU+quant 0.944 -> 0.957, block tail 0.938 -> 0.968 of the HBM peak,
// channels_last step 0.559 -> 0.552 ms), 5 for the run-time-CTA-size ones (DYN: they spill at 6 and lose 2-4 %).
// FQ_MINB / FQ_MINB_CL / FQ_MINB_CL_DYN override for tuning builds.
template <int KMODE, int PRE, bool DYN = false>
struct StreamMinBlocks {
#ifdef FQ_MINB
  static constexpr int value = FQ_MINB;
#else
#ifndef FQ_MINB_CL
#define FQ_MINB_CL 6
#endif
#ifndef FQ_MINB_CL_DYN
#define FQ_MINB_CL_DYN 4   // x kMaxDynThreads = 320 threads: the same 48-register budget as 5 x 256
#endif
  static constexpr int value =
      PreTraits<PRE>::kCL ? (DYN ? FQ_MINB_CL_DYN : FQ_MINB_CL)
      : ((PRE == PRE_PLAIN || PRE == PRE_AFFINE || PRE == PRE_AFFINE_PL || PRE == PRE_AFFINE_G) && KMODE != 1) ? 6 : 5;
#endif
};

// DYN: the CTA size is a launch parameter (channel-innermost variants whose channel count does not divide
// kThreads * VEC); everywhere else it is the compile-time kThreads, which keeps the index arithmetic constant-folded.
// (up to kMaxDynThreads = 320 threads, so that MobileNetV2's 1280-channel layer -- 320 four-channel lanes -- gets a
// fitted CTA as well; the register budget, 65536 / (320 * 4) -> 48, is that of 5 CTAs of 256 threads)
constexpr int kMaxDynThreads = 320;
template <int KMODE, int PRE, int VEC, bool CODES, int BNM, bool DYN = false>
__global__ void __launch_bounds__(DYN ? kMaxDynThreads : kThreads, StreamMinBlocks<KMODE, PRE, DYN>::value)
fq_stream_kernel(const StreamArgs a) {
  constexpr bool kTail = PreTraits<PRE>::kTail;
  constexpr bool kPerLane = PreTraits<PRE>::kPerLane;
  constexpr bool kLocalRows = PreTraits<PRE>::kLocalRows;
  constexpr bool kTwoIn = PreTraits<PRE>::kTwoIn;
  constexpr bool kCL = PreTraits<PRE>::kCL;
  constexpr bool kBnAct = PreTraits<PRE>::kBnAct;
  constexpr int kUnroll = StreamUnroll<PRE>::value;

  // the channel-innermost variants run with the largest CTA size whose pass stride (threads * VEC) is a multiple of the
  // channel count, so that a thread meets the same channels in every vector (bn_act_quant_nhwc_impl).  Their tensors
  // have fewer than 2^31 elements (checked by the launcher), so all of their index arithmetic is 32-bit: the run-time
  // CTA size of the DYN instantiations otherwise costs a 64-bit multiply-add chain and two-instruction bounds tests
  // per vector (round 2, static SASS: 23 -> 9 address / bounds instructions per vector).
  using idx_t = typename std::conditional<kCL, uint32_t, int64_t>::type;
  const int nthr = DYN ? (int)blockDim.x : kThreads;
  const idx_t kTile = (idx_t)nthr * VEC * kUnroll;
  const idx_t nvec_elems = (idx_t)a.n - (idx_t)(a.n % VEC);
  const idx_t ntiles = (idx_t)a.ntiles;
  pdl_prologue();
  // per-CTA parameters: uniform loads, no barrier; issued first, consumed only after the data loads went out
  ElemCtx<KMODE> ctx, ctx2;
  load_ctx_direct<KMODE>(ctx, a.table, a.K);
  if (kTail) load_ctx_direct<KMODE>(ctx2, a.table2, a.K2);
  if (KMODE == 0 && FP8FQ_PIN_SEL) {
    if (a.K >= 2) { pin_vreg(ctx.rt.s2, a.table + off_sr(a.K) + 4); pin_vreg(ctx.rt.r2, a.table + off_sr(a.K) + 5); }
    if (kTail && a.K2 >= 2) { pin_vreg(ctx2.rt.s2, a.table2 + off_sr(a.K2) + 4); pin_vreg(ctx2.rt.r2, a.table2 + off_sr(a.K2) + 5); }
  }
#if FP8FQ_FOLD_ACT
  // the activation (if any) feeds the LAST quantiser of the launch: ctx2 in the block tail, ctx otherwise
  constexpr bool kFoldAct = KMODE != 2 && (kBnAct || PRE == PRE_ADD || kTail);
  if (kFoldAct) fold_act(kTail ? ctx2 : ctx, a.act);
  const int act_left = kFoldAct ? FP8FQ_ACT_NONE : a.act;
#define FQ_ACT act_left
#else
#define FQ_ACT a.act   // (-DFP8FQ_FOLD_ACT=0 -DFP8FQ_FULL_TILE=0: token for token the code of the round-1 profiles)
#endif

  // FP8FQ_MAGIC_HOIST: which element path the launch's table takes is decided HERE, once, and the tile loop is
  // instantiated per case (quant_vec's MM), so that the per-vector code is straight-line: the scaled-domain loop of that
  // table kind with its own finish for tie-guard lanes, or the look-up path -- not all three behind uniform branches in
  // every vector.  (The block tails, with two tables per element, keep the run-time form.)
  constexpr bool kHoist = FP8FQ_MAGIC_HOIST && FP8FQ_MAGIC && !CODES && !kTail && (KMODE == 1 || (KMODE == 0 && FP8FQ_MAGIC_K0));
  auto run_tiles = [&](auto mm_tag) {
  constexpr int MM = decltype(mm_tag)::value;
  for (idx_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const idx_t tile0 = tile * kTile;
#if FP8FQ_FULL_TILE
    // the tile body is instantiated twice: every tile but the last one is full and needs no bounds predicates
    auto tile_body = [&](auto full_tag) {
    constexpr bool kFullTile = decltype(full_tag)::value;
#else
    constexpr bool kFullTile = false;
#endif
    const idx_t base = tile0 + (idx_t)threadIdx.x * VEC;
    // 1. all of this thread's loads go out before any of them is used
    Pack<VEC> in[kUnroll], in2[kUnroll];
    bool ok[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const idx_t i = base + (idx_t)u * nthr * VEC;
      ok[u] = kFullTile || i < nvec_elems;
      if (ok[u]) {
        in[u].load(a.x + i);
        if (kTwoIn) in2[u].load(a.x2 + i);
      }
    }
    uint32_t col0 = 0, ch0 = 0;
    if (kLocalRows) {
      // tile-local row arithmetic: one division per tile, then one multiply-high per vector
      const uint32_t row0 = fdiv((uint32_t)tile0, a.hw_div);
      col0 = (uint32_t)tile0 - row0 * a.hw;
      ch0 = row0 - fdiv(row0, a.c_div) * a.Cbn;
    }
    // 3. math + stores
    if (kCL) {
      // channel-innermost: lanes are channels ch .. ch + VEC - 1 (Cbn % VEC == 0, checked by the launcher).  Batch norm
      // is applied to ALL of the thread's vectors first, in place, so that the parameter registers are dead before
      // the quantiser's table registers come alive.
      // One channel's parameters are live at a time (lane k of every vector, then lane k + 1, ...).
      if (a.cl_same) {       // Cbn divides the pass stride: every vector of this thread has the same channels
        const uint32_t i32 = (uint32_t)base;
        const uint32_t ch = i32 - fdiv(i32, a.c_div) * a.Cbn;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const BnParams bp = bn_load<BNM>(a, ch + k);
#pragma unroll
          for (int u = 0; u < kUnroll; ++u)
            if (ok[u]) in[u].v[k] = bn_apply<BNM>(in[u].v[k], bp);
        }
      } else {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          if (!ok[u]) continue;
          const uint32_t i32 = (uint32_t)(base + (idx_t)u * nthr * VEC);
          const uint32_t ch = i32 - fdiv(i32, a.c_div) * a.Cbn;
#pragma unroll
          for (int k = 0; k < VEC; ++k) in[u].v[k] = bn_apply<BNM>(in[u].v[k], bn_load<BNM>(a, ch + k));
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (!ok[u]) continue;
      const idx_t i = base + (idx_t)u * nthr * VEC;
      float v[VEC], yv[VEC];
      int32_t cd[VEC];
      if (kCL) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[k] = in[u].v[k];
      } else if (kLocalRows) {
        const uint32_t p = col0 + (uint32_t)(i - tile0);
        if (!kPerLane) {
          uint32_t ch = ch0 + __umulhi(p, a.hw_rcp);
          ch = ch >= a.Cbn ? ch - a.Cbn : ch;
          const BnParams bp = bn_load<BNM>(a, ch);
          bn_apply_vec<BNM, VEC>(in[u].v, bp, v);
        } else {
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            uint32_t ch = ch0 + __umulhi(p + k, a.hw_rcp);
            ch = ch >= a.Cbn ? ch - a.Cbn : ch;
            v[k] = bn_apply<BNM>(in[u].v[k], bn_load<BNM>(a, ch));
          }
        }
      } else if (PRE == PRE_AFFINE_G) {
        // all VEC lanes of a vector share a row because hw % VEC == 0 (checked by the launcher)
        const uint32_t row = fdiv((uint32_t)i, a.hw_div);
        const uint32_t ch = row - fdiv(row, a.c_div) * a.Cbn;
        const BnParams bp = bn_load<BNM>(a, ch);
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[k] = bn_apply<BNM>(in[u].v[k], bp);
      } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[k] = in[u].v[k];
      }
      if (kBnAct) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[k] = apply_act(v[k], FQ_ACT);
      } else if (PRE == PRE_ADD) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[k] = apply_act(add_rn(v[k], in2[u].v[k]), FQ_ACT);
      } else if (kTail) {
        float t[VEC];
        quant_vec<KMODE, false, VEC>(v, ctx, t, cd);  // inner quantiser (output of the block's last BN)
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[k] = apply_act(add_rn(t[k], in2[u].v[k]), FQ_ACT);
      }
      if (kTail) quant_vec<KMODE, CODES, VEC>(v, ctx2, yv, cd);
      else quant_vec<KMODE, CODES, VEC, false, true, true, MM>(v, ctx, yv, cd);
      Pack<VEC> out;
      IPack<VEC> co;
#pragma unroll
      for (int k = 0; k < VEC; ++k) { out.v[k] = yv[k]; co.v[k] = cd[k]; }
      out.store(a.y + i);
      if (CODES) co.store(a.codes + i);
    }
#if FP8FQ_FULL_TILE
    };
    // (not for the channel-innermost variants: with the predicates gone the compiler hoists all their batch-norm
    // parameter loads and spills -- static SASS, profiles/static_build_options_r01.json)
#ifndef FP8FQ_FULL_TILE_CL
#define FP8FQ_FULL_TILE_CL 0
#endif
    if ((!kCL || FP8FQ_FULL_TILE_CL) && tile0 + kTile <= nvec_elems) tile_body(std::true_type{});
    else tile_body(std::false_type{});
#endif
  }
  };
  if (kHoist) {
    if (!ctx.magic) run_tiles(std::integral_constant<int, 3>{});
    else if (ctx.mc.two) run_tiles(std::integral_constant<int, 2>{});
    else run_tiles(std::integral_constant<int, 1>{});
  } else {
    run_tiles(std::integral_constant<int, 0>{});
  }
  // scalar tail (n % VEC elements), VEC == 4 only
  if (VEC > 1 && blockIdx.x == 0 && threadIdx.x < (int)(a.n - nvec_elems)) {
    const int64_t i = nvec_elems + threadIdx.x;
    float v = a.x[i];
    if (kCL) {
      const uint32_t ch = (uint32_t)i - fdiv((uint32_t)i, a.c_div) * a.Cbn;
      v = bn_apply<BNM>(v, bn_load<BNM>(a, ch));
    } else if (kLocalRows || PRE == PRE_AFFINE_G) {
      const uint32_t row = fdiv((uint32_t)i, a.hw_div);
      const uint32_t ch = row - fdiv(row, a.c_div) * a.Cbn;
      v = bn_apply<BNM>(v, bn_load<BNM>(a, ch));
    }
    int32_t cd;
    if (kBnAct) v = apply_act(v, FQ_ACT);
    else if (PRE == PRE_ADD) v = apply_act(add_rn(v, a.x2[i]), FQ_ACT);
    else if (kTail) v = apply_act(add_rn(quant_elem<KMODE, false>(v, ctx, &cd), a.x2[i]), FQ_ACT);
    a.y[i] = kTail ? quant_elem<KMODE, CODES>(v, ctx2, &cd) : quant_elem<KMODE, CODES>(v, ctx, &cd);
    if (CODES) a.codes[i] = cd;
  }
}

#undef FQ_ACT

// ------------------------------------------------------------------------------------------------
// K1 per-channel: x is [C, inner]; one CTA per (row, chunk) work item, row table staged in smem
// ------------------------------------------------------------------------------------------------
constexpr int kMaxMulti = 48;  // tensors per launch (descriptors travel in the kernel parameter space)

struct RowsTensor {
  const float* x;
  float* y;
  int32_t* codes;
  const float* table;
  int64_t C, inner;
  int64_t chunks_per_row;   // ceil(inner / kRowsWarpChunk)
  int64_t work0;            // first work item of this tensor
  int vec_ok;               // inner % 4 == 0 and pointers 16B aligned
};

struct RowsArgs {
  RowsTensor t[kMaxMulti];
  int count;
  int64_t nwork;
  int K;
  int stride;               // floats per channel table
};

// One WARP per (tensor, row, 1024-element chunk) work item, all of the lane's (up to 8) 128-bit loads issued before
// the first is used.  Several weight tensors of the same format share one launch (a model's 21..53 weight tensors
// are each far too small to fill the GPU or to amortise a launch on their own).  Weight rows are short (ResNet-18:
// 147 .. 4608 elements, MobileNetV2: 9 .. 1280), so memory-level parallelism has to come from within the thread, and
// a warp is the unit that matches a row.  Measured on B200 (tools/bench_kernels.py, ResNet-18's 21 tensors, 93 MB,
// L2 flushed): 20.2 us = 4.6 TB/s; the earlier CTA-per-(row, 4096-chunk) kernel with one load in flight per thread
// took 27.0 us.
constexpr int kRowsWarpChunk = 1024;
constexpr int kRowsWarps = 4;
// FQ_ROWS_MINB: resident CTAs per SM the K > 3 row kernel is compiled for (6 x 128 threads -> at most 80 registers; with
// both element paths inlined it would otherwise take 95: MobileNetV2's 53-tensor weight call 29.1 -> 23.2 us)
#ifndef FQ_ROWS_MINB
#define FQ_ROWS_MINB 6
#endif
template <int KMODE, bool CODES>
__device__ __forceinline__ void fq_rows_body(const RowsArgs& a);
// (two entry points because a launch bound cannot be left out per instantiation: the K <= 3 / INT row kernels keep the
// unconstrained 71 registers of round 1 -- any stated minimum, even one their 71 registers satisfy, made ResNet-18's
// weight call 2 us slower)
template <int KMODE, bool CODES>
__global__ void __launch_bounds__(kRowsWarps * 32) fq_rows_kernel(const __grid_constant__ RowsArgs a) {
  fq_rows_body<KMODE, CODES>(a);
}
template <int KMODE, bool CODES>
__global__ void __launch_bounds__(kRowsWarps * 32, FQ_ROWS_MINB) fq_rows_kernel_k1(const __grid_constant__ RowsArgs a) {
  fq_rows_body<KMODE, CODES>(a);
}
template <int KMODE, bool CODES>
__device__ __forceinline__ void fq_rows_body(const RowsArgs& a) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * kRowsWarps + (threadIdx.x >> 5);
  if (w >= a.nwork) return;
  int ti = 0;
  while (ti + 1 < a.count && w >= a.t[ti + 1].work0) ++ti;
  const RowsTensor& T = a.t[ti];
  const int64_t lw = w - T.work0;
  const int64_t row = lw / T.chunks_per_row;
  const int64_t beg = (lw - row * T.chunks_per_row) * kRowsWarpChunk;
  const int n = (int)((T.inner - beg < kRowsWarpChunk) ? T.inner - beg : kRowsWarpChunk);  // elements of this item
  const float* xr = T.x + row * T.inner + beg;
  float* yr = T.y + row * T.inner + beg;
  int32_t* cr = CODES ? T.codes + row * T.inner + beg : nullptr;
  ElemCtx<KMODE> ctx;
  if (T.vec_ok) {
    constexpr int kU = kRowsWarpChunk / 128;  // vectors per lane
    Pack<4> in[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if ((u * 32 + lane) * 4 < n) in[u].load(xr + (u * 32 + lane) * 4);
    load_ctx_direct<KMODE>(ctx, T.table + row * a.stride, a.K);
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int i = (u * 32 + lane) * 4;
      if (i >= n) break;
      Pack<4> out;
      IPack<4> cd;
      quant_vec<KMODE, CODES, 4>(in[u].v, ctx, out.v, cd.v);
      out.store(yr + i);
      if (CODES) cd.store(cr + i);
    }
  } else {
    load_ctx_direct<KMODE>(ctx, T.table + row * a.stride, a.K);
    constexpr int kU = 8;  // scalar loads in flight per lane
    for (int i0 = 0; i0 < n; i0 += 32 * kU) {
      float v[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int i = i0 + u * 32 + lane;
        v[u] = i < n ? __ldcs(xr + i) : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int i = i0 + u * 32 + lane;
        if (i >= n) break;
        int32_t cd;
        yr[i] = quant_elem<KMODE, CODES>(v[u], ctx, &cd);
        if (CODES) cr[i] = cd;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// STE backward of the fake-quantiser.  With round_ste and the detached exponent code
// (fp8_quantizer.py:128-132), y = round(xc / s) * s gives   dy/dxc = 1,   dy/ds = q - xc/s,   and s depends on
// maxval and M only through the bias (grad_x = ((g * s) / s) * clamp weight, rounded as ATen does): ds/dmaxval = s / maxval, ds/dM = s * ln2 * (-1 - dbias/dM).  The clamp
// (:112-113, torch.max / torch.min) passes the gradient to x inside the range, to +-maxval outside, and splits it
// evenly on exact ties.  The kernel writes grad_x and accumulates per channel
//   acc[2c]   = sum g * (d xc / d maxval)            (clipping term)
//   acc[2c+1] = sum g * (q - xc/s) * s               (scale term; grad_maxval += acc / maxval, grad_M = coef * sum)
// ------------------------------------------------------------------------------------------------
struct BwdArgs {
  const float* g;
  const float* x;
  float* gx;
  const float* table;
  double* acc;
  int64_t C, inner, chunks_per_row, chunk;
  int K, stride, sign_bits;
};

// N elements of one channel.  grad_x reproduces autograd's roundings: mul backward g * s, div backward (g * s) / s.
// When every scale of the channel is an exact power of two (FLAG_POW2, the common case) the division is a
// multiplication by the exactly representable 2^-k, which rounds identically; otherwise IEEE division.
// One range test per vector: lanes strictly inside (lo, hi) have clamp weight 1 and no clipping term.
template <int KMODE, int N>
__device__ __forceinline__ void bwd_vec(const float (&g)[N], const float (&x)[N], const ElemCtx<KMODE>& ctx,
                                        int sign_bits, bool pow2, float (&gx)[N], float& a1, float& a2) {
  float y[N], s[N];
  int32_t cd[N];
  quant_vec<KMODE, false, N>(x, ctx, y, cd, s);
  float vmin = x[0], vmax = x[0];
#pragma unroll
  for (int k = 1; k < N; ++k) { vmin = min_nan(vmin, x[k]); vmax = max_nan(vmax, x[k]); }
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const float t = mul_rn(g[k], s[k]);
    gx[k] = pow2 ? mul_rn(t, u2f(0x7f000000u - f2u(s[k]))) : div_rn(t, s[k]);
  }
  if (vmin > ctx.lo && vmax < ctx.hi) {   // false when any lane is NaN
#pragma unroll
    for (int k = 0; k < N; ++k) a2 += g[k] * (y[k] - x[k]);   // = g * (q - xc / s) * s
    return;
  }
#pragma unroll
  for (int k = 0; k < N; ++k) {
    // torch.max(x, minval): gradient to x where x > minval, half on a tie, else to minval (d minval/d maxval = -1)
    float wx = x[k] < ctx.lo ? 0.0f : (x[k] == ctx.lo ? 0.5f : 1.0f);
    float clip = sign_bits ? wx - 1.0f : 0.0f;
    const float tt = max_nan(x[k], ctx.lo);
    const float wt = tt > ctx.hi ? 0.0f : (tt == ctx.hi ? 0.5f : 1.0f);  // torch.min(t, maxval)
    clip = clip * wt + (1.0f - wt);
    wx *= wt;
    const float xc = min_nan(tt, ctx.hi);
    const bool nan_x = !(x[k] == x[k]);   // the reference's s is NaN there: every gradient it touches becomes NaN
    gx[k] = nan_x ? x[k] : mul_rn(gx[k], wx);
    a1 += nan_x ? x[k] : g[k] * clip;
    a2 += g[k] * (y[k] - xc);             // NaN for a NaN input
  }
}

#ifndef BWD_MINB
#define BWD_MINB 4
#endif
template <int KMODE, bool VEC>
__global__ void __launch_bounds__(256, BWD_MINB) fq_backward_kernel(const BwdArgs a) {
  __shared__ float s_a1[8], s_a2[8];
  const int64_t nwork = a.C * a.chunks_per_row;
  for (int64_t w = blockIdx.x; w < nwork; w += gridDim.x) {
    const int64_t row = w / a.chunks_per_row;
    const int64_t ck = w - row * a.chunks_per_row;
    const int64_t beg = ck * a.chunk;
    const int64_t end = (beg + a.chunk < a.inner) ? beg + a.chunk : a.inner;
    ElemCtx<KMODE> ctx;
    load_ctx_direct<KMODE>(ctx, a.table + row * a.stride, a.K);
    const bool pow2 = (f2u(__ldg(a.table + row * a.stride + H_FLAGS)) & FLAG_POW2) != 0;
    const float* xr = a.x + row * a.inner;
    const float* gr = a.g + row * a.inner;
    float* gxr = a.gx + row * a.inner;
    float a1 = 0.0f, a2 = 0.0f;
    if (VEC) {   // inner % 4 == 0 and 16-byte aligned bases (checked by the host), chunk % 4 == 0
      const int64_t step = 4 * (int64_t)blockDim.x;
      int64_t i = beg + 4 * (int64_t)threadIdx.x;
      for (; i + step < end; i += 2 * step) {
        Pack<4> g0, x0, g1, x1, o0, o1;
        g0.load(gr + i); x0.load(xr + i);
        g1.load(gr + i + step); x1.load(xr + i + step);
        bwd_vec<KMODE, 4>(g0.v, x0.v, ctx, a.sign_bits, pow2, o0.v, a1, a2);
        bwd_vec<KMODE, 4>(g1.v, x1.v, ctx, a.sign_bits, pow2, o1.v, a1, a2);
        o0.store(gxr + i);
        o1.store(gxr + i + step);
      }
      if (i < end) {
        Pack<4> g0, x0, o0;
        g0.load(gr + i); x0.load(xr + i);
        bwd_vec<KMODE, 4>(g0.v, x0.v, ctx, a.sign_bits, pow2, o0.v, a1, a2);
        o0.store(gxr + i);
      }
    } else {
      for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
        float gi[1] = {gr[i]}, xi[1] = {xr[i]}, oi[1];
        bwd_vec<KMODE, 1>(gi, xi, ctx, a.sign_bits, pow2, oi, a1, a2);
        gxr[i] = oi[0];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { s_a1[threadIdx.x >> 5] = a1; s_a2[threadIdx.x >> 5] = a2; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double t1 = 0.0, t2 = 0.0;
      for (int k = 0; k < (int)((blockDim.x + 31) >> 5); ++k) { t1 += s_a1[k]; t2 += s_a2[k]; }
      atomicAdd(a.acc + 2 * row, t1);
      atomicAdd(a.acc + 2 * row + 1, t2);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K2a: min/max
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_minmax(float& mn, float& mx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = min_nan(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = max_nan(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
}
// result valid in thread 0 (all threads must call)
__device__ __forceinline__ void block_minmax(float& mn, float& mx) {
  __shared__ float s_mn[32], s_mx[32];
  warp_minmax(mn, mx);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) { s_mn[wid] = mn; s_mx[wid] = mx; }
  __syncthreads();
  if (wid == 0) {
    mn = lane < nw ? s_mn[lane] : __int_as_float(0x7f800000);
    mx = lane < nw ? s_mx[lane] : __int_as_float(0xff800000);
    warp_minmax(mn, mx);
  }
}

struct EstArgs {
  float* cur_min;
  float* cur_max;
  int est_mode;
  int initialized;
  float w_new;   // (float)(1 - momentum), the Python double expression rounded to fp32
  float w_old;   // (float)momentum
  // optional fused set_quant_range + prologue
  float* maxval_out;
  float* table;  // nullptr -> no fused prologue
  int M, E, K, sign_bits;
  int64_t C;     // channels (FP8FQ_EST_DP_STATS: cur_max holds 2C entries, the second half the NaN flags)
  // data-parallel exchange over peer memory (world > 1): every rank's exchange buffer, this rank, the call's epoch
  unsigned long long* const* peers;
  int rank, world;
  unsigned int epoch;
};

// ---- data-parallel statistics exchange over NVLink peer memory ------------------------------------------------------
// The calibration step of a batch-sharded run is "local statistics -> MAX over the ranks -> estimator rule ->
// set_quant_range -> table" (dependency order of quantization_manager.py:114-122).  With NCCL that is
// statistics kernel + all-reduce + finishing kernel, ~33 us per site of which the reduction of 12 bytes is almost all
// latency.  Here the LAST CTA of the statistics kernel does the exchange itself: it stores its three values
// (-min, max, NaN flag) into every peer's exchange buffer with 64-bit system-scope stores -- value in the low word, the
// call's epoch in the high word, so data and "ready" flag arrive in one single-copy-atomic access, as in NCCL's LL
// protocol -- polls its own buffer until every peer's words carry this epoch, takes the MAX and goes on to the estimator
// rule and the table build: ONE launch per site, no collective call.  Buffers: kXchgSlots x world x 3 words per rank,
// slot = epoch % kXchgSlots (a rank can run at most one call ahead of the slowest peer, so 2 slots would do), allocated
// as symmetric memory by the host (torch.distributed._symmetric_memory) and zeroed once; epochs start at 1.
constexpr int kXchgSlots = 4;
constexpr int kXchgMaxWorld = 16;
constexpr long long kXchgTimeoutClocks = 6000000000ll;   // ~3 s at 1.9 GHz: a peer that never arrives poisons the range with NaN

__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
#ifndef FP8FQ_HOST_SIM
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#else
  *reinterpret_cast<volatile unsigned long long*>(p) = v;
#endif
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
#ifndef FP8FQ_HOST_SIM
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
#else
  return *reinterpret_cast<const volatile unsigned long long*>(p);
#endif
}

// All threads of the (last) CTA call it; thread 0 brings this rank's (mn, mx) and leaves with the global ones.
__device__ __forceinline__ void peer_exchange_minmax(const EstArgs& e, float& mn, float& mx) {
  __shared__ float s_loc[3];
  __shared__ float s_all[3 * kXchgMaxWorld];
  const int tid = threadIdx.x;
  if (tid == 0) {
    const bool nan = !(mn == mn) || !(mx == mx);
    const float ninf = __int_as_float(0xff800000);
    s_loc[0] = nan ? ninf : -mn;
    s_loc[1] = nan ? ninf : mx;
    s_loc[2] = nan ? 1.0f : 0.0f;
  }
  __syncthreads();
  const int nw = e.world * 3;
  if (tid < nw) {
    const int peer = tid / 3, j = tid - peer * 3;
    const size_t slot = (size_t)(e.epoch % kXchgSlots) * (size_t)e.world * 3;
    st_sys_u64(e.peers[peer] + slot + (size_t)e.rank * 3 + j,
               ((unsigned long long)e.epoch << 32) | (unsigned long long)f2u(s_loc[j]));
    const unsigned long long* src = e.peers[e.rank] + slot + (size_t)peer * 3 + j;
    unsigned long long got = ld_sys_u64(src);
#ifndef FP8FQ_HOST_SIM
    const long long t0 = clock64();
    while ((unsigned int)(got >> 32) != e.epoch && clock64() - t0 < kXchgTimeoutClocks) got = ld_sys_u64(src);
#endif
    s_all[tid] = (unsigned int)(got >> 32) == e.epoch ? u2f((uint32_t)got) : __int_as_float(0x7fc00000);
  }
  __syncthreads();
  if (tid == 0) {
    float a = s_all[0], b = s_all[1], f = s_all[2];
    for (int p = 1; p < e.world; ++p) {
      a = max_nan(a, s_all[3 * p]);
      b = max_nan(b, s_all[3 * p + 1]);
      f = max_nan(f, s_all[3 * p + 2]);
    }
    if (!(f <= 0.0f)) a = b = __int_as_float(0x7fc00000);   // a NaN somewhere (or a peer that never arrived)
    mn = -a;
    mx = b;
  }
}

// estimator update rule for one channel; returns the updated (min, max)
__device__ __forceinline__ void est_update(const EstArgs& e, int64_t c, float& mn, float& mx) {
  if (e.est_mode == FP8FQ_EST_DP_STATS) {
    // batch statistics for a data-parallel exchange: [-min | max | NaN flag], no state rule.  A MAX all-reduce does not
    // promise to propagate NaN (torch.min / torch.max do), so a NaN statistic travels as a flag and the value slots
    // carry the neutral element of MAX; fp8fq_dp_finish_prepare_f32 turns the flag back into NaN.
    const bool nan = !(mn == mn) || !(mx == mx);
    const float ninf = __int_as_float(0xff800000);
    e.cur_min[c] = nan ? ninf : -mn;
    e.cur_max[c] = nan ? ninf : mx;
    e.cur_max[e.C + c] = nan ? 1.0f : 0.0f;
    return;
  }
  if (e.initialized && e.est_mode == FP8FQ_EST_ALL) {           // range_estimators.py:96-98
    mn = min_nan(e.cur_min[c], mn);
    mx = max_nan(e.cur_max[c], mx);
  } else if (e.initialized && e.est_mode == FP8FQ_EST_RUNNING) { // :121-123
    mn = add_rn(mul_rn(e.w_new, mn), mul_rn(e.w_old, e.cur_min[c]));
    mx = add_rn(mul_rn(e.w_new, mx), mul_rn(e.w_old, e.cur_max[c]));
  }
  e.cur_min[c] = mn;
  e.cur_max[c] = mx;
}

constexpr int kMMThreads = 256;
constexpr int kMMUnroll = 4;

// Second stage of the per-tensor reduction (all threads of every CTA call it with their running min / max): block
// reduce, publish the CTA's partial, and the LAST CTA to arrive reduces the partials, applies the estimator's update
// rule, set_quant_range and -- optionally -- builds the quantiser table.  Leaves the ticket counter at zero.
__device__ __forceinline__ void minmax_tensor_finish(float mn, float mx, float* __restrict__ partial,
                                                     unsigned int* __restrict__ counter, const EstArgs& est) {
  block_minmax(mn, mx);
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = mn;
    partial[2 * blockIdx.x + 1] = mx;
    __threadfence();
    const unsigned int ticket = atomicAdd(counter, 1u);
    s_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  mn = __int_as_float(0x7f800000);
  mx = __int_as_float(0xff800000);
  for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
    mn = min_nan(mn, __ldcg(partial + 2 * b));
    mx = max_nan(mx, __ldcg(partial + 2 * b + 1));
  }
  block_minmax(mn, mx);
  __shared__ float s_mv;
  if (est.world > 1) peer_exchange_minmax(est, mn, mx);   // data-parallel: MAX over the ranks, inside this launch
  if (threadIdx.x == 0) {
    *counter = 0u;  // leave the workspace ready for the next call
    est_update(est, 0, mn, mx);
    const float mv = range_to_maxval(mn, mx);
    if (est.maxval_out != nullptr) est.maxval_out[0] = mv;
    s_mv = mv;
  }
  __syncthreads();
  if (est.table != nullptr) prepare_channel(s_mv, est.M, est.E, est.K, est.sign_bits, est.table);
}


__global__ void __launch_bounds__(kMMThreads) minmax_tensor_kernel(const float* __restrict__ x, int64_t n,
                                                                   int vec_ok, float* __restrict__ partial,
                                                                   unsigned int* __restrict__ counter,
                                                                   const EstArgs est) {
  float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
  if (vec_ok) {
    const int64_t nv = n >> 2;
    const float4* xv = reinterpret_cast<const float4*>(x);
    const int64_t step = (int64_t)gridDim.x * kMMThreads;
    int64_t i = (int64_t)blockIdx.x * kMMThreads + threadIdx.x;
    for (; i + (kMMUnroll - 1) * step < nv; i += kMMUnroll * step) {
      float4 t[kMMUnroll];
#pragma unroll
      for (int u = 0; u < kMMUnroll; ++u) t[u] = __ldcs(xv + i + u * step);
#pragma unroll
      for (int u = 0; u < kMMUnroll; ++u) {
        mn = min_nan(min_nan(mn, t[u].x), min_nan(t[u].y, min_nan(t[u].z, t[u].w)));
        mx = max_nan(max_nan(mx, t[u].x), max_nan(t[u].y, max_nan(t[u].z, t[u].w)));
      }
    }
    for (; i < nv; i += step) {
      const float4 t = __ldcs(xv + i);
      mn = min_nan(min_nan(mn, t.x), min_nan(t.y, min_nan(t.z, t.w)));
      mx = max_nan(max_nan(mx, t.x), max_nan(t.y, max_nan(t.z, t.w)));
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {
      const float t = x[(nv << 2) + threadIdx.x];
      mn = min_nan(mn, t);
      mx = max_nan(mx, t);
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * kMMThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kMMThreads) {
      const float t = x[i];
      mn = min_nan(mn, t);
      mx = max_nan(mx, t);
    }
  }
  minmax_tensor_finish(mn, mx, partial, counter, est);
}

// Calibration epilogue of a BN-fused layer (quantized_folded_bn.py:39-55 in state estimate_ranges): min / max of
// act(bn(x)) WITHOUT materialising it -- the reference (and the op-by-op path) writes F.batch_norm's output, reads
// and writes it again for the activation and reads it a third time for the estimator (20 B/element before the
// quantiser even starts); this reads the convolution output once (4 B/element).  Same tile / channel arithmetic as
// fq_stream_kernel's PRE_AFFINE (LAYOUT 0: NCHW rows, H*W % 4 == 0) and PRE_AFFINE_CL (LAYOUT 1) variants, same
// batch-norm arithmetic (bn_mode), same NaN-propagating activation; the finish is minmax_tensor_kernel's.
template <int LAYOUT, int BNM>
__global__ void __launch_bounds__(kMMThreads) minmax_bn_act_kernel(const StreamArgs a, float* __restrict__ partial,
                                                                   unsigned int* __restrict__ counter,
                                                                   const EstArgs est) {
  constexpr int kTile = kMMThreads * 4 * kMMUnroll;
  const int64_t ntiles = (a.n + kTile - 1) / kTile;   // a.n % 4 == 0 (checked by the launcher)
  float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t tile0 = tile * kTile;
    const int64_t base = tile0 + (int64_t)threadIdx.x * 4;
    Pack<4> in[kMMUnroll];
    bool ok[kMMUnroll];
#pragma unroll
    for (int u = 0; u < kMMUnroll; ++u) {
      const int64_t i = base + (int64_t)u * kMMThreads * 4;
      ok[u] = i < a.n;
      if (ok[u]) in[u].load(a.x + i);
    }
    if (LAYOUT == 1) {
      if (a.cl_same) {
        const uint32_t i32 = (uint32_t)base;
        const uint32_t ch = i32 - fdiv(i32, a.c_div) * a.Cbn;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const BnParams bp = bn_load<BNM>(a, ch + k);
#pragma unroll
          for (int u = 0; u < kMMUnroll; ++u)
            if (ok[u]) in[u].v[k] = bn_apply<BNM>(in[u].v[k], bp);
        }
      } else {
#pragma unroll
        for (int u = 0; u < kMMUnroll; ++u) {
          if (!ok[u]) continue;
          const uint32_t i32 = (uint32_t)(base + (int64_t)u * kMMThreads * 4);
          const uint32_t ch = i32 - fdiv(i32, a.c_div) * a.Cbn;
#pragma unroll
          for (int k = 0; k < 4; ++k) in[u].v[k] = bn_apply<BNM>(in[u].v[k], bn_load<BNM>(a, ch + k));
        }
      }
    } else {
      const uint32_t row0 = fdiv((uint32_t)tile0, a.hw_div);
      const uint32_t col0 = (uint32_t)tile0 - row0 * a.hw;
      const uint32_t ch0 = row0 - fdiv(row0, a.c_div) * a.Cbn;
#pragma unroll
      for (int u = 0; u < kMMUnroll; ++u) {
        if (!ok[u]) continue;
        const uint32_t p = col0 + (uint32_t)(threadIdx.x * 4 + u * kMMThreads * 4);
        uint32_t ch = ch0 + __umulhi(p, a.hw_rcp);
        ch = ch >= a.Cbn ? ch - a.Cbn : ch;
        const BnParams bp = bn_load<BNM>(a, ch);
#pragma unroll
        for (int k = 0; k < 4; ++k) in[u].v[k] = bn_apply<BNM>(in[u].v[k], bp);
      }
    }
#pragma unroll
    for (int u = 0; u < kMMUnroll; ++u) {
      if (!ok[u]) continue;
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = apply_act(in[u].v[k], a.act);
      mn = min_nan(min_nan(mn, v[0]), min_nan(v[1], min_nan(v[2], v[3])));
      mx = max_nan(max_nan(mx, v[0]), max_nan(v[1], max_nan(v[2], v[3])));
    }
  }
  minmax_tensor_finish(mn, mx, partial, counter, est);
}

// per-channel: one CTA per row
__global__ void __launch_bounds__(128) minmax_rows_kernel(const float* __restrict__ x, int64_t C, int64_t inner,
                                                          int vec_ok, const EstArgs est) {
  __shared__ float s_mv;
  const int stride = est.table != nullptr ? table_stride(est.K) : 0;
  for (int64_t row = blockIdx.x; row < C; row += gridDim.x) {
    const float* xr = x + row * inner;
    float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
    if (vec_ok) {
      for (int64_t i = (int64_t)threadIdx.x * 4; i < inner; i += (int64_t)blockDim.x * 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(xr + i));
        mn = min_nan(min_nan(mn, t.x), min_nan(t.y, min_nan(t.z, t.w)));
        mx = max_nan(max_nan(mx, t.x), max_nan(t.y, max_nan(t.z, t.w)));
      }
    } else {
      for (int64_t i = threadIdx.x; i < inner; i += blockDim.x) {
        const float t = __ldg(xr + i);
        mn = min_nan(mn, t);
        mx = max_nan(mx, t);
      }
    }
    block_minmax(mn, mx);
    if (threadIdx.x == 0) {
      est_update(est, row, mn, mx);
      const float mv = range_to_maxval(mn, mx);
      if (est.maxval_out != nullptr) est.maxval_out[row] = mv;
      s_mv = mv;
    }
    __syncthreads();
    if (est.table != nullptr) prepare_channel(s_mv, est.M, est.E, est.K, est.sign_bits, est.table + row * stride);
    __syncthreads();
  }
}

// Data-parallel calibration, second half: `packed` = [-min (C) | max (C) | NaN flag (C)] of the GLOBAL batch (after the MAX all-reduce
// of every rank's fp8fq_minmax / bn_act_estimate statistics in mode FP8FQ_EST_DP_STATS).  One CTA per channel: the
// estimator's update rule on (cur_min, cur_max), set_quant_range and the quantiser table -- what the single-GPU fused
// calibration launch does after its own reduction.
__global__ void dp_finish_kernel(const float* __restrict__ packed, int64_t C, const EstArgs est) {
  __shared__ float s_mv;
  const int stride = est.table != nullptr ? table_stride(est.K) : 0;
  for (int64_t c = blockIdx.x; c < C; c += gridDim.x) {
    if (threadIdx.x == 0) {
      float mn = -packed[c], mx = packed[C + c];
      if (packed[2 * C + c] > 0.0f) mn = mx = __int_as_float(0x7fc00000);   // some rank saw a NaN
      est_update(est, c, mn, mx);
      const float mv = range_to_maxval(mn, mx);
      if (est.maxval_out != nullptr) est.maxval_out[c] = mv;
      s_mv = mv;
    }
    __syncthreads();
    if (est.table != nullptr) prepare_channel(s_mv, est.M, est.E, est.K, est.sign_bits, est.table + c * stride);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// BN fold
// ------------------------------------------------------------------------------------------------
__global__ void bn_fold_kernel(const float* mean, const float* var, const float* gamma, const float* beta, float eps,
                               int64_t C, float* scale, float* shift) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = div_rn(1.0f, sqrtf(add_rn(var[c], eps)));
  const float g = gamma != nullptr ? gamma[c] : 1.0f;
  const float b = beta != nullptr ? beta[c] : 0.0f;
  const float sc = mul_rn(g, invstd);
  scale[c] = sc;
  shift[c] = sub_rn(b, mul_rn(mean[c], sc));
}

// packed parameters of the exact mode: [C][4] = {mean, gamma, rsqrtf(var + eps), beta} per channel
__global__ void bn_pack_kernel(const float* mean, const float* var, const float* gamma, const float* beta, float eps,
                               int64_t C, float* packed) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  reinterpret_cast<float4*>(packed)[c] = make_float4(mean[c], gamma != nullptr ? gamma[c] : 1.0f,
                                                     rsqrtf(add_rn(var[c], eps)), beta != nullptr ? beta[c] : 0.0f);
}

// ------------------------------------------------------------------------------------------------
// 2x2 space-to-depth of an NCHW image into channel-innermost memory, with the convolution's zero padding applied:
//   y[n, Y, X, c*4 + p*2 + q] = xpad[n, c, 2Y + p, 2X + q],   xpad[r, t] = x[r - pad, t - pad] or 0,
// channels >= 4C zero.  With it a stride-2 k x k stem convolution (k odd, padding k/2) over C <= 4 input channels
// becomes a stride-1 (k+1)/2 x (k+1)/2 convolution over 16 channels -- the same sum re-indexed -- which is a shape
// cuDNN's tensor-core kernels handle (ResNet-18 stem at batch 128: 720 -> 294 us, tools/s2d_probe.py).
// One thread per output pixel: 4C scalar loads (coalesced across the warp: consecutive X read consecutive column
// pairs), four 128-bit stores (64 contiguous bytes per thread).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) s2d_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t npix,
                                                       int C, int H, int W, int pad, int Hs, int Ws) {
  const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const int X = (int)(pix % Ws);
  const int64_t t = pix / Ws;
  const int Y = (int)(t % Hs);
  const int64_t n = t / Hs;
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = 0.0f;
  const float* xn = x + n * (int64_t)C * H * W;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (c >= C) break;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int r = 2 * Y + p - pad;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int col = 2 * X + q - pad;
        if (r >= 0 && r < H && col >= 0 && col < W) v[c * 4 + p * 2 + q] = __ldg(xn + ((int64_t)c * H + r) * W + col);
      }
    }
  }
  float4* o = reinterpret_cast<float4*>(y + pix * 16);
#pragma unroll
  for (int j = 0; j < 4; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}

// ------------------------------------------------------------------------------------------------
// Max-pool over channel-innermost memory ([N, H, W, C], C % 4 == 0): the op that consumes the stem's quantised
// activation (models/resnet_quantized.py:73-78 keeps torchvision's nn.MaxPool2d).  ATen's max_pool_forward_nhwc takes
// 440 us for [128,64,112,112] -> [128,64,56,56] on B200 (5.5x its HBM bound; profiles/forward_r01_summary.json); this
// is one thread per (output pixel, 4 channels): kh*kw independent 128-bit loads (window overlap served by L1/L2), one
// 128-bit store.  Same selection rule as ATen (`val > max || isnan(val)`, row-major window order), so NaN
// propagation and the sign of a zero result are ATen's.
// ------------------------------------------------------------------------------------------------
struct PoolArgs {
  const float* x;
  float* y;
  int64_t nvec;   // N * Ho * Wo * (C / 4)
  int H, W, C4, Ho, Wo, kh, kw, sh, sw, ph, pw;
};

__global__ void __launch_bounds__(256) maxpool_nhwc_kernel(const PoolArgs a) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.nvec) return;
  const int c4 = (int)(t % a.C4);
  int64_t r = t / a.C4;
  const int ow = (int)(r % a.Wo);
  r /= a.Wo;
  const int oh = (int)(r % a.Ho);
  const int64_t n = r / a.Ho;
  const int h0 = oh * a.sh - a.ph, w0 = ow * a.sw - a.pw;
  const float4* xn = reinterpret_cast<const float4*>(a.x) + n * (int64_t)a.H * a.W * a.C4 + c4;
  const float ninf = __int_as_float(0xff800000);
  float4 m = make_float4(ninf, ninf, ninf, ninf);
  for (int i = 0; i < a.kh; ++i) {
    const int h = h0 + i;
    if (h < 0 || h >= a.H) continue;
#pragma unroll 3
    for (int j = 0; j < a.kw; ++j) {
      const int w = w0 + j;
      if (w < 0 || w >= a.W) continue;
      const float4 v = __ldg(xn + ((int64_t)h * a.W + w) * a.C4);
      m.x = (v.x > m.x || v.x != v.x) ? v.x : m.x;
      m.y = (v.y > m.y || v.y != v.y) ? v.y : m.y;
      m.z = (v.z > m.z || v.z != v.z) ? v.z : m.z;
      m.w = (v.w > m.w || v.w != v.w) ? v.w : m.w;
    }
  }
  reinterpret_cast<float4*>(a.y)[t] = m;
}

// ------------------------------------------------------------------------------------------------
// uint8 image -> normalised fp32 image, NCHW: the ToTensor + Normalize steps of the reference's input pipeline
// (utils/imagenet_dataloaders.py:66-81: x / 255, then (x - mean[c]) / std[c]) applied on the device, so that the host
// link carries 1 byte per pixel instead of 4.  The three arithmetic steps are tabulated per (channel, byte value) by
// the caller with the reference's own fp32 operations (256 entries per channel), which makes the result bit-identical
// to torchvision's by construction.  One thread per 16 pixels (one 128-bit load, four 128-bit stores); the table is
// staged in shared memory.  HBM-bound: 5 B / pixel.
// ------------------------------------------------------------------------------------------------
constexpr int kU8MaxC = 8;
__global__ void __launch_bounds__(256) u8_normalize_kernel(const uint8_t* __restrict__ x, const float* __restrict__ lut,
                                                           float* __restrict__ y, int64_t n, int C, int64_t hw, int vec_ok) {
  __shared__ float s_lut[kU8MaxC * 256];
  for (int i = threadIdx.x; i < C * 256; i += blockDim.x) s_lut[i] = __ldg(lut + i);
  __syncthreads();
  if (vec_ok) {   // hw % 16 == 0 and 16-byte aligned pointers: the 16 pixels of a thread share a channel
    const int64_t ngroups = n >> 4;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += (int64_t)gridDim.x * blockDim.x) {
      const int64_t i = g << 4;
      const int c = (int)((i / hw) % C);
      const float* t = s_lut + c * 256;
      const uint4 v = __ldcs(reinterpret_cast<const uint4*>(x + i));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      float4* o = reinterpret_cast<float4*>(y + i);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        o[j] = make_float4(t[w[j] & 0xffu], t[(w[j] >> 8) & 0xffu], t[(w[j] >> 16) & 0xffu], t[w[j] >> 24]);
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
      y[i] = s_lut[(int)((i / hw) % C) * 256 + x[i]];
  }
}

// ------------------------------------------------------------------------------------------------
// INT uniform quantisers: set_quant_range (uniform_quantizers.py:224-246, 303-314) + channel tables, one CTA
// ------------------------------------------------------------------------------------------------
__global__ void uq_prepare_kernel(const float* __restrict__ xmin, const float* __restrict__ xmax, int64_t C,
                                  int n_bits, int symmetric, int recip_div, float eps, float* __restrict__ delta_out,
                                  float* __restrict__ zero_float_out, float* __restrict__ signed_out,
                                  float* __restrict__ table) {
  __shared__ float s_min;
  // x_min = min(x_min, 0) over all channels decides `signed` (uniform_quantizers.py:306): x_min.min() < 0
  float mn = __int_as_float(0x7f800000), dummy = __int_as_float(0xff800000);
  for (int64_t c = threadIdx.x; c < C; c += blockDim.x) mn = min_nan(mn, min_nan(xmin[c], 0.0f));
  block_minmax(mn, dummy);
  if (threadIdx.x == 0) s_min = mn;
  __syncthreads();
  const bool is_signed = s_min < 0.0f;  // NaN -> false, like `tensor(nan) < 0`
  float int_min = 0.0f, int_max = ldexpf(1.0f, n_bits) - 1.0f;
  if (symmetric) {
    int_min = is_signed ? -ldexpf(1.0f, n_bits - 1) : 0.0f;
    int_max = ldexpf(1.0f, n_bits - (is_signed ? 1 : 0)) - 1.0f;
  }
  if (threadIdx.x == 0 && signed_out != nullptr) signed_out[0] = is_signed ? 1.0f : 0.0f;
  for (int64_t c = threadIdx.x; c < C; c += blockDim.x) {
    const float xm = min_nan(xmin[c], 0.0f);   // torch.min(x_min, zeros)
    const float xM = max_nan(xmax[c], eps);    // torch.max(x_max, ones * eps)
    // `tensor / python_scalar` (uniform_quantizers.py:239, 309): ATen divides on the CPU but multiplies by the
    // fp32 reciprocal of the scalar on CUDA -- recip_div selects which of the reference's two behaviours to match
    const float num = symmetric ? max_nan(fabsf(xm), xM) : sub_rn(xM, xm);
    const float delta = recip_div ? mul_rn(num, div_rn(1.0f, int_max)) : div_rn(num, int_max);
    const float zf = symmetric ? 0.0f : div_rn(-xm, delta);
    if (delta_out != nullptr) delta_out[c] = delta;
    if (zero_float_out != nullptr) zero_float_out[c] = zf;
    uq_build(table + c * kUStride, delta, zf, int_min, int_max, eps, n_bits, symmetric != 0);
  }
}

// ------------------------------------------------------------------------------------------------
// K2b: MSE grid (kernel below mse_finish_kernel)
// ------------------------------------------------------------------------------------------------
__global__ void mse_finish_kernel(const double* __restrict__ acc, int64_t GC, double inv_count,
                                  float* __restrict__ mses) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= GC) return;
  mses[i] = add_rn(mses[i], (float)(acc[i] * inv_count));
}

// K2b: MSE grid.  grid = (chunks, C).  Each CTA keeps its slice of one channel row in registers (EPT = 16 elements
// per thread, 4 for short rows) and sweeps the G candidate tables of the current mantissa width:
//   * candidate tables are staged in shared memory a GROUP at a time (as many as fit in ~40 KB: all 111 for the
//     8-bit formats with M >= 2), so the candidate loop runs without a barrier per candidate;
//   * per-warp partial sums go to a [warps][G] shared array (plain stores, fixed summation order), then one double
//     atomicAdd per (CTA, candidate) into the global accumulators;
//   * formats with <= 3 exponent codes (M >= 5) select scales with compares on registers (KMODE 0), the others read
//     the (s, 1/s) pair with ld.shared;
//   * x is padded with zeros, not masked: Q(0) = 0 contributes nothing (and a degenerate candidate whose table is
//     NaN poisons the sum through the real elements already, like the reference).
// Measured on B200 (tools/bench_mse.py): 1.47 T candidate evaluations/s on [64,64,56,56] x 666 candidates; the first
// version (table double-buffered per candidate behind a barrier, 8 elements per thread, shared double atomics) 0.75 T.
constexpr int kMseThreads = 256;
// ABS (signed formats): the elements are kept as |x| -- the squared error only needs |x| - |y|, and |clamp(x)| is then
// one min per candidate instead of a max and a min.
template <int KMODE, int EPT, bool ABS>
__global__ void __launch_bounds__(kMseThreads) mse_grid_kernel(const float* __restrict__ x, int64_t inner, int64_t C,
                                                                  const float* __restrict__ tables, int G, int K,
                                                                  int gpb, double* __restrict__ acc) {
#ifdef FP8FQ_HOST_SIM
  float* s_dyn = fp8fq_sim::dynamic_smem();
#else
  extern __shared__ __align__(16) float s_dyn[];
#endif
  constexpr int kWarps = kMseThreads / 32;
  const int stride = table_stride(K);
  const int strideP = (stride + 3) & ~3;
  float* s_tab = s_dyn;                       // [gpb][strideP]
  float* s_part = s_dyn + (size_t)gpb * strideP;  // [kWarps][G]
  const int64_t c = blockIdx.y;
  const int64_t beg = (int64_t)blockIdx.x * kMseThreads * EPT;
  const float* xr = x + c * inner;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float v[EPT];
#pragma unroll
  for (int u = 0; u < EPT; ++u) {
    const int64_t i = beg + (int64_t)u * kMseThreads + threadIdx.x;
    v[u] = i < inner ? __ldg(xr + i) : 0.0f;
    if (ABS) v[u] = fabsf(v[u]);
  }
  for (int g0 = 0; g0 < G; g0 += gpb) {
    const int ng = (G - g0 < gpb) ? G - g0 : gpb;
    __syncthreads();  // the previous group's tables are no longer read
    // table of candidate g for channel c lives at tables[(g * C + c) * stride]
    for (int g = warp; g < ng; g += kWarps) {
      const float* src = tables + ((int64_t)(g0 + g) * C + c) * stride;
      for (int j = lane; j < stride; j += 32) s_tab[g * strideP + j] = __ldg(src + j);
    }
    __syncthreads();
    for (int g = 0; g < ng; ++g) {
      ElemCtx<KMODE> ctx;
      load_ctx<KMODE>(ctx, s_tab + g * strideP, K, [](const float* p) { return *p; });
      // A rounding tie of an UNCLIPPED x resolved the other way moves y to the neighbouring code on the other side
      // of x: |x - y| is unchanged to ~1e-6 relative, so the tie guard of the reciprocal multiply is skipped here (3
      // of ~20 instructions per evaluation).  It is kept where it matters: tables with an unusable reciprocal, and
      // the E = 0 formats (K == 1), whose top code maxval / s = 2^M - 1/2 puts every CLIPPED element exactly on a
      // tie while x itself is elsewhere.
      const bool guarded = K == 1 || (f2u(s_tab[g * strideP + H_FLAGS]) & FLAG_RSNAN) != 0;
      float err = 0.0f;
#pragma unroll
      for (int u = 0; u < EPT; u += 4) {
        float vi[4] = {v[u], v[u + 1], v[u + 2], v[u + 3]}, yo[4];
        int32_t cd[4];
        if (guarded) quant_vec<KMODE, false, 4, true, true, false, 0, ABS>(vi, ctx, yo, cd);
        else quant_vec<KMODE, false, 4, true, false, false, 0, ABS>(vi, ctx, yo, cd);
        // y has the sign of x or is zero, so |x| - |y| is x - y up to that sign: the same square (and the scaled-domain
        // path need not restore the sign)
        float d[4];
        sub2_rn(fabsf(vi[0]), fabsf(vi[1]), fabsf(yo[0]), fabsf(yo[1]), d[0], d[1]);   // (two-wide under FP8FQ_PACK2; same
        sub2_rn(fabsf(vi[2]), fabsf(vi[3]), fabsf(yo[2]), fabsf(yo[3]), d[2], d[3]);   //  roundings and order either way)
#pragma unroll
        for (int k = 0; k < 4; ++k) err = fmaf(d[k], d[k], err);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) err += __shfl_xor_sync(0xffffffffu, err, o);
      if (lane == 0) s_part[warp * G + g0 + g] = err;
    }
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += kMseThreads) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) t += (double)s_part[w * G + g];
    atomicAdd(&acc[(int64_t)g * C + c], t);
  }
}

// ------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool aligned4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3u) == 0; }

template <int KMODE, int PRE, int VEC, bool CODES, int BNM, bool DYN = false>
int launch_stream_t(const StreamArgs& a, cudaStream_t st) {
  const int threads = DYN ? a.threads : kThreads;
  const int64_t tile = (int64_t)threads * VEC * StreamUnroll<PRE>::value;
  int64_t ntiles = (a.n + tile - 1) / tile;
  if (ntiles < 1) ntiles = 1;
  // Tiles per CTA.  Measured (tools/bench_kernels.py, [128,64,112,112]): the plain and residual-add kernels are
  // fastest with one tile per CTA (6.56 / 6.68 TB/s); the batch-norm variants, which are instruction-issue bound
  // (ncu: 78 % issue slots busy), gain from amortising their longer per-CTA prologue over 4 tiles (5.64 -> 6.16 and
  // 4.97 -> 5.84 TB/s) -- as long as that still leaves >= 2 full waves of CTAs.  FP8FQ_TILES_PER_CTA overrides.
  constexpr bool kHasBn = PreTraits<PRE>::kHasBn;
  static const int tpc_env = [] {
    const char* e = getenv("FP8FQ_TILES_PER_CTA");
    const int v = e ? atoi(e) : 0;
    return v >= 1 && v <= 64 ? v : 0;
  }();
  // (K > 3 formats: 2 tiles per CTA for the fitted-CTA instantiations, whose per-CTA prologue is the longest -- measured
  // +1.5..5 % at [128,96..576,14..56], -1..2 % for the compile-time-CTA ones; profiles/tiles_per_cta_k1_r02.json)
  int tpc = tpc_env ? tpc_env : ((kHasBn && KMODE == 0) ? 4 : ((kHasBn && KMODE == 1 && DYN) ? 2 : 1));
  while (tpc > 1 && !tpc_env && ntiles / tpc < (int64_t)sm_count() * 12) tpc >>= 1;
  int64_t grid = (ntiles + tpc - 1) / tpc;
  // FP8FQ_WAVE_GRID=1 (experiment): round the grid up to a whole number of waves (SMs x resident CTAs of this
  // instantiation), so that the last wave is as full as the others; the kernel's tile loop strides by the grid size.
  static const bool wave_env = [] {
    const char* e = getenv("FP8FQ_WAVE_GRID");
    return e && e[0] == '1';
  }();
#ifndef FP8FQ_HOST_SIM
  if (wave_env) {
    static const int resident = [] {
      int nb = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fq_stream_kernel<KMODE, PRE, VEC, CODES, BNM, DYN>, kThreads, 0) !=
          cudaSuccess || nb < 1)
        nb = StreamMinBlocks<KMODE, PRE, DYN>::value;
      return nb;
    }();
    const int64_t wave = (int64_t)sm_count() * resident;
    if (grid > wave) {
      const int64_t g2 = ((grid + wave - 1) / wave) * wave;
      if (g2 <= ntiles) grid = g2;
    }
  }
#endif
  if (grid > 0x7fffffffll) grid = 0x7fffffffll;  // the kernel strides over the remaining tiles
  StreamArgs b = a;
  const int64_t nvec = a.n - a.n % VEC;
  b.ntiles = (nvec + tile - 1) / tile;
  launch_kernel(fq_stream_kernel<KMODE, PRE, VEC, CODES, BNM, DYN>, dim3((unsigned)grid), dim3(threads), 0, st, b);
  return launch_status();
}

template <int PRE, int VEC>
int launch_stream(const StreamArgs& a, cudaStream_t st) {
  constexpr bool kHasBn = PreTraits<PRE>::kHasBn;
  const bool small = a.K <= 3 && (!PreTraits<PRE>::kTail || a.K2 <= 3);
  if (PRE == PRE_PLAIN && a.codes != nullptr)  // the code planes exist for the parity tests of the plain quantiser
    return small ? launch_stream_t<0, PRE_PLAIN, VEC, true, 0>(a, st) : launch_stream_t<1, PRE_PLAIN, VEC, true, 0>(a, st);
  if (PreTraits<PRE>::kCL && a.threads > 0 && a.threads != kThreads) {   // CTA size fitted to the channel count
    constexpr bool kDyn = PreTraits<PRE>::kCL;
    if (a.bn_mode == 1)
      return small ? launch_stream_t<0, PRE, VEC, false, 1, kDyn>(a, st) : launch_stream_t<1, PRE, VEC, false, 1, kDyn>(a, st);
    return small ? launch_stream_t<0, PRE, VEC, false, 0, kDyn>(a, st) : launch_stream_t<1, PRE, VEC, false, 0, kDyn>(a, st);
  }
  if (kHasBn && a.bn_mode == 1)
    return small ? launch_stream_t<0, PRE, VEC, false, kHasBn ? 1 : 0>(a, st)
                 : launch_stream_t<1, PRE, VEC, false, kHasBn ? 1 : 0>(a, st);
  return small ? launch_stream_t<0, PRE, VEC, false, 0>(a, st) : launch_stream_t<1, PRE, VEC, false, 0>(a, st);
}

// Fills the row-geometry part of StreamArgs; true if the tile-local-rows variant applies, false if the generic
// (flat division per vector) variant must be used.
bool setup_affine(StreamArgs& a, int64_t hw, int64_t Cbn) {
  a.hw = (uint32_t)hw;
  a.Cbn = (uint32_t)Cbn;
  a.hw_div = make_fastdiv((uint32_t)hw);
  a.c_div = make_fastdiv((uint32_t)Cbn);
  a.hw_rcp = hw > 1 ? (uint32_t)(((1ull << 32) + (uint64_t)hw - 1) / (uint64_t)hw) : 0u;
  const int64_t max_rows = 2 + 4095 / hw;  // rows a tile of <= 4096 elements can advance: one channel wrap at most
  const bool exact = (uint64_t)(hw + 4096) * (uint64_t)hw < (1ull << 32);  // umulhi(p, hw_rcp) == p / hw
  return hw > 1 && exact && max_rows <= Cbn;
}

int check_format(float mantissa_bits, int n_bits, int sign_bits, int* M, int* E, int* K) {
  if (sign_bits != 0 && sign_bits != 1) return FP8FQ_ERR_BAD_ARG;
  int r = format_split(mantissa_bits, n_bits, sign_bits, M, E, K);
  if (r == -1) return FP8FQ_ERR_BAD_ARG;
  if (r == -2) return FP8FQ_ERR_UNSUPPORTED;
  return FP8FQ_OK;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int fp8fq_version(void) { return 100; }

const char* fp8fq_build_info(void) { return "fp8fq 0.1 sm_100a " __DATE__ " " __TIME__; }

int64_t fp8fq_launch_count(void) { return g_launches.load(); }

int fp8fq_format_split(float mantissa_bits, int n_bits, int sign_bits, int* M, int* E, int* K) {
  int m, e, k;
  int r = check_format(mantissa_bits, n_bits, sign_bits, &m, &e, &k);
  if (r != FP8FQ_OK) return r;
  if (M) *M = m;
  if (E) *E = e;
  if (K) *K = k;
  return FP8FQ_OK;
}

int64_t fp8fq_table_stride(float mantissa_bits, int n_bits, int sign_bits) {
  int M, E, K;
  int r = check_format(mantissa_bits, n_bits, sign_bits, &M, &E, &K);
  if (r != FP8FQ_OK) return r;
  return table_stride(K);
}

int64_t fp8fq_table_floats(float mantissa_bits, int n_bits, int sign_bits, int64_t C) {
  int64_t s = fp8fq_table_stride(mantissa_bits, n_bits, sign_bits);
  if (s < 0) return s;
  if (C < 1) return FP8FQ_ERR_BAD_ARG;
  return s * C;
}

static int prepare_impl(const float* maxval, const float* xmin, const float* xmax, float* maxval_out, int64_t C,
                        float mantissa_bits, int n_bits, int sign_bits, float* table, void* stream) {
  int M, E, K;
  int r = check_format(mantissa_bits, n_bits, sign_bits, &M, &E, &K);
  if (r != FP8FQ_OK) return r;
  if (table == nullptr || C < 1) return FP8FQ_ERR_BAD_ARG;
  if (maxval == nullptr && (xmin == nullptr || xmax == nullptr)) return FP8FQ_ERR_BAD_ARG;
  int threads = K <= 32 ? 32 : (K <= 64 ? 64 : 128);
  int64_t grid = C < 4096 ? C : 4096;
  launch_plain(prepare_kernel, dim3((unsigned)grid), dim3(threads), 0, (cudaStream_t)stream, maxval, xmin, xmax, maxval_out, C, M, E, K,
                                                                       sign_bits, table);
  return launch_status();
}

int fp8fq_prepare_f32(const float* maxval, int64_t C, float mantissa_bits, int n_bits, int sign_bits, float* table,
                      void* stream) {
  if (maxval == nullptr) return FP8FQ_ERR_BAD_ARG;
  return prepare_impl(maxval, nullptr, nullptr, nullptr, C, mantissa_bits, n_bits, sign_bits, table, stream);
}

int fp8fq_set_range_prepare_f32(const float* xmin, const float* xmax, int64_t C, float* maxval_out,
                                float mantissa_bits, int n_bits, int sign_bits, float* table, void* stream) {
  if (xmin == nullptr || xmax == nullptr) return FP8FQ_ERR_BAD_ARG;
  return prepare_impl(nullptr, xmin, xmax, maxval_out, C, mantissa_bits, n_bits, sign_bits, table, stream);
}

static int launch_rows(const fp8fq_tensor_desc* d, int32_t* codes0, int count, int K, cudaStream_t st,
                       bool uniform = false) {
  RowsArgs a{};
  a.count = count;
  a.K = K;
  a.stride = uniform ? kUStride : table_stride(K);
  int64_t work = 0;
  for (int i = 0; i < count; ++i) {
    RowsTensor& t = a.t[i];
    t.x = d[i].x; t.y = d[i].y; t.table = d[i].table; t.C = d[i].C; t.inner = d[i].inner;
    t.codes = i == 0 ? codes0 : nullptr;
    t.vec_ok = (d[i].inner % 4 == 0) && aligned16(d[i].x) && aligned16(d[i].y) && (t.codes == nullptr || aligned16(t.codes));
    t.chunks_per_row = (d[i].inner + kRowsWarpChunk - 1) / kRowsWarpChunk;
    t.work0 = work;
    work += d[i].C * t.chunks_per_row;
  }
  a.nwork = work;
  const int64_t grid = (work + kRowsWarps - 1) / kRowsWarps;
  if (grid > 0x7fffffffll) return FP8FQ_ERR_UNSUPPORTED;
  const dim3 g((unsigned)grid), b(kRowsWarps * 32);
  const bool codes = codes0 != nullptr;
  if (uniform) {
    launch_kernel(fq_rows_kernel<2, false>, g, b, 0, st, a);
  } else if (K <= 3) {
    if (codes) launch_kernel(fq_rows_kernel<0, true>, g, b, 0, st, a);
    else launch_kernel(fq_rows_kernel<0, false>, g, b, 0, st, a);
  } else {
    if (codes) launch_kernel(fq_rows_kernel_k1<1, true>, g, b, 0, st, a);
    else launch_kernel(fq_rows_kernel_k1<1, false>, g, b, 0, st, a);
  }
  return launch_status();
}

static int fake_quant_impl(const float* x, float* y, int32_t* codes, const float* table, int64_t n, int64_t C,
                           int64_t inner, float mantissa_bits, int n_bits, int sign_bits, void* stream) {
  int M, E, K;
  int r = check_format(mantissa_bits, n_bits, sign_bits, &M, &E, &K);
  if (r != FP8FQ_OK) return r;
  if (n < 0 || C < 1 || inner < 0 || n != C * inner) return FP8FQ_ERR_BAD_ARG;
  if (n == 0) return FP8FQ_OK;
  if (x == nullptr || y == nullptr || table == nullptr) return FP8FQ_ERR_BAD_ARG;
  if (!aligned4(x) || !aligned4(y) || (codes && !aligned4(codes))) return FP8FQ_ERR_ALIGNMENT;
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 1) {
    StreamArgs a{};
    a.x = x; a.y = y; a.codes = codes; a.table = table; a.n = n; a.K = K; a.act = 0;
    const bool vec = aligned16(x) && aligned16(y) && (codes == nullptr || aligned16(codes));
    return vec ? launch_stream<PRE_PLAIN, 4>(a, st) : launch_stream<PRE_PLAIN, 1>(a, st);
  }
  fp8fq_tensor_desc d{x, y, table, C, inner};
  return launch_rows(&d, codes, 1, K, st);
}

int fp8fq_fake_quant_f32(const float* x, float* y, const float* table, int64_t n, int64_t C, int64_t inner,
                         float mantissa_bits, int n_bits, int sign_bits, void* stream) {
  return fake_quant_impl(x, y, nullptr, table, n, C, inner, mantissa_bits, n_bits, sign_bits, stream);
}

int fp8fq_fake_quant_codes_f32(const float* x, float* y, int32_t* codes, const float* table, int64_t n, int64_t C,
                               int64_t inner, float mantissa_bits, int n_bits, int sign_bits, void* stream) {
  if (codes == nullptr) return FP8FQ_ERR_BAD_ARG;
  return fake_quant_impl(x, y, codes, table, n, C, inner, mantissa_bits, n_bits, sign_bits, stream);
}

int fp8fq_fake_quant_multi_f32(const fp8fq_tensor_desc* descs_host, int count, float mantissa_bits, int n_bits,
                               int sign_bits, void* stream) {
  int M, E, K;
  int r = check_format(mantissa_bits, n_bits, sign_bits, &M, &E, &K);
  if (r != FP8FQ_OK) return r;
  if (count < 0 || (count > 0 && descs_host == nullptr)) return FP8FQ_ERR_BAD_ARG;
  fp8fq_tensor_desc group[kMaxMulti];
  int g = 0;
  for (int i = 0; i < count; ++i) {
    const fp8fq_tensor_desc& d = descs_host[i];
    if (d.C < 1 || d.inner < 0) return FP8FQ_ERR_BAD_ARG;
    if (d.C * d.inner == 0) continue;
    if (d.x == nullptr || d.y == nullptr || d.table == nullptr) return FP8FQ_ERR_BAD_ARG;
    if (!aligned4(d.x) || !aligned4(d.y)) return FP8FQ_ERR_ALIGNMENT;
    group[g++] = d;
    if (g == kMaxMulti) {
      r = launch_rows(group, nullptr, g, K, (cudaStream_t)stream);
      if (r != FP8FQ_OK) return r;
      g = 0;
    }
  }
  if (g > 0) return launch_rows(group, nullptr, g, K, (cudaStream_t)stream);
  return FP8FQ_OK;
}

int fp8fq_fake_quant_backward_f32(const float* grad_y, const float* x, float* grad_x, const float* table, int64_t n,
                                  int64_t C, int64_t inner, float mantissa_bits, int n_bits, int sign_bits,
                                  double* acc, void* stream) {
  int M, E, K;
  int r = check_format(mantissa_bits, n_bits, sign_bits, &M, &E, &K);
  if (r != FP8FQ_OK) return r;
  if (n < 0 || C < 1 || inner < 0 || n != C * inner) return FP8FQ_ERR_BAD_ARG;
  if (acc == nullptr) return FP8FQ_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t ce = cudaMemsetAsync(acc, 0, sizeof(double) * 2 * C, st);
  if (ce != cudaSuccess) return (int)ce;
  if (n == 0) return FP8FQ_OK;
  if (grad_y == nullptr || x == nullptr || grad_x == nullptr || table == nullptr) return FP8FQ_ERR_BAD_ARG;
  if (!aligned4(grad_y) || !aligned4(x) || !aligned4(grad_x) || (reinterpret_cast<uintptr_t>(acc) & 7u)) return FP8FQ_ERR_ALIGNMENT;
  BwdArgs a{};
  a.g = grad_y; a.x = x; a.gx = grad_x; a.table = table; a.acc = acc; a.C = C; a.inner = inner;
  const bool vec = (inner % 4 == 0) && aligned16(grad_y) && aligned16(x) && aligned16(grad_x);
  // 4 or 8 vectors per thread and two double atomics per chunk (tools/bench_backward.py: 8192 is the best of
  // 2048..65536 at the ResNet-18 activation shapes; FP8FQ_BWD_CHUNK overrides for experiments)
  a.chunk = inner >= 4 * 8192 ? 8192 : 4096;
  if (const char* e = getenv("FP8FQ_BWD_CHUNK")) { const long v = atol(e); if (v >= 1024 && v % 1024 == 0) a.chunk = v; }
  a.chunks_per_row = (inner + a.chunk - 1) / a.chunk;
  a.K = K; a.stride = table_stride(K); a.sign_bits = sign_bits;
  int64_t grid = C * a.chunks_per_row;
  if (grid > (int64_t)1 << 30) grid = (int64_t)1 << 30;
  if (K <= 3) {
    if (vec) launch_plain(fq_backward_kernel<0, true>, dim3((unsigned)grid), dim3(256), 0, st, a);
    else launch_plain(fq_backward_kernel<0, false>, dim3((unsigned)grid), dim3(256), 0, st, a);
  } else {
    if (vec) launch_plain(fq_backward_kernel<1, true>, dim3((unsigned)grid), dim3(256), 0, st, a);
    else launch_plain(fq_backward_kernel<1, false>, dim3((unsigned)grid), dim3(256), 0, st, a);
  }
  return launch_status();
}

int64_t fp8fq_uniform_table_floats(int64_t C) { return C < 1 ? FP8FQ_ERR_BAD_ARG : C * kUStride; }

int fp8fq_uniform_prepare_f32(const float* xmin, const float* xmax, int64_t C, int n_bits, int symmetric,
                              int aten_cuda_scalar_div, float eps, float* delta_out, float* zero_float_out,
                              float* signed_out, float* table, void* stream) {
  if (xmin == nullptr || xmax == nullptr || table == nullptr || C < 1) return FP8FQ_ERR_BAD_ARG;
  if (n_bits < 1 || n_bits > 16) return FP8FQ_ERR_UNSUPPORTED;
  launch_plain(uq_prepare_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, xmin, xmax, C, n_bits, symmetric ? 1 : 0,
                                                         aten_cuda_scalar_div ? 1 : 0, eps, delta_out, zero_float_out,
                                                         signed_out, table);
  return launch_status();
}

int fp8fq_uniform_quant_f32(const float* x, float* y, const float* table, int64_t n, int64_t C, int64_t inner,
                            void* stream) {
  if (n < 0 || C < 1 || inner < 0 || n != C * inner) return FP8FQ_ERR_BAD_ARG;
  if (n == 0) return FP8FQ_OK;
  if (x == nullptr || y == nullptr || table == nullptr) return FP8FQ_ERR_BAD_ARG;
  if (!aligned4(x) || !aligned4(y)) return FP8FQ_ERR_ALIGNMENT;
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 1) {
    StreamArgs a{};
    a.x = x; a.y = y; a.table = table; a.n = n; a.K = 1;
    const bool vec = aligned16(x) && aligned16(y);
    return vec ? launch_stream_t<2, PRE_PLAIN, 4, false, 0>(a, st) : launch_stream_t<2, PRE_PLAIN, 1, false, 0>(a, st);
  }
  fp8fq_tensor_desc d{x, y, table, C, inner};
  return launch_rows(&d, nullptr, 1, 1, st, true);
}

int fp8fq_bn_fold_f32(const float* mean, const float* var, const float* gamma, const float* beta, float eps,
                      int64_t Cbn, float* bn_scale, float* bn_shift, void* stream) {
  if (mean == nullptr || var == nullptr || bn_scale == nullptr || bn_shift == nullptr || Cbn < 1)
    return FP8FQ_ERR_BAD_ARG;
  const int threads = 128;
  launch_plain(bn_fold_kernel, dim3((unsigned)((Cbn + threads - 1) / threads)), dim3(threads), 0, (cudaStream_t)stream, 
      mean, var, gamma, beta, eps, Cbn, bn_scale, bn_shift);
  return launch_status();
}

static int bn_act_quant_impl(const float* x, const float* residual, float* y, const float* bn_scale,
                             const float* bn_shift, int64_t rows, int64_t hw, int64_t Cbn, int act, int bn_mode,
                             const float* table, float mb, int nb, int sb, const float* table2, float mb2, int nb2,
                             int sb2, void* stream) {
  int M, E, K, K2 = 0;
  int r = check_format(mb, nb, sb, &M, &E, &K);
  if (r != FP8FQ_OK) return r;
  if (table2 != nullptr) {
    r = check_format(mb2, nb2, sb2, &M, &E, &K2);
    if (r != FP8FQ_OK) return r;
  }
  if (rows < 0 || hw < 1 || Cbn < 1 || act < 0 || act > 2 || bn_mode < 0 || bn_mode > 1) return FP8FQ_ERR_BAD_ARG;
  const int64_t n = rows * hw;
  if (n == 0) return FP8FQ_OK;
  if (x == nullptr || y == nullptr || table == nullptr || bn_scale == nullptr || (bn_mode == 0 && bn_shift == nullptr))
    return FP8FQ_ERR_BAD_ARG;
  if (n >= (1ll << 32) || hw >= (1ll << 31) || Cbn >= (1ll << 31)) return FP8FQ_ERR_UNSUPPORTED;
  if (!aligned4(x) || !aligned4(y) || (residual && !aligned4(residual))) return FP8FQ_ERR_ALIGNMENT;
  StreamArgs a{};
  a.x = x; a.x2 = residual; a.y = y; a.table = table; a.table2 = table2; a.n = n; a.K = K; a.K2 = K2; a.act = act;
  if (bn_mode == 1 && !aligned16(bn_scale)) return FP8FQ_ERR_ALIGNMENT;
  a.bn_p[0] = bn_scale; a.bn_p[1] = bn_shift; a.bn_mode = bn_mode;
  const bool al = aligned16(x) && aligned16(y) && (residual == nullptr || aligned16(residual));
  cudaStream_t st = (cudaStream_t)stream;
  const bool tail = table2 != nullptr;
  if (setup_affine(a, hw, Cbn)) {
    const bool v4 = al && (hw % 4 == 0);
    if (tail) {
      if (v4) return launch_stream<PRE_BNQ_ADD, 4>(a, st);
      if (al) return launch_stream<PRE_BNQ_ADD_PL, 4>(a, st);  // 128-bit accesses, per-lane rows
      return launch_stream<PRE_BNQ_ADD, 1>(a, st);
    }
    if (v4) return launch_stream<PRE_AFFINE, 4>(a, st);
    if (al) return launch_stream<PRE_AFFINE_PL, 4>(a, st);
    return launch_stream<PRE_AFFINE, 1>(a, st);
  }
  if (tail) return FP8FQ_ERR_UNSUPPORTED;  // caller composes the two unfused kernels instead
  const bool v4 = al && (hw % 4 == 0);
  return v4 ? launch_stream<PRE_AFFINE_G, 4>(a, st) : launch_stream<PRE_AFFINE_G, 1>(a, st);
}

// channel-innermost twin of bn_act_quant_impl: x is [pixels, Cbn] (channels_last activations, Linear outputs)
static int bn_act_quant_nhwc_impl(const float* x, const float* residual, float* y, const float* bn_scale,
                                  const float* bn_shift, int64_t pixels, int64_t Cbn, int act, int bn_mode,
                                  const float* table, float mb, int nb, int sb, const float* table2, float mb2,
                                  int nb2, int sb2, void* stream) {
  int M, E, K, K2 = 0;
  int r = check_format(mb, nb, sb, &M, &E, &K);
  if (r != FP8FQ_OK) return r;
  if (table2 != nullptr) {
    r = check_format(mb2, nb2, sb2, &M, &E, &K2);
    if (r != FP8FQ_OK) return r;
  }
  if (pixels < 0 || Cbn < 1 || act < 0 || act > 2 || bn_mode < 0 || bn_mode > 1) return FP8FQ_ERR_BAD_ARG;
  // (the channel-innermost kernels index with 32 bits: fewer than 2^31 elements -- an 8 GB tensor)
  if (Cbn >= (1ll << 31) || pixels >= (1ll << 31) || pixels * Cbn >= (1ll << 31)) return FP8FQ_ERR_UNSUPPORTED;
  const int64_t n = pixels * Cbn;
  if (n == 0) return FP8FQ_OK;
  if (x == nullptr || y == nullptr || table == nullptr || bn_scale == nullptr || (bn_mode == 0 && bn_shift == nullptr))
    return FP8FQ_ERR_BAD_ARG;
  if (!aligned4(x) || !aligned4(y) || (residual && !aligned4(residual))) return FP8FQ_ERR_ALIGNMENT;
  if (bn_mode == 1 && !aligned16(bn_scale)) return FP8FQ_ERR_ALIGNMENT;
  StreamArgs a{};
  a.x = x; a.x2 = residual; a.y = y; a.table = table; a.table2 = table2; a.n = n; a.K = K; a.K2 = K2; a.act = act;
  a.bn_p[0] = bn_scale; a.bn_p[1] = bn_shift; a.bn_mode = bn_mode;
  a.Cbn = (uint32_t)Cbn;
  a.c_div = make_fastdiv((uint32_t)Cbn);
  const bool v4 = (Cbn % 4 == 0) && aligned16(x) && aligned16(y) && (residual == nullptr || aligned16(residual));
  // CTA size: the largest multiple of (channels per pass-lane group) = Cbn / VEC that fits in kThreads, so that the
  // pass stride threads * VEC is a multiple of Cbn (MobileNetV2: C = 96 -> 240 threads, 144 -> 252, 576 -> 144, ...;
  // 1280 -> 320, the one size above kThreads the DYN instantiations are compiled for);
  // wider layers (Cbn / VEC > kMaxDynThreads) keep kThreads and look the channel up per vector
  const int64_t lanes = Cbn / (v4 ? 4 : 1);
  a.threads = lanes <= kThreads ? (int)((kThreads / lanes) * lanes) : (lanes <= kMaxDynThreads ? (int)lanes : kThreads);
  // ... preferring one that is also a whole number of warps when such a size of at least 128 threads exists (C = 96 or
  // 192 -> 192 threads instead of 240: no half-empty warp; measured +3..5 % at those sites, profiles/cl_shapes_r02.json;
  // FP8FQ_CL_WARP_THREADS=0 restores the largest fitting size)
  static const bool warp_env = [] {
    const char* e = getenv("FP8FQ_CL_WARP_THREADS");
    return !(e && e[0] == '0');
  }();
  if (warp_env && lanes <= kThreads && a.threads % 32 != 0) {
    for (int64_t t = (kThreads / lanes) * lanes; t >= 128; t -= lanes)
      if (t % 32 == 0) { a.threads = (int)t; break; }
  }
  a.cl_same = ((int64_t)a.threads * (v4 ? 4 : 1)) % Cbn == 0 ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (table2 != nullptr) return v4 ? launch_stream<PRE_BNQ_ADD_CL, 4>(a, st) : launch_stream<PRE_BNQ_ADD_CL, 1>(a, st);
  return v4 ? launch_stream<PRE_AFFINE_CL, 4>(a, st) : launch_stream<PRE_AFFINE_CL, 1>(a, st);
}

int fp8fq_bn_act_quant_nhwc_f32(const float* x, float* y, const float* bn_scale, const float* bn_shift,
                                int64_t pixels, int64_t Cbn, int act, int bn_mode, const float* table,
                                float mantissa_bits, int n_bits, int sign_bits, void* stream) {
  return bn_act_quant_nhwc_impl(x, nullptr, y, bn_scale, bn_shift, pixels, Cbn, act, bn_mode, table, mantissa_bits,
                                n_bits, sign_bits, nullptr, 0.0f, 0, 0, stream);
}

int fp8fq_bn_quant_add_act_quant_nhwc_f32(const float* x, const float* residual, float* y, const float* bn_scale,
                                          const float* bn_shift, int64_t pixels, int64_t Cbn, int act, int bn_mode,
                                          const float* table_inner, float mantissa_bits_inner, int n_bits_inner,
                                          int sign_bits_inner, const float* table_outer, float mantissa_bits_outer,
                                          int n_bits_outer, int sign_bits_outer, void* stream) {
  if (residual == nullptr || table_outer == nullptr) return FP8FQ_ERR_BAD_ARG;
  return bn_act_quant_nhwc_impl(x, residual, y, bn_scale, bn_shift, pixels, Cbn, act, bn_mode, table_inner,
                                mantissa_bits_inner, n_bits_inner, sign_bits_inner, table_outer,
                                mantissa_bits_outer, n_bits_outer, sign_bits_outer, stream);
}

int fp8fq_bn_pack_f32(const float* mean, const float* var, const float* gamma, const float* beta, float eps,
                      int64_t Cbn, float* packed, void* stream) {
  if (mean == nullptr || var == nullptr || packed == nullptr || Cbn < 1) return FP8FQ_ERR_BAD_ARG;
  if (!aligned16(packed)) return FP8FQ_ERR_ALIGNMENT;
  const int threads = 128;
  launch_plain(bn_pack_kernel, dim3((unsigned)((Cbn + threads - 1) / threads)), dim3(threads), 0, (cudaStream_t)stream, 
      mean, var, gamma, beta, eps, Cbn, packed);
  return launch_status();
}

int fp8fq_bn_act_quant_f32(const float* x, float* y, const float* bn_scale, const float* bn_shift, int64_t rows,
                           int64_t hw, int64_t Cbn, int act, int bn_mode, const float* table, float mantissa_bits,
                           int n_bits, int sign_bits, void* stream) {
  return bn_act_quant_impl(x, nullptr, y, bn_scale, bn_shift, rows, hw, Cbn, act, bn_mode, table, mantissa_bits,
                           n_bits, sign_bits, nullptr, 0.0f, 0, 0, stream);
}

int fp8fq_bn_quant_add_act_quant_f32(const float* x, const float* residual, float* y, const float* bn_scale,
                                     const float* bn_shift, int64_t rows, int64_t hw, int64_t Cbn, int act,
                                     int bn_mode, const float* table_inner, float mantissa_bits_inner,
                                     int n_bits_inner, int sign_bits_inner, const float* table_outer,
                                     float mantissa_bits_outer, int n_bits_outer, int sign_bits_outer,
                                     void* stream) {
  if (residual == nullptr || table_outer == nullptr) return FP8FQ_ERR_BAD_ARG;
  return bn_act_quant_impl(x, residual, y, bn_scale, bn_shift, rows, hw, Cbn, act, bn_mode, table_inner,
                           mantissa_bits_inner, n_bits_inner, sign_bits_inner, table_outer, mantissa_bits_outer,
                           n_bits_outer, sign_bits_outer, stream);
}

int fp8fq_add_act_quant_f32(const float* a_in, const float* b_in, float* y, int64_t n, int act, const float* table,
                            float mantissa_bits, int n_bits, int sign_bits, void* stream) {
  int M, E, K;
  int r = check_format(mantissa_bits, n_bits, sign_bits, &M, &E, &K);
  if (r != FP8FQ_OK) return r;
  if (n < 0 || act < 0 || act > 2) return FP8FQ_ERR_BAD_ARG;
  if (n == 0) return FP8FQ_OK;
  if (a_in == nullptr || b_in == nullptr || y == nullptr || table == nullptr) return FP8FQ_ERR_BAD_ARG;
  if (!aligned4(a_in) || !aligned4(b_in) || !aligned4(y)) return FP8FQ_ERR_ALIGNMENT;
  StreamArgs a{};
  a.x = a_in; a.x2 = b_in; a.y = y; a.table = table; a.n = n; a.K = K; a.act = act;
  const bool vec = aligned16(a_in) && aligned16(b_in) && aligned16(y);
  cudaStream_t st = (cudaStream_t)stream;
  return vec ? launch_stream<PRE_ADD, 4>(a, st) : launch_stream<PRE_ADD, 1>(a, st);
}

int fp8fq_space_to_depth2_nhwc_f32(const float* x, float* y, int64_t N, int64_t C, int64_t H, int64_t W, int64_t pad,
                                   int64_t Hs, int64_t Ws, void* stream) {
  if (N < 0 || C < 1 || C > 4 || H < 1 || W < 1 || pad < 0 || Hs < 1 || Ws < 1) return FP8FQ_ERR_BAD_ARG;
  if (H >= (1ll << 30) || W >= (1ll << 30) || Hs >= (1ll << 30) || Ws >= (1ll << 30) || pad >= (1ll << 20))
    return FP8FQ_ERR_UNSUPPORTED;
  const int64_t npix = N * Hs * Ws;
  if (npix == 0) return FP8FQ_OK;
  if (x == nullptr || y == nullptr) return FP8FQ_ERR_BAD_ARG;
  if (!aligned4(x) || !aligned16(y)) return FP8FQ_ERR_ALIGNMENT;
  const int64_t grid = (npix + 255) / 256;
  if (grid > 0x7fffffffll) return FP8FQ_ERR_UNSUPPORTED;
  launch_plain(s2d_nhwc_kernel, dim3((unsigned)grid), dim3(256), 0, (cudaStream_t)stream, x, y, npix, (int)C, (int)H, (int)W, (int)pad,
                                                                     (int)Hs, (int)Ws);
  return launch_status();
}

int fp8fq_max_pool2d_nhwc_f32(const float* x, float* y, int64_t N, int64_t H, int64_t W, int64_t C, int kh, int kw,
                              int sh, int sw, int ph, int pw, void* stream) {
  if (N < 0 || H < 1 || W < 1 || C < 1 || kh < 1 || kw < 1 || sh < 1 || sw < 1 || ph < 0 || pw < 0) return FP8FQ_ERR_BAD_ARG;
  if (2 * ph > kh || 2 * pw > kw) return FP8FQ_ERR_BAD_ARG;            // torch: pad <= kernel / 2
  if (C % 4 != 0 || H >= (1ll << 30) || W >= (1ll << 30) || C >= (1ll << 30)) return FP8FQ_ERR_UNSUPPORTED;
  const int64_t Ho = (H + 2 * ph - kh) / sh + 1, Wo = (W + 2 * pw - kw) / sw + 1;   // floor mode, dilation 1
  if (Ho < 1 || Wo < 1) return FP8FQ_ERR_BAD_ARG;
  const int64_t nvec = N * Ho * Wo * (C / 4);
  if (nvec == 0) return FP8FQ_OK;
  if (x == nullptr || y == nullptr) return FP8FQ_ERR_BAD_ARG;
  if (!aligned16(x) || !aligned16(y)) return FP8FQ_ERR_ALIGNMENT;
  const int64_t grid = (nvec + 255) / 256;
  if (grid > 0x7fffffffll) return FP8FQ_ERR_UNSUPPORTED;
  PoolArgs a{x, y, nvec, (int)H, (int)W, (int)(C / 4), (int)Ho, (int)Wo, kh, kw, sh, sw, ph, pw};
  launch_plain(maxpool_nhwc_kernel, dim3((unsigned)grid), dim3(256), 0, (cudaStream_t)stream, a);
  return launch_status();
}

int fp8fq_u8_normalize_nchw_f32(const uint8_t* x, const float* lut, float* y, int64_t N, int64_t C, int64_t HW,
                                void* stream) {
  if (N < 0 || C < 1 || HW < 1) return FP8FQ_ERR_BAD_ARG;
  if (C > kU8MaxC) return FP8FQ_ERR_UNSUPPORTED;
  const int64_t n = N * C * HW;
  if (n == 0) return FP8FQ_OK;
  if (x == nullptr || lut == nullptr || y == nullptr) return FP8FQ_ERR_BAD_ARG;
  if (!aligned4(lut) || !aligned4(y)) return FP8FQ_ERR_ALIGNMENT;
  const int vec_ok = (HW % 16 == 0) && aligned16(x) && aligned16(y) ? 1 : 0;
  const int64_t work = vec_ok ? (n >> 4) : n;
  int64_t grid = (work + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 32;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  launch_plain(u8_normalize_kernel, dim3((unsigned)grid), dim3(256), 0, (cudaStream_t)stream, x, lut, y, n, (int)C, HW, vec_ok);
  return launch_status();
}

int64_t fp8fq_minmax_workspace_bytes(void) { return 16384; }

namespace {
struct Xchg {   // data-parallel exchange over peer memory (peer_exchange_minmax); world <= 1: none
  const void* peers = nullptr;
  int rank = 0, world = 1;
  unsigned int epoch = 0;
};
int check_xchg(const Xchg& xc, int64_t C) {
  if (xc.world <= 1) return FP8FQ_OK;
  if (xc.peers == nullptr || xc.rank < 0 || xc.rank >= xc.world || xc.epoch == 0) return FP8FQ_ERR_BAD_ARG;
  if (xc.world > kXchgMaxWorld || C != 1) return FP8FQ_ERR_UNSUPPORTED;   // per-tensor statistics only
  return FP8FQ_OK;
}
void set_xchg(EstArgs& e, const Xchg& xc) {
  e.peers = reinterpret_cast<unsigned long long* const*>(xc.peers);
  e.rank = xc.rank;
  e.world = xc.world > 1 ? xc.world : 1;
  e.epoch = xc.epoch;
}
}  // namespace

static int minmax_impl(const float* x, int64_t n, int64_t C, int64_t inner, float* cur_min, float* cur_max,
                       int est_mode, int initialized, double momentum, float* maxval_out, bool fuse,
                       float mantissa_bits, int n_bits, int sign_bits, float* table, void* workspace, void* stream,
                       const Xchg& xc = Xchg()) {
  if (n < 1 || C < 1 || inner < 1 || n != C * inner) return FP8FQ_ERR_BAD_ARG;
  if (x == nullptr || cur_min == nullptr || cur_max == nullptr) return FP8FQ_ERR_BAD_ARG;
  if (est_mode < 0 || est_mode > FP8FQ_EST_DP_STATS) return FP8FQ_ERR_BAD_ARG;
  if (!aligned4(x)) return FP8FQ_ERR_ALIGNMENT;
  EstArgs e{};
  e.cur_min = cur_min; e.cur_max = cur_max; e.est_mode = est_mode; e.initialized = initialized;
  e.w_new = (float)(1.0 - momentum);
  e.w_old = (float)momentum;
  e.maxval_out = maxval_out;
  e.table = nullptr;
  e.C = C;
  {
    const int r = check_xchg(xc, C);
    if (r != FP8FQ_OK) return r;
    if (xc.world > 1 && est_mode == FP8FQ_EST_DP_STATS) return FP8FQ_ERR_BAD_ARG;
    set_xchg(e, xc);
  }
  if (fuse) {
    int r = check_format(mantissa_bits, n_bits, sign_bits, &e.M, &e.E, &e.K);
    if (r != FP8FQ_OK) return r;
    if (table == nullptr) return FP8FQ_ERR_BAD_ARG;
    e.table = table;
    e.sign_bits = sign_bits;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 1) {
    if (workspace == nullptr) return FP8FQ_ERR_WORKSPACE;
    float* partial = reinterpret_cast<float*>(workspace) + 4;
    unsigned int* counter = reinterpret_cast<unsigned int*>(workspace);
    const int max_grid = (int)((fp8fq_minmax_workspace_bytes() - 16) / 8);
    int64_t grid = (int64_t)sm_count() * 8;
    const int64_t need = (n / 4 + kMMThreads * kMMUnroll - 1) / (kMMThreads * kMMUnroll);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    if (grid > max_grid) grid = max_grid;
    launch_plain(minmax_tensor_kernel, dim3((unsigned)grid), dim3(kMMThreads), 0, st, x, n, aligned16(x) ? 1 : 0, partial, counter, e);
    return launch_status();
  }
  int threads = inner >= 512 ? 128 : (inner >= 128 ? 64 : 32);
  if (fuse && e.K > threads) threads = e.K <= 64 ? 64 : 128;
  int64_t grid = C < (int64_t)sm_count() * 16 ? C : (int64_t)sm_count() * 16;
  launch_plain(minmax_rows_kernel, dim3((unsigned)grid), dim3(threads), 0, st, x, C, inner, (inner % 4 == 0 && aligned16(x)) ? 1 : 0, e);
  return launch_status();
}

int fp8fq_minmax_f32(const float* x, int64_t n, int64_t C, int64_t inner, float* cur_min, float* cur_max,
                     int est_mode, int initialized, double momentum, void* workspace, void* stream) {
  return minmax_impl(x, n, C, inner, cur_min, cur_max, est_mode, initialized, momentum, nullptr, false, 0.f, 0, 0,
                     nullptr, workspace, stream);
}

int fp8fq_estimate_prepare_f32(const float* x, int64_t n, int64_t C, int64_t inner, float* cur_min, float* cur_max,
                               int est_mode, int initialized, double momentum, float* maxval_out,
                               float mantissa_bits, int n_bits, int sign_bits, float* table, void* workspace,
                               void* stream) {
  return minmax_impl(x, n, C, inner, cur_min, cur_max, est_mode, initialized, momentum, maxval_out, true,
                     mantissa_bits, n_bits, sign_bits, table, workspace, stream);
}

static int bn_act_estimate_impl(const float* x, int64_t outer, int64_t hw, int64_t Cbn, int nhwc,
                                const float* bn_scale, const float* bn_shift, int bn_mode, int act,
                                float* cur_min, float* cur_max, int est_mode, int initialized, double momentum,
                                float* maxval_out, float mantissa_bits, int n_bits, int sign_bits, float* table,
                                void* workspace, void* stream, const Xchg& xc) {
  if (outer < 1 || hw < 1 || Cbn < 1 || act < 0 || act > 2 || bn_mode < 0 || bn_mode > 1 || est_mode < 0 ||
      est_mode > FP8FQ_EST_DP_STATS)
    return FP8FQ_ERR_BAD_ARG;
  if (x == nullptr || cur_min == nullptr || cur_max == nullptr || bn_scale == nullptr ||
      (bn_mode == 0 && bn_shift == nullptr))
    return FP8FQ_ERR_BAD_ARG;
  if (workspace == nullptr) return FP8FQ_ERR_WORKSPACE;
  if (nhwc) hw = 1;                               // [pixels, Cbn]
  if (Cbn >= (1ll << 31) || hw >= (1ll << 31) || outer >= (1ll << 32)) return FP8FQ_ERR_UNSUPPORTED;
  const int64_t n = nhwc ? outer * Cbn : outer * hw;   // NCHW: outer = N * Cbn rows of hw elements
  if (n >= (1ll << 32)) return FP8FQ_ERR_UNSUPPORTED;
  if (!aligned16(x) || (bn_mode == 1 && !aligned16(bn_scale))) return FP8FQ_ERR_UNSUPPORTED;  // 128-bit accesses only
  EstArgs e{};
  e.cur_min = cur_min; e.cur_max = cur_max; e.est_mode = est_mode; e.initialized = initialized;
  e.w_new = (float)(1.0 - momentum);
  e.w_old = (float)momentum;
  e.maxval_out = maxval_out;
  e.table = nullptr;
  e.C = 1;
  {
    const int r = check_xchg(xc, 1);
    if (r != FP8FQ_OK) return r;
    if (xc.world > 1 && est_mode == FP8FQ_EST_DP_STATS) return FP8FQ_ERR_BAD_ARG;
    set_xchg(e, xc);
  }
  if (table != nullptr) {
    int r = check_format(mantissa_bits, n_bits, sign_bits, &e.M, &e.E, &e.K);
    if (r != FP8FQ_OK) return r;
    e.table = table;
    e.sign_bits = sign_bits;
  }
  StreamArgs a{};
  a.x = x; a.n = n; a.act = act;
  a.bn_p[0] = bn_scale; a.bn_p[1] = bn_shift; a.bn_mode = bn_mode;
  if (nhwc) {
    if (Cbn % 4 != 0) return FP8FQ_ERR_UNSUPPORTED;
    a.Cbn = (uint32_t)Cbn;
    a.c_div = make_fastdiv((uint32_t)Cbn);
    a.cl_same = ((int64_t)kMMThreads * 4) % Cbn == 0 ? 1 : 0;
  } else {
    if (hw % 4 != 0 || !setup_affine(a, hw, Cbn)) return FP8FQ_ERR_UNSUPPORTED;
  }
  float* partial = reinterpret_cast<float*>(workspace) + 4;
  unsigned int* counter = reinterpret_cast<unsigned int*>(workspace);
  const int max_grid = (int)((fp8fq_minmax_workspace_bytes() - 16) / 8);
  int64_t grid = (int64_t)sm_count() * 8;
  const int64_t ntiles = (n + (int64_t)kMMThreads * 4 * kMMUnroll - 1) / ((int64_t)kMMThreads * 4 * kMMUnroll);
  if (grid > ntiles) grid = ntiles;
  if (grid > max_grid) grid = max_grid;
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 g((unsigned)grid), b(kMMThreads);
  if (nhwc) {
    if (bn_mode == 1) launch_plain(minmax_bn_act_kernel<1, 1>, g, b, 0, st, a, partial, counter, e);
    else launch_plain(minmax_bn_act_kernel<1, 0>, g, b, 0, st, a, partial, counter, e);
  } else {
    if (bn_mode == 1) launch_plain(minmax_bn_act_kernel<0, 1>, g, b, 0, st, a, partial, counter, e);
    else launch_plain(minmax_bn_act_kernel<0, 0>, g, b, 0, st, a, partial, counter, e);
  }
  return launch_status();
}

int fp8fq_bn_act_estimate_prepare_f32(const float* x, int64_t outer, int64_t hw, int64_t Cbn, int nhwc,
                                      const float* bn_scale, const float* bn_shift, int bn_mode, int act,
                                      float* cur_min, float* cur_max, int est_mode, int initialized, double momentum,
                                      float* maxval_out, float mantissa_bits, int n_bits, int sign_bits, float* table,
                                      void* workspace, void* stream) {
  return bn_act_estimate_impl(x, outer, hw, Cbn, nhwc, bn_scale, bn_shift, bn_mode, act, cur_min, cur_max, est_mode,
                              initialized, momentum, maxval_out, mantissa_bits, n_bits, sign_bits, table, workspace, stream,
                              Xchg());
}

int64_t fp8fq_dp_exchange_words(int world) {
  return (world < 1 || world > kXchgMaxWorld) ? FP8FQ_ERR_UNSUPPORTED : (int64_t)kXchgSlots * world * 3;
}

int fp8fq_estimate_prepare_p2p_f32(const float* x, int64_t n, float* cur_min, float* cur_max, int est_mode, int initialized,
                                   double momentum, float* maxval_out, float mantissa_bits, int n_bits, int sign_bits,
                                   float* table, void* workspace, const void* peer_bufs, int rank, int world,
                                   unsigned int epoch, void* stream) {
  Xchg xc;
  xc.peers = peer_bufs; xc.rank = rank; xc.world = world; xc.epoch = epoch;
  if (world < 2) return FP8FQ_ERR_BAD_ARG;
  return minmax_impl(x, n, 1, n, cur_min, cur_max, est_mode, initialized, momentum, maxval_out, table != nullptr,
                     mantissa_bits, n_bits, sign_bits, table, workspace, stream, xc);
}

int fp8fq_bn_act_estimate_prepare_p2p_f32(const float* x, int64_t outer, int64_t hw, int64_t Cbn, int nhwc,
                                          const float* bn_scale, const float* bn_shift, int bn_mode, int act,
                                          float* cur_min, float* cur_max, int est_mode, int initialized, double momentum,
                                          float* maxval_out, float mantissa_bits, int n_bits, int sign_bits, float* table,
                                          void* workspace, const void* peer_bufs, int rank, int world, unsigned int epoch,
                                          void* stream) {
  Xchg xc;
  xc.peers = peer_bufs; xc.rank = rank; xc.world = world; xc.epoch = epoch;
  if (world < 2) return FP8FQ_ERR_BAD_ARG;
  return bn_act_estimate_impl(x, outer, hw, Cbn, nhwc, bn_scale, bn_shift, bn_mode, act, cur_min, cur_max, est_mode,
                              initialized, momentum, maxval_out, mantissa_bits, n_bits, sign_bits, table, workspace, stream,
                              xc);
}

int fp8fq_dp_finish_prepare_f32(const float* packed, int64_t C, float* cur_min, float* cur_max, int est_mode,
                                int initialized, double momentum, float* maxval_out, float mantissa_bits, int n_bits,
                                int sign_bits, float* table, void* stream) {
  if (packed == nullptr || cur_min == nullptr || cur_max == nullptr || C < 1) return FP8FQ_ERR_BAD_ARG;
  if (est_mode < 0 || est_mode > 2) return FP8FQ_ERR_BAD_ARG;
  EstArgs e{};
  e.cur_min = cur_min; e.cur_max = cur_max; e.est_mode = est_mode; e.initialized = initialized;
  e.w_new = (float)(1.0 - momentum);
  e.w_old = (float)momentum;
  e.maxval_out = maxval_out;
  e.table = nullptr;
  e.C = C;
  set_xchg(e, Xchg());
  int threads = 32;
  if (table != nullptr) {
    int r = check_format(mantissa_bits, n_bits, sign_bits, &e.M, &e.E, &e.K);
    if (r != FP8FQ_OK) return r;
    e.table = table;
    e.sign_bits = sign_bits;
    threads = e.K <= 32 ? 32 : (e.K <= 64 ? 64 : 128);
  }
  const int64_t grid = C < 4096 ? C : 4096;
  launch_plain(dp_finish_kernel, dim3((unsigned)grid), dim3(threads), 0, (cudaStream_t)stream, packed, C, e);
  return launch_status();
}

// ---- MSE grid -----------------------------------------------------------------------------------
int64_t fp8fq_mse_table_floats(const float* mbits_host, int Mn, int n_bits, int sign_bits, int64_t G, int64_t C) {
  if (mbits_host == nullptr || Mn < 1 || G < 1 || C < 1) return FP8FQ_ERR_BAD_ARG;
  int64_t mx = 0;
  for (int m = 0; m < Mn; ++m) {
    int64_t s = fp8fq_table_stride(mbits_host[m], n_bits, sign_bits);
    if (s < 0) return s;
    if (s > mx) mx = s;
  }
  // tables of one mantissa width at a time + double accumulators [G*C]
  return mx * G * C + 2 * G * C + 4;
}

int fp8fq_mse_grid_f32(const float* x, int64_t n, int64_t C, int64_t inner, const float* grid, int64_t G,
                       const float* mbits_host, int Mn, int n_bits, int sign_bits, float* mses, float* tables,
                       void* stream) {
  if (x == nullptr || grid == nullptr || mbits_host == nullptr || mses == nullptr || tables == nullptr)
    return FP8FQ_ERR_BAD_ARG;
  if (n < 1 || C < 1 || inner < 1 || n != C * inner || G < 1 || Mn < 1 || C > 65535) return FP8FQ_ERR_BAD_ARG;
  int64_t tf = fp8fq_mse_table_floats(mbits_host, Mn, n_bits, sign_bits, G, C);
  if (tf < 0) return (int)tf;
  cudaStream_t st = (cudaStream_t)stream;
  // accumulators live at the (8-byte aligned) tail of the scratch buffer
  const int64_t acc_off = (tf - 2 * G * C - 2) & ~1ll;
  double* acc = reinterpret_cast<double*>(tables + acc_off);
  if ((reinterpret_cast<uintptr_t>(tables) & 7u) != 0) return FP8FQ_ERR_ALIGNMENT;
  for (int m = 0; m < Mn; ++m) {
    int M, E, K;
    int r = check_format(mbits_host[m], n_bits, sign_bits, &M, &E, &K);
    if (r != FP8FQ_OK) return r;
    // candidate tables: "channel" index = g * C + c, maxval = grid[g, c]
    r = prepare_impl(grid, nullptr, nullptr, nullptr, G * C, mbits_host[m], n_bits, sign_bits, tables, stream);
    if (r != FP8FQ_OK) return r;
    cudaError_t ce = cudaMemsetAsync(acc, 0, sizeof(double) * G * C, st);
    if (ce != cudaSuccess) return (int)ce;
    const int stride = table_stride(K);
    // candidates are processed in slices of <= 1024 so that the [warps][G] partial sums fit in shared memory
    const int strideP = (stride + 3) & ~3;
    const int ept = inner > 1024 ? 16 : 4;
    const int64_t chunks = (inner + (int64_t)kMseThreads * ept - 1) / ((int64_t)kMseThreads * ept);
    if (chunks > 2147483647ll) return FP8FQ_ERR_UNSUPPORTED;
    dim3 gdim((unsigned)chunks, (unsigned)C);
    for (int64_t gs = 0; gs < G; gs += 1024) {
      const int Gs = (int)((G - gs < 1024) ? G - gs : 1024);
      const size_t part_bytes = sizeof(float) * (kMseThreads / 32) * (size_t)Gs;
      int gpb = (int)((48 * 1024 - part_bytes) / (sizeof(float) * strideP));   // >= 9: strideP <= 392 floats
      if (gpb > Gs) gpb = Gs;
      const size_t smem = sizeof(float) * (size_t)gpb * strideP + part_bytes;
      const float* tb = tables + gs * C * stride;
      double* ac = acc + gs * C;
      auto go = [&](auto kern) { launch_plain(kern, gdim, dim3(kMseThreads), smem, st, x, inner, C, tb, Gs, K, gpb, ac); };
      if (sign_bits != 0) {
        if (K <= 3) { if (ept == 16) go(mse_grid_kernel<0, 16, true>); else go(mse_grid_kernel<0, 4, true>); }
        else { if (ept == 16) go(mse_grid_kernel<1, 16, true>); else go(mse_grid_kernel<1, 4, true>); }
      } else {
        if (K <= 3) { if (ept == 16) go(mse_grid_kernel<0, 16, false>); else go(mse_grid_kernel<0, 4, false>); }
        else { if (ept == 16) go(mse_grid_kernel<1, 16, false>); else go(mse_grid_kernel<1, 4, false>); }
      }
      r = launch_status();
      if (r != FP8FQ_OK) return r;
    }
    const int64_t GC = G * C;
    launch_plain(mse_finish_kernel, dim3((unsigned)((GC + 127) / 128)), dim3(128), 0, st, acc, GC, 1.0 / (double)inner, mses + m * GC);
    r = launch_status();
    if (r != FP8FQ_OK) return r;
  }
  return FP8FQ_OK;
}

// ---- host-buffer end-to-end entry point -------------------------------------------------------------
namespace {
struct HostPipe {
  static constexpr int kStreams = 3;
  static constexpr int64_t kChunk = 8ll << 20;  // elements per chunk (32 MiB)
  cudaStream_t st[kStreams] = {};
  float* pin_in[kStreams] = {};
  float* pin_out[kStreams] = {};
  float* dev[kStreams] = {};
  float* d_maxval = nullptr;
  float* d_table = nullptr;
  int64_t maxval_cap = 0, table_cap = 0;
  int device = -1;
  bool ready = false;
};
constexpr int kMaxPipeDevices = 64;
HostPipe g_pipes[kMaxPipeDevices];      // one pipeline (streams + staging buffers) per device, created on first use
std::mutex g_pipe_mu[kMaxPipeDevices];  // calls for the same device are serialised; different devices run concurrently

struct DeviceGuard {  // the entry point selects `device`; the caller's current device is restored on every exit
  int prev = -1;
  DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
}  // namespace

int fp8fq_fake_quant_host_f32(const float* x_host, float* y_host, const float* maxval_host, int64_t n, int64_t C,
                              int64_t inner, float mantissa_bits, int n_bits, int sign_bits, int device) {
  int M, E, K;
  int r = check_format(mantissa_bits, n_bits, sign_bits, &M, &E, &K);
  if (r != FP8FQ_OK) return r;
  if (x_host == nullptr || y_host == nullptr || maxval_host == nullptr) return FP8FQ_ERR_BAD_ARG;
  if (n < 0 || C < 1 || inner < 0 || n != C * inner) return FP8FQ_ERR_BAD_ARG;
  if (n == 0) return FP8FQ_OK;
  if (device < 0 || device >= kMaxPipeDevices) return FP8FQ_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(g_pipe_mu[device]);
  cudaError_t ce;
#define FQ_CK(call) do { ce = (call); if (ce != cudaSuccess) return (int)ce; } while (0)
  DeviceGuard restore_device;
  FQ_CK(cudaSetDevice(device));
  HostPipe& p = g_pipes[device];
  if (!p.ready) {
    for (int i = 0; i < HostPipe::kStreams; ++i) {
      FQ_CK(cudaStreamCreateWithFlags(&p.st[i], cudaStreamNonBlocking));
      FQ_CK(cudaMallocHost(&p.pin_in[i], HostPipe::kChunk * sizeof(float)));
      FQ_CK(cudaMallocHost(&p.pin_out[i], HostPipe::kChunk * sizeof(float)));
      FQ_CK(cudaMalloc(&p.dev[i], HostPipe::kChunk * sizeof(float)));
    }
    p.device = device;
    p.ready = true;
  }
  const int64_t tfl = (int64_t)table_stride(K) * C;
  if (p.maxval_cap < C) {
    if (p.d_maxval) cudaFree(p.d_maxval);
    FQ_CK(cudaMalloc(&p.d_maxval, C * sizeof(float)));
    p.maxval_cap = C;
  }
  if (p.table_cap < tfl) {
    if (p.d_table) cudaFree(p.d_table);
    FQ_CK(cudaMalloc(&p.d_table, tfl * sizeof(float)));
    p.table_cap = tfl;
  }
  FQ_CK(cudaMemcpyAsync(p.d_maxval, maxval_host, C * sizeof(float), cudaMemcpyHostToDevice, p.st[0]));
  r = fp8fq_prepare_f32(p.d_maxval, C, mantissa_bits, n_bits, sign_bits, p.d_table, p.st[0]);
  if (r != FP8FQ_OK) return r;
  FQ_CK(cudaStreamSynchronize(p.st[0]));
  // chunks are whole rows when per-channel, so each chunk sees a contiguous range of channel tables; rows longer than a
  // chunk are cut into pieces, each piece a per-tensor call with its row's table
  int64_t rows_per_chunk = 1, chunk_elems = HostPipe::kChunk;
  const bool long_rows = C > 1 && inner > HostPipe::kChunk;
  if (C > 1 && !long_rows) {
    rows_per_chunk = HostPipe::kChunk / inner;
    chunk_elems = rows_per_chunk * inner;
  }
  const int stride = table_stride(K);
  // Page-locked caller buffers (cudaHostAlloc / cudaHostRegister, e.g. torch's pin_memory()) are DMA'd directly;
  // pageable ones go through the internal pinned staging buffers with a host memcpy on either side.
  auto is_pinned = [](const void* p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
  };
  const bool pin_x = is_pinned(x_host), pin_y = is_pinned(y_host);
  int64_t done = 0;
  int64_t k = 0;
  int64_t pend_off[HostPipe::kStreams] = {-1, -1, -1}, pend_len[HostPipe::kStreams] = {0, 0, 0};
  while (done < n) {
    const int s = (int)(k % HostPipe::kStreams);
    int64_t len = (n - done) < chunk_elems ? (n - done) : chunk_elems;
    if (long_rows) {   // stay inside the current row
      const int64_t row_left = inner - done % inner;
      len = len < row_left ? len : row_left;
    }
    if (!pin_x || !pin_y) {  // a staging buffer of this stream is about to be reused
      FQ_CK(cudaStreamSynchronize(p.st[s]));
      if (!pin_y && pend_off[s] >= 0) memcpy(y_host + pend_off[s], p.pin_out[s], pend_len[s] * sizeof(float));
    }
    const float* src = x_host + done;
    if (!pin_x) {
      memcpy(p.pin_in[s], x_host + done, len * sizeof(float));
      src = p.pin_in[s];
    }
    FQ_CK(cudaMemcpyAsync(p.dev[s], src, len * sizeof(float), cudaMemcpyHostToDevice, p.st[s]));
    if (C == 1) {
      r = fp8fq_fake_quant_f32(p.dev[s], p.dev[s], p.d_table, len, 1, len, mantissa_bits, n_bits, sign_bits, p.st[s]);
    } else if (long_rows) {
      r = fp8fq_fake_quant_f32(p.dev[s], p.dev[s], p.d_table + (done / inner) * stride, len, 1, len, mantissa_bits, n_bits,
                               sign_bits, p.st[s]);
    } else {
      const int64_t row0 = done / inner, nrows = len / inner;
      r = fp8fq_fake_quant_f32(p.dev[s], p.dev[s], p.d_table + row0 * stride, len, nrows, inner, mantissa_bits,
                               n_bits, sign_bits, p.st[s]);
    }
    if (r != FP8FQ_OK) return r;
    FQ_CK(cudaMemcpyAsync(pin_y ? y_host + done : p.pin_out[s], p.dev[s], len * sizeof(float), cudaMemcpyDeviceToHost,
                          p.st[s]));
    pend_off[s] = done;
    pend_len[s] = len;
    done += len;
    ++k;
  }
  for (int s = 0; s < HostPipe::kStreams; ++s) {
    FQ_CK(cudaStreamSynchronize(p.st[s]));
    if (!pin_y && pend_off[s] >= 0) memcpy(y_host + pend_off[s], p.pin_out[s], pend_len[s] * sizeof(float));
  }
#undef FQ_CK
  return FP8FQ_OK;
}

}  // extern "C"
