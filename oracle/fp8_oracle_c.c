/* oracle/fp8_oracle_c.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the reference's per-element arithmetic, evaluated DIRECTLY (log2f and
 * powf per element, IEEE division), i.e. without any of the product's tables or shortcuts:
 *   quantize_to_fp8_ste_MM, quantization/quantizers/fp8_quantizer.py:105-133.
 *
 * Uses: (1) tests/test_host_emul.py proves that the product's table/bucket/tie-guard algorithm
 * (fp8_quantization_b200/csrc/fp8fq_core.h, compiled for the host with the same libm) returns the
 * same bits as this direct evaluation; (2) bench.py's cpu_baseline times it with OpenMP as a
 * second, compiled CPU baseline next to the torch-eager one.
 *
 * Note on libm: the reference's fp32 results depend on the log2/pow implementation of the backend
 * it runs on (Sleef in ATen's vectorised CPU loops, glibc in their scalar tails, libdevice on
 * CUDA); they agree to <= 1 ulp, not bit for bit.  This file uses the C library's log2f/powf.
 *
 * Build: gcc -O2 -fno-fast-math -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static float f_max_nan(float a, float b) { if (a != a) return a; if (b != b) return b; return a > b ? a : b; }
static float f_min_nan(float a, float b) { if (a != a) return a; if (b != b) return b; return a < b ? a : b; }

/* fp8_quantizer.py:105-106 */
int oracle_c_format(float mantissa_bits, int n_bits, int sign_bits, int* M, int* E) {
  float r = nearbyintf(mantissa_bits);
  float hi = (float)(n_bits - sign_bits);
  if (r < 1.0f) r = 1.0f;
  if (r > hi) r = hi;
  *M = (int)r;
  *E = n_bits - sign_bits - *M;
  return 0;
}

/* fp8_quantizer.py:110 */
float oracle_c_bias(float maxval, int M, int E) {
  volatile float t = powf(2.0f, (float)E) - log2f(maxval);
  volatile float c = 2.0f - powf(2.0f, -(float)M);
  t = t + log2f(c);
  t = t - 1.0f;
  return t;
}

/* fp8_quantizer.py:112-132 for one element; returns y, writes e (:128) and q (:132) */
static float quant_one(float x, float maxval, float bias, int M, int sign_bits, float* e_out, float* q_out) {
  float minval = sign_bits == 1 ? -maxval : 0.0f;
  float xc = f_min_nan(f_max_nan(x, minval), maxval);
  volatile float l = log2f(fabsf(xc)) + bias;
  float ls = floorf(l);
  if (ls != ls) { /* torch.clamp propagates NaN */ } else if (ls < 1.0f) ls = 1.0f;
  volatile float ex = ls - (float)M;
  ex = ex - bias;
  float scale = powf(2.0f, ex);
  volatile float t = xc / scale;
  float q = nearbyintf(t);
  volatile float y = q * scale;
  *e_out = ls;
  *q_out = q;
  return y;
}

/* x: [C, inner]; maxval: [C] (C == 1: per tensor).  e_out/q_out may be NULL. */
void oracle_c_fake_quant(const float* x, float* y, float* e_out, float* q_out, const float* maxval, int64_t C,
                         int64_t inner, float mantissa_bits, int n_bits, int sign_bits) {
  int M, E;
  oracle_c_format(mantissa_bits, n_bits, sign_bits, &M, &E);
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < C; ++c) {
    float bias = oracle_c_bias(maxval[c], M, E);
    for (int64_t i = 0; i < inner; ++i) {
      float e, q;
      int64_t k = c * inner + i;
      y[k] = quant_one(x[k], maxval[c], bias, M, sign_bits, &e, &q);
      if (e_out) e_out[k] = e;
      if (q_out) q_out[k] = q;
    }
  }
}

/* per-tensor variant parallelised over elements (bench cpu_baseline) */
void oracle_c_fake_quant_tensor(const float* x, float* y, float maxval, int64_t n, float mantissa_bits, int n_bits,
                                int sign_bits) {
  int M, E;
  oracle_c_format(mantissa_bits, n_bits, sign_bits, &M, &E);
  float bias = oracle_c_bias(maxval, M, E);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    float e, q;
    y[i] = quant_one(x[i], maxval, bias, M, sign_bits, &e, &q);
  }
}

/* range_estimators.py:73-74 / 85-91: NaN-propagating min and max of a row */
void oracle_c_minmax(const float* x, int64_t C, int64_t inner, float* mn, float* mx) {
  for (int64_t c = 0; c < C; ++c) {
    float lo = INFINITY, hi = -INFINITY;
    for (int64_t i = 0; i < inner; ++i) {
      lo = f_min_nan(lo, x[c * inner + i]);
      hi = f_max_nan(hi, x[c * inner + i]);
    }
    mn[c] = lo;
    mx[c] = hi;
  }
}

/* quantized_folded_bn.py:39-51: eval-mode F.batch_norm followed by the fused activation, per element, with the plainest
 * possible indexing (the CUDA kernels' tile / channel arithmetic is what this checks).
 *   layout 0: x is [N, C, hw] (NCHW), channel = (i / hw) % C;  layout 1: x is [pixels, C], channel = i % C.
 *   bn_mode 0: y = fma(x, p0[c], p1[c])  (p0 = gamma / sqrt(var + eps), p1 = beta - mean * p0, computed by the caller);
 *   bn_mode 1: y = fma(gamma * (x - mean), invstd, beta) with p0 = [C][4] = {mean, gamma, invstd, beta} -- the arithmetic
 *              ATen's eval batch norm performs on CUDA (measured).
 *   act: 0 none, 1 relu, 2 relu6 (torch.relu / hardtanh propagate NaN). */
void oracle_c_bn_act(const float* x, float* y, int64_t n, int64_t hw, int64_t C, int layout, int bn_mode,
                     const float* p0, const float* p1, int act) {
  for (int64_t i = 0; i < n; ++i) {
    int64_t c = layout == 1 ? i % C : (i / hw) % C;
    float v;
    if (bn_mode == 0) {
      v = fmaf(x[i], p0[c], p1[c]);
    } else {
      volatile float d = x[i] - p0[4 * c];
      volatile float g = p0[4 * c + 1] * d;
      v = fmaf(g, p0[4 * c + 2], p0[4 * c + 3]);
    }
    if (act >= 1) v = f_max_nan(v, 0.0f);
    if (act == 2) v = f_min_nan(v, 6.0f);
    y[i] = v;
  }
}
