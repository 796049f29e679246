// tests/host_emul/emul.cpp -- host build of the product's table/bucket/tie-guard algorithm
// (fp8_quantization_b200/csrc/fp8fq_core.h) so its logic can be tested without a GPU.  TEST ONLY:
// the product has no CPU path; this file is never part of libfp8fq.so.
#include <stdint.h>
#include <vector>
#include "../../fp8_quantization_b200/csrc/fp8fq_core.h"

using namespace fp8fq;

extern "C" {

int emul_table_stride(float mantissa_bits, int n_bits, int sign_bits) {
  int M, E, K;
  if (format_split(mantissa_bits, n_bits, sign_bits, &M, &E, &K) != 0) return -1;
  return table_stride(K);
}

// serial equivalent of prepare_kernel
int emul_prepare(const float* maxval, int64_t C, float mantissa_bits, int n_bits, int sign_bits, float* table) {
  int M, E, K;
  if (format_split(mantissa_bits, n_bits, sign_bits, &M, &E, &K) != 0) return -1;
  const int stride = table_stride(K);
  for (int64_t c = 0; c < C; ++c) {
    float* tab = table + c * stride;
    const float bias = prep_header(tab, maxval[c], M, E, K, sign_bits);
    for (int k = 1; k <= K; ++k) prep_entry(tab, k, M, K, bias);
    prep_finish(tab, M, K, maxval[c]);
  }
  return 0;
}

// serial equivalent of quant_elem<1, true> over a [C, inner] tensor; force_irregular exercises the
// linear-search lookup; codes as in fp8fq_fake_quant_codes_f32
int emul_fake_quant(const float* x, float* y, int32_t* codes, const float* table, int64_t C, int64_t inner,
                    float mantissa_bits, int n_bits, int sign_bits, int force_irregular, int64_t* slow_count) {
  int M, E, K;
  if (format_split(mantissa_bits, n_bits, sign_bits, &M, &E, &K) != 0) return -1;
  const int stride = table_stride(K);
  int64_t slow = 0;
  for (int64_t c = 0; c < C; ++c) {
    const float* tab = table + c * stride;
    const float hi = tab[H_HI], lo = tab[H_LO], guard = tab[H_GUARD];
    const uint32_t base = f2u(tab[H_BASE]);
    const bool irregular = force_irregular == 1 || (f2u(tab[H_FLAGS]) & FLAG_IRREGULAR);
    const uint32_t ref = f2u(tab[H_REF]), band = force_irregular == 1 ? 0x7fffffu : flags_band(f2u(tab[H_FLAGS]));
    // force_irregular: 0 = as the kernels run (FLAG_MAGIC tables take the scaled-domain path, quant_vec<1, false>),
    // 1 = linear threshold scan, 2 = the table path even for FLAG_MAGIC tables (what the code-plane variant runs)
    const bool magic = force_irregular == 0 && codes == nullptr && (f2u(tab[H_FLAGS]) & FLAG_MAGIC) != 0;
    const MagicConsts mc = magic_consts(tab, K, f2u(tab[H_FLAGS]), [](const float* p) { return *p; });
    for (int64_t i = 0; i < inner; ++i) {
      const float v = x[c * inner + i];
      const float xc = min_nan(max_nan(v, lo), hi);
      if (magic) {
        bool ok;
        const float ya = quant_magic(xc, mc, &ok);
        if (ok) {
          y[c * inner + i] = u2f((f2u(ya) & 0x7fffffffu) | (f2u(xc) & 0x80000000u));
          continue;
        }
        ++slow;
      }
      const float a = fabsf(xc);
      bool amb;
      int e = lookup_code_fast(a, ref, band, K, &amb);  // same two-level lookup as quant_vec<1, ...>
      if (amb) e = lookup_code(a, tab, K, base, irregular, [](const float* p) { return *p; });
      const float s = tab[off_sr(K) + 2 * e], rs = tab[off_sr(K) + 2 * e + 1];
      e = e < 1 ? 1 : e;
      float q;
      {  // count how often the exact-division fallback fires
        float r = mul_rn(xc, rs);
        float qq = nearbyintf(r);
        if (!(fabsf(r - qq) < guard)) ++slow;
      }
      const float yy = quant_core(xc, s, rs, guard, &q);
      y[c * inner + i] = yy;
      if (codes) {
        if (yy != yy) codes[c * inner + i] = 0x7fffffff;
        else codes[c * inner + i] = (int32_t)((f2u(yy) & 0x80000000u) | ((uint32_t)e << 16) | (uint32_t)fabsf(q));
      }
    }
  }
  if (slow_count) *slow_count = slow;
  return 0;
}

// serial equivalents of uq_prepare_kernel / quant_vec<2, ...> (INT uniform quantisers)
int emul_uq_prepare(const float* xmin, const float* xmax, int64_t C, int n_bits, int symmetric, float eps, float* delta,
                    float* table) {
  float mn = INFINITY;
  for (int64_t c = 0; c < C; ++c) mn = min_nan(mn, min_nan(xmin[c], 0.0f));
  const bool is_signed = mn < 0.0f;
  float int_min = 0.0f, int_max = ldexpf(1.0f, n_bits) - 1.0f;
  if (symmetric) {
    int_min = is_signed ? -ldexpf(1.0f, n_bits - 1) : 0.0f;
    int_max = ldexpf(1.0f, n_bits - (is_signed ? 1 : 0)) - 1.0f;
  }
  for (int64_t c = 0; c < C; ++c) {
    const float xm = min_nan(xmin[c], 0.0f), xM = max_nan(xmax[c], eps);
    float d, zf = 0.0f;
    if (symmetric) d = div_rn(max_nan(fabsf(xm), xM), int_max);
    else { d = div_rn(sub_rn(xM, xm), int_max); zf = div_rn(-xm, d); }
    if (delta) delta[c] = d;
    uq_build(table + c * kUStride, d, zf, int_min, int_max, eps, n_bits, symmetric != 0);
  }
  return 0;
}

int emul_uq_quant(const float* x, float* y, const float* table, int64_t C, int64_t inner) {
  for (int64_t c = 0; c < C; ++c) {
    const float* t = table + c * kUStride;
    for (int64_t i = 0; i < inner; ++i)
      y[c * inner + i] = uq_quant(x[c * inner + i], t[U_SCALE], t[U_RS], t[U_ZP], t[U_IMIN], t[U_IMAX], t[U_GUARD], t[U_SAT]);
  }
  return 0;
}

int emul_table_flags(const float* table, int64_t c, float mantissa_bits, int n_bits, int sign_bits) {
  int M, E, K;
  if (format_split(mantissa_bits, n_bits, sign_bits, &M, &E, &K) != 0) return -1;
  return (int)(f2u(table[c * table_stride(K) + H_FLAGS]) & 0xff);
}

// mantissa band of the exponent-arithmetic fast path (ulps); 0x7fffff = fast path disabled for this channel
// 0: one scale group (or not a FLAG_MAGIC table); kb >= 1: codes 1..kb / kb+1..K are the two groups
int emul_table_break(const float* table, int64_t c, float mantissa_bits, int n_bits, int sign_bits) {
  int M, E, K;
  if (format_split(mantissa_bits, n_bits, sign_bits, &M, &E, &K) != 0) return -1;
  return (int)(f2u(table[c * table_stride(K) + H_K]) >> 8);
}

int emul_table_band(const float* table, int64_t c, float mantissa_bits, int n_bits, int sign_bits) {
  int M, E, K;
  if (format_split(mantissa_bits, n_bits, sign_bits, &M, &E, &K) != 0) return -1;
  return (int)flags_band(f2u(table[c * table_stride(K) + H_FLAGS]));
}
}
