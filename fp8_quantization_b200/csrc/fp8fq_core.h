// fp8fq_core.h -- the arithmetic of the FP8 fake-quantiser, shared by the CUDA kernels and by a
// host build used for CPU-side logic tests (tests/host_emul).  Everything here is a pure function.
//
// Reference being reproduced: quantize_to_fp8_ste_MM, quantization/quantizers/fp8_quantizer.py:91-133
//
//   M    = clamp(round(mantissa_bits), 1, n_bits - sign_bits)                        (:105)
//   E    = n_bits - sign_bits - M                                                    (:106)
//   bias = 2^E - log2(maxval) + log2(2 - 2^-M) - 1                                   (:110)
//   xc   = min(max(x, minval), maxval),  minval = -maxval | 0                        (:112-113)
//   e    = clamp_min(floor(log2|xc| + bias), 1)                                      (:128)
//   s    = 2^(e - M - bias)                                                          (:130)
//   y    = round_half_even(xc / s) * s                                               (:132)
//
// Design: e and s depend on x only through which of K = max(1, 2^E - 1) intervals |xc| falls in.
// A prologue ("prepare") evaluates, with the SAME fp32 operations and the same libm entry points the
// reference's ATen kernels call (log2f, powf; no FMA contraction across the reference's op
// boundaries), the bias, the K scales and the K-1 switching points of (:128), found by bisection on
// the float ordering.  The streaming kernel then needs per element: two NaN-propagating min/max, a
// table lookup, one multiply by the pre-computed reciprocal with an exactness guard that falls back
// to IEEE division near rounding ties, one round-to-nearest-even and one multiply.  No log2, no pow,
// no division on the fast path -- and bit-identical results to evaluating (:128)-(:132) directly.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define FQ_HD __host__ __device__ __forceinline__
#else
#define FQ_HD inline
#endif

namespace fp8fq {

constexpr int kHdr = 8;          // header floats per channel
constexpr int kMaxE = 7;         // at most 127 exponent codes
constexpr int kMaxM = 12;        // fast-path tie guard validated up to here
constexpr int kMaxK = (1 << kMaxE) - 1;

enum : int { H_HI = 0, H_LO = 1, H_BASE = 2, H_BIAS = 3, H_FLAGS = 4, H_K = 5, H_GUARD = 6, H_REF = 7 };
// flags = FLAG_* | (band field << BAND_SHIFT) | (M << M_SHIFT); band field = band (< 2^20 - 1), or kBandMask for
// "every mantissa is ambiguous" (band = 0x7fffff); M <= kMaxM < 16
enum : int { FLAG_IRREGULAR = 1, FLAG_POW2 = 2, FLAG_RSNAN = 4, FLAG_SDOUBLE = 8, FLAG_MAGIC = 16, FLAG_TWO = 32, BAND_SHIFT = 8,
             M_SHIFT = 28 };
constexpr uint32_t kBandMask = 0xfffffu;
// FLAG_MAGIC: how far (ulps) a switching point T_k may sit from its ideal position s_k * 2^M (see prep_finish)
constexpr int kMagicMaxDev = 128;
// ... and how far (ulps) the normalised scales / reciprocals of the two groups of a two-group table may differ
constexpr int kMagicGroupUlps = 16;   // (11 ulps occur: bias in [8, 16), |y| in [16, 32))

FQ_HD int k_codes(int E) { return E <= 0 ? 1 : ((1 << E) - 1); }
FQ_HD int k_pad(int K) { return (K + 2) & ~1; }                       // K+1 rounded up to even
constexpr int kExt = 4;           // trailing floats: constants of the scaled-domain path at a FIXED offset (see off_ext)
FQ_HD int table_stride(int K) { return kHdr + k_pad(K) + 2 * (K + 1) + kExt; }
FQ_HD int off_thr(int) { return kHdr; }
FQ_HD int off_sr(int K) { return kHdr + k_pad(K); }
// [0] switching point that selects scale group b (NaN: one group), [1] normalised scale of group b (s_1 if one group),
// [2] the tie guard of the scaled-domain path, [3] unused.  At an offset that depends on K only, so that a kernel's
// prologue reads them with loads independent of each other (a look-up through the break code stored in H_K was a chain
// of two dependent loads in front of every CTA's data loads: measured 3-10 % on one-tile CTAs, round 2 call u).
FQ_HD int off_ext(int K) { return kHdr + k_pad(K) + 2 * (K + 1); }
enum : int { X_TB = 0, X_SB = 1, X_GUARD = 2 };

FQ_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
FQ_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}

// ---- fp32 primitives with pinned rounding (never fused) ------------------------------------
FQ_HD float add_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b; return r;
#endif
}
FQ_HD float sub_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  volatile float r = a - b; return r;
#endif
}
FQ_HD float mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b; return r;
#endif
}
FQ_HD float div_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  volatile float r = a / b; return r;
#endif
}
FQ_HD float rcp_rn(float a) {
#if defined(__CUDA_ARCH__)
  return __frcp_rn(a);
#else
  volatile float r = 1.0f / a; return r;
#endif
}
// torch.max / torch.min semantics: NaN in either operand propagates.
FQ_HD float max_nan(float a, float b) {
#if defined(__CUDA_ARCH__)
  float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r;
#else
  // host builds (tests): PTX max.NaN semantics incl. the signed-zero order -0.0 < +0.0
  if (a != a) return a; if (b != b) return b;
  if (a == b) return u2f(f2u(a) & f2u(b));   // equal values: +0.0 unless both are -0.0
  return a > b ? a : b;
#endif
}
FQ_HD float min_nan(float a, float b) {
#if defined(__CUDA_ARCH__)
  float r; asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r;
#else
  if (a != a) return a; if (b != b) return b;
  if (a == b) return u2f(f2u(a) | f2u(b));   // equal values: -0.0 if either is -0.0 (PTX min.NaN)
  return a < b ? a : b;
#endif
}
FQ_HD bool is_normal_pos(float f) {
  uint32_t u = f2u(f);
  return u >= 0x00800000u && u < 0x7f800000u;
}

// ---- format split (:105-106); host side --------------------------------------------------------
inline int format_split(float mantissa_bits, int n_bits, int sign_bits, int* M, int* E, int* K) {
  if (!(mantissa_bits == mantissa_bits)) return -1;
  float r = nearbyintf(mantissa_bits);  // torch.round: half to even
  float hi = (float)(n_bits - sign_bits);
  if (hi < 1.0f) return -1;             // torch.clamp with min > max returns max; not a valid format
  if (r < 1.0f) r = 1.0f;
  if (r > hi) r = hi;
  int m = (int)r, e = n_bits - sign_bits - m;
  if (e > kMaxE || m > kMaxM) return -2;
  *M = m; *E = e; *K = k_codes(e);
  return 0;
}

// ---- prologue pieces ---------------------------------------------------------------------------
// bias (:110), evaluated left to right like the Python expression.
FQ_HD float ref_bias(float maxval, int M, int E) {
  float p2E = powf(2.0f, (float)E);           // 2**E      (E is a float tensor in the reference)
  float t = sub_rn(p2E, log2f(maxval));       // - log2(maxval)
  float p2nM = powf(2.0f, -(float)M);         // 2 ** (-M)
  float l2 = log2f(sub_rn(2.0f, p2nM));       // log2(2 - 2**-M)
  t = add_rn(t, l2);
  return sub_rn(t, 1.0f);
}
// scale of exponent code e (:130): 2.0 ** ((e - M) - bias)
FQ_HD float ref_scale(int e, int M, float bias) {
  float y = sub_rn(sub_rn((float)e, (float)M), bias);
  return powf(2.0f, y);
}
// the reference's exponent code before the clamp (:128) reaches k?
FQ_HD bool ref_code_ge(float a, float bias, float k) {
  float v = floorf(add_rn(log2f(a), bias));
  return v >= k;  // NaN -> false
}
// Smallest positive float a with ref_code_ge(a, bias, k), by bisection on the float ordering
// (positive floats order like their bit patterns).  NaN ("a >= NaN" is never true) if there is none.
FQ_HD float find_threshold(float bias, int k) {
  const float kf = (float)k;
  uint32_t lo = 0u;            // invariant: code(lo) < k   (a = +0 -> log2 = -inf)
  uint32_t hi = 0x7f800000u;   // invariant: code(hi) >= k, or hi == +inf
  if (!ref_code_ge(u2f(hi), bias, kf)) return u2f(0x7fc00000u);  // never reached: NaN compares false
  if (ref_code_ge(u2f(1u), bias, kf)) return u2f(1u);
  lo = 1u;
  // narrow bracket around the analytic switching point 2^(k - bias) first
  float g = exp2f(sub_rn(kf, bias));
  if (is_normal_pos(g)) {
    uint32_t gi = f2u(g);
    uint32_t span = 1u << 12;
    uint32_t l2 = gi > span + 1u ? gi - span : 1u;
    uint32_t h2 = gi + span < 0x7f800000u ? gi + span : 0x7f800000u;
    if (!ref_code_ge(u2f(l2), bias, kf) && l2 > lo) lo = l2;
    if (ref_code_ge(u2f(h2), bias, kf) && h2 < hi) hi = h2;
  }
  while (hi - lo > 1u) {
    uint32_t mid = lo + ((hi - lo) >> 1);
    if (ref_code_ge(u2f(mid), bias, kf)) hi = mid; else lo = mid;
  }
  // guard against a locally non-monotone log2f: take the lowest switching point within 4 ulps
  for (uint32_t t = 1; t <= 4u && hi > t; ++t)
    if (ref_code_ge(u2f(hi - t), bias, kf)) { hi -= t; t = 0; }
  return u2f(hi);
}

// Writes header + thr/sr entries of one channel given bias and the already computed thresholds
// T[k] (k = 2..K, T[k] stored at thr_in[k]) -- used by the host build; the device does the same
// work spread over a CTA (see prepare kernel).
FQ_HD float tie_guard(int M) {
  // |xc/s| <= 2^(M+1)(1+eps); multiplying by RN(1/s) instead of dividing by s perturbs the quotient
  // by < 1.6 * 2^-23 relative.  If the fast quotient is further than this from a .5 tie, rounding
  // it gives the same integer as rounding the IEEE quotient.  2^(M-20) >= 2.6x that bound.
  return 0.5f - ldexpf(1.0f, M - 20);
}

// One channel's table is built in three steps so that a CTA can spread step 2 over its threads
// while the host emulation (tests/host_emul) runs the very same code serially.
// step 1 (one thread): header
FQ_HD float prep_header(float* tab, float mv, int M, int E, int K, int sign_bits) {
  const float bias = ref_bias(mv, M, E);
  float* thr = tab + kHdr;
  tab[H_HI] = mv;
  tab[H_LO] = sign_bits ? -mv : 0.0f;  // (:112) -maxval or zeros_like(maxval)
  tab[H_BIAS] = bias;
  tab[H_K] = u2f((uint32_t)K);
  tab[H_GUARD] = tie_guard(M);
  tab[H_REF] = 0.0f;
  thr[0] = u2f(0x7fc00000u);  // NaN: "a >= thr" is false for every a, including +inf
  for (int j = K; j < k_pad(K); ++j) thr[j] = u2f(0x7fc00000u);
  return bias;
}
// step 2 (any thread, k = 1..K): scale pair of code k and the threshold at which code k starts
FQ_HD void prep_entry(float* tab, int k, int M, int K, float bias) {
  float* thr = tab + kHdr;
  float* sr = tab + off_sr(K);
  const float s = ref_scale(k, M, bias);
  float rs = rcp_rn(s);
  if (!(is_normal_pos(s) && is_normal_pos(rs))) rs = u2f(0x7fc00000u);  // forces the exact-division path
  sr[2 * k] = s;
  sr[2 * k + 1] = rs;
  if (k == 1) { sr[0] = s; sr[1] = rs; }
  if (k >= 2) thr[k - 1] = find_threshold(bias, k);
}
// step 3 (one thread, after all entries are visible): bucket base + regularity check
FQ_HD void prep_finish(float* tab, int M, int K, float mv) {
  const float* thr = tab + kHdr;
  int flags = 0;
  uint32_t base = 0x00800000u;
  if (K >= 2) {
    const float t2 = thr[1];
    const uint32_t t2i = f2u(t2);
    if (is_normal_pos(t2) && t2i > 0x00C00000u + 0x00800000u) {
      base = t2i - 0x00C00000u;  // 1.5 binades below T_2 on the float bit-pattern axis
      for (int k = 2; k <= K; ++k) {
        const uint32_t ti = f2u(thr[k - 1]);
        if (!is_normal_pos(thr[k - 1]) || ti < base || (int)((ti - base) >> 23) != k - 1) flags |= FLAG_IRREGULAR;
      }
      // the largest clamped input must not index past entry K
      const float hi = fabsf(mv);
      if (is_normal_pos(hi) && f2u(hi) >= base && (int)((f2u(hi) - base) >> 23) > K) flags |= FLAG_IRREGULAR;
    } else {
      flags |= FLAG_IRREGULAR;
    }
  }
  // Exponent-arithmetic fast path (lookup_code_fast): the thresholds are almost exact doublings of each other,
  // T_k = T_2 * 2^(k-2) up to a few ulps d_k.  With ref = bits(T_2) + min d_k and band = max d_k - min d_k, every
  // |xc| whose distance to ref is not within `band` ulps above a binade multiple has its code determined by
  // integer arithmetic alone; the (band + 1) / 2^23 rest takes the table lookup.
  uint32_t ref = 0, band = 0x7fffffu;  // band = all mantissas: always use the table
  if (K >= 2 && !(flags & FLAG_IRREGULAR)) {
    const int64_t t2 = (int64_t)f2u(thr[1]);
    int64_t dmin = 0, dmax = 0;
    for (int k = 3; k <= K; ++k) {
      const int64_t d = (int64_t)f2u(thr[k - 1]) - (t2 + ((int64_t)(k - 2) << 23));
      dmin = d < dmin ? d : dmin;
      dmax = d > dmax ? d : dmax;
    }
    if (dmax - dmin < (1 << 20) - 1 && t2 + dmin > 0) {
      ref = (uint32_t)(t2 + dmin);
      band = (uint32_t)(dmax - dmin);
    }
  } else if (K == 1) {
    ref = 0x7f800000u;  // every finite |xc| is below: code 1
    band = 0;
  }
  // FLAG_POW2: every scale is an exact power of two 2^k with 2^-k representable, so x / s == x * 2^-k bit for bit
  // (used by the STE backward kernel, which reproduces autograd's (g * s) / s)
  {
    const float* sr = tab + off_sr(K);
    bool pow2 = true;
    for (int k = 1; k <= K; ++k) {
      const uint32_t sb = f2u(sr[2 * k]);
      pow2 = pow2 && (sb & 0x007fffffu) == 0u && sb >= 0x01000000u && sb <= 0x7e000000u;
    }
    if (pow2) flags |= FLAG_POW2;
    // FLAG_RSNAN: some 1/s is not usable (stored as NaN): the multiply-by-reciprocal path must keep its guard, which
    // then routes every element to the IEEE division (consumers that otherwise skip the guard: the MSE kernel)
    bool rsnan = false;
    for (int k = 1; k <= K; ++k) rsnan = rsnan || !(sr[2 * k + 1] == sr[2 * k + 1]);
    if (rsnan) flags |= FLAG_RSNAN;
    // FLAG_SDOUBLE: the scales the reference's powf produced are EXACT doublings of each other, s_k = s_1 * 2^(k-1) bit
    // for bit (and therefore 1/s_k = (1/s_1) * 2^-(k-1)), all normal -- the usual case (the exponents (k - M) - bias of
    // neighbouring codes differ by exactly 1 unless the subtraction rounds differently across a binade of |y|: measured
    // 87 % of random ranges for E3M4, 96 % for E4M3).  The element path then derives (s, 1/s) from the exponent code by
    // integer arithmetic on the bit patterns instead of loading them (lookup_scale_fast).
    bool dbl = K >= 2 && !(flags & FLAG_IRREGULAR) && !rsnan;
    for (int k = 1; k <= K && dbl; ++k) {
      const uint32_t sh = (uint32_t)(k - 1) << 23;
      dbl = is_normal_pos(sr[2 * k]) && is_normal_pos(sr[2 * k + 1]) && f2u(sr[2 * k]) == f2u(sr[2]) + sh &&
            f2u(sr[2 * k + 1]) + sh == f2u(sr[3]);
    }
    if (dbl) flags |= FLAG_SDOUBLE;
    // FLAG_MAGIC: the whole element path can run in the SCALED domain u = |xc| / s_1 (quant_magic below): with exact
    // doublings the reference's grid is s_1 * {q * 2^(e-1)}, i.e. u rounded to M + 1 significant bits with a fixed
    // spacing of 1 below 2^(M+1) -- an ordinary floating-point rounding, done by adding and subtracting
    // 2^(p+23), p = max(exponent(u) - M, 0) (<= K - 1 because |xc| <= maxval).  That places the code boundaries at the powers of two
    // 2^(M+k-1) of u instead of at the reference's switching points T_k; the two differ by the fp32 noise of the
    // reference's log2 / pow (a few ulps), and an element between them is within that distance of the boundary value
    // itself, which BOTH spacings contain -- it rounds to it either way as long as the distance stays far below half
    // the finer spacing (2^(21-M) ulps).  Checked here per table: every T_k within kMagicMaxDev ulps of s_k * 2^M.
    // K == 1 formats qualify too (p is always 0).
    // Two-group tables: where |(k - M) - bias| crosses into a binade coarser than the one bias was rounded in, the
    // reference's subtraction rounds and the scales of the codes on either side of that point, each group an exact
    // doubling sequence in itself, differ by a few ulps after normalisation (s_k * 2^-(k-1)).  E3M4 with maxval in
    // [2, 8) -- the usual range behind a ReLU6 -- is the typical case: code 1 against codes 2..7.  Such a table
    // qualifies as well: the rounding runs with group a's reciprocal for every element (a quotient perturbed by the
    // <= kMagicGroupUlps difference, covered by a wider tie guard, magic_consts), and the final multiply takes the scale
    // of the element's own group, decided by comparing |xc| with the reference's own switching point T_{kb+1} -- an
    // exact classification, so the one boundary whose two sides are different floats is resolved as the reference does.
    int kb = 0;   // last code of group a (0: one group)
    bool magic = !rsnan && !(flags & FLAG_IRREGULAR) && M >= 1 && M <= kMaxM;
    for (int k = 1; k <= K && magic; ++k) magic = is_normal_pos(sr[2 * k]) && is_normal_pos(sr[2 * k + 1]);
    if (magic && K >= 2) {
      auto sn = [&](int k) { return (int64_t)f2u(sr[2 * k]) - ((int64_t)(k - 1) << 23); };
      auto rn = [&](int k) { return (int64_t)f2u(sr[2 * k + 1]) + ((int64_t)(k - 1) << 23); };
      int g = 1;
      while (g < K && sn(g + 1) == sn(1) && rn(g + 1) == rn(1)) ++g;
      if (g < K) {
        kb = g;
        for (int k = kb + 2; k <= K && magic; ++k) magic = sn(k) == sn(kb + 1) && rn(k) == rn(kb + 1);
        const int64_t ds = sn(kb + 1) - sn(1), dr = rn(kb + 1) - rn(1);
        magic = magic && ds >= -kMagicGroupUlps && ds <= kMagicGroupUlps && dr >= -kMagicGroupUlps && dr <= kMagicGroupUlps &&
                sn(kb + 1) >= 0x00800000 && sn(kb + 1) < 0x7f800000 && M <= 8;   // (M: the wider guard of magic_consts)
      }
    }
    for (int k = 2; k <= K && magic; ++k) {
      const float ideal = ldexpf(sr[2 * k], M);
      const int64_t dev = (int64_t)f2u(thr[k - 1]) - (int64_t)f2u(ideal);
      magic = is_normal_pos(ideal) && is_normal_pos(thr[k - 1]) && dev >= -kMagicMaxDev && dev <= kMagicMaxDev;
    }
    // no upper clamp of p in the element path: the largest clamped input must lie inside the top binade of the scaled
    // domain, maxval / s_1 < 2^(M+K) with room for the perturbations above; and nothing may overflow
    magic = magic && is_normal_pos(ldexpf(sr[2], M + K + 1)) && is_normal_pos(mv) &&
            mul_rn(mul_rn(mv, sr[3]), 1.0f + 1.0f / 65536.0f) < ldexpf(1.0f, M + K) && M + K + 24 < 127;
    if (magic) flags |= FLAG_MAGIC;
    if (!magic) kb = 0;
    if (kb) flags |= FLAG_TWO;   // (a bit of the flags word: the kernels branch on it with a uniform predicate)
    tab[H_K] = u2f((uint32_t)K | ((uint32_t)kb << 8));
    float* ext = tab + off_ext(K);
    ext[X_TB] = kb ? thr[kb] : u2f(0x7fc00000u);
    ext[X_SB] = kb ? u2f(f2u(sr[2 * K]) - ((uint32_t)(K - 1) << 23)) : sr[2];
    // tie guard: 1/2 - 2^(M-20) (tie_guard); two groups: the quotient of a group-b element is taken with group a's
    // reciprocal, i.e. perturbed by up to kMagicGroupUlps more ulps -- (16 + 1.5) * 2^-23 * 2^(M+1) < 2^(M-17)
    ext[X_GUARD] = kb ? 0.5f - ldexpf(1.0f, M - 17) : tab[H_GUARD];
    ext[3] = 0.0f;
  }
  tab[H_BASE] = u2f(base);
  tab[H_REF] = u2f(ref);
  tab[H_FLAGS] = u2f((uint32_t)flags | ((band < kBandMask ? band : kBandMask) << BAND_SHIFT) | ((uint32_t)M << M_SHIFT));
}

FQ_HD uint32_t flags_band(uint32_t fl) {
  const uint32_t b = (fl >> BAND_SHIFT) & kBandMask;
  return b == kBandMask ? 0x7fffffu : b;
}
FQ_HD int flags_M(uint32_t fl) { return (int)(fl >> M_SHIFT); }

// ---- element path ------------------------------------------------------------------------------
// Generic lookup of e' (index into the (scale, rcp) pairs) for a = |xc|, from a channel table.
template <typename Ld>
FQ_HD int lookup_code(float a, const float* tab, int K, uint32_t base, bool irregular, Ld ld) {
  const float* thr = tab + kHdr;
  if (!irregular) {
    // bucket j = binade of |xc| relative to `base` (a float 1.5 binades below T_2); bucket j
    // (1 <= j <= K-1) contains exactly the threshold T_{j+1}.
    float ac = fmaxf(a, u2f(base));              // non-NaN max: NaN -> base, j = 0
    int j = (int)((f2u(ac) - base) >> 23);
    j = j < K ? j : K;
    return j + (a >= ld(thr + j) ? 1 : 0);
  }
  int e = 1;
  for (int k = 2; k <= K; ++k) e += (a >= ld(thr + (k - 1)) ? 1 : 0);
  return e;
}

// Exponent-arithmetic code lookup (see prep_finish).  Returns the code in [1, K]; *ambiguous is set when |xc| lies
// in the band where the thresholds' mantissas differ and the caller must use lookup_code instead.
FQ_HD int lookup_code_fast(float a, uint32_t ref, uint32_t band, int K, bool* ambiguous) {
  const int32_t i = (int32_t)(f2u(a) - ref);
  const int32_t q = i >> 23;  // arithmetic shift: floor(i / 2^23)
  *ambiguous = (uint32_t)(i & 0x7fffff) <= band;
  int e = q + 2;
  e = e < 1 ? 1 : e;
  return e > K ? K : e;
}

// Scale pair of the element's exponent code for FLAG_SDOUBLE tables, by integer arithmetic alone: with
// i = bits(|xc|) - ref, the code is e = clamp(floor(i / 2^23) + 2, 1, K) (lookup_code_fast), so
// t = (e - 1) << 23 = clamp((i & ~0x7fffff) + 2^23, 0, (K - 1) << 23), s = s_1 * 2^(e-1) has the bits s1b + t and
// 1/s the bits r1b - t.  *ambiguous as in lookup_code_fast.  Returns t.
FQ_HD uint32_t lookup_scale_fast(float a, uint32_t ref, uint32_t band, uint32_t tmax, uint32_t s1b, uint32_t r1b,
                                 float* s, float* rs, bool* ambiguous) {
  const int32_t i = (int32_t)(f2u(a) - ref);
  *ambiguous = (uint32_t)(i & 0x7fffff) <= band;
  int32_t t = (int32_t)((uint32_t)i & 0xff800000u) + (1 << 23);   // floor(i / 2^23) * 2^23 + 2^23 (two's complement)
  t = t < 0 ? 0 : t;
  t = t > (int32_t)tmax ? (int32_t)tmax : t;
  *s = u2f(s1b + (uint32_t)t);
  *rs = u2f(r1b - (uint32_t)t);
  return (uint32_t)t;
}

// ---- INT uniform quantisers (quantization/quantizers/uniform_quantizers.py:107-164) ------------------------
//   scale = clamp(delta, min=eps); zp = clamp(round(zero_float), int_min, int_max) | 0 (symmetric)
//   x_int = clamp(round(x / scale) + zp, int_min, int_max);  y = scale * (x_int - zp)
// Channel table (kUStride floats): scale, 1/scale, zp, int_min, int_max, tie guard, saturation bound.
constexpr int kUStride = 8;
enum : int { U_SCALE = 0, U_RS = 1, U_ZP = 2, U_IMIN = 3, U_IMAX = 4, U_GUARD = 5, U_SAT = 6 };

FQ_HD void uq_build(float* tab, float delta, float zero_float, float int_min, float int_max, float eps, int n_bits,
                    bool symmetric) {
  const float scale = delta < eps ? eps : delta;  // torch.clamp(delta, min=eps): NaN stays NaN
  float rs = rcp_rn(scale);
  if (!(is_normal_pos(scale) && is_normal_pos(rs))) rs = u2f(0x7fc00000u);
  float zp = 0.0f;
  if (!symmetric) zp = min_nan(max_nan(nearbyintf(zero_float), int_min), int_max);
  tab[U_SCALE] = scale;
  tab[U_RS] = rs;
  tab[U_ZP] = zp;
  tab[U_IMIN] = int_min;
  tab[U_IMAX] = int_max;
  // |x/scale| < sat = 2^(n_bits+2): multiplying by RN(1/scale) perturbs the quotient by < 1.6 * 2^(n_bits-21), the
  // guard leaves 2^(n_bits-19); beyond sat the clamp saturates whatever the rounding was
  tab[U_GUARD] = 0.5f - ldexpf(1.0f, n_bits - 19);
  tab[U_SAT] = ldexpf(1.0f, n_bits + 2);
  tab[7] = 0.0f;
}

FQ_HD float uq_quant(float x, float scale, float rs, float zp, float imin, float imax, float guard, float sat) {
  const float r = mul_rn(x, rs);
  float q = nearbyintf(r);
  if (!(fabsf(r - q) < guard) && !(fabsf(r) >= sat)) q = nearbyintf(div_rn(x, scale));
  const float xi = min_nan(max_nan(add_rn(q, zp), imin), imax);
  return mul_rn(scale, sub_rn(xi, zp));
}

// Quantise xc (already clamped) with the selected (s, rs).  Returns y; *q_out = round(xc / s).
// ---- FLAG_MAGIC tables: the element path in the scaled domain (see prep_finish) ----------------------------------
// Per-table constants, all derived from (M, tie guard, s_1, 1/s_1).
struct MagicConsts {
  float s1, r1;        // scale of code 1 and its reciprocal (group a)
  float sb, tb;        // two-group tables: normalised scale of group b and the switching point |xc| >= tb that selects
                       // it; one group: sb == s1, tb = NaN (never)
  float kap;           // guard * 2^p == C * kap for the magic constant C = 2^(p+23)
  uint32_t lo, add;    // bits(C) = max(bits(u) & 0x7f800000, lo) + add
  bool two;
};
// tab: the channel table; ld reads one float of it (global or shared memory).  Independent loads at fixed offsets.
template <typename Ld>
FQ_HD MagicConsts magic_consts(const float* tab, int K, uint32_t flags, Ld ld) {
  const int M = flags_M(flags);
  MagicConsts m;
  m.s1 = ld(tab + off_sr(K) + 2);
  m.r1 = ld(tab + off_sr(K) + 3);
  m.tb = ld(tab + off_ext(K) + X_TB);
  m.sb = ld(tab + off_ext(K) + X_SB);
  m.two = (flags & FLAG_TWO) != 0;
  m.kap = ld(tab + off_ext(K) + X_GUARD) * (1.0f / 8388608.0f);   // 2^-23
  m.lo = (uint32_t)(127 + M) << 23;                           // exponent field of 2^M: p = 0 up to u < 2^(M+1)
  m.add = (uint32_t)(23 - M) << 23;                           // -> exponent p + 23 (u >= 0: u + C stays in C's binade)
  return m;
}
// One element.  Returns |y| (the caller restores the sign of xc: a negative value that rounds to zero is -0.0 in the
// reference); *ok is false when the reciprocal multiply landed within the guard band of a rounding tie (or xc is NaN):
// the caller must then take the exact path.  Instruction count: FMUL, LOP, integer max, integer add, 3 FADD, FMUL,
// FSETP, FMUL (+ FSETP, FSEL for a two-group table) -- no table access, no FRND.  There is no upper clamp of p: prep_finish sets FLAG_MAGIC only when the
// largest clamped input, maxval / s_1, lies inside the top binade [2^(M+K-1), 2^(M+K)).
FQ_HD float quant_magic(float xc, const MagicConsts& m, bool* ok) {
  const float u = mul_rn(fabsf(xc), m.r1);
  uint32_t cb = f2u(u) & 0x7f800000u;
  cb = (cb < m.lo ? m.lo : cb) + m.add;
  const float C = u2f(cb);
  const float qu = sub_rn(add_rn(u, C), C);                   // u rounded half-to-even to a multiple of 2^p
  *ok = fabsf(sub_rn(u, qu)) < mul_rn(C, m.kap);
  return mul_rn(qu, fabsf(xc) >= m.tb ? m.sb : m.s1);
}

FQ_HD float quant_core(float xc, float s, float rs, float guard, float* q_out) {
  float r = mul_rn(xc, rs);
  float q = nearbyintf(r);
  float d = r - q;
  if (!(fabsf(d) < guard)) q = nearbyintf(div_rn(xc, s));  // near a tie, or rs unusable (NaN)
  *q_out = q;
  return mul_rn(q, s);
}

}  // namespace fp8fq
