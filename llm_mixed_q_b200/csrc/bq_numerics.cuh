// bq_numerics.cuh — element arithmetic of the four block formats, written so that every
// fp32 operation happens in the SAME ORDER and with the SAME ROUNDING as the reference's
// torch emulation (paths relative to /root/reference/src/llm_mixed_q/models/quantize/quantizers/).
//
// Rules that make the results bit-identical to torch-CUDA:
//   * no -use_fast_math, -fmad=false; every rounding step is an explicit __f*_rn intrinsic;
//   * log2f is libdevice's precise routine (the one torch.log2 lowers to) — never __log2f,
//     never an integer exponent extract (ceil/floor/round(log2f) cliffs);
//   * 2**e on an integer-valued fp32 is exact in [-149,127], +inf at >= 128, 0 below -149;
//   * x / 2**e is IEEE division; it is replaced by a multiplication with the exact
//     reciprocal only when 2**-e is a normal power of two (identical correctly-rounded result);
//   * torch.clamp propagates NaN, torch.sign(NaN) = 0, torch.round = rintf (half to even);
//   * bool*float blends ((~c)*q + c*x) are kept as multiplications: 0*inf = NaN poisons
//     exactly like the reference.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bq {

enum Kind : int { kBlockFP = 0, kBlockMinifloat = 1, kBlockLog = 2, kMinifloatDenorm = 3, kMinifloatIEEE = 4,
                  kInteger = 5, kNone = 6 };

// Host-prepared scalars of one operand format (all small integers, exact in fp32).
struct FmtParams {
  int kind;
  int fold_zero;     // +0.0f after the element math (col2im accumulation of F.fold)
  float emin, emax;  // block_fp / minifloat_*: exponent clamp (scalars). integer: int_min / int_max
  float shift;       // 2^mantissa_bits (integer: 2^frac_width)
  float inv_shift;   // exact reciprocal of shift (q / shift == q * inv_shift bit for bit)
  float qmax;        // 2^mantissa_bits - 1
  float bias_hi;     // block_minifloat / block_log: 2^exponent_bias_width - 1 (upper clamp of the shared bias)
  float eb_top;      // block_minifloat: 2^exponent_width - 1 ; block_log: 2^(width-1) - 1
  int mbits;         // mantissa bits (log2 of shift)
  int fast_fmt;      // host: the format's static ranges allow the fast path (see fast_state)
  // integer copies of emin / emax / bias_hi / eb_top (valid when fast_fmt): fast_state() runs once per block, and a float -> int
  // conversion there is an XU-pipe instruction with ~20 cycles of latency in front of a dependent chain (3 % of the attention
  // kernel's stall samples, ncu source view)
  int emin_i, emax_i, bias_hi_i, eb_top_i;
};

__device__ __forceinline__ float clamp_t(float x, float lo, float hi) {   // torch.clamp: NaN in -> NaN out
  return (x != x) ? x : fminf(fmaxf(x, lo), hi);
}
__device__ __forceinline__ float sign_t(float w) { return (float)((w > 0.f) - (w < 0.f)); }  // torch.sign, NaN -> 0

// 2**e for an integer-valued float e (torch: pow(2.0f, e))
__device__ __forceinline__ float pow2_t(float ef) {
  if (ef != ef) return ef;
  if (ef >= 128.f) return __int_as_float(0x7f800000);
  if (ef < -149.5f) return 0.f;          // 2^-150 ties-to-even -> 0
  int e = (int)ef;
  return e >= -126 ? __int_as_float((e + 127) << 23) : __int_as_float(1 << (e + 149));
}
// v / 2**e with one correctly rounded result
__device__ __forceinline__ float div_pow2(float v, float ef, float p2e) {
  if (ef >= -126.f && ef <= 126.f) return __fmul_rn(v, __int_as_float((127 - (int)ef) << 23));
  return __fdiv_rn(v, p2e);
}
__device__ __forceinline__ bool isclose_t(float a, float b) {   // torch.isclose(a, b), rtol 1e-5, atol 1e-8, fp32
  if (a == b) return true;
  float err = fabsf(__fsub_rn(a, b));
  float allowed = __fadd_rn(1e-8f, fabsf(__fmul_rn(1e-5f, b)));
  return (err <= allowed) && (err - err == 0.f);   // isfinite(err)
}
__device__ __forceinline__ float blend_t(bool c, float q, float x) {   // (~c)*q + c*x
  return __fadd_rn(__fmul_rn(c ? 0.f : 1.f, q), __fmul_rn(c ? 1.f : 0.f, x));
}

// ------------------------------------------------------------------------------------------
// Per-block state (from the block's max |x| after the reference's zero substitution)
// ------------------------------------------------------------------------------------------
struct BlockState {
  float a;   // block_fp: 2^E          block_minifloat: emin (= -b)     block_log: emin
  float b;   // block_fp: E            block_minifloat: emax            block_log: emax
  float c;   //                                                          block_log: delta = 2^emin * 0.1
};

template <int KIND>
__device__ __forceinline__ BlockState block_state(float mx, const FmtParams& p) {
  BlockState s;
  s.a = s.b = s.c = 0.f;
  if (KIND == kBlockFP) {                                  // block_fp.py:72-73
    float e = clamp_t(ceilf(log2f(mx)), p.emin, p.emax);
    s.b = e;
    s.a = pow2_t(e);
  } else if (KIND == kBlockMinifloat) {                    // block_minifloat.py:57-59, minifloat.py:164-165
    float b = clamp_t(floorf(log2f(mx)), 0.f, p.bias_hi);
    s.a = -b;
    s.b = __fsub_rn(p.eb_top, b);
  } else if (KIND == kBlockLog) {                          // block_log.py:55-58, log.py:47-52
    float me = ceilf(log2f(mx));
    float b = clamp_t(__fsub_rn(p.eb_top, me), 0.f, p.bias_hi);
    s.a = -b;
    s.b = __fsub_rn(p.eb_top, b);
    s.c = __fmul_rn(pow2_t(s.a), 0.1f);
  }
  return s;
}

// ------------------------------------------------------------------------------------------
// Element math
// ------------------------------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ float quant_elem(float x, const BlockState& s, const FmtParams& p) {
  float out;
  if (KIND == kBlockFP) {                                  // block_fp.py:69-94
    float sg = sign_t(__fadd_rn(x, 1e-9f));
    float v = __fadd_rn(fabsf(x), 1e-9f);
    float t = div_pow2(v, s.b, s.a);
    float q = clamp_t(rintf(__fmul_rn(t, p.shift)), 0.f, p.qmax);
    float mant = __fmul_rn(q, p.inv_shift);
    float y = __fmul_rn(__fmul_rn(sg, s.a), mant);
    if (p.fold_zero) y = __fadd_rn(y, 0.f);
    return blend_t(fabsf(x) <= 1e-8f, y, x);
  } else if (KIND == kBlockMinifloat || KIND == kMinifloatIEEE) {   // minifloat.py:172-194
    float emin = (KIND == kBlockMinifloat) ? s.a : p.emin;
    float emax = (KIND == kBlockMinifloat) ? s.b : p.emax;
    float sg = sign_t(__fadd_rn(x, 1e-9f));
    float v = fabsf(x);
    float e = clamp_t(floorf(log2f(__fadd_rn(v, 1e-9f))), emin, emax);
    float p2e = pow2_t(e);
    float t = div_pow2(v, e, p2e);
    bool normal = !isclose_t(e, emin);
    float ts = __fmul_rn(t, p.shift);
    float qa = clamp_t(rintf(__fsub_rn(ts, p.shift)), 0.f, p.qmax);
    float qb = clamp_t(rintf(__fmul_rn(ts, 0.5f)), 0.f, p.qmax);
    float nf = normal ? 1.f : 0.f, sf = normal ? 0.f : 1.f;
    float sm = __fadd_rn(__fmul_rn(nf, qa), __fmul_rn(sf, qb));
    float frac = __fmul_rn(sm, p.inv_shift);
    float mant = __fadd_rn(__fmul_rn(nf, __fadd_rn(1.0f, frac)), __fmul_rn(sf, __fmul_rn(frac, 2.f)));
    float y = __fmul_rn(__fmul_rn(sg, p2e), mant);
    out = blend_t(v <= 1e-8f, y, x);
  } else if (KIND == kBlockLog) {                          // log.py:51-56
    float sg = sign_t(__fadd_rn(x, s.c));
    float v = __fadd_rn(fabsf(x), s.c);
    float e = clamp_t(rintf(log2f(v)), s.a, s.b);
    out = __fmul_rn(sg, pow2_t(e));
  } else if (KIND == kMinifloatDenorm) {                   // minifloat.py:60-80
    float sg = sign_t(__fadd_rn(x, 1e-9f));
    float v = fabsf(x);
    float e = clamp_t(ceilf(log2f(__fadd_rn(v, 1e-9f))), p.emin, p.emax);
    float p2e = pow2_t(e);
    float t = div_pow2(v, e, p2e);
    float q = clamp_t(rintf(__fmul_rn(t, p.shift)), 0.f, p.qmax);
    float mant = __fmul_rn(q, p.inv_shift);
    float y = __fmul_rn(__fmul_rn(sg, p2e), mant);
    out = blend_t(v <= 1e-8f, y, x);
  } else if (KIND == kInteger) {                           // integer.py:52
    out = __fmul_rn(clamp_t(rintf(__fmul_rn(x, p.shift)), p.emin, p.emax), p.inv_shift);
  } else {
    out = x;
  }
  if (p.fold_zero) out = __fadd_rn(out, 0.f);
  return out;
}

// ------------------------------------------------------------------------------------------
// Fast path.  Same results bit for bit, fewer instructions:
//   * ceil/floor/rint(log2f(x)) come from the exponent field unless the mantissa sits in a narrow zone
//     around the rounding cliff (2^-11 of all values), where the libdevice log2f is evaluated as before.
//     tests/test_gpu_numerics.py checks the three helpers against libdevice on EVERY positive fp32 pattern.
//   * powers of two are built from integer exponents; x / 2^e is a multiplication by 2^-e; the
//     bool*float blends collapse to selects.  All of this is valid only while every intermediate is a
//     finite normal number — FastState::ok says so per block, otherwise the literal path above runs.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kZone = 0x200u;           // 2^9 ulps around a cliff (needed: ~0.7 |k| + a few ulps of log2f error, |k| <= 128)
constexpr uint32_t kSqrt2Mant = 0x3504f3u;   // mantissa field of sqrt(2)

// cold fall-backs (inputs within 2^-11 of a rounding cliff, denormals): kept out of line — every inlined copy of the
// libdevice log2f polynomial costs ~40 instructions of code, and the fused kernels instantiate these helpers dozens of
// times per loop body (instruction-cache footprint, see DESIGN.md "code size")
static __device__ __noinline__ int ceil_log2_slow(float x) { return (int)ceilf(log2f(x)); }
static __device__ __noinline__ int floor_log2_slow(float x) { return (int)floorf(log2f(x)); }
static __device__ __noinline__ int rint_log2_slow(float x) { return (int)rintf(log2f(x)); }

// One-sided cliffs (checked exhaustively by bq_selftest_log2): log2f(x) can only land ON an integer k by rounding when x sits
// just above 2^k (result k instead of k+tiny: matters for ceil) or just below it (k instead of k-tiny: matters for floor).
__device__ __forceinline__ int ceil_log2_i(float x) {        // == (int)ceilf(log2f(x)) for x > 0 finite
  const uint32_t b = __float_as_uint(x), ex = b >> 23;
  if (ex - 1u >= 254u || (b & 0x7fffffu) < kZone) return ceil_log2_slow(x);
  return (int)ex - 126;
}
__device__ __forceinline__ int floor_log2_i(float x) {       // == (int)floorf(log2f(x))
  const uint32_t b = __float_as_uint(x), ex = b >> 23;
  if (ex - 1u >= 254u || (b | 0xff800000u) >= 0u - kZone) return floor_log2_slow(x);
  return (int)ex - 127;
}
__device__ __forceinline__ int rint_log2_i(float x) {        // == (int)rintf(log2f(x))
  const uint32_t b = __float_as_uint(x), ex = b >> 23, f = b & 0x7fffffu;
  if (ex - 1u >= 254u || (f - kSqrt2Mant + kZone) < 2 * kZone) return rint_log2_slow(x);
  return (int)ex - 127 + (f > kSqrt2Mant ? 1 : 0);
}
// ---- per-element variants for the fast paths.  NC = false: checked (cliff zone -> libdevice, out of line).  NC = true ("no check"):
// always the shortcut, and the element's distance from the cliff is min-accumulated into `zacc`; the caller tests
// zone_hit(zacc) ONCE per block and redoes the block through the checked path if any element was close (rare: 16 * 2^-14).
// All return the BIASED value (+127) and require a finite argument.
template <int KIND> __device__ __forceinline__ constexpr uint32_t zone_thr() { return KIND == kBlockLog ? 2 * kZone : kZone; }
template <int KIND> __device__ __forceinline__ bool zone_hit(uint32_t zacc) { return zacc < zone_thr<KIND>(); }
template <bool NC> __device__ __forceinline__ int floor_log2_biased_nf(float x, uint32_t& zacc) {   // normal x
  const uint32_t b = __float_as_uint(x), key = ~b & 0x7fffffu;       // ulps below the next power of two, minus one
  if (NC) { zacc = min(zacc, key); return (int)(b >> 23); }
  if (key < kZone) return floor_log2_slow(x) + 127;
  return (int)(b >> 23);
}
template <bool NC> __device__ __forceinline__ int ceil_log2_biased_nf(float x, uint32_t& zacc) {    // normal x
  const uint32_t b = __float_as_uint(x), key = b & 0x7fffffu;        // ulps above the power of two
  if (NC) { zacc = min(zacc, key); return (int)(b >> 23) + 1; }
  if (key < kZone) return ceil_log2_slow(x) + 127;
  return (int)(b >> 23) + 1;
}
// denormal x returns <= 1 (callers clamp from below).  A mantissa above sqrt(2)'s carries into the exponent.
template <bool NC> __device__ __forceinline__ int rint_log2_biased_f(float x, uint32_t& zacc) {
  const uint32_t b = __float_as_uint(x), key = (b & 0x7fffffu) - kSqrt2Mant + kZone;
  if (NC) zacc = min(zacc, key);
  else if (key < 2 * kZone) return (b >> 23) ? rint_log2_slow(x) + 127 : 0;
  return (int)((b + (0x7fffffu - kSqrt2Mant)) >> 23);
}
// rint(log2f(x)) + 127 for ANY finite x >= 0 of a block whose maximum is below 2^100, denormals included: the argument is scaled by
// 2^24 (exact), which turns every denormal into a normal number with the same significand; the shortcut then reads the exponent field
// and the sqrt(2) comparison as before.  Near the cliff (NC = false) libdevice's log2f is evaluated on the ORIGINAL value, like the
// reference does.  x == 0 returns -24 + 0 (callers clamp from below and select on the sign word).
template <bool NC> __device__ __forceinline__ int rint_log2_biased_deep(float x, uint32_t& zacc) {
  const uint32_t b = __float_as_uint(__fmul_rn(x, 16777216.0f)), key = (b & 0x7fffffu) - kSqrt2Mant + kZone;
  if (NC) zacc = min(zacc, key);
  else if (key < 2 * kZone) return rint_log2_slow(x) + 127;
  return (int)((b + (0x7fffffu - kSqrt2Mant)) >> 23) - 24;
}
__device__ __forceinline__ float pow2_i(int e) { return __int_as_float((e + 127) << 23); }   // e in [-126, 127]
// rintf(t) for |t| < 2^22 as two full-rate FADDs (FRND runs on the quarter-rate conversion pipe): adding 1.5 * 2^23 moves t into
// a binade whose ulp is 1, the addition itself rounds to nearest-even, the subtraction is exact.
constexpr float kRintMagic = 12582912.0f;
__device__ __forceinline__ float rint_small(float t) { return __fsub_rn(__fadd_rn(t, kRintMagic), kRintMagic); }

struct FastState {
  bool ok;
  bool deep;      // block_log: emin < -126 — outputs and the log2 argument reach into the denormal range (block maxima <= 1 at width 8)
  float f0, f1;   // block_fp: scale 2^(m-E), step 2^(E-m)            block_log: delta, 2^emin
  int i0, i1, i2; // block_log: emin, emax (integers); minifloat: biased step-exponent clamp [i0, i1], normal threshold i2
  float c0, c1, hi;   // block_fp: 1e-9f * f0, -kRintMagic * f1, kRintMagic + qmax  (all exact)
};

// host-known part of the validity check lives in FmtParams::fast_fmt (mantissa width / scalar exponent range)
template <int KIND>
__device__ __forceinline__ FastState fast_state(uint32_t mbits, const FmtParams& p) {
  FastState s;
  s.ok = false;
  s.deep = false;
  s.f0 = s.f1 = 0.f;
  s.i0 = s.i1 = s.i2 = 0;
  s.c0 = s.c1 = s.hi = 0.f;
  if (!p.fast_fmt || mbits >= 0x7f800000u) return s;           // inf / NaN block max -> literal path
  const float mx = __uint_as_float(mbits);
  if (KIND == kBlockFP) {
    int E = ceil_log2_i(mx);
    E = min(max(E, p.emin_i), p.emax_i);
    if (E < -100 || E > 100) return s;
    s.f0 = pow2_i(p.mbits - E);
    s.f1 = pow2_i(E - p.mbits);
    s.c0 = __fmul_rn(1e-9f, s.f0);
    s.c1 = -__fmul_rn(kRintMagic, s.f1);
    s.hi = __fadd_rn(kRintMagic, p.qmax);
    s.ok = true;
  } else if (KIND == kBlockMinifloat || KIND == kMinifloatIEEE) {
    int emin, emax;
    if (KIND == kBlockMinifloat) {
      int b = floor_log2_i(mx);
      b = min(max(b, 0), p.bias_hi_i);
      emin = -b;
      emax = p.eb_top_i - b;
    } else {
      emin = p.emin_i;
      emax = p.emax_i;
    }
    // see quant_elem_fast: biased clamp range of the step exponent, "normal" threshold, clamp bounds in the shifted domain
    s.i0 = emin + 1 + 127;
    s.i1 = emax + 127;
    s.i2 = (emax > emin) ? emin + 127 : 0x7fffffff;
    s.f0 = __fadd_rn(kRintMagic, p.shift);
    s.f1 = __fadd_rn(s.f0, p.qmax);
    s.hi = __fadd_rn(kRintMagic, p.qmax);
    s.ok = (emax >= emin) && (s.i0 - p.mbits >= 1) && (max(s.i1, s.i0) <= 253);
  } else if (KIND == kBlockLog) {
    int b = p.eb_top_i - ceil_log2_i(mx);
    b = min(max(b, 0), p.bias_hi_i);
    s.i0 = -b;
    s.i1 = p.eb_top_i - b;
    // deep: 2^emin is a denormal (or 0 below 2^-149) and so may be |x| + delta and the output; handled by scaling the log2 argument
    // by 2^24 (exact; needs the block maximum below 2^100) and building denormal results from integers — softmax probabilities
    // (block maxima < 1) used to take the literal path here: 1.98 TB/s against 4.7 for maxima above 1
    s.deep = s.i0 < -126;
    s.ok = (s.i1 <= 127 && s.i1 >= s.i0 && (!s.deep || (s.i0 >= -400 && mbits < ((100u + 127u) << 23))));
    s.f1 = s.deep ? pow2_t((float)s.i0) : pow2_i(s.i0);
    s.f0 = __fmul_rn(s.f1, 0.1f);
  } else if (KIND == kMinifloatDenorm) {
    s.i0 = p.emin_i + 127;                                    // format-level range check: FmtParams::fast_fmt
    s.i1 = p.emax_i + 127;
    s.hi = __fadd_rn(kRintMagic, p.qmax);
    s.ok = true;
  } else {
    s.ok = true;
  }
  return s;
}

template <int KIND, bool NC>
__device__ __forceinline__ float quant_elem_fast_impl(float x, const FastState& s, const FmtParams& p, uint32_t& zacc) {
  const float ax = fabsf(x);
  float out;
  if (KIND == kBlockFP) {
    // (|x| + 1e-9f) * 2^(m-E) == fma(|x|, f0, 1e-9f * f0): scaling by a power of two commutes with rounding (no under/overflow:
    // |m - E| <= 122).  Rounding and clamping happen in the magic-shifted domain; (tm - magic) * f1 is exact, so is the fma.
    const float t = __fmaf_rn(ax, s.f0, s.c0);
    const float tm = fminf(__fadd_rn(t, kRintMagic), s.hi);
    float y = copysignf(__fmaf_rn(tm, s.f1, s.c1), x);
    if (p.fold_zero) y = __fadd_rn(y, 0.f);
    return (ax <= 1e-8f) ? __fadd_rn(x, 0.f) : y;
  } else if (KIND == kBlockMinifloat || KIND == kMinifloatIEEE) {
    // minifloat.py:172-194 collapsed.  With e = clamp(floor(log2(|x|+1e-9)), emin, emax):
    //   normal    (e > emin):  y = 2^(e-M)      * clamp(rint(|x| * 2^(M-e)),      2^M, 2^(M+1)-1)   [rint(ts - 2^M) == rint(ts) - 2^M: the
    //                                                                                 subtraction is exact wherever the clamp does not decide]
    //   subnormal (e == emin): y = 2^(emin+1-M) * clamp(rint(|x| * 2^(M-emin-1)), 0,   2^M-1)
    // i.e. one step exponent es = max(e, emin+1), rounded and clamped in the magic-shifted domain; every scaling is by a power of two.
    const int eb = floor_log2_biased_nf<NC>(__fadd_rn(ax, 1e-9f), zacc);
    const bool normal = eb > s.i2;
    const int es = max(min(eb, s.i1), s.i0);
    const float inv_step = __int_as_float((254 + p.mbits - es) << 23), step = __int_as_float((es - p.mbits) << 23);
    const float tm = fminf(fmaxf(__fadd_rn(__fmul_rn(ax, inv_step), kRintMagic), normal ? s.f0 : kRintMagic), normal ? s.f1 : s.hi);
    const float y = copysignf(__fmul_rn(__fsub_rn(tm, kRintMagic), step), x);
    out = (ax <= 1e-8f) ? __fadd_rn(x, 0.f) : y;
  } else if (KIND == kBlockLog) {
    const float w = __fadd_rn(x, s.f0);
    const float v = __fadd_rn(ax, s.f0);
    // v is finite (block max is) and every v below 2^emin — zeros (v = delta), denormals — clamps to emin whatever the shortcut
    // returns for it (its biased result is <= 1 there), so neither an exponent-range test nor a v < 2^emin select is needed.
    if (!s.deep) {
      const int eb = min(max(rint_log2_biased_f<NC>(v, zacc), s.i0 + 127), s.i1 + 127);
      out = (w == 0.f) ? 0.f : copysignf(__int_as_float(eb << 23), w);
    } else {
      const int eb = min(max(rint_log2_biased_deep<NC>(v, zacc), s.i0 + 127), s.i1 + 127);      // biased exponent, may be <= 0
      // 2^(eb - 127): normal for eb >= 1, the denormal 1 << (eb + 22) down to eb = -22 (2^-149), 0 below (torch: pow(2, e) underflows)
      const uint32_t bits = eb >= 1 ? (uint32_t)eb << 23 : (eb >= -22 ? 1u << (eb + 22) : 0u);
      out = (w == 0.f) ? 0.f : copysignf(__uint_as_float(bits), w);
    }
  } else if (KIND == kMinifloatDenorm) {
    // minifloat.py:60-80 with e = clamp(ceil(log2(|x|+1e-9)), emin, emax):  y = 2^(e-M) * min(rint(|x| * 2^(M-e)), 2^M - 1)
    const int eb = min(max(ceil_log2_biased_nf<NC>(__fadd_rn(ax, 1e-9f), zacc), s.i0), s.i1);
    const float inv_step = __int_as_float((254 + p.mbits - eb) << 23), step = __int_as_float((eb - p.mbits) << 23);
    const float tm = fminf(__fadd_rn(__fmul_rn(ax, inv_step), kRintMagic), s.hi);
    const float y = copysignf(__fmul_rn(__fsub_rn(tm, kRintMagic), step), x);
    out = (ax <= 1e-8f) ? __fadd_rn(x, 0.f) : y;
  } else {
    out = x;
  }
  if (p.fold_zero) out = __fadd_rn(out, 0.f);
  return out;
}

template <int KIND>
__device__ __forceinline__ float quant_elem_fast(float x, const FastState& s, const FmtParams& p) {
  uint32_t unused = 0xffffffffu;
  return quant_elem_fast_impl<KIND, false>(x, s, p, unused);
}

template <int KIND> struct IsBlocked { static constexpr bool value = (KIND == kBlockFP || KIND == kBlockMinifloat || KIND == kBlockLog); };

}  // namespace bq
