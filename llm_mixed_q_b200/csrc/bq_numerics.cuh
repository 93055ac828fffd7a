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

template <int KIND> struct IsBlocked { static constexpr bool value = (KIND == kBlockFP || KIND == kBlockMinifloat || KIND == kBlockLog); };

}  // namespace bq
