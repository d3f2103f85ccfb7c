/*
 * eaz_math.h -- the fp32 arithmetic contract of the E-MCTS hot path.
 *
 * Every transcendental the search / network path needs (exp for the softmaxes
 * of emctx's action selection, tanh for the value and UBE heads of
 * /root/reference/src/network/fully_connected.py:54,63) is written here as an
 * explicit sequence of IEEE-754 binary32 operations (add, mul, fma, div, sqrt,
 * round-to-nearest-even), so that a host C compiler (gcc -ffp-contract=off) and
 * nvcc (-fmad=false, default -prec-div/-prec-sqrt/-ftz=false) produce the same
 * bits.  That is what lets tests demand bit-exact actions, visit counts and
 * tree topology between the CUDA path and the CPU oracle even though argmax
 * flips on 1-ulp differences (SURVEY.md section 7 "Hard parts").
 *
 * Neither libm's expf nor CUDA's expf is used anywhere on the path.
 * The header is plain C99 and is also valid CUDA device code.
 */
#ifndef EAZ_MATH_H_
#define EAZ_MATH_H_

#include <stdint.h>

#if defined(__CUDACC__)
#define EAZ_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#include <string.h>
#define EAZ_HD static inline
#endif

/* jnp.finfo(jnp.float32).min / .tiny, used by mask_invalid_actions
 * (/root/reference/src/reanalyze.py:16-29) and _compute_mixed_value. */
#define EAZ_F32_MIN (-3.4028234663852886e38f)
#define EAZ_F32_TINY (1.1754943508222875e-38f)

EAZ_HD float eaz_bits_to_f32(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

EAZ_HD uint32_t eaz_f32_to_bits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
#endif
}

/* Single-rounding fused multiply-add on both sides. */
EAZ_HD float eaz_fma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return fmaf(a, b, c);
#endif
}

/* Separately rounded mul / add (never contracted into an fma). */
EAZ_HD float eaz_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
EAZ_HD float eaz_add(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
EAZ_HD float eaz_sub(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}
EAZ_HD float eaz_div(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
EAZ_HD float eaz_sqrt(float a) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(a);
#else
  return sqrtf(a);
#endif
}
EAZ_HD float eaz_max(float a, float b) { return a > b ? a : (b > a ? b : (a == a ? a : b)); }
EAZ_HD float eaz_min(float a, float b) { return a < b ? a : (b < a ? b : (a == a ? a : b)); }

/* exp(x), <= 1 ulp-class accuracy (Cephes-style degree-5 kernel after
 * Cody-Waite reduction).  exp(-inf) = exp(EAZ_F32_MIN - m) = 0. */
EAZ_HD float eaz_exp(float x) {
  if (!(x > -104.0f)) return (x != x) ? x : 0.0f;
  if (x > 88.72283f) return eaz_bits_to_f32(0x7f800000u);
  const float kf =
#if defined(__CUDA_ARCH__)
      rintf(__fmul_rn(x, 1.44269504088896341f));
#else
      rintf(x * 1.44269504088896341f);
#endif
  float r = eaz_fma(kf, -0.693359375f, x);
  r = eaz_fma(kf, 2.12194440e-4f, r);
  float p = 1.9875691500e-4f;
  p = eaz_fma(p, r, 1.3981999507e-3f);
  p = eaz_fma(p, r, 8.3334519073e-3f);
  p = eaz_fma(p, r, 4.1665795894e-2f);
  p = eaz_fma(p, r, 1.6666665459e-1f);
  p = eaz_fma(p, r, 5.0000001201e-1f);
  const float r2 = eaz_mul(r, r);
  p = eaz_fma(p, r2, r);
  p = eaz_add(p, 1.0f);
  int k = (int)kf;
  if (k > 127) { /* only k == 128 reaches here */
    p = eaz_mul(p, 2.0f);
    k -= 1;
  }
  if (k >= -126) return eaz_mul(p, eaz_bits_to_f32((uint32_t)(k + 127) << 23));
  /* subnormal result: two exact power-of-two scalings, second one rounds */
  const float s = eaz_bits_to_f32((uint32_t)(k + 100 + 127) << 23);
  return eaz_mul(eaz_mul(p, s), eaz_bits_to_f32((uint32_t)(127 - 100) << 23));
}

/* tanh(x): odd polynomial for |x| < 0.625, 1 - 2/(exp(2|x|)+1) beyond. */
EAZ_HD float eaz_tanh(float x) {
  const float ax = x < 0.0f ? -x : x;
  if (!(ax == ax)) return x;
  if (ax < 0.625f) {
    const float z = eaz_mul(x, x);
    float p = -5.70498872745e-3f;
    p = eaz_fma(p, z, 2.06390887954e-2f);
    p = eaz_fma(p, z, -5.37397155531e-2f);
    p = eaz_fma(p, z, 1.33314422036e-1f);
    p = eaz_fma(p, z, -3.33332819422e-1f);
    p = eaz_mul(p, z);
    return eaz_fma(p, x, x);
  }
  float t;
  if (ax > 9.02f) {
    t = 1.0f;
  } else {
    const float e = eaz_exp(eaz_add(ax, ax));
    t = eaz_sub(1.0f, eaz_div(2.0f, eaz_add(e, 1.0f)));
  }
  return x < 0.0f ? -t : t;
}

/* ln(x) for finite x > 0 (normal or subnormal), <= 1 ulp-class accuracy: x = m * 2^e with m in [sqrt(1/2), sqrt(2)),
 * Cephes-style degree-8 kernel in f = m - 1.  ln(0) = -inf, ln(x < 0) = NaN, ln(inf) = inf.  Used by the PUCT exploration
 * constant (mctx action_selection.muzero_action_selection) and the visit-count logits of muzero_policy. */
EAZ_HD float eaz_log(float x) {
  if (x != x || x < 0.0f) return eaz_bits_to_f32(0x7fc00000u);
  if (x == 0.0f) return eaz_bits_to_f32(0xff800000u);
  uint32_t u = eaz_f32_to_bits(x);
  if (u == 0x7f800000u) return x;
  int e = 0;
  if (u < 0x00800000u) { /* subnormal: scale by 2^23 (exact) */
    x = eaz_mul(x, 8388608.0f);
    u = eaz_f32_to_bits(x);
    e = -23;
  }
  e += (int)(u >> 23) - 127;
  float m = eaz_bits_to_f32((u & 0x007fffffu) | 0x3f800000u); /* [1, 2) */
  if (m > 1.41421356237f) {
    m = eaz_mul(m, 0.5f);
    e += 1;
  }
  const float f = eaz_sub(m, 1.0f);
  const float z = eaz_mul(f, f);
  float p = 7.0376836292e-2f;
  p = eaz_fma(p, f, -1.1514610310e-1f);
  p = eaz_fma(p, f, 1.1676998740e-1f);
  p = eaz_fma(p, f, -1.2420140846e-1f);
  p = eaz_fma(p, f, 1.4249322787e-1f);
  p = eaz_fma(p, f, -1.6668057665e-1f);
  p = eaz_fma(p, f, 2.0000714765e-1f);
  p = eaz_fma(p, f, -2.4999993993e-1f);
  p = eaz_fma(p, f, 3.3333331174e-1f);
  float y = eaz_mul(eaz_mul(f, z), p);
  const float fe = (float)e;
  y = eaz_fma(fe, -2.12194440e-4f, y);
  y = eaz_fma(-0.5f, z, y);
  y = eaz_add(f, y);
  return eaz_fma(fe, 0.693359375f, y);
}

/* softplus(x) = log(1 + exp(x)) = max(x, 0) + log(1 + exp(-|x|)) (jax.nn.softplus = logaddexp(x, 0); the UBE head of
 * /root/reference/src/network/minatar.py:91).  Absolute error <= 1e-7-class: the sum 1 + t rounds for tiny t. */
EAZ_HD float eaz_softplus(float x) {
  if (x != x) return x;
  const float ax = x < 0.0f ? -x : x;
  const float t = eaz_exp(-ax);
  return eaz_add(eaz_max(x, 0.0f), eaz_log(eaz_add(1.0f, t)));
}

#endif /* EAZ_MATH_H_ */
