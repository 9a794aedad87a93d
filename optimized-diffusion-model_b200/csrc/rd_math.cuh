// rd_math.cuh -- scalar building blocks shared by the element-wise kernels.
// Every function restates a piece of the reference's arithmetic (cited per function) in the
// same fp32 operation order, with explicit round-to-nearest intrinsics so the compiler cannot
// contract mul+add pairs the reference evaluates as two rounded ATen ops.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rd {

// cube.reflect (reference cube.py:47-49):  m = x % 2 (torch floor-mod) ; m > 1 -> 2 - m.
// torch.remainder(x, 2) == fmod(x,2) (+2 when negative and non-zero).  x - 2*floor(x/2) evaluated
// with a single rounding (fma) is the same real number rounded once, hence bit-identical, except
// for corner cases handled explicitly: x == +-0 and exact multiples of 2 (fmod keeps the sign of the
// dividend on a zero result: reflect(-2) = -0.0) and negative
// denormals whose half rounds to -0 (floor must still be -1).  NaN/Inf -> NaN like torch.
__device__ __forceinline__ float reflect1(float x) {
  if (x == 0.0f) return x;
  float k = floorf(x * 0.5f);
  if (x < 0.0f && k == 0.0f) k = -1.0f;
  float m = __fmaf_rn(-2.0f, k, x);
  if (m == 0.0f) m = copysignf(0.0f, x);  // fmod keeps the dividend's sign on exact multiples of 2
  return (m > 1.0f) ? (2.0f - m) : m;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator (Salmon et al., SC'11) -- stands in for torch.randn_like
// (reference sampling.py:200,224) in production mode.  The stream is OURS (it does not reproduce
// torch's offsets); parity runs inject a noise tape instead, or dump this stream with
// rd_philox_normal_f32 and replay it through the oracle.
struct Philox4 {
  uint32_t c[4];
};
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  Philox4 o;
  o.c[0] = c0; o.c[1] = c1; o.c[2] = c2; o.c[3] = c3;
  return o;
}

// Four N(0,1) draws for the 4-element group `quad` of noise tensor number `draw` under `seed`.
// Box-Muller on (0,1] x (-pi,pi] uniforms evaluated on the special-function unit (lg2 / sqrt / sin / cos
// .approx: absolute error ~2^-21 on the angle functions, ~2e-7 on -2 ln u): the stream is OURS, every consumer
// (fused step kernels, rd_philox_normal_f32 dump) calls this one function, and the accurate libm versions cost
// three times the instructions of the whole update they feed.  |z| <= 6.76 (u1 >= 2^-33).
__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint32_t draw, uint64_t quad, float (&z)[4]) {
  Philox4 r = philox4x32_10(static_cast<uint32_t>(quad), static_cast<uint32_t>(quad >> 32), draw, 0x5eedu,
                            static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  const float S = 2.3283064365386963e-10f;  // 2^-32
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    float u1 = fminf((static_cast<float>(r.c[2 * p]) + 0.5f) * S, 1.0f);                 // (0,1]
    const float ang = (static_cast<float>(r.c[2 * p + 1]) + 0.5f) * (S * 6.283185307179586f) - 3.141592653589793f;
    const float rad = sqrt_approx(fmaxf(-1.3862943611198906f * __log2f(u1), 0.0f));     // sqrt(-2 ln u1)
    z[2 * p] = rad * __cosf(ang);
    z[2 * p + 1] = rad * __sinf(ang);
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace rd
