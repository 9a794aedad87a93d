// elementwise.cu -- HBM-bound kernels of the sampling hot path:
//   cube.reflect / cube.inside / cube.score_hk, the CFG combine, the Philox noise dump and the
//   fused Langevin-corrector / Euler-Maruyama-predictor updates (reflection fused in).
// All are coalesced 128-bit streaming kernels; reductions use warp shuffles + a fixed-order
// second-stage sum (no atomics => run-to-run deterministic).
#include "rd_common.h"
#include "rd_math.cuh"

namespace rd {

// ------------------------------------------------------------------------------------------------
// error plumbing
static thread_local char g_err[512] = "";
char* err_buf() { return g_err; }
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

__device__ __forceinline__ float4 ld_stream4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st_stream4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

// ------------------------------------------------------------------------------------------------
// cube.reflect : 8 B / element
__global__ void __launch_bounds__(256) reflect_kernel(const float* __restrict__ x, float* __restrict__ out, size_t n) {
  const size_t n4 = n >> 2;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  // 4 independent 128-bit loads in flight per thread
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = ld_stream4(x + 4 * (i + u * stride));
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v[u].x = reflect1(v[u].x); v[u].y = reflect1(v[u].y); v[u].z = reflect1(v[u].z); v[u].w = reflect1(v[u].w);
      st_stream4(out + 4 * (i + u * stride), v[u]);
    }
  }
  for (; i < n4; i += stride) {
    float4 v = ld_stream4(x + 4 * i);
    v.x = reflect1(v.x); v.y = reflect1(v.y); v.z = reflect1(v.z); v.w = reflect1(v.w);
    st_stream4(out + 4 * i, v);
  }
  // tail (< 4 elements)
  size_t t = (n4 << 2) + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t < n) out[t] = reflect1(x[t]);
}

// cube.inside (cube.py:17-31): one warp per sample
__global__ void __launch_bounds__(256) inside_kernel(const float* __restrict__ x, uint8_t* __restrict__ ok, size_t B,
                                                     size_t D) {
  size_t w = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (w >= B) return;
  bool good = true;
  for (size_t j = lane; j < D; j += 32) {
    float v = x[w * D + j];
    good = good && (v >= 0.0f) && (v <= 1.0f);  // NaN fails both, like torch
  }
  good = __all_sync(0xffffffffu, good);
  if (lane == 0) ok[w] = good ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// cube.score_hk (cube.py:149-193): score of the heat kernel on [0,1] with reflecting walls.  Per sample
// t = sigma^2/2.  The reference evaluates  t > min_cutoff  with the eigenfunction series (_score_hk_ef,
// cube.py:73-107, `efs` cosine modes) and the rest with the method of images (_score_hk_refl, cube.py:110-146,
// images 2m+x and 2m-x for |m| <= refls).  Both are the SAME function (Poisson summation); the reference's
// fp32 eigen-series loses up to 1e-5 of the score range just above the cutoff because its denominator
// 1 + 2 sum(...) cancels, while the image sum has only positive terms.  This kernel therefore
//   * truncates both sums to the terms that can still change an fp32 result (e^-17.5 relative):
//       modes  Kc(t) = ceil(sqrt(17.5 / (pi^2 t))) + 1,   images |m| <= Mc(t) = ceil(sqrt((70 t + 1) / 4))
//   * uses the image sum wherever it is the cheaper converged form (t <= 0.214, at most 10 images), also above
//     the reference's cutoff -- but only when the reference's own series is converged there (efs >= Kc); a
//     deliberately short series (efs < Kc) is reproduced term by term;
//   * matches the reference's epsilons: score = num / (den + 1e-12) in the units of whichever form the
//     REFERENCE would have used (the image sum is sqrt(4 pi t) times the eigen-series density).
// Arithmetic: exponentials on the SFU (ex2.approx, 2 ulp), the common factor -2/(4t) applied once, harmonics
// by the rotation recurrence from one accurate sincospi per coordinate.  Measured error against the fp64
// reference is at or below the reference's own fp32 error everywhere (tests/test_gpu_elementwise.py).
// Layout: one block owns HK_SPB whole samples and visits them in an order sorted by (form, term count), so
// that the threads of a warp run the same loop for the same number of iterations even when sigma differs
// from sample to sample.
constexpr int HK_SPB = 32;      // samples per block
constexpr int HK_MAX_EFS = 64;  // table capacity per sample
constexpr float PI_F = 3.14159265358979323846f;
constexpr float PI2_F = 9.869604401089358f;  // python float pi**2 -> fp32 scalar
constexpr float HK_T_IMAGES = 0.214f;        // above this the eigen-series (<= 5 modes) is the cheaper form
// Below this t the far image -2 - x (|d| >= 2) is under e^-17.5 of the sum: den >= exp(-1/4t) (the image x itself has
// |d| <= 1) against exp(-4/4t).  It covers every t for which one image pair per side suffices (Mc == 1 <=> t <= 0.04286).
constexpr float HK_T_FIVE = 0.0429f;

struct HkSampleTab {
  float sc[HK_SPB];      // images: sqrt(log2(e) / (4t)); distances are carried in these units so that e = ex2(-d'^2)
  float scale[HK_SPB];   // images: (-2 / (4t)) / sc
  float eps[HK_SPB];     // denominator epsilon in this form's units
  int count[HK_SPB];     // images: M (|m| <= M; 0 = the five-image form), modes: K
  int modes[HK_SPB];     // 1: eigen-series, 0: images
  int order[HK_SPB];     // order[slot] = sample visited at that slot
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Eigen-series (cube.py:92-107) for t > HK_T_IMAGES: at most five modes survive, harmonics by the rotation recurrence.
// sin / cos(pi x) come from the special-function unit: x, x0 lie in [0,1], so the argument is inside [0, pi] where
// sin.approx / cos.approx are accurate to 2^-21 absolute -- 4e-7 of a score whose scale is set by the same e_1.
__device__ __forceinline__ float hk_modes(float x, float x0, const float2* __restrict__ e, int K, float eps) {
  const float s1 = __sinf(PI_F * x), c1 = __cosf(PI_F * x);
  const float s01 = __sinf(PI_F * x0), c01 = __cosf(PI_F * x0);
  float s = s1, c = c1, s0 = s01, c0 = c01;
  float2 ek = e[0];                          // (exp(-t k^2 pi^2), k * exp(-t k^2 pi^2))
  float den = ek.x * (c * c0), num = ek.y * (s * c0);
  for (int k = 1; k < K; ++k) {
    const float cn = fmaf(c, c1, -s * s1), sn = fmaf(s, c1, c * s1);
    const float c0n = fmaf(c0, c01, -s0 * s01), s0n = fmaf(s0, c01, c0 * s01);
    c = cn; s = sn; c0 = c0n; s0 = s0n;
    ek = e[k];
    den = fmaf(ek.x, c * c0, den);           // cos(k pi x) cos(k pi x0)
    num = fmaf(ek.y, s * c0, num);           // k sin(k pi x) cos(k pi x0)
  }
  return (-2.0f * PI_F * num) * rcp_approx(fmaf(2.0f, den, 1.0f) + eps);
}

// Accurate-trig variant for arguments outside [0,1] (score_hk is only specified on the cube, but the reference
// evaluates whatever it is given).
__device__ __forceinline__ float hk_modes_any(float x, float x0, const float2* __restrict__ e, int K, float eps) {
  float s1, c1, s01, c01;
  sincospif(x, &s1, &c1);
  sincospif(x0, &s01, &c01);
  float s = s1, c = c1, s0 = s01, c0 = c01, num = 0.0f, den = 0.0f;
  for (int k = 0; k < K; ++k) {
    const float2 ek = e[k];
    den = fmaf(ek.x, c * c0, den);
    num = fmaf(ek.y, s * c0, num);
    const float cn = fmaf(c, c1, -s * s1), sn = fmaf(s, c1, c * s1);
    const float c0n = fmaf(c0, c01, -s0 * s01), s0n = fmaf(s0, c01, c0 * s01);
    c = cn; s = sn; c0 = c0n; s0 = s0n;
  }
  return __fdividef(-2.0f * PI_F * num, fmaf(2.0f, den, 1.0f) + eps);
}

// Method of images (cube.py:129-146).  Distances are formed in x units exactly as the reference forms them --
// (2m + x) - x0 and (2m - x) - x0, so the near-wall images whose distance is a small difference of O(1) numbers keep
// it exact -- and only then scaled by sc = sqrt(log2(e) / 4t): the Gaussian of an image is ex2(-d'^2), and
// sum(sign d e) / sum(e) picks up the common factor (-2/4t)/sc once at the end.
//   FIVE: images x-2, x, x+2, -x, 2-x (the converged sum for t <= HK_T_FIVE);  MC: |m| <= MC on both families.
template <int MC, bool FIVE>
__device__ __forceinline__ float hk_images_fixed(float x, float x0, float sc, float scale, float eps) {
  float numa = 0.0f, numb = 0.0f, dena = 0.0f, denb = 0.0f;
#pragma unroll
  for (int m = -MC; m <= MC; ++m) {
    const float r = static_cast<float>(2 * m);
    const float da = ((r + x) - x0) * sc;
    const float ea = ex2_approx(-(da * da));
    numa = fmaf(da, ea, numa);
    dena += ea;
    if (!(FIVE && m == -MC)) {
      const float db = ((r - x) - x0) * sc;
      const float eb = ex2_approx(-(db * db));
      numb = fmaf(db, eb, numb);
      denb += eb;
    }
  }
  return (scale * (numa - numb)) * rcp_approx((dena + denb) + eps);
}

__device__ __forceinline__ float hk_images(float x, float x0, float sc, float scale, int M, float eps) {
  float num = 0.0f, den = 0.0f;
  for (int m = -M; m <= M; ++m) {
    const float r = static_cast<float>(2 * m);
    const float da = ((r + x) - x0) * sc, db = ((r - x) - x0) * sc;   // images 2m + x (sign +) and 2m - x (sign -)
    const float ea = ex2_approx(-(da * da)), eb = ex2_approx(-(db * db));
    num = fmaf(da, ea, num);
    num = fmaf(-db, eb, num);
    den += ea;
    den += eb;
  }
  return __fdividef(scale * num, den + eps);
}

// Which form / how many terms a sample needs (shared by the single-kernel path and the streaming path's pre-pass).
//   form 0: images |m| <= count   1: eigen-series, count modes   2: empty series (efs == 0)   3: the five-image form
enum { HK_IMAGES = 0, HK_MODES = 1, HK_EMPTY = 2, HK_FIVE = 3 };
struct HkClass {
  int form, count;
  float t, sc, scale, eps;
};
__device__ __forceinline__ HkClass hk_classify(float sg, int efs, int refls, float min_cutoff) {
  HkClass c;
  const float t = __fdiv_rn(__fmul_rn(sg, sg), 2.0f);  // sigma ** 2 / 2
  const int Kc = static_cast<int>(fminf(ceilf(sqrtf(17.5f / (PI2_F * t))) + 1.0f, 1.0e6f));
  const int Mc = static_cast<int>(fminf(ceilf(sqrtf(0.25f * fmaf(70.0f, t, 1.0f))), 1.0e6f));
  c.t = t;
  c.eps = 1e-12f;
  if (t > min_cutoff) {            // the reference's ef_cond (cube.py:176)
    if (efs >= Kc && t <= HK_T_IMAGES) {
      c.form = HK_IMAGES; c.count = Mc;       // converged series == converged image sum
      c.eps = 1e-12f * sqrtf(4.0f * PI_F * t);
    } else if (efs == 0) {
      c.form = HK_EMPTY; c.count = 0;          // num = 0, den = 1
    } else {
      c.form = HK_MODES; c.count = min(efs, min(Kc, HK_MAX_EFS));
    }
  } else {
    c.form = HK_IMAGES; c.count = min(refls, Mc);
  }
  if (c.form == HK_IMAGES && c.count == 1 && Mc == 1 && t <= HK_T_FIVE) c.form = HK_FIVE;
  // (t != t: NaN sigma) -> NaN out through the image form
  c.sc = sqrtf(1.4426950408889634f / (4.0f * t));
  c.scale = (-2.0f / (4.0f * t)) / c.sc;
  return c;
}

// VEC consecutive elements of one sample per thread and pass (8: 256-bit accesses, two independent chains per form).
template <int VEC>
__global__ void __launch_bounds__(288) score_hk_kernel(const float* __restrict__ x, const float* __restrict__ x0,
                                                       const float* __restrict__ sigma, float sigma_scalar,
                                                       float* __restrict__ out, size_t B, int D, int efs, int refls,
                                                       float min_cutoff) {
  __shared__ HkSampleTab tab;
  __shared__ float2 e_tab[HK_SPB][HK_MAX_EFS];
  const size_t s0 = static_cast<size_t>(blockIdx.x) * HK_SPB;
  const int ns = static_cast<int>(min(static_cast<size_t>(HK_SPB), B - s0));

  if (threadIdx.x < 32) {
    const int i = threadIdx.x;
    int key = 0x7fffffff;
    if (i < ns) {
      const HkClass c = hk_classify(sigma ? sigma[s0 + i] : sigma_scalar, efs, refls, min_cutoff);
      tab.sc[i] = c.sc;
      tab.scale[i] = c.scale;
      tab.eps[i] = c.eps;
      tab.count[i] = c.count;
      tab.modes[i] = c.form;
      key = (c.form << 20) | min(c.count, (1 << 20) - 1);
      if (c.form == HK_MODES) {
        for (int k = 1; k <= c.count; ++k) {
          const float kf = static_cast<float>(k);
          const float ev = expf(__fmul_rn(__fmul_rn(-c.t, __fmul_rn(kf, kf)), PI2_F));  // exp(-t k^2 pi^2) (cube.py:103-104)
          e_tab[i][k - 1] = make_float2(ev, ev * kf);
        }
      }
    }
    // stable rank of (form, term count) among the block's samples
    int rank = 0;
#pragma unroll 8
    for (int j = 0; j < 32; ++j) {
      const int kj = __shfl_sync(0xffffffffu, key, j);
      rank += (kj < key || (kj == key && j < i)) ? 1 : 0;
    }
    if (i < ns) tab.order[rank] = i;
  }
  __syncthreads();

  const size_t base = s0 * D;
  const int nelem = ns * D;
  for (int e = threadIdx.x * VEC; e < nelem; e += blockDim.x * VEC) {
    const int slot = e / D;  // VEC > 1 requires D % VEC == 0 so a vector never straddles samples
    const int s = tab.order[slot];
    const size_t off = base + static_cast<size_t>(s) * D + (e - slot * D);
    float xv[VEC], x0v[VEC], r[VEC];
    if (VEC == 1) {
      xv[0] = x[off];
      x0v[0] = x0[off];
    } else {
#pragma unroll
      for (int h = 0; h < VEC / 4; ++h) {
        const float4 a = ld_stream4(x + off + 4 * h), b = ld_stream4(x0 + off + 4 * h);
        xv[(4 * h) % VEC] = a.x; xv[(4 * h + 1) % VEC] = a.y; xv[(4 * h + 2) % VEC] = a.z; xv[(4 * h + 3) % VEC] = a.w;
        x0v[(4 * h) % VEC] = b.x; x0v[(4 * h + 1) % VEC] = b.y; x0v[(4 * h + 2) % VEC] = b.z; x0v[(4 * h + 3) % VEC] = b.w;
      }
    }
    const int cnt = tab.count[s], form = tab.modes[s];
    const float eps = tab.eps[s];
    if (form == HK_MODES) {
      bool in_cube = true;
#pragma unroll
      for (int v = 0; v < VEC; ++v) in_cube = in_cube && xv[v] >= 0.0f && xv[v] <= 1.0f && x0v[v] >= 0.0f && x0v[v] <= 1.0f;
      if (in_cube) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) r[v] = hk_modes(xv[v], x0v[v], e_tab[s], cnt, eps);
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) r[v] = hk_modes_any(xv[v], x0v[v], e_tab[s], cnt, eps);
      }
    } else if (form == HK_EMPTY) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) r[v] = -0.0f;  // -2 pi * (empty sum) / (1 + eps), whatever x is
    } else {
      const float sc = tab.sc[s], scl = tab.scale[s];
      if (form == HK_FIVE) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) r[v] = hk_images_fixed<1, true>(xv[v], x0v[v], sc, scl, eps);
      } else if (cnt == 1) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) r[v] = hk_images_fixed<1, false>(xv[v], x0v[v], sc, scl, eps);
      } else if (cnt == 2) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) r[v] = hk_images_fixed<2, false>(xv[v], x0v[v], sc, scl, eps);
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) r[v] = hk_images(xv[v], x0v[v], sc, scl, cnt, eps);
      }
    }
    if (VEC == 1) {
      out[off] = r[0];
    } else {
#pragma unroll
      for (int h = 0; h < VEC / 4; ++h)
        st_stream4(out + off + 4 * h, make_float4(r[(4 * h) % VEC], r[(4 * h + 1) % VEC], r[(4 * h + 2) % VEC], r[(4 * h + 3) % VEC]));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Streaming path of cube.score_hk (used whenever the caller provides a workspace): the per-sample classification is a
// pre-pass over sigma alone (17 bytes per sample: a 16-byte record and the sample's slot in its 256-sample tile, sorted
// by form), and the element kernel is a pure grid-stride stream -- no shared memory, no barrier, the next vector's
// loads in flight while the current one is evaluated.
struct HkRec {
  float a;    // images: sc                 modes: e1 = exp(-t pi^2) (e_k = e1^(k^2))    slow modes: t
  float b;    // images: (-2/4t)/sc
  float eps;
  int fc;     // form | count << 8 ;  forms as HK_*, plus HK_MODES_SLOW for series longer than 4 modes
};
enum { HK_MODES_SLOW = 4 };

constexpr int HK_TILE = 256;   // samples per sorting tile of the streaming path (slot index fits one byte)
constexpr int HK_BINS = 5 * (HK_MAX_EFS + 1);

// One block = one tile of HK_TILE samples: classify, then counting-sort the tile by (form, term count) so that the
// threads of a warp of the element kernel evaluate the same series for the same number of terms even when sigma
// differs from sample to sample (order within a key is arbitrary and irrelevant).
__global__ void __launch_bounds__(HK_TILE) hk_prepare_kernel(const float* __restrict__ sigma, float sigma_scalar, size_t B, int efs,
                                                             int refls, float min_cutoff, HkRec* __restrict__ rec,
                                                             unsigned char* __restrict__ order) {
  __shared__ int hist[HK_BINS];
  const size_t base = static_cast<size_t>(blockIdx.x) * HK_TILE;
  const size_t i = base + threadIdx.x;
  for (int k = threadIdx.x; k < HK_BINS; k += HK_TILE) hist[k] = 0;
  __syncthreads();
  int bin = -1, pos = 0;
  if (i < B) {
    const HkClass c = hk_classify(sigma ? sigma[i] : sigma_scalar, efs, refls, min_cutoff);
    HkRec r;
    int form = c.form;
    r.a = c.sc; r.b = c.scale; r.eps = c.eps;
    if (form == HK_MODES) {
      if (c.count <= 4) r.a = expf(__fmul_rn(-c.t, PI2_F));
      else { form = HK_MODES_SLOW; r.a = c.t; }
    }
    r.fc = form | (c.count << 8);
    rec[i] = r;
    bin = form * (HK_MAX_EFS + 1) + min(c.count, HK_MAX_EFS);
  }
  {
    // warp-aggregated histogram insert: with one sigma for the whole batch every sample lands in the same bin, and 256
    // serialised shared-memory atomics on one address cost more than the classification itself
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    int start = 0;
    if (lane == leader && bin >= 0) start = atomicAdd(&hist[bin], __popc(peers));
    start = __shfl_sync(0xffffffffu, start, leader);
    pos = start + __popc(peers & ((1u << lane) - 1u));
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // exclusive prefix over the (few hundred, mostly empty) bins
    int run = 0;
    for (int k = 0; k < HK_BINS; ++k) { const int n = hist[k]; hist[k] = run; run += n; }
  }
  __syncthreads();
  if (i < B) order[base + hist[bin] + pos] = static_cast<unsigned char>(threadIdx.x);
}

// Eigen-series with at most four modes, e_k = e1^(k^2) by repeated squaring (the relative error of the higher powers,
// ~k^2 2^-24, sits on terms that are already e1^3 .. e1^15 times smaller than the first).
__device__ __forceinline__ float hk_modes4(float x, float x0, float e1, int K, float eps) {
  // sin / cos(pi x) on the special-function unit, accurate to 2^-21 absolute for arguments in [-pi, pi]: x is first
  // reduced to [-1, 1] by its exact period 2 (a no-op on the cube, where score_hk is specified)
  x = fmaf(-2.0f, rintf(0.5f * x), x);
  x0 = fmaf(-2.0f, rintf(0.5f * x0), x0);
  const float s1 = __sinf(PI_F * x), c1 = __cosf(PI_F * x), s01 = __sinf(PI_F * x0), c01 = __cosf(PI_F * x0);
  const float e1_2 = e1 * e1, e2 = e1_2 * e1_2, e1_8 = e2 * e2;
  const float ek[4] = {e1, e2, e1_8 * e1, e1_8 * e1_8};
  float s = s1, c = c1, s0 = s01, c0 = c01;
  float den = ek[0] * (c * c0), num = ek[0] * (s * c0);
#pragma unroll
  for (int k = 1; k < 4; ++k) {
    if (k < K) {
      const float cn = fmaf(c, c1, -s * s1), sn = fmaf(s, c1, c * s1);
      const float c0n = fmaf(c0, c01, -s0 * s01), s0n = fmaf(s0, c01, c0 * s01);
      c = cn; s = sn; c0 = c0n; s0 = s0n;
      den = fmaf(ek[k], c * c0, den);
      num = fmaf(ek[k] * static_cast<float>(k + 1), s * c0, num);
    }
  }
  return (-2.0f * PI_F * num) * rcp_approx(fmaf(2.0f, den, 1.0f) + eps);
}

__device__ __forceinline__ float hk_modes_slow(float x, float x0, float t, int K, float eps) {
  float s1, c1, s01, c01;
  sincospif(x, &s1, &c1);
  sincospif(x0, &s01, &c01);
  float s = s1, c = c1, s0 = s01, c0 = c01, num = 0.0f, den = 0.0f;
  for (int k = 1; k <= K; ++k) {
    const float kf = static_cast<float>(k);
    const float ev = expf(__fmul_rn(__fmul_rn(-t, __fmul_rn(kf, kf)), PI2_F));
    den = fmaf(ev, c * c0, den);
    num = fmaf(ev * kf, s * c0, num);
    const float cn = fmaf(c, c1, -s * s1), sn = fmaf(s, c1, c * s1);
    const float c0n = fmaf(c0, c01, -s0 * s01), s0n = fmaf(s0, c01, c0 * s01);
    c = cn; s = sn; c0 = c0n; s0 = s0n;
  }
  return __fdividef(-2.0f * PI_F * num, fmaf(2.0f, den, 1.0f) + eps);
}

template <int VEC>
__device__ __forceinline__ void hk_eval(const float (&xv)[VEC], const float (&x0v)[VEC], const HkRec rc, float (&r)[VEC]) {
  const int form = rc.fc & 0xff, cnt = rc.fc >> 8;
  if (form == HK_FIVE) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) r[v] = hk_images_fixed<1, true>(xv[v], x0v[v], rc.a, rc.b, rc.eps);
  } else if (form == HK_IMAGES) {
    if (cnt == 1) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) r[v] = hk_images_fixed<1, false>(xv[v], x0v[v], rc.a, rc.b, rc.eps);
    } else if (cnt == 2) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) r[v] = hk_images_fixed<2, false>(xv[v], x0v[v], rc.a, rc.b, rc.eps);
    } else {
#pragma unroll
      for (int v = 0; v < VEC; ++v) r[v] = hk_images(xv[v], x0v[v], rc.a, rc.b, cnt, rc.eps);
    }
  } else if (form == HK_MODES) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) r[v] = hk_modes4(xv[v], x0v[v], rc.a, cnt, rc.eps);
  } else if (form == HK_MODES_SLOW) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) r[v] = hk_modes_slow(xv[v], x0v[v], rc.a, cnt, rc.eps);
  } else {
#pragma unroll
    for (int v = 0; v < VEC; ++v) r[v] = -0.0f;  // -2 pi * (empty sum) / (1 + eps), whatever x is
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) score_hk_stream_kernel(const float* __restrict__ x, const float* __restrict__ x0,
                                                              float* __restrict__ out, const HkRec* __restrict__ rec,
                                                              const unsigned char* __restrict__ order, unsigned int nvec,
                                                              unsigned int vps /* vectors per sample */, int D) {
  const unsigned int vpt = HK_TILE * vps;  // vectors per sorting tile
  const unsigned int stride = gridDim.x * blockDim.x;
  unsigned int v = blockIdx.x * blockDim.x + threadIdx.x;
  struct Item { size_t off; HkRec rc; float xv[VEC], x0v[VEC]; };
  auto fetch = [&](unsigned int vi, Item& it) {
    const unsigned int tile = vi / vpt, w = vi - tile * vpt;
    const unsigned int slot = w / vps, j = w - slot * vps;
    const size_t s = static_cast<size_t>(tile) * HK_TILE + order[static_cast<size_t>(tile) * HK_TILE + slot];
    it.off = s * D + j * VEC;
    const int4 rr = __ldg(reinterpret_cast<const int4*>(rec + s));
    it.rc.a = __int_as_float(rr.x); it.rc.b = __int_as_float(rr.y); it.rc.eps = __int_as_float(rr.z); it.rc.fc = rr.w;
    if (VEC == 1) {
      it.xv[0] = x[it.off];
      it.x0v[0] = x0[it.off];
    } else {
#pragma unroll
      for (int h = 0; h < VEC / 4; ++h) {
        const float4 a = ld_stream4(x + it.off + 4 * h), b = ld_stream4(x0 + it.off + 4 * h);
        it.xv[(4 * h) % VEC] = a.x; it.xv[(4 * h + 1) % VEC] = a.y; it.xv[(4 * h + 2) % VEC] = a.z; it.xv[(4 * h + 3) % VEC] = a.w;
        it.x0v[(4 * h) % VEC] = b.x; it.x0v[(4 * h + 1) % VEC] = b.y; it.x0v[(4 * h + 2) % VEC] = b.z; it.x0v[(4 * h + 3) % VEC] = b.w;
      }
    }
  };
  if (v >= nvec) return;
  Item cur, nxt;
  fetch(v, cur);
  for (;;) {
    const unsigned int vn = v + stride;
    const bool more = vn < nvec && vn > v;
    if (more) fetch(vn, nxt);
    float r[VEC];
    hk_eval<VEC>(cur.xv, cur.x0v, cur.rc, r);
    if (VEC == 1) {
      out[cur.off] = r[0];
    } else {
#pragma unroll
      for (int h = 0; h < VEC / 4; ++h)
        st_stream4(out + cur.off + 4 * h, make_float4(r[(4 * h) % VEC], r[(4 * h + 1) % VEC], r[(4 * h + 2) % VEC], r[(4 * h + 3) % VEC]));
    }
    if (!more) break;
    cur = nxt;
    v = vn;
  }
}

// ------------------------------------------------------------------------------------------------
// noise dump: exactly the stream the fused step kernels consume
__global__ void __launch_bounds__(256) philox_normal_kernel(float* __restrict__ out, size_t n, uint64_t seed,
                                                            uint32_t draw, int vec_ok) {
  const size_t nquad = (n + 3) >> 2;
  size_t q = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; q < nquad; q += stride) {
    float z[4];
    philox_normal4(seed, draw, q, z);
    if (vec_ok && 4 * q + 4 <= n) {
      st_stream4(out + 4 * q, make_float4(z[0], z[1], z[2], z[3]));
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (4 * q + u < n) out[4 * q + u] = z[u];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// classifier-free guidance combine (models/utils.py:120-138): (1 + w) * s_c - w * s_u
__global__ void __launch_bounds__(256) cfg_combine_kernel(const float* __restrict__ s, const float* __restrict__ w,
                                                          float w_scalar, float* __restrict__ out, size_t B, size_t D) {
  const size_t n = B * D;
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) {
    const float wv = w ? w[i / D] : w_scalar;
    out[i] = __fsub_rn(__fmul_rn(__fadd_rn(1.0f, wv), s[i]), __fmul_rn(wv, s[n + i]));
  }
}

// ------------------------------------------------------------------------------------------------
// Langevin corrector, stage 1 (sampling.py:225-226): per-sample L2 norms of grad and noise.
// One warp per sample, 8 warps per block, persistent over the batch (grid <= 8 CTAs per SM): every warp sums
// the norms of its samples in a fixed order, the block adds its 8 warps in a fixed order, so the result is
// run-to-run deterministic and the number of partials stays small at any batch size.
// Philox quads are indexed over the FLAT tensor (element 4q..4q+3), so a sample whose D is not a multiple of 4
// (the shipped 9x9 latents) simply shares its boundary quads with its neighbours.
constexpr int PC_SPB = 8;

__device__ __forceinline__ uint32_t draw_index(uint32_t draw_base, const int32_t* step_ctr, int which) {
  const int32_t step = step_ctr ? *step_ctr : 0;
  return draw_base + 2u * static_cast<uint32_t>(step) + static_cast<uint32_t>(which);
}

__global__ void __launch_bounds__(32 * PC_SPB) pc_norms_kernel(const float* __restrict__ grad,
                                                               const float* __restrict__ noise,
                                                               float* __restrict__ partial, size_t B, int D,
                                                               uint64_t seed, uint32_t draw_base,
                                                               const int32_t* __restrict__ step_ctr,
                                                               size_t noise_step_stride, int vec_ok) {
  __shared__ float sg[PC_SPB], sn[PC_SPB];
  if (noise && step_ctr) noise += static_cast<size_t>(*step_ctr) * noise_step_stride;
  if (noise && (reinterpret_cast<uintptr_t>(noise) & 15)) vec_ok = 0;  // tape slices of odd-sized tensors
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t draw = draw_index(draw_base, step_ctr, 0);
  float acc_g = 0.0f, acc_n = 0.0f;
  for (size_t b = static_cast<size_t>(blockIdx.x) * PC_SPB + warp; b < B; b += static_cast<size_t>(gridDim.x) * PC_SPB) {
    const size_t e0 = b * D, e1 = e0 + D;
    float g2 = 0.0f, n2 = 0.0f;
    for (size_t q = (e0 >> 2) + lane; 4 * q < e1; q += 32) {
      const size_t i0 = 4 * q;
      float gv[4], z[4];
      const bool whole = i0 >= e0 && i0 + 4 <= e1;
      if (whole && vec_ok) {
        const float4 g = *reinterpret_cast<const float4*>(grad + i0);
        gv[0] = g.x; gv[1] = g.y; gv[2] = g.z; gv[3] = g.w;
        if (noise) {
          const float4 nz = *reinterpret_cast<const float4*>(noise + i0);
          z[0] = nz.x; z[1] = nz.y; z[2] = nz.z; z[3] = nz.w;
        }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool in = i0 + u >= e0 && i0 + u < e1;
          gv[u] = in ? grad[i0 + u] : 0.0f;
          z[u] = (in && noise) ? noise[i0 + u] : 0.0f;
        }
      }
      if (!noise) {
        philox_normal4(seed, draw, q, z);
        if (!whole) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (!(i0 + u >= e0 && i0 + u < e1)) z[u] = 0.0f;
        }
      }
      g2 += gv[0] * gv[0] + gv[1] * gv[1] + gv[2] * gv[2] + gv[3] * gv[3];
      n2 += z[0] * z[0] + z[1] * z[1] + z[2] * z[2] + z[3] * z[3];
    }
    g2 = warp_sum(g2);
    n2 = warp_sum(n2);
    acc_g += sqrtf(g2);
    acc_n += sqrtf(n2);
  }
  if (lane == 0) { sg[warp] = acc_g; sn[warp] = acc_n; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.0f, c = 0.0f;
#pragma unroll
    for (int i = 0; i < PC_SPB; ++i) { a += sg[i]; c += sn[i]; }
    partial[2 * blockIdx.x] = a;
    partial[2 * blockIdx.x + 1] = c;
  }
}

// Large batches (D % 4 == 0): one warp per sample leaves 14 of 32 lanes idle at D = 72 (18 quads), and the kernel is
// bound by the Philox / Box-Muller arithmetic, not by its 4 B/element of traffic.  Here a warp takes PCB_SC samples at
// a time, walks their quads with all 32 lanes, parks the per-quad sums of squares in shared memory and lets one
// lane per sample add them in a fixed order.
constexpr int PCB_SC = 16;
constexpr int PCB_MAXQ = 32;  // D <= 128

__global__ void __launch_bounds__(32 * PC_SPB) pc_norms_batched_kernel(const float* __restrict__ grad,
                                                                       const float* __restrict__ noise,
                                                                       float* __restrict__ partial, size_t B, int Q,
                                                                       uint64_t seed, uint32_t draw_base,
                                                                       const int32_t* __restrict__ step_ctr,
                                                                       size_t noise_step_stride) {
  __shared__ float2 part[PC_SPB][PCB_SC * PCB_MAXQ];
  __shared__ float sg[PC_SPB], sn[PC_SPB];
  if (noise && step_ctr) noise += static_cast<size_t>(*step_ctr) * noise_step_stride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t draw = draw_index(draw_base, step_ctr, 0);
  float acc_g = 0.0f, acc_n = 0.0f;
  const size_t nchunks = (B + PCB_SC - 1) / PCB_SC;
  for (size_t ch = static_cast<size_t>(blockIdx.x) * PC_SPB + warp; ch < nchunks; ch += static_cast<size_t>(gridDim.x) * PC_SPB) {
    const size_t b0 = ch * PCB_SC;
    const int ns = static_cast<int>(min(static_cast<size_t>(PCB_SC), B - b0));
    const size_t q0 = b0 * Q;
    for (int it = lane; it < ns * Q; it += 32) {
      const size_t q = q0 + it;
      const float4 g = *reinterpret_cast<const float4*>(grad + 4 * q);
      float z[4];
      if (noise) {
        const float4 nz = *reinterpret_cast<const float4*>(noise + 4 * q);
        z[0] = nz.x; z[1] = nz.y; z[2] = nz.z; z[3] = nz.w;
      } else {
        philox_normal4(seed, draw, q, z);
      }
      part[warp][it] = make_float2(g.x * g.x + g.y * g.y + g.z * g.z + g.w * g.w,
                                   z[0] * z[0] + z[1] * z[1] + z[2] * z[2] + z[3] * z[3]);
    }
    __syncwarp();
    if (lane < ns) {
      float g2 = 0.0f, n2 = 0.0f;
      for (int k = 0; k < Q; ++k) { const float2 v = part[warp][lane * Q + k]; g2 += v.x; n2 += v.y; }
      acc_g += sqrtf(g2);
      acc_n += sqrtf(n2);
    }
    __syncwarp();
  }
  acc_g = warp_sum(acc_g);
  acc_n = warp_sum(acc_n);
  if (lane == 0) { sg[warp] = acc_g; sn[warp] = acc_n; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.0f, c = 0.0f;
#pragma unroll
    for (int i = 0; i < PC_SPB; ++i) { a += sg[i]; c += sn[i]; }
    partial[2 * blockIdx.x] = a;
    partial[2 * blockIdx.x + 1] = c;
  }
}

// fixed-order reduction of the per-block partials by one warp; every block recomputes it (<= 10 KB from L2)
__device__ __forceinline__ void reduce_partials(const float* __restrict__ partial, int nblk, float& gsum, float& nsum) {
  const int lane = threadIdx.x & 31;
  float a = 0.0f, c = 0.0f;
  for (int i = lane; i < nblk; i += 32) {
    a += partial[2 * i];
    c += partial[2 * i + 1];
  }
  gsum = warp_sum(a);
  nsum = warp_sum(c);
}

// Shared body of the two fused updates:  x_mean = x + a * v ;  x = reflect(x_mean + b * z) ; x_mean = reflect(x_mean)
// over quads of the flat tensor; whole, aligned quads move as 128-bit vectors, the rest element by element.
template <bool MEAN, bool PER_SAMPLE, bool PRED>
__device__ __forceinline__ void pc_update_loop(const float* __restrict__ x, const float* __restrict__ v,
                                               const float* __restrict__ zt, float a, float b,
                                               const float* __restrict__ g_table, float dt, float sqrt_dt,
                                               float* __restrict__ x_out, float* __restrict__ x_mean_out, size_t n,
                                               size_t D, uint64_t seed, uint32_t draw, bool vec_ok) {
  const size_t nq = (n + 3) >> 2;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  auto load = [&](size_t i, float (&xs)[4], float (&vs)[4], float (&z)[4]) {
    const size_t i0 = 4 * i;
    if (vec_ok && i0 + 4 <= n) {
      const float4 xv = *reinterpret_cast<const float4*>(x + i0), vv = *reinterpret_cast<const float4*>(v + i0);
      xs[0] = xv.x; xs[1] = xv.y; xs[2] = xv.z; xs[3] = xv.w;
      vs[0] = vv.x; vs[1] = vv.y; vs[2] = vv.z; vs[3] = vv.w;
      if (zt) {
        const float4 nz = ld_stream4(zt + i0);
        z[0] = nz.x; z[1] = nz.y; z[2] = nz.z; z[3] = nz.w;
      }
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool in = i0 + u < n;
        xs[u] = in ? x[i0 + u] : 0.0f;
        vs[u] = in ? v[i0 + u] : 0.0f;
        z[u] = (in && zt) ? zt[i0 + u] : 0.0f;
      }
    }
  };
  auto finish = [&](size_t i, const float (&xs)[4], const float (&vs)[4], float (&z)[4]) {
    const size_t i0 = 4 * i;
    if (!zt) philox_normal4(seed, draw, i, z);
    float xm[4], xn[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float au = a, bu = b;
      if (PER_SAMPLE) {  // predictor with one diffusion coefficient per sample (update_fn API, arbitrary t[B])
        const float g = g_table[min(i0 + u, n - 1) / D];
        au = -__fmul_rn(g, g);
        bu = __fmul_rn(g, sqrt_dt);
      }
      // predictor: drift = -(g^2) * score ; x_mean = x + drift * dt        corrector: x_mean = x + step * grad
      const float m = PRED ? __fadd_rn(xs[u], __fmul_rn(__fmul_rn(au, vs[u]), dt)) : __fadd_rn(xs[u], __fmul_rn(au, vs[u]));
      xn[u] = reflect1(__fadd_rn(m, __fmul_rn(bu, z[u])));
      if (MEAN) xm[u] = reflect1(m);
    }
    if (vec_ok && i0 + 4 <= n) {
      *reinterpret_cast<float4*>(x_out + i0) = make_float4(xn[0], xn[1], xn[2], xn[3]);
      if (MEAN) *reinterpret_cast<float4*>(x_mean_out + i0) = make_float4(xm[0], xm[1], xm[2], xm[3]);
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i0 + u < n) {
          x_out[i0 + u] = xn[u];
          if (MEAN) x_mean_out[i0 + u] = xm[u];
        }
    }
  };
  // two quads per iteration: both threads' loads are in flight before either quad's Philox / reflect arithmetic starts
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nq; i += 2 * stride) {
    const size_t j = i + stride;
    float xa[4], va[4], za[4], xb[4], vb[4], zb[4];
    load(i, xa, va, za);
    if (j < nq) load(j, xb, vb, zb);
    finish(i, xa, va, za);
    if (j < nq) finish(j, xb, vb, zb);
  }
}

// Langevin corrector, stage 2 (sampling.py:227-231)
template <bool MEAN>
__global__ void __launch_bounds__(256) pc_corrector_apply_kernel(
    const float* __restrict__ x, const float* __restrict__ grad, const float* __restrict__ noise,
    const float* __restrict__ partial, int nblk, float snr, float* __restrict__ x_out, float* __restrict__ x_mean_out,
    float* __restrict__ stats_out, size_t B, size_t n, uint64_t seed, uint32_t draw_base,
    const int32_t* __restrict__ step_ctr, size_t noise_step_stride, int vec_ok) {
  __shared__ float s_step, s_noise_c;
  if (noise && step_ctr) noise += static_cast<size_t>(*step_ctr) * noise_step_stride;
  if (noise && (reinterpret_cast<uintptr_t>(noise) & 15)) vec_ok = 0;
  if (threadIdx.x < 32) {
    float gsum, nsum;
    reduce_partials(partial, nblk, gsum, nsum);
    if (threadIdx.x == 0) {
      const float gbar = gsum / static_cast<float>(B), nbar = nsum / static_cast<float>(B);
      const float r = __fdiv_rn(__fmul_rn(snr, nbar), gbar);      // target_snr * noise_norm / grad_norm
      const float step = __fmul_rn(__fmul_rn(r, r), 2.0f);        // (...) ** 2 * 2 * alpha(=1)
      s_step = step;
      s_noise_c = sqrtf(__fmul_rn(step, 2.0f));                   // sqrt(step_size * 2)
      if (stats_out && blockIdx.x == 0) { stats_out[0] = gbar; stats_out[1] = nbar; stats_out[2] = step; }
    }
  }
  __syncthreads();
  // x_mean = x + step * grad ; x = x_mean + sqrt(2 step) * noise
  pc_update_loop<MEAN, false, false>(x, grad, noise, s_step, s_noise_c, nullptr, 0.0f, 0.0f, x_out, x_mean_out, n, 1, seed,
                              draw_index(draw_base, step_ctr, 0), vec_ok != 0);
}

// Euler-Maruyama predictor on the reverse reflected VE-SDE (sampling.py:198-207, sde_lib.py:93-101,135-140):
//   drift = 0 - g^2 * score ; x_mean = x + drift * dt ; x = x_mean + (g * sqrt(-dt)) * z ; reflect both
template <bool MEAN, bool PER_SAMPLE>
__global__ void __launch_bounds__(256) pc_predictor_kernel(const float* __restrict__ x, const float* __restrict__ score,
                                                           const float* __restrict__ zt, const float* __restrict__ g_table,
                                                           float dt, float sqrt_dt, float* __restrict__ x_out,
                                                           float* __restrict__ x_mean_out, size_t n, uint64_t seed,
                                                           uint32_t draw_base, const int32_t* __restrict__ step_ctr,
                                                           size_t noise_step_stride, size_t D, int vec_ok) {
  const int32_t step = step_ctr ? *step_ctr : 0;
  if (zt) zt += static_cast<size_t>(step) * noise_step_stride;
  if (zt && (reinterpret_cast<uintptr_t>(zt) & 15)) vec_ok = 0;
  // PER_SAMPLE: g_table holds one diffusion coefficient per sample; otherwise one per sampler step
  const float g = PER_SAMPLE ? 0.0f : g_table[step];
  pc_update_loop<MEAN, PER_SAMPLE, true>(x, score, zt, -__fmul_rn(g, g), __fmul_rn(g, sqrt_dt), g_table, dt, sqrt_dt, x_out,
                                   x_mean_out, n, D, seed, draw_base + 2u * static_cast<uint32_t>(step) + 1u, vec_ok != 0);
}

__global__ void step_advance_kernel(int32_t* ctr) { *ctr += 1; }

// Grid of a grid-stride streaming kernel: exactly one resident wave (SMs x blocks that fit per SM), so that there is
// no partial second wave (1184 blocks on 148 x 6 resident ones ran 1.33 waves).
template <typename K>
static inline int resident_grid(K kernel, size_t work_items, int threads) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) {
    (void)cudaGetLastError();
    per_sm = 4;
  }
  size_t blocks = (work_items + threads - 1) / threads;
  const size_t cap = static_cast<size_t>(kNumSMs) * per_sm;
  if (blocks > cap) blocks = cap;
  return static_cast<int>(blocks < 1 ? 1 : blocks);
}

static inline int stream_grid(size_t work_items, int threads, int max_waves = 8) {
  size_t blocks = (work_items + threads - 1) / threads;
  size_t cap = static_cast<size_t>(kNumSMs) * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace rd

using namespace rd;

extern "C" {

const char* rd_last_error(void) { return err_buf(); }
int rd_version(void) { return 100; }
int rd_device_cc(void) {
  int dev = 0, maj = 0, min = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev);
  if (e != cudaSuccess) return -fail(static_cast<int>(e), "rd_device_cc: %s", cudaGetErrorString(e));
  return maj * 10 + min;
}

int rd_reflect_f32(const float* x, float* out, size_t n, void* stream) {
  if (n == 0) return RD_OK;
  RD_REQUIRE(x && out, "rd_reflect_f32: null pointer");
  RD_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             "rd_reflect_f32: pointers must be 16-byte aligned");
  // 148 SMs x 8 resident CTAs of 256 threads, 4 x 128-bit loads in flight per thread
  int grid = stream_grid((n + 3) / 4, 256, 8);
  reflect_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, out, n);
  return check_launch("reflect_kernel");
}

int rd_inside_f32(const float* x, uint8_t* ok, size_t B, size_t D, void* stream) {
  if (B == 0) return RD_OK;
  RD_REQUIRE(x && ok, "rd_inside_f32: null pointer");
  size_t blocks = (B * 32 + 255) / 256;
  inside_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ok, B, D);
  return check_launch("inside_kernel");
}

int rd_score_hk_f32(const float* x, const float* x_orig, const float* sigma, float sigma_scalar, float* out,
                    size_t B, size_t D, int efs, int refls, float min_cutoff, void* stream) {
  if (B == 0 || D == 0) return RD_OK;
  RD_REQUIRE(x && x_orig && out, "rd_score_hk_f32: null pointer");
  RD_REQUIRE(efs >= 0 && efs <= HK_MAX_EFS, "rd_score_hk_f32: efs must be in [0,%d]", HK_MAX_EFS);
  RD_REQUIRE(refls >= 0, "rd_score_hk_f32: refls must be >= 0");
  RD_REQUIRE(D <= (1u << 20), "rd_score_hk_f32: D too large");
  size_t blocks = (B + HK_SPB - 1) / HK_SPB;
  const bool aligned16 = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(x_orig) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  const int vec = (aligned16 && D % 8 == 0) ? 8 : ((aligned16 && D % 4 == 0) ? 4 : 1);
  // block size: the one (multiple of 32, 128..288) that leaves the fewest idle threads in the last pass over the
  // block's HK_SPB * D elements (D = 72, 8 elements per thread: 288 threads, exactly one pass)
  const size_t items = static_cast<size_t>(HK_SPB) * D / vec;
  int threads = 256;
  double best = 0.0;
  for (int t = 288; t >= 128; t -= 32) {
    const double eff = static_cast<double>(items) / (static_cast<double>((items + t - 1) / t) * t);
    if (eff > best + 1e-9) { best = eff; threads = t; }
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (vec == 8)
    score_hk_kernel<8><<<static_cast<unsigned>(blocks), threads, 0, st>>>(x, x_orig, sigma, sigma_scalar, out, B,
                                                                           static_cast<int>(D), efs, refls, min_cutoff);
  else if (vec == 4)
    score_hk_kernel<4><<<static_cast<unsigned>(blocks), threads, 0, st>>>(x, x_orig, sigma, sigma_scalar, out, B,
                                                                           static_cast<int>(D), efs, refls, min_cutoff);
  else
    score_hk_kernel<1><<<static_cast<unsigned>(blocks), threads, 0, st>>>(x, x_orig, sigma, sigma_scalar, out, B,
                                                                           static_cast<int>(D), efs, refls, min_cutoff);
  return check_launch("score_hk_kernel");
}

size_t rd_score_hk_workspace_bytes(size_t B) { return (B + HK_TILE - 1) / HK_TILE * HK_TILE * (sizeof(HkRec) + 1) + 256; }

int rd_score_hk_ws_f32(const float* x, const float* x_orig, const float* sigma, float sigma_scalar, float* out, size_t B,
                       size_t D, int efs, int refls, float min_cutoff, void* workspace, size_t workspace_bytes, void* stream) {
  if (B == 0 || D == 0) return RD_OK;
  const bool aligned16 = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(x_orig) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  const int vec = (aligned16 && D % 8 == 0) ? 8 : ((aligned16 && D % 4 == 0) ? 4 : 1);
  const size_t nvec = B * (D / vec);
  if (!workspace || workspace_bytes < rd_score_hk_workspace_bytes(B) || nvec >= (1ull << 31) ||
      (reinterpret_cast<uintptr_t>(workspace) & 15))
    return rd_score_hk_f32(x, x_orig, sigma, sigma_scalar, out, B, D, efs, refls, min_cutoff, stream);  // single-kernel path
  RD_REQUIRE(x && x_orig && out, "rd_score_hk_ws_f32: null pointer");
  RD_REQUIRE(efs >= 0 && efs <= HK_MAX_EFS, "rd_score_hk_ws_f32: efs must be in [0,%d]", HK_MAX_EFS);
  RD_REQUIRE(refls >= 0, "rd_score_hk_ws_f32: refls must be >= 0");
  RD_REQUIRE(D <= (1u << 20), "rd_score_hk_ws_f32: D too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HkRec* rec = static_cast<HkRec*>(workspace);
  unsigned char* order = reinterpret_cast<unsigned char*>(rec + (B + HK_TILE - 1) / HK_TILE * HK_TILE);
  hk_prepare_kernel<<<static_cast<unsigned>((B + HK_TILE - 1) / HK_TILE), HK_TILE, 0, st>>>(sigma, sigma_scalar, B, efs, refls, min_cutoff, rec,
                                                                                      order);
  int rc = check_launch("hk_prepare_kernel");
  if (rc != RD_OK) return rc;
  // one resident wave: 148 SMs x 8 blocks of 256 threads, each thread walks its vectors with the next one in flight
  size_t blocks = (nvec + 255) / 256;
  if (blocks > static_cast<size_t>(kNumSMs) * 8) blocks = static_cast<size_t>(kNumSMs) * 8;
  const unsigned int nv = static_cast<unsigned int>(nvec), vps = static_cast<unsigned int>(D / vec);
  if (vec == 8)
    score_hk_stream_kernel<8><<<static_cast<unsigned>(blocks), 256, 0, st>>>(x, x_orig, out, rec, order, nv, vps, static_cast<int>(D));
  else if (vec == 4)
    score_hk_stream_kernel<4><<<static_cast<unsigned>(blocks), 256, 0, st>>>(x, x_orig, out, rec, order, nv, vps, static_cast<int>(D));
  else
    score_hk_stream_kernel<1><<<static_cast<unsigned>(blocks), 256, 0, st>>>(x, x_orig, out, rec, order, nv, vps, static_cast<int>(D));
  return check_launch("score_hk_stream_kernel");
}

int rd_philox_normal_f32(float* out, size_t n, uint64_t seed, uint32_t draw, void* stream) {
  if (n == 0) return RD_OK;
  RD_REQUIRE(out, "rd_philox_normal_f32: null pointer");
  int grid = stream_grid((n + 3) / 4, 256, 8);
  philox_normal_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(out, n, seed, draw,
                                                                             (reinterpret_cast<uintptr_t>(out) & 15) == 0);
  return check_launch("philox_normal_kernel");
}

int rd_cfg_combine_f32(const float* s, const float* w, float w_scalar, float* out, size_t B, size_t D,
                       void* stream) {
  if (B * D == 0) return RD_OK;
  RD_REQUIRE(s && out, "rd_cfg_combine_f32: null pointer");
  int grid = stream_grid(B * D, 256, 8);
  cfg_combine_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(s, w, w_scalar, out, B, D);
  return check_launch("cfg_combine_kernel");
}

static inline int all_aligned16(const void* a, const void* b, const void* c, const void* d) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
           reinterpret_cast<uintptr_t>(d)) & 15) == 0;
}

int rd_pc_norms(const float* grad, const float* noise, float* partial, int* nblk, size_t B, size_t D,
                uint64_t seed, uint32_t draw_base, const int32_t* step_ctr, size_t noise_step_stride,
                void* stream) {
  RD_REQUIRE(grad && partial && B > 0 && D > 0, "rd_pc_norms: bad arguments");
  size_t want = (B + PC_SPB - 1) / PC_SPB, cap = static_cast<size_t>(kNumSMs) * 8;
  if (D % 4 == 0 && D / 4 <= PCB_MAXQ && B >= cap * PC_SPB * PCB_SC && all_aligned16(grad, noise, nullptr, nullptr) &&
      (!noise || noise_step_stride % 4 == 0)) {
    // enough samples to give every warp of a full grid whole chunks of PCB_SC samples
    if (nblk) *nblk = static_cast<int>(cap);
    pc_norms_batched_kernel<<<static_cast<int>(cap), 32 * PC_SPB, 0, static_cast<cudaStream_t>(stream)>>>(
        grad, noise, partial, B, static_cast<int>(D / 4), seed, draw_base, step_ctr, noise_step_stride);
    return check_launch("pc_norms_batched_kernel");
  }
  int blocks = static_cast<int>(want < cap ? want : cap);
  if (nblk) *nblk = blocks;
  pc_norms_kernel<<<blocks, 32 * PC_SPB, 0, static_cast<cudaStream_t>(stream)>>>(
      grad, noise, partial, B, static_cast<int>(D), seed, draw_base, step_ctr, noise_step_stride,
      all_aligned16(grad, noise, nullptr, nullptr));
  return check_launch("pc_norms_kernel");
}

int rd_pc_corrector_apply(const float* x, const float* grad, const float* noise, const float* partial,
                          int nblk, float snr, float* x_out, float* x_mean_out, float* stats_out, size_t B,
                          size_t D, uint64_t seed, uint32_t draw_base, const int32_t* step_ctr,
                          size_t noise_step_stride, void* stream) {
  RD_REQUIRE(x && grad && partial && x_out && B > 0 && D > 0 && nblk > 0, "rd_pc_corrector_apply: bad arguments");
  const size_t n = B * D;
  const int vec_ok = all_aligned16(x, grad, x_out, x_mean_out) && all_aligned16(noise, nullptr, nullptr, nullptr);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t half = ((n + 3) / 4 + 1) / 2;  // two quads per thread and iteration
  int grid = x_mean_out ? resident_grid(pc_corrector_apply_kernel<true>, half, 256) : resident_grid(pc_corrector_apply_kernel<false>, half, 256);
  if (x_mean_out)
    pc_corrector_apply_kernel<true><<<grid, 256, 0, st>>>(x, grad, noise, partial, nblk, snr, x_out, x_mean_out, stats_out,
                                                          B, n, seed, draw_base, step_ctr, noise_step_stride, vec_ok);
  else
    pc_corrector_apply_kernel<false><<<grid, 256, 0, st>>>(x, grad, noise, partial, nblk, snr, x_out, x_mean_out, stats_out,
                                                           B, n, seed, draw_base, step_ctr, noise_step_stride, vec_ok);
  return check_launch("pc_corrector_apply_kernel");
}

int rd_pc_predictor_step(const float* x, const float* score, const float* z, const float* g_table, float dt,
                         float sqrt_dt, float* x_out, float* x_mean_out, size_t B, size_t D, uint64_t seed,
                         uint32_t draw_base, int32_t* step_ctr, size_t noise_step_stride, int advance_ctr,
                         int g_per_sample, void* stream) {
  RD_REQUIRE(x && score && g_table && x_out && B > 0 && D > 0, "rd_pc_predictor_step: bad arguments");
  const size_t n = B * D;
  RD_REQUIRE(!advance_ctr || step_ctr, "rd_pc_predictor_step: advance_ctr needs step_ctr");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int vec_ok = all_aligned16(x, score, x_out, x_mean_out) && all_aligned16(z, nullptr, nullptr, nullptr);
  const size_t half = ((n + 3) / 4 + 1) / 2;
#define RD_PRED(M, P)                                                                                              \
  pc_predictor_kernel<M, P><<<resident_grid(pc_predictor_kernel<M, P>, half, 256), 256, 0, st>>>(x, score, z, g_table, dt, sqrt_dt, x_out, x_mean_out, n, seed, \
                                                  draw_base, step_ctr, noise_step_stride, D, vec_ok)
  if (x_mean_out) { if (g_per_sample) RD_PRED(true, true); else RD_PRED(true, false); }
  else            { if (g_per_sample) RD_PRED(false, true); else RD_PRED(false, false); }
#undef RD_PRED
  int rc = check_launch("pc_predictor_kernel");
  if (rc != RD_OK) return rc;
  if (advance_ctr) {
    step_advance_kernel<<<1, 1, 0, st>>>(step_ctr);
    rc = check_launch("step_advance_kernel");
  }
  return rc;
}

}  // extern "C"
