// attn_tc.cu -- the fused attention block of the bf16 plan on the 5th-generation tensor cores (AttnBlockpp.forward,
// reference models/layerspp.py:80-96: GroupNorm -> q, k, v NIN -> softmax(q k^T / sqrt(C)) v -> output NIN -> (x + h) / sqrt 2).
//
// One CTA (two warpgroups) = one sample at a time, two CTAs per SM.  A sample's T <= 128 tokens are the 128 rows of an
// M = 128 tcgen05 tile: token r is TMEM lane r, and the two threads that may read lane r (thread r of either warpgroup)
// split the accumulator COLUMNS between them, so a query row's softmax is two thread-local halves joined by one
// exchange through shared memory -- no shuffles, no score tile.  Five tensor-core products per sample, all M = 128,
// bf16 x bf16 -> fp32 in TMEM, both operands K-major in shared memory (SWIZZLE_NONE canonical layout, the descriptor form
// of conv_gemm.cu):
//
//   v^T     = [Wv ; Wp] Xn^T       N = TP,  K = 64     TMEM columns 128..128+TP-1, lanes 0..63 (lanes 64..127: unused by-product)
//   [q | k] = Xn  [Wq ; Wk]^T      N = 128, K = 64     TMEM columns   0..127
//   S       = Q K^T                N = TP,  K = 64     TMEM columns   0..TP-1      (q | k are in shared memory by then)
//   O       = P V                  N = 64,  K = TP     TMEM columns   0..63
//   Y       = O Wp^T + X I         N = 64,  K = 64+64  TMEM columns  64..127
//
// (TP = T rounded up to 16.)  The values come out of the tensor core already transposed -- computed as Wv Xn^T -- so that the
// B operand of P V is written with 16-byte stores by the threads that hold the channel rows.  The residual is accumulated
// by the tensor core as well (raw rows times a 64 x 64 identity: exact in fp32), so the last epilogue is bias, scale, store.
// Between two products the accumulator goes TMEM -> registers (tcgen05.ld) -> bias / softmax / rounding -> shared memory as
// the next operand; the activations never leave the SM.  Rows and keys >= T are zero in every operand that feeds a valid
// row (zeroed once, never written), so padded products are exact zeros and only the softmax has to mask.  The next
// sample's rows arrive as ONE bulk copy while this sample is computed.  Rounding points (q, k, v, un-normalised P,
// normalised O in bf16) are those of the mma.sync kernel this one replaces (attn_core.cu, kept for A/B runs).
#include "rd_common.h"
#include "rd_ptx.cuh"
#include <cuda_bf16.h>
#include <cstdlib>

namespace rd {
namespace {

constexpr int TC_C = 64;
constexpr int TC_THREADS = 256;
constexpr uint32_t TC_TMEM_COLS = 256;
// Operand images: element (row, k) at row * 16 B + (k / 8) * LBO + (k % 8) * 2 B.  The k-segment stride LBO is the row count
// times 16 B plus 16 B of padding, so that a quarter-warp writing the eight segments of one row hits eight different banks.
constexpr int W_ROWS = 5 * TC_C;              // [Wq ; Wk ; Wv ; Wp ; I]
constexpr uint32_t LBO_W = W_ROWS * 16 + 16;
constexpr uint32_t LBO_A = 128 * 16 + 16;     // A-side image (Xn, then Q, then P, then O): 128 rows
constexpr uint32_t LBO_V = 64 * 16 + 16;      // v^T: 64 channel rows, TP keys along k

template <int T16>
struct TcLayout {
  static constexpr int TP = 16 * T16;
  static constexpr int NSEG = TP / 8;                   // 8-key segments
  static constexpr int NSEG_A = NSEG > 8 ? NSEG : 8;
  static constexpr uint32_t LBO_K = TP * 16 + 16;       // k image and raw-row image: TP rows (an M = 128 read of the raw-row
  static constexpr uint32_t w_off = 0;                  //  image runs into the bytes that follow: rows >= TP, never used)
  static constexpr uint32_t a_off = w_off + 8 * LBO_W;
  static constexpr uint32_t k_off = a_off + NSEG_A * LBO_A;
  static constexpr uint32_t v_off = k_off + 8 * LBO_K;
  static constexpr uint32_t xr_off = v_off + NSEG * LBO_V;
  static constexpr uint32_t stg_off = xr_off + 8 * LBO_K;
  static constexpr uint32_t stg_bytes = (TP * 128 > 128 * 16 ? TP * 128 : 128 * 16);  // (also the landing zone of the over-read)
  static constexpr uint32_t par_off = stg_off + stg_bytes;
  // floats: bq[64] bp[64] gamma[64] beta[64] | part[8 warps][2][64] | coef[2][64] | mx[2][128] | l[2][128]
  static constexpr uint32_t n_par = 4 * 64 + 8 * 128 + 128 + 256 + 256;
  static constexpr uint32_t bar_off = par_off + n_par * 4;
  static constexpr uint32_t total = bar_off + 32;  // mbarriers: products, input rows
};

__device__ __forceinline__ uint32_t tc_pack(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ float tc_ex2(float x) {
#ifdef RD_EXACT_ACT  // error-budget builds only
  return exp2f(x);
#else
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// TMEM -> registers, 8 consecutive fp32 columns of this thread's lane
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait8(uint32_t (&v)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7])
               :
               : "memory");
}

// K / 16 products  D (+)= A[128 x 16] B[N x 16]^T, k-step images 2 * LBO apart
template <int KSTEPS>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint32_t a_addr, uint32_t a_lbo, uint32_t b_addr, uint32_t b_lbo, uint32_t idesc,
                                       bool accumulate = false) {
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks)
    umma_bf16_ss(tmem_d, umma_desc_kmajor(a_addr + ks * 2 * a_lbo, a_lbo, 128), umma_desc_kmajor(b_addr + ks * 2 * b_lbo, b_lbo, 128), idesc,
                 (ks > 0 || accumulate) ? 1u : 0u);
}

#ifdef RD_TCA_PROF  // per-phase cycle counts of thread 0 of CTA 0, printed at exit (tools/run_tcaprof.sh)
#define TCA_MARK(i) do { const long long t_ = clock64(); prof[i] += t_ - t_last; t_last = t_; } while (0)
#else
#define TCA_MARK(i) do { } while (0)
#endif

template <int T16>
__global__ void __launch_bounds__(TC_THREADS, 2)
    attn_block_tc_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ wqkv_t,  // [192][72]
                         const __nv_bfloat16* __restrict__ wproj_t,                                                                     // [64][72]
                         const float* __restrict__ bqkv, const float* __restrict__ bproj, const float* __restrict__ gamma,
                         const float* __restrict__ beta, int B2, int T, int groups, float eps, float scale, float out_scale) {
  using L = TcLayout<T16>;
  constexpr int TP = L::TP, C = TC_C, NSEG = L::NSEG;
  constexpr int NR = (TP + 31) / 32;    // rows per thread in the GroupNorm pass
  constexpr int NCH = (TP + 31) / 32;   // 32-column TMEM reads that cover the padded key count
  constexpr int SH = (NSEG + 1) / 2;    // key segments of warpgroup 0 in the softmax (warpgroup 1: the rest)
  constexpr int ALL_VALID = TP - 16;    // keys below this index exist for every T this instantiation serves
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint32_t tmem_slot;
  // the conv launch that follows may start its prologue as SMs free up (it waits for this grid before reading `out`)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wg = tid >> 7;   // warpgroup: which half of the accumulator columns this thread handles
  const int r = tid & 127;   // this thread's token = TMEM lane
  unsigned char* Stg = sm + L::stg_off;
  float* s_bq = reinterpret_cast<float*>(sm + L::par_off);
  float* s_bp = s_bq + C;
  float* s_ga = s_bp + C;
  float* s_be = s_ga + C;
  float* s_pt = s_be + C;          // [8][2][C]
  float* s_cf = s_pt + 8 * 2 * C;  // [2][C]: y = x * a + b
  float* s_mx = s_cf + 2 * C;      // [2][128] row maxima of the two column halves
  float* s_l = s_mx + 256;         // [2][128] row sums of the two column halves
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + L::bar_off);
  uint64_t* bar_x = bar + 1;
  const uint32_t Ws = smem_u32(sm + L::w_off), As = smem_u32(sm + L::a_off), Ks = smem_u32(sm + L::k_off), Vs = smem_u32(sm + L::v_off),
                 Xr = smem_u32(sm + L::xr_off);

  auto prefetch = [&](int b) {  // one thread: sample b's T rows (contiguous) -> staging
    mbar_arrive_expect_tx(bar_x, static_cast<uint32_t>(T) * 128u);
    bulk_g2s(Stg, x + static_cast<size_t>(b) * T * C, static_cast<uint32_t>(T) * 128u, bar_x);
  };

  // ------------------------------------------------------------------ setup
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar_x, 1);
    fence_mbar_init();
    if (static_cast<int>(blockIdx.x) < B2) prefetch(blockIdx.x);
  }
  if (warp == 0) tmem_alloc(&tmem_slot, TC_TMEM_COLS);
  for (int i = tid; i < W_ROWS * 8; i += TC_THREADS) {  // weights: row n = output channel of q | k | v | proj, then the identity
    const int n = i >> 3, j = i & 7;
    uint4 w;
    if (n < 4 * C) {
      w = __ldg(reinterpret_cast<const uint4*>(n < 3 * C ? wqkv_t + n * 72 + j * 8 : wproj_t + (n - 3 * C) * 72 + j * 8));
    } else {
      const int d = n - 4 * C - 8 * j;  // position of the 1.0 inside this 8-element segment, if any
      const uint32_t one = 0x3F80u;
      w = make_uint4(d == 0 ? one : d == 1 ? one << 16 : 0u, d == 2 ? one : d == 3 ? one << 16 : 0u, d == 4 ? one : d == 5 ? one << 16 : 0u,
                     d == 6 ? one : d == 7 ? one << 16 : 0u);
    }
    *reinterpret_cast<uint4*>(sm + L::w_off + n * 16 + j * LBO_W) = w;
  }
  // rows / keys >= T of the A-side image and of the k image stay zero for the kernel's lifetime (nothing below writes them)
  for (int i = tid; i < static_cast<int>(L::NSEG_A * LBO_A + 8 * L::LBO_K) / 16; i += TC_THREADS)
    *reinterpret_cast<uint4*>(sm + L::a_off + i * 16) = make_uint4(0u, 0u, 0u, 0u);
  if (tid < C) { s_bq[tid] = bqkv[tid]; s_bp[tid] = bproj[tid]; s_ga[tid] = gamma[tid]; s_be[tid] = beta[tid]; }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;

  const bool row_ok = r < T;
  const bool warp_ok = (warp & 3) * 32 < T;  // the warp holds at least one token
  const uint32_t tm_lane = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  const int seg = tid & 7, rsub = tid >> 3;  // GroupNorm pass: thread = 8-channel segment of rows rsub, rsub + 32, ...
  const int cpg = C / groups;
  const float inv_n = 1.0f / static_cast<float>(cpg * T);
  const float cexp = scale * 1.4426950408889634f;
  const int sg0 = wg ? SH : 0, nsg = wg ? NSEG - SH : SH;  // this thread's key segments in the softmax
  constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128), idesc_tp = umma_idesc_bf16(128, TP), idesc_c = umma_idesc_bf16(128, C);
  uint32_t par = 0, xpar = 0;
#ifdef RD_TCA_PROF
  long long prof[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, t_last = clock64();
  int nsamp = 0;
#endif

  for (int b = blockIdx.x; b < B2; b += gridDim.x) {
    mbar_wait(bar_x, xpar);  // this sample's rows have landed
    xpar ^= 1;
    TCA_MARK(0);

    // ---- GroupNorm statistics (per-channel sums -> per-group affine), then the normalised rows as the A image
    uint4 raw[NR];
    {
      float s1[8], s2[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { s1[j] = 0.0f; s2[j] = 0.0f; }
#pragma unroll
      for (int k = 0; k < NR; ++k) {
        const int row = rsub + 32 * k;
        raw[k] = make_uint4(0u, 0u, 0u, 0u);
        if (row < T) raw[k] = *reinterpret_cast<const uint4*>(Stg + row * 128 + seg * 16);
        const uint32_t w[4] = {raw[k].x, raw[k].y, raw[k].z, raw[k].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float lo = __uint_as_float(w[j] << 16), hi = __uint_as_float(w[j] & 0xffff0000u);
          s1[2 * j] += lo; s2[2 * j] = fmaf(lo, lo, s2[2 * j]);
          s1[2 * j + 1] += hi; s2[2 * j + 1] = fmaf(hi, hi, s2[2 * j + 1]);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 8); s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 16);
        s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 8); s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 16);
      }
      if (lane < 8) {
        float4* d1 = reinterpret_cast<float4*>(s_pt + warp * 2 * C + seg * 8);
        float4* d2 = reinterpret_cast<float4*>(s_pt + warp * 2 * C + C + seg * 8);
        d1[0] = make_float4(s1[0], s1[1], s1[2], s1[3]); d1[1] = make_float4(s1[4], s1[5], s1[6], s1[7]);
        d2[0] = make_float4(s2[0], s2[1], s2[2], s2[3]); d2[1] = make_float4(s2[4], s2[5], s2[6], s2[7]);
      }
    }
    __syncthreads();
    // every thread holds its rows in registers: the staging buffer is free for the next sample
    if (tid == 0 && b + static_cast<int>(gridDim.x) < B2) prefetch(b + gridDim.x);
    if (tid < C) {  // warps 0 and 1: channel tid; its group = cpg adjacent lanes (cpg is a power of two: it divides 64)
      float a1 = 0.0f, a2 = 0.0f;
#pragma unroll
      for (int w = 0; w < 8; ++w) { a1 += s_pt[w * 2 * C + tid]; a2 += s_pt[w * 2 * C + C + tid]; }
      for (int o = 1; o < cpg; o <<= 1) { a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o); }
      const float mean = a1 * inv_n;
      const float var = fmaxf(a2 * inv_n - mean * mean, 0.0f);
      const float ca = s_ga[tid] * rsqrtf(var + eps);
      s_cf[tid] = ca;
      s_cf[C + tid] = fmaf(-mean, ca, s_be[tid]);
    }
    __syncthreads();
    TCA_MARK(1);
    {
      const float4 a0 = *reinterpret_cast<const float4*>(s_cf + seg * 8), a1 = *reinterpret_cast<const float4*>(s_cf + seg * 8 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(s_cf + C + seg * 8), b1 = *reinterpret_cast<const float4*>(s_cf + C + seg * 8 + 4);
      const float ca[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float cb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int k = 0; k < NR; ++k) {
        const int row = rsub + 32 * k;
        if (row < T) {
          const uint32_t w[4] = {raw[k].x, raw[k].y, raw[k].z, raw[k].w};
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            o[j] = tc_pack(fmaf(__uint_as_float(w[j] << 16), ca[2 * j], cb[2 * j]),
                           fmaf(__uint_as_float(w[j] & 0xffff0000u), ca[2 * j + 1], cb[2 * j + 1]));
          st_shared_v4(As + row * 16 + seg * LBO_A, o[0], o[1], o[2], o[3]);
          st_shared_v4(Xr + row * 16 + seg * L::LBO_K, w[0], w[1], w[2], w[3]);  // raw rows: the residual operand
        }
      }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    TCA_MARK(2);
    if (tid == 0) {
      tc_fence_after_sync();
      tc_mma<4>(tmem + 128, Ws + 2 * C * 16, LBO_W, As, LBO_A, idesc_tp);  // v^T (+ by-product rows) = [Wv ; Wp] Xn^T
      tc_mma<4>(tmem, As, LBO_A, Ws, LBO_W, idesc_qk);                     // [q | k] = Xn [Wq ; Wk]^T
      umma_commit(bar);
    }
    mbar_wait(bar, par);
    TCA_MARK(3);
    par ^= 1;
    tc_fence_after_sync();

    // ---- warpgroup 0: q (+ bias) -> A image; warpgroup 1: k -> B image; both: v^T -> B image of P V.  No bias on k and v:
    // a key bias adds the same q.b_k to every score of a row, which softmax cancels, and the value bias passes through
    // P V unchanged -- the host folds it into the projection bias (rdb200/pack.py `proj.bias_fused`).
    if (warp_ok) {
      uint32_t v[2][32];
      tmem_ld32(tm_lane + 64 * wg, v[0]);
      tmem_ld32(tm_lane + 64 * wg + 32, v[1]);
      tmem_ld_wait32(v[0]);
      tmem_ld_wait32(v[1]);
      if (row_ok) {
        const uint32_t dst = wg ? Ks + r * 16 : As + r * 16;
        const uint32_t lbo = wg ? L::LBO_K : LBO_A;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = 32 * h + 8 * j + 2 * e;
              float f0 = __uint_as_float(v[h][8 * j + 2 * e]), f1 = __uint_as_float(v[h][8 * j + 2 * e + 1]);
              if (wg == 0) { f0 += s_bq[c]; f1 += s_bq[c + 1]; }
              o[e] = tc_pack(f0, f1);
            }
            st_shared_v4(dst + (4 * h + j) * lbo, o[0], o[1], o[2], o[3]);
          }
      }
    }
    if ((warp & 3) < 2) {  // TMEM lanes 0..63 = channels; 32-key chunk h belongs to warpgroup h & 1
#pragma unroll
      for (int h = 0; h < NCH; ++h) {
        if ((h & 1) == wg) {
          uint32_t v[32];
          tmem_ld32(tm_lane + 128 + 32 * h, v);
          tmem_ld_wait32(v);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int sg = 4 * h + j;
            if (sg < NSEG) {
              uint32_t o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) o[e] = tc_pack(__uint_as_float(v[8 * j + 2 * e]), __uint_as_float(v[8 * j + 2 * e + 1]));
              st_shared_v4(Vs + r * 16 + sg * LBO_V, o[0], o[1], o[2], o[3]);
            }
          }
        }
      }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    TCA_MARK(4);
    if (tid == 0) {
      tc_fence_after_sync();
      tc_mma<4>(tmem, As, LBO_A, Ks, L::LBO_K, idesc_tp);  // S = Q K^T
      umma_commit(bar);
    }
    mbar_wait(bar, par);
    TCA_MARK(5);
    par ^= 1;
    tc_fence_after_sync();

    // ---- softmax_j(scale * s_ij) over the T valid keys: maximum on the raw scores (scale > 0), exp(scale (s - m)) as one
    // FMA + ex2; the probabilities stay un-normalised (1 / l is applied to O).  This thread: key segments sg0 .. sg0+nsg-1.
    uint32_t sv[SH][8];
    if (warp_ok) {
#pragma unroll
      for (int j = 0; j < SH; ++j)
        if (j < nsg) tmem_ld8(tm_lane + 8 * (sg0 + j), sv[j]);
#pragma unroll
      for (int j = 0; j < SH; ++j)
        if (j < nsg) tmem_ld_wait8(sv[j]);
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // four independent chains
#pragma unroll
      for (int j = 0; j < SH; ++j)
        if (j < nsg) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int key = 8 * (sg0 + j) + e;
            const bool ok = (8 * (SH + j) + e < ALL_VALID) || key < T;  // (compile-time true except in the last segments)
            m4[e & 3] = fmaxf(m4[e & 3], ok ? __uint_as_float(sv[j][e]) : -INFINITY);
          }
        }
      s_mx[wg * 128 + r] = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
    }
    __syncthreads();
    if (warp_ok) {
      const float m = fmaxf(s_mx[r], s_mx[128 + r]);
      const float mc = -m * cexp;
      float l4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int j = 0; j < SH; ++j)
        if (j < nsg) {
          float p[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int key = 8 * (sg0 + j) + e;
            const bool ok = (8 * (SH + j) + e < ALL_VALID) || key < T;
            p[e] = ok ? tc_ex2(fmaf(__uint_as_float(sv[j][e]), cexp, mc)) : 0.0f;
            l4[e & 3] += p[e];
          }
          if (row_ok) st_shared_v4(As + r * 16 + (sg0 + j) * LBO_A, tc_pack(p[0], p[1]), tc_pack(p[2], p[3]), tc_pack(p[4], p[5]), tc_pack(p[6], p[7]));
        }
      s_l[wg * 128 + r] = (l4[0] + l4[1]) + (l4[2] + l4[3]);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    TCA_MARK(6);
    if (tid == 0) {
      tc_fence_after_sync();
      tc_mma<T16>(tmem, As, LBO_A, Vs, LBO_V, idesc_c);  // O = P V
      umma_commit(bar);
    }
    mbar_wait(bar, par);
    TCA_MARK(7);
    par ^= 1;
    tc_fence_after_sync();

    // ---- normalised attention output -> A image of the output projection (32 channels per thread)
    if (warp_ok) {
      const float il = 1.0f / (s_l[r] + s_l[128 + r]);
      uint32_t v[32];
      tmem_ld32(tm_lane + 32 * wg, v);
      tmem_ld_wait32(v);
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = tc_pack(__uint_as_float(v[8 * j + 2 * e]) * il, __uint_as_float(v[8 * j + 2 * e + 1]) * il);
          st_shared_v4(As + r * 16 + (4 * wg + j) * LBO_A, o[0], o[1], o[2], o[3]);
        }
      }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    TCA_MARK(8);
    if (tid == 0) {
      tc_fence_after_sync();
      tc_mma<4>(tmem + C, As, LBO_A, Ws + 3 * C * 16, LBO_W, idesc_c);               // Y = O Wp^T
      tc_mma<4>(tmem + C, Xr, L::LBO_K, Ws + 4 * C * 16, LBO_W, idesc_c, true);      //   + X I  (the residual)
      umma_commit(bar);
    }
    mbar_wait(bar, par);
    TCA_MARK(9);
    par ^= 1;
    tc_fence_after_sync();

    // ---- (x + y + bias) * out_scale: 32 channels = 64 bytes of this thread's output row
    if (warp_ok) {
      uint32_t v[32];
      tmem_ld32(tm_lane + C + 32 * wg, v);
      tmem_ld_wait32(v);
      if (row_ok) {
        __nv_bfloat16* orow = out + (static_cast<size_t>(b) * T + r) * C + 32 * wg;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          u32x8 o;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = 32 * wg + 16 * j + 2 * e;
            o.v[e] = tc_pack((__uint_as_float(v[16 * j + 2 * e]) + s_bp[c]) * out_scale, (__uint_as_float(v[16 * j + 2 * e + 1]) + s_bp[c + 1]) * out_scale);
          }
          st_global_256(orow + 16 * j, o);
        }
      }
    }
    TCA_MARK(10);
#ifdef RD_TCA_PROF
    ++nsamp;
#endif
  }

#ifdef RD_TCA_PROF
  if (blockIdx.x == 0 && tid == 0 && nsamp > 0)
    printf("attn_tc T=%d samples=%d cycles/sample: wait_x %lld | stats+coef %lld | xn %lld | mma_qkv %lld | epi_qkv %lld | mma_s %lld | softmax %lld | mma_o %lld | epi_o %lld | mma_y %lld | epi_y %lld\n",
           T, nsamp, prof[0] / nsamp, prof[1] / nsamp, prof[2] / nsamp, prof[3] / nsamp, prof[4] / nsamp, prof[5] / nsamp, prof[6] / nsamp,
           prof[7] / nsamp, prof[8] / nsamp, prof[9] / nsamp, prof[10] / nsamp);
#endif
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TC_TMEM_COLS);
}

}  // namespace

// RD_E_UNSUPPORTED when the shape is outside this kernel (the caller falls back to the mma.sync kernel)
int attn_block_tc_launch(const rd_op_attn_block& op, cudaStream_t st) {
  if (op.C != TC_C || op.T < 1 || op.T > 128) return RD_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(op.x) | reinterpret_cast<uintptr_t>(op.out)) & 31) return RD_E_UNSUPPORTED;
  const float scale = 1.0f / sqrtf(static_cast<float>(op.C));
  const int t16 = (op.T + 15) / 16;
  int grid = 2 * kNumSMs;
  if (grid > op.B2) grid = op.B2;
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(op.x);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(op.out);
  const __nv_bfloat16* wq = static_cast<const __nv_bfloat16*>(op.wqkv_t);
  const __nv_bfloat16* wp = static_cast<const __nv_bfloat16*>(op.wproj_t);
  static bool configured_dev[64][9] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  bool* configured = configured_dev[dev];
  switch (t16) {
#define RD_TC_CASE(n)                                                                                                                   \
  case n:                                                                                                                               \
    if (!configured[n]) {                                                                                                               \
      cudaError_t e = cudaFuncSetAttribute(attn_block_tc_kernel<n>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcLayout<n>::total);   \
      if (e != cudaSuccess) return fail(static_cast<int>(e), "attn_block (tcgen05): %s", cudaGetErrorString(e));                        \
      configured[n] = true;                                                                                                             \
    }                                                                                                                                   \
    attn_block_tc_kernel<n><<<grid, TC_THREADS, TcLayout<n>::total, st>>>(x, out, wq, wp, op.bqkv, op.bproj, op.gamma, op.beta, op.B2,  \
                                                                           op.T, op.groups, op.eps, scale, op.out_scale);               \
    break;
    RD_TC_CASE(1) RD_TC_CASE(2) RD_TC_CASE(3) RD_TC_CASE(4) RD_TC_CASE(5) RD_TC_CASE(6) RD_TC_CASE(7) RD_TC_CASE(8)
#undef RD_TC_CASE
    default: return RD_E_UNSUPPORTED;
  }
  return check_launch("attn_block_tc_kernel");
}

}  // namespace rd
