// attn_tc.cu -- the fused attention block of the bf16 plan on the 5th-generation tensor cores (AttnBlockpp.forward,
// reference models/layerspp.py:80-96: GroupNorm -> q, k, v NIN -> softmax(q k^T / sqrt(C)) v -> output NIN -> (x + h) / sqrt 2).
//
// One CTA per SM runs NP independent sample pipelines, one per warpgroup (NP = 4 for T <= 80 tokens); the pipelines share
// only the weights in shared memory.  A sample's T <= 128 tokens are the 128 rows of an M = 128 tcgen05 tile: thread r of
// the warpgroup owns token r, which is also TMEM lane r, so a query row's softmax (max, exp, sum) is thread-local -- no
// shuffles, no score tile.  Four tensor-core products per sample, all M = 128, bf16 x bf16 -> fp32 in the pipeline's 128
// TMEM columns, both operands K-major in shared memory (SWIZZLE_NONE canonical layout, the descriptor form of conv_gemm.cu):
//
//   u^T     = [Wpv ; 0] Xn^T       N = TP,  K = 64     columns 0..TP-1, lanes 0..63      Wpv = Wp Wv (64 x 64)
//   [q | k] = Xn  [Wq ; Wk]^T      N = 128, K = 64     columns 0..127
//   S       = Q K^T                N = TP,  K = 64     columns 0..TP-1
//   Y       = P U                  N = 64,  K = TP     columns 0..63                      P = exp(scale (s - max)), rows normalised in the epilogue
//
// (TP = T rounded up to 16.)  The output projection is folded into the value projection -- softmax(.) (Xn Wv^T) Wp^T =
// softmax(.) (Xn (Wp Wv)^T) -- which removes one product, one accumulator read and one operand write per sample; the
// 64 x 64 product Wp Wv is formed once per CTA in fp32.  u comes out of the tensor core already transposed (computed as
// Wpv Xn^T), so the B operand of the last product is written with 16-byte stores by the threads that hold the channel
// rows.  Between two products the accumulator goes TMEM -> registers (tcgen05.ld) -> bias / softmax / rounding -> shared
// memory as the next operand; the activations never leave the SM.  What bounds the kernel is the TMEM read port (64 B per
// clock per SM, ~125 KB of accumulator per sample); the pipelines exist to keep that port busy while each of them waits
// for its own product / barrier round trips.  Rows >= T of the A-side image are zero (zeroed once, never written), so u
// is zero for keys >= T; score columns >= T are masked by the softmax.  Shared memory per pipeline: the A-side image (Xn,
// then Q, then Pn), the u^T image, and the k image, which doubles as the landing buffer of the next sample's rows (one
// bulk copy, issued as soon as S = Q K^T has consumed the keys).  The residual is re-read from global memory (L2) by the
// thread that owns the row.  Rounding points: q, k, u, P (un-normalised, <= 1) in bf16 (the mma.sync kernel in attn_core.cu, kept for A/B runs
// and for shapes this kernel does not cover, rounds v, P and the normalised attention output instead).
#include "rd_common.h"
#include "rd_ptx.cuh"
#include <cuda_bf16.h>
#include <cstdlib>

namespace rd {
namespace {

constexpr int TC_C = 64;
// Operand images: element (row, k) at row * 16 B + (k / 8) * LBO + (k % 8) * 2 B.  The k-segment stride LBO is the row count
// times 16 B plus 16 B of padding, so that a quarter-warp writing the eight segments of one row hits eight different banks.
constexpr int W_ROWS = 4 * TC_C;              // [Wq ; Wk ; Wpv ; 0]
constexpr uint32_t LBO_W = W_ROWS * 16 + 16;
constexpr uint32_t LBO_A = 128 * 16 + 16;     // A-side image: 128 rows
constexpr uint32_t LBO_V = 64 * 16 + 16;      // u^T: 64 channel rows, TP keys along k

template <int T16>
struct TcLayout {
  static constexpr int TP = 16 * T16;
  static constexpr int NSEG = TP / 8;                   // 8-key segments
  static constexpr int NSEG_A = NSEG > 8 ? NSEG : 8;
  static constexpr uint32_t LBO_K = TP * 16 + 16;       // k image: TP rows
  // per pipeline: A image | k image (= input landing buffer, T x 128 B) | u^T image | floats part[4 warps][2][64]
  static constexpr uint32_t a_off = 0;
  static constexpr uint32_t k_off = a_off + NSEG_A * LBO_A;
  static constexpr uint32_t v_off = k_off + 8 * LBO_K;
  static constexpr uint32_t f_off = v_off + NSEG * LBO_V;
  static constexpr uint32_t pipe_bytes = f_off + 4 * 128 * 4;
  static constexpr uint32_t w_bytes = 8 * LBO_W;
  static constexpr uint32_t shared_floats = 4 * 64;  // bq bp gamma beta
  static constexpr uint32_t fixed_bytes = w_bytes + shared_floats * 4 + 128;  // + mbarriers
  static constexpr int NP_FIT = (227 * 1024 - static_cast<int>(fixed_bytes)) / static_cast<int>(pipe_bytes);
  static constexpr int NP = NP_FIT >= 4 ? 4 : NP_FIT;  // pipelines = warpgroups per CTA
  static_assert(NP >= 1, "attention tile does not fit");
  static constexpr uint32_t pipes_off = w_bytes;
  static constexpr uint32_t par_off = pipes_off + NP * pipe_bytes;
  static constexpr uint32_t bar_off = par_off + shared_floats * 4;
  static constexpr uint32_t total = bar_off + 128;
  static constexpr uint32_t tmem_cols = NP > 2 ? 512 : NP * 128;
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

// 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// TP consecutive columns -> v[TP] (32-column reads, then a 16-column one when TP % 32 == 16); one wait at the end
template <int TP>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[TP]) {
#pragma unroll
  for (int h = 0; h < TP / 32; ++h) tmem_ld32(taddr + 32 * h, *reinterpret_cast<uint32_t(*)[32]>(&v[32 * h]));
  if constexpr (TP % 32 != 0) tmem_ld16(taddr + (TP / 32) * 32, *reinterpret_cast<uint32_t(*)[16]>(&v[(TP / 32) * 32]));
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < TP; ++i) asm volatile("" : "+r"(v[i]));  // no use of v may be scheduled above the wait
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
__global__ void __launch_bounds__(128 * TcLayout<T16>::NP, 1)
    attn_block_tc_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ wqkv_t,  // [192][72]
                         const __nv_bfloat16* __restrict__ wproj_t,                                                                     // [64][72]
                         const float* __restrict__ bqkv, const float* __restrict__ bproj, const float* __restrict__ gamma,
                         const float* __restrict__ beta, int B2, int T, int groups, float eps, float scale, float out_scale) {
  using L = TcLayout<T16>;
  constexpr int TP = L::TP, C = TC_C, NSEG = L::NSEG, NP = L::NP, NTHR = 128 * NP;
  constexpr int ALL_VALID = TP - 16;  // keys below this index exist for every T this instantiation serves
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint32_t tmem_slot;
  // the conv launch that follows may start its prologue as SMs free up (it waits for this grid before reading `out`)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int tid = threadIdx.x, lane = tid & 31;
  const int pipe = tid >> 7;      // this thread's pipeline (warpgroup)
  const int r = tid & 127;        // this thread's token = TMEM lane
  const int warp = r >> 5;        // warp within the warpgroup = TMEM lane quarter
  unsigned char* P0 = sm + L::pipes_off + pipe * L::pipe_bytes;
  unsigned char* Stg = P0 + L::k_off;  // the next sample's rows land in the k image once S = Q K^T is done
  float* s_ab = reinterpret_cast<float*>(P0 + L::f_off);  // [2][C]: GroupNorm affine (a, b) of the pipeline's current / next sample
  float* s_bq = reinterpret_cast<float*>(sm + L::par_off);
  float* s_bp = s_bq + C;
  float* s_ga = s_bp + C;
  float* s_be = s_ga + C;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + L::bar_off) + 2 * pipe;  // this pipeline's products
  uint64_t* bar_x = bar + 1;                                                //                 input rows
  const uint32_t Ws = smem_u32(sm), As = smem_u32(P0 + L::a_off), Ks = smem_u32(P0 + L::k_off), Vs = smem_u32(P0 + L::v_off);

  auto pipe_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + pipe) : "memory"); };
  auto prefetch = [&](int b) {  // one thread: sample b's T rows (contiguous in global memory) -> landing buffer
    mbar_arrive_expect_tx(bar_x, static_cast<uint32_t>(T) * 128u);
    bulk_g2s(Stg, x + static_cast<size_t>(b) * T * C, static_cast<uint32_t>(T) * 128u, bar_x);
  };
  const int b0 = blockIdx.x * NP + pipe, bstep = gridDim.x * NP;

  // ------------------------------------------------------------------ setup
  if (r == 0) {
    mbar_init(bar, 1);
    mbar_init(bar_x, 1);
    fence_mbar_init();
    if (b0 < B2) prefetch(b0);
  }
  if (tid < 32) tmem_alloc(&tmem_slot, L::tmem_cols);
  for (int i = tid; i < 2 * C * 8; i += NTHR) {  // Wq ; Wk: row n = output channel, 8 input channels per 16-byte segment
    const int n = i >> 3, j = i & 7;
    *reinterpret_cast<uint4*>(sm + n * 16 + j * LBO_W) = __ldg(reinterpret_cast<const uint4*>(wqkv_t + n * 72 + j * 8));
  }
  for (int i = tid; i < C * 8; i += NTHR) {  // Wpv = Wp Wv (fp32 accumulation, then bf16): thread = output row o, input segment j
    const int o = i >> 3, j = i & 7;
    float acc[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    for (int c = 0; c < C; ++c) {
      const float wp = __bfloat162float(wproj_t[o * 72 + c]);
      const uint4 wv = __ldg(reinterpret_cast<const uint4*>(wqkv_t + (2 * C + c) * 72 + j * 8));
      const uint32_t w[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc[2 * e] = fmaf(wp, __uint_as_float(w[e] << 16), acc[2 * e]);
        acc[2 * e + 1] = fmaf(wp, __uint_as_float(w[e] & 0xffff0000u), acc[2 * e + 1]);
      }
    }
    *reinterpret_cast<uint4*>(sm + (2 * C + o) * 16 + j * LBO_W) =
        make_uint4(tc_pack(acc[0], acc[1]), tc_pack(acc[2], acc[3]), tc_pack(acc[4], acc[5]), tc_pack(acc[6], acc[7]));
    *reinterpret_cast<uint4*>(sm + (3 * C + o) * 16 + j * LBO_W) = make_uint4(0u, 0u, 0u, 0u);  // rows 192..255: M = 128 padding
  }
  // rows >= T of the A-side image stay zero for the kernel's lifetime (nothing below writes them)
  for (int i = r; i < static_cast<int>(L::k_off) / 16; i += 128) *reinterpret_cast<uint4*>(P0 + i * 16) = make_uint4(0u, 0u, 0u, 0u);
  if (tid < C) { s_bq[tid] = bqkv[tid]; s_bp[tid] = bproj[tid] * out_scale; s_ga[tid] = gamma[tid]; s_be[tid] = beta[tid]; }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();  // the only CTA-wide barrier: from here on the pipelines run independently
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot + 128 * pipe;

  const bool row_ok = r < T;
  const bool warp_ok = warp * 32 < T;  // the warp holds at least one token
  const uint32_t tm_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const int seg = r & 7, rsub = r >> 3;  // GroupNorm pass: thread = 8-channel segment of rows rsub, rsub + 16, ...
  const int cpg = C / groups;            // a power of two (it divides 64)
  const float inv_n = 1.0f / static_cast<float>(cpg * T);
  const float cexp = scale * 1.4426950408889634f;
  constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128), idesc_tp = umma_idesc_bf16(128, TP), idesc_c = umma_idesc_bf16(128, C);
  uint32_t par = 0, xpar = 0;
  auto product_done = [&]() {  // every thread of the pipeline: wait for the products committed last
    mbar_wait(bar, par);
    par ^= 1;
    tc_fence_after_sync();
  };
  auto operands_ready = [&]() {  // shared-memory operands written, accumulator columns read: hand over to the issuing thread
    fence_proxy_async_smem();
    tc_fence_before_sync();
    pipe_sync();
  };
  // GroupNorm statistics of the sample whose rows are landing, by warp 3 alone: with T <= 96 tokens this warp owns no query
  // row, so the statistics run on the otherwise idle fourth scheduler, beside the softmax / last product / epilogue of the
  // sample before -- off the pipeline's critical path.  Lane = (8-channel segment, row slot 0..3); result: the table
  // s_ab[0..63] = a, s_ab[64..127] = b with y = x a + b.
  auto sample_stats = [&]() {
    mbar_wait(bar_x, xpar);
    xpar ^= 1;
    const int sseg = lane & 7, rs4 = lane >> 3;
    float c1[8], c2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { c1[j] = 0.0f; c2[j] = 0.0f; }
#pragma unroll 4
    for (int row = rs4; row < T; row += 4) {
      const uint4 raw = *reinterpret_cast<const uint4*>(Stg + row * 128 + sseg * 16);
      const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float lo = __uint_as_float(w[j] << 16), hi = __uint_as_float(w[j] & 0xffff0000u);
        c1[2 * j] += lo; c2[2 * j] = fmaf(lo, lo, c2[2 * j]);
        c1[2 * j + 1] += hi; c2[2 * j + 1] = fmaf(hi, hi, c2[2 * j + 1]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      c1[j] += __shfl_xor_sync(0xffffffffu, c1[j], 8); c1[j] += __shfl_xor_sync(0xffffffffu, c1[j], 16);
      c2[j] += __shfl_xor_sync(0xffffffffu, c2[j], 8); c2[j] += __shfl_xor_sync(0xffffffffu, c2[j], 16);
    }
    // group sums: butterflies inside the lane's 8 channels, then across the lanes that hold the group's other segments
    if (cpg >= 2) {
      float t1[8], t2[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { t1[j] = c1[j] + c1[j ^ 1]; t2[j] = c2[j] + c2[j ^ 1]; }
#pragma unroll
      for (int j = 0; j < 8; ++j) { c1[j] = t1[j]; c2[j] = t2[j]; }
    }
    if (cpg >= 4) {
      float t1[8], t2[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { t1[j] = c1[j] + c1[j ^ 2]; t2[j] = c2[j] + c2[j ^ 2]; }
#pragma unroll
      for (int j = 0; j < 8; ++j) { c1[j] = t1[j]; c2[j] = t2[j]; }
    }
    if (cpg >= 8) {
      float t1[8], t2[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { t1[j] = c1[j] + c1[j ^ 4]; t2[j] = c2[j] + c2[j ^ 4]; }
#pragma unroll
      for (int j = 0; j < 8; ++j) { c1[j] = t1[j]; c2[j] = t2[j]; }
    }
    for (int o = 1; o < cpg / 8; o <<= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { c1[j] += __shfl_xor_sync(0xffffffffu, c1[j], o); c2[j] += __shfl_xor_sync(0xffffffffu, c2[j], o); }
    }
    if (lane < 8) {
      const float4 g0 = *reinterpret_cast<const float4*>(s_ga + sseg * 8), g1 = *reinterpret_cast<const float4*>(s_ga + sseg * 8 + 4);
      const float4 e0 = *reinterpret_cast<const float4*>(s_be + sseg * 8), e1 = *reinterpret_cast<const float4*>(s_be + sseg * 8 + 4);
      const float ga[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float be[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
      float ca[8], cb[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float mean = c1[j] * inv_n;
        const float var = fmaxf(c2[j] * inv_n - mean * mean, 0.0f);
        ca[j] = ga[j] * rsqrtf(var + eps);
        cb[j] = fmaf(-mean, ca[j], be[j]);
      }
      float4* da = reinterpret_cast<float4*>(s_ab + sseg * 8);
      float4* db = reinterpret_cast<float4*>(s_ab + C + sseg * 8);
      da[0] = make_float4(ca[0], ca[1], ca[2], ca[3]); da[1] = make_float4(ca[4], ca[5], ca[6], ca[7]);
      db[0] = make_float4(cb[0], cb[1], cb[2], cb[3]); db[1] = make_float4(cb[4], cb[5], cb[6], cb[7]);
    }
  };
  const bool stats_async = warp == 3 && !warp_ok;  // warp 3 holds no token: its statistics overlap the sample before
  if (warp == 3 && b0 < B2) sample_stats();
#ifdef RD_TCA_PROF
  long long prof[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, t_last = clock64();
  int nsamp = 0;
#endif

  for (int b = b0; b < B2; b += bstep) {
    if (warp != 3) {  // (warp 3 waited for these rows when it formed their statistics)
      mbar_wait(bar_x, xpar);  // this sample's rows have landed
      xpar ^= 1;
    }
    TCA_MARK(0);
    pipe_sync();  // the affine table of this sample is complete
    TCA_MARK(1);
    // ---- GroupNorm: y = x a + b with the per-channel (a, b) of this sample; thread = 8-channel segment of rows rsub, rsub + 16, ...
    {
      const float4 a0 = *reinterpret_cast<const float4*>(s_ab + seg * 8), a1 = *reinterpret_cast<const float4*>(s_ab + seg * 8 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(s_ab + C + seg * 8), b1 = *reinterpret_cast<const float4*>(s_ab + C + seg * 8 + 4);
      const float ca[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float cb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int k = 0; k < T16; ++k) {
        const int row = rsub + 16 * k;
        if (row < T) {
          const uint4 raw = *reinterpret_cast<const uint4*>(Stg + row * 128 + seg * 16);
          const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            o[j] = tc_pack(fmaf(__uint_as_float(w[j] << 16), ca[2 * j], cb[2 * j]),
                           fmaf(__uint_as_float(w[j] & 0xffff0000u), ca[2 * j + 1], cb[2 * j + 1]));
          st_shared_v4(As + row * 16 + seg * LBO_A, o[0], o[1], o[2], o[3]);
        }
      }
    }
    operands_ready();
    TCA_MARK(2);
    if (r == 0) {
      tc_fence_after_sync();
      tc_mma<4>(tmem, Ws + 2 * C * 16, LBO_W, As, LBO_A, idesc_tp);  // u^T = [Wpv ; 0] Xn^T
      umma_commit(bar);
    }
    product_done();
    TCA_MARK(3);

    // ---- u^T -> B image of the last product (TMEM lanes 0..63 = channels).  No bias on k and v: a key bias adds the same
    // q.b_k to every score of a row, which softmax cancels, and the value bias passes through the normalised probabilities
    // unchanged -- the host folds it into the projection bias (rdb200/pack.py `proj.bias_fused`).
    if (warp < 2) {
      uint32_t v[TP];
      tmem_ld_cols<TP>(tm_lane, v);
#pragma unroll
      for (int sg = 0; sg < NSEG; ++sg) {
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = tc_pack(__uint_as_float(v[8 * sg + 2 * e]), __uint_as_float(v[8 * sg + 2 * e + 1]));
        st_shared_v4(Vs + r * 16 + sg * LBO_V, o[0], o[1], o[2], o[3]);
      }
    }
    tc_fence_before_sync();
    pipe_sync();
    TCA_MARK(4);
    if (r == 0) {
      tc_fence_after_sync();
      tc_mma<4>(tmem, As, LBO_A, Ws, LBO_W, idesc_qk);  // [q | k] = Xn [Wq ; Wk]^T
      umma_commit(bar);
    }
    product_done();
    TCA_MARK(5);

    // ---- q (+ bias) -> A image, k -> B image (over the landing buffer: every thread has its raw rows in registers... no longer needed)
    if (warp_ok) {
#pragma unroll
      for (int qk = 0; qk < 2; ++qk) {
        uint32_t v[64];
        tmem_ld_cols<64>(tm_lane + 64 * qk, v);
        if (row_ok) {
#pragma unroll
          for (int sg = 0; sg < 8; ++sg) {
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = 8 * sg + 2 * e;
              float f0 = __uint_as_float(v[c]), f1 = __uint_as_float(v[c + 1]);
              if (qk == 0) { f0 += s_bq[c]; f1 += s_bq[c + 1]; }
              o[e] = tc_pack(f0, f1);
            }
            st_shared_v4(qk == 0 ? As + r * 16 + sg * LBO_A : Ks + r * 16 + sg * L::LBO_K, o[0], o[1], o[2], o[3]);
          }
        }
      }
    }
    operands_ready();
    TCA_MARK(6);
    if (r == 0) {
      tc_fence_after_sync();
      tc_mma<4>(tmem, As, LBO_A, Ks, L::LBO_K, idesc_tp);  // S = Q K^T
      umma_commit(bar);
    }
    // the residual: this thread's raw row, again (the bulk copy left it in L2), in flight while the products run.  (Requested
    // after the softmax instead -- 32 fewer live registers there -- it was measured 3 % slower: 0.285 vs 0.276 ms per block.)
    u32x8 xres[4];
    if (row_ok) {
      const __nv_bfloat16* xrow = x + (static_cast<size_t>(b) * T + r) * C;
#pragma unroll
      for (int j = 0; j < 4; ++j) xres[j] = ld_global_256(xrow + 16 * j);
    }
    product_done();
    TCA_MARK(7);
    // the keys are consumed: their image is the landing buffer of the next sample's rows
    const bool more = b + bstep < B2;
    if (r == 0 && more) prefetch(b + bstep);
    if (warp == 3 && !stats_async && more) sample_stats();  // (T > 96: warp 3 has query rows of its own, no overlap)

    // ---- softmax_j(scale * s_ij) over the T valid keys of this thread's row: maximum on the raw scores (scale > 0),
    // exp(scale (s - m)) as one FMA + ex2; the row sum (of the unrounded values) divides Y in the epilogue
    float y_scale = 0.0f;
    if (warp_ok) {
      uint32_t v[TP];
      tmem_ld_cols<TP>(tm_lane, v);
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // four independent chains
#pragma unroll
      for (int i = 0; i < TP; ++i) m4[i & 3] = fmaxf(m4[i & 3], (i < ALL_VALID || i < T) ? __uint_as_float(v[i]) : -INFINITY);
      const float mc = -fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * cexp;
      float l4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int i = 0; i < TP; ++i) {
        const float p = (i < ALL_VALID || i < T) ? tc_ex2(fmaf(__uint_as_float(v[i]), cexp, mc)) : 0.0f;
        l4[i & 3] += p;
        v[i] = __float_as_uint(p);
      }
      y_scale = out_scale / ((l4[0] + l4[1]) + (l4[2] + l4[3]));  // 1 / row sum: applied to Y in the epilogue
      if (row_ok) {
#pragma unroll
        for (int sg = 0; sg < NSEG; ++sg) {
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = tc_pack(__uint_as_float(v[8 * sg + 2 * e]), __uint_as_float(v[8 * sg + 2 * e + 1]));
          st_shared_v4(As + r * 16 + sg * LBO_A, o[0], o[1], o[2], o[3]);
        }
      }
    }
    if (stats_async) {
      // hand over without waiting (this warp wrote no operand), then the next sample's statistics while the others run
      // the last product and the epilogue; the table is published by the pipeline barrier at the top of the next sample
      fence_proxy_async_smem();
      tc_fence_before_sync();
      asm volatile("bar.arrive %0, 128;" ::"r"(1 + pipe) : "memory");
      if (more) sample_stats();
    } else {
      operands_ready();
    }
    TCA_MARK(8);
    if (r == 0) {
      tc_fence_after_sync();
      tc_mma<T16>(tmem, As, LBO_A, Vs, LBO_V, idesc_c);  // Y = P U
      umma_commit(bar);
    }
    product_done();
    TCA_MARK(9);

    // ---- (x + y / l + bias) * out_scale -> this thread's output row (4 x 32 bytes)
    if (warp_ok) {
      uint32_t v[64];
      tmem_ld_cols<64>(tm_lane, v);
      if (row_ok) {
        __nv_bfloat16* orow = out + (static_cast<size_t>(b) * T + r) * C;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          u32x8 o;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = 16 * j + 2 * e;
            const uint32_t xr = xres[j].v[e];
            o.v[e] = tc_pack(fmaf(__uint_as_float(xr << 16), out_scale, fmaf(__uint_as_float(v[c]), y_scale, s_bp[c])),
                             fmaf(__uint_as_float(xr & 0xffff0000u), out_scale, fmaf(__uint_as_float(v[c + 1]), y_scale, s_bp[c + 1])));
          }
          st_global_256(orow + 16 * j, o);
        }
      }
    }
    tc_fence_before_sync();  // (the next sample's first product overwrites these columns after its own barriers)
    TCA_MARK(10);
#ifdef RD_TCA_PROF
    ++nsamp;
#endif
  }

#ifdef RD_TCA_PROF
  if (blockIdx.x == 0 && tid == 0 && nsamp > 0)
    printf("attn_tc T=%d NP=%d samples=%d cycles/sample: wait_x %lld | stats %lld | coef+xn %lld | mma_u %lld | epi_u %lld | mma_qk %lld | epi_qk %lld | mma_s %lld | softmax %lld | mma_y %lld | epi_y %lld\n",
           T, NP, nsamp, prof[0] / nsamp, prof[1] / nsamp, prof[2] / nsamp, prof[3] / nsamp, prof[4] / nsamp, prof[5] / nsamp, prof[6] / nsamp,
           prof[7] / nsamp, prof[8] / nsamp, prof[9] / nsamp, prof[10] / nsamp);
#endif
  tc_fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem_slot, L::tmem_cols);
}

}  // namespace

// RD_E_UNSUPPORTED when the shape is outside this kernel (the caller falls back to the mma.sync kernel)
int attn_block_tc_launch(const rd_op_attn_block& op, cudaStream_t st) {
  if (op.C != TC_C || op.T < 1 || op.T > 128) return RD_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(op.x) | reinterpret_cast<uintptr_t>(op.out)) & 31) return RD_E_UNSUPPORTED;
  const float scale = 1.0f / sqrtf(static_cast<float>(op.C));
  const int t16 = (op.T + 15) / 16;
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
    {                                                                                                                                   \
      constexpr int np = TcLayout<n>::NP;                                                                                               \
      const int grid = kNumSMs * np > op.B2 ? (op.B2 + np - 1) / np : kNumSMs;                                                          \
      attn_block_tc_kernel<n><<<grid, 128 * np, TcLayout<n>::total, st>>>(x, out, wq, wp, op.bqkv, op.bproj, op.gamma, op.beta, op.B2,  \
                                                                          op.T, op.groups, op.eps, scale, op.out_scale);                \
    }                                                                                                                                   \
    break;
    RD_TC_CASE(1) RD_TC_CASE(2) RD_TC_CASE(3) RD_TC_CASE(4) RD_TC_CASE(5) RD_TC_CASE(6) RD_TC_CASE(7) RD_TC_CASE(8)
#undef RD_TC_CASE
    default: return RD_E_UNSUPPORTED;
  }
  return check_launch("attn_block_tc_kernel");
}

}  // namespace rd
