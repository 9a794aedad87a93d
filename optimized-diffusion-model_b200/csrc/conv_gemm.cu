// conv_gemm.cu -- the NCSN++ 3x3 convolutions / 1x1 NINs as implicit GEMMs on the 5th-gen tensor
// cores (tcgen05.mma, fp32 accumulators in TMEM), with GroupNorm + SiLU fused into the operand
// staging and bias / time-embedding / residual / (x+h)/sqrt(2) fused into the epilogue.
//
// Reference semantics (all under /root/reference/Reflected-Diffusion/models):
//   layers.ddpm_conv3x3 layers.py:103-109, layers.NIN layers.py:531-540,
//   ResnetBlockDDPMpp.forward layerspp.py:198-214 (GN -> SiLU -> conv, + Dense_0(temb), skip/sqrt2),
//   AttnBlockpp q/k/v/out projections layerspp.py:82-96, Downsample layerspp.py:157-159 (pad (0,1,0,1),
//   stride 2), Upsample layerspp.py:122-124 and the ragged nearest fix-up ncsnpp.py:319-320.
//
// Design (B200-first, nothing like the reference's cuDNN calls):
//   * A CTA owns S whole samples, so GroupNorm statistics never leave the CTA.
//   * The (normalised, activated, bf16) input image of those samples is staged ONCE into shared
//     memory in the un-swizzled K-major UMMA layout  [C/8][R rows][8 ch]  with one shared zero
//     column / zero row of padding between image rows / samples.  Rows are 16 B apart, so the A
//     operand of tap (dy,dx) is the same buffer with the descriptor start address advanced by
//     (dy*Wp+dx)*16 B: the nine taps of a 3x3 conv cost no data movement at all (verified on B200
//     by tools/probe_umma.cu).  Outputs at padding positions are computed and dropped.
//   * Weights are pre-packed on the host into the matching B layout, one 64-channel x tap slab
//     (C_out*128 B) per pipeline stage, and streamed from L2 by 1-D bulk TMA copies
//     (cp.async.bulk + mbarrier tx-count) issued by one thread; tcgen05.commit recycles the slots.
//   * One elected thread issues all MMAs (M=128 tiles, N=C_out, K=16 per instruction).
//   * All 8 warps run the epilogue straight out of TMEM (tcgen05.ld 32x32b).
#include "rd_common.h"
#include "rd_ptx.cuh"

namespace rd {

constexpr int CONV_THREADS = 256;
constexpr int CONV_MAX_STAGES = 4;
constexpr int MAX_HW = 32;  // gather maps

struct ConvParams {
  const __nv_bfloat16* src[2];
  int C[2], Hs[2], Ws[2];
  int nsrc;
  int H, W;          // logical (gathered) input image
  int Wp, rps;       // padded row width, rows per sample in the staged image
  int pad;           // pixel (y,x) is staged at (y+pad, x+pad)
  int stride;        // output (oy,ox) = accumulator row (oy*stride, ox*stride)
  int Ho, Wo;
  int ntaps;         // 9 or 1
  int Cin, KC;       // total input channels, Cin/8
  int nchunks;       // Cin/64
  int N;             // C_out
  int S, n_tiles, R; // samples per CTA, 128-row accumulator tiles, staged rows (odd)
  int nstages;       // weight ring depth
  int tmem_cols;     // power of two >= n_tiles*N
  int groups, cpg, silu;
  float eps;
  const float* gamma;
  const float* beta;
  const __nv_bfloat16* w;  // [nchunks][ntaps][8][N][8]
  const float* bias;
  const float* tproj;
  int tproj_stride, tproj_off;
  const __nv_bfloat16* residual;
  float out_scale;
  __nv_bfloat16* out;
  int B2;
  unsigned char ymap[2][MAX_HW], xmap[2][MAX_HW];
};

// dynamic shared memory carve-up (offsets in bytes, all 128-B aligned)
struct ConvSmemLayout {
  int a_off, a_bytes;        // staged operand  [KC][R][16 B]
  int w_off, w_stage_bytes;  // weight ring     [nstages][N*128 B]
  int tab_off;               // gamma[Cin], beta[Cin], gidx[Cin](int), mean[S*G], rstd[S*G]
  int total;
};

__host__ __device__ inline ConvSmemLayout conv_smem_layout(int Cin, int R, int N, int nstages, int S, int groups) {
  ConvSmemLayout L;
  L.a_off = 0;
  L.a_bytes = (Cin / 8) * R * 16;
  L.w_off = (L.a_bytes + 127) / 128 * 128;
  L.w_stage_bytes = N * 128;
  L.tab_off = L.w_off + nstages * L.w_stage_bytes;
  int tab = Cin * 12 + (groups > 0 ? S * groups * 8 : 0);
  L.total = L.tab_off + (tab + 127) / 128 * 128;
  return L;
}

__device__ __forceinline__ const __nv_bfloat16* src_pixel(const ConvParams& p, int which, int sample, int y, int x) {
  const int sy = p.ymap[which][y], sx = p.xmap[which][x];
  return p.src[which] + ((static_cast<size_t>(sample) * p.Hs[which] + sy) * p.Ws[which] + sx) * p.C[which];
}

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }

__global__ void __launch_bounds__(CONV_THREADS, 2) conv_gemm_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar_full[CONV_MAX_STAGES], bar_empty[CONV_MAX_STAGES], bar_acc;
  __shared__ uint32_t tmem_slot;

  const ConvSmemLayout L = conv_smem_layout(p.Cin, p.R, p.N, p.nstages, p.S, p.groups);
  unsigned char* As = smem + L.a_off;
  unsigned char* Ws = smem + L.w_off;
  float* s_gamma = reinterpret_cast<float*>(smem + L.tab_off);
  float* s_beta = s_gamma + p.Cin;
  int* s_gidx = reinterpret_cast<int*>(s_beta + p.Cin);
  float* s_mean = reinterpret_cast<float*>(s_gidx + p.Cin);
  float* s_rstd = s_mean + p.S * p.groups;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sample0 = blockIdx.x * p.S;
  const int S_act = min(p.S, p.B2 - sample0);
  const int P = p.H * p.W;
  const int total_it = p.nchunks * p.ntaps;

  // ------------------------------------------------------------------ setup
  if (tid == 0) {
    for (int i = 0; i < p.nstages; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    mbar_init(&bar_acc, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, p.tmem_cols);
  __syncthreads();

  // weight producer: prime the ring right away (weights do not depend on the staging below)
  if (warp == 1 && lane == 0) {
    const int prime = min(p.nstages, total_it);
    for (int it = 0; it < prime; ++it) {
      mbar_arrive_expect_tx(&bar_full[it], L.w_stage_bytes);
      bulk_g2s(Ws + it * L.w_stage_bytes, reinterpret_cast<const unsigned char*>(p.w) + static_cast<size_t>(it) * L.w_stage_bytes,
               L.w_stage_bytes, &bar_full[it]);
    }
  }

  // ------------------------------------------------------------------ GroupNorm statistics (fp32, in-CTA)
  if (p.groups > 0) {
    for (int c = tid; c < p.Cin; c += CONV_THREADS) {
      s_gamma[c] = p.gamma[c];
      s_beta[c] = p.beta[c];
      s_gidx[c] = c / p.cpg;
    }
    // pass 1: per-(sample, 8-channel chunk, pixel slice) partial sums, scratch aliased onto the A buffer
    float* part = reinterpret_cast<float*>(As);
    const int pairs = S_act * p.KC;
    int PS = CONV_THREADS / max(pairs, 1);
    PS = max(1, min(PS, P));
    for (int item = tid; item < pairs * PS; item += CONV_THREADS) {
      const int pair = item % pairs, slice = item / pairs;
      const int s = pair / p.KC, kc = pair % p.KC;
      const int which = (kc * 8 < p.C[0]) ? 0 : 1;
      const int coff = kc * 8 - (which ? p.C[0] : 0);
      float sum[8], sq[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { sum[j] = 0.0f; sq[j] = 0.0f; }
      for (int px = slice; px < P; px += PS) {
        const int y = px / p.W, x = px % p.W;
        const uint4 raw = *reinterpret_cast<const uint4*>(src_pixel(p, which, sample0 + s, y, x) + coff);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(h2[j]);
          sum[2 * j] += f.x; sq[2 * j] += f.x * f.x;
          sum[2 * j + 1] += f.y; sq[2 * j + 1] += f.y * f.y;
        }
      }
      float* dst = part + (static_cast<size_t>(slice) * pairs + pair) * 16;
#pragma unroll
      for (int j = 0; j < 8; ++j) { dst[j] = sum[j]; dst[8 + j] = sq[j]; }
    }
    __syncthreads();
    // per-(sample, group) mean / rstd
    for (int sg = tid; sg < S_act * p.groups; sg += CONV_THREADS) {
      const int s = sg / p.groups, g = sg % p.groups;
      float sum = 0.0f, sq = 0.0f;
      for (int c = g * p.cpg; c < (g + 1) * p.cpg; ++c) {
        const int pair = s * p.KC + (c >> 3), j = c & 7;
        for (int slice = 0; slice < PS; ++slice) {
          const float* src = part + (static_cast<size_t>(slice) * pairs + pair) * 16;
          sum += src[j];
          sq += src[8 + j];
        }
      }
      const float inv_n = 1.0f / static_cast<float>(p.cpg * P);
      const float mean = sum * inv_n;
      const float var = fmaxf(sq * inv_n - mean * mean, 0.0f);
      s_mean[sg] = mean;
      s_rstd[sg] = 1.0f / sqrtf(var + p.eps);
    }
    __syncthreads();
  }

  // ------------------------------------------------------------------ stage the operand image
  {
    uint4* a4 = reinterpret_cast<uint4*>(As);
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < p.KC * p.R; i += CONV_THREADS) a4[i] = zero;
    __syncthreads();
    const int items = S_act * P * p.KC;
    for (int item = tid; item < items; item += CONV_THREADS) {
      const int kc = item % p.KC;
      const int sp = item / p.KC;
      const int s = sp / P, px = sp % P;
      const int y = px / p.W, x = px % p.W;
      const int which = (kc * 8 < p.C[0]) ? 0 : 1;
      const int coff = kc * 8 - (which ? p.C[0] : 0);
      uint4 raw = *reinterpret_cast<const uint4*>(src_pixel(p, which, sample0 + s, y, x) + coff);
      if (p.groups > 0) {
        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&raw);
        const int c0 = kc * 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 f = __bfloat1622float2(h2[j]);
          const int ca = c0 + 2 * j, cb = ca + 1;
          const int ga = s * p.groups + s_gidx[ca], gb = s * p.groups + s_gidx[cb];
          f.x = (f.x - s_mean[ga]) * s_rstd[ga] * s_gamma[ca] + s_beta[ca];
          f.y = (f.y - s_mean[gb]) * s_rstd[gb] * s_gamma[cb] + s_beta[cb];
          if (p.silu) { f.x = silu_f(f.x); f.y = silu_f(f.y); }
          h2[j] = __floats2bfloat162_rn(f.x, f.y);
        }
      }
      const int row = s * p.rps + (y + p.pad) * p.Wp + (x + p.pad);
      a4[kc * p.R + row] = raw;
    }
  }
  fence_proxy_async_smem();  // st.shared above must be visible to the tensor core's async-proxy reads
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;

  // ------------------------------------------------------------------ MMA issue + weight ring
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, p.N);
      const uint32_t a_base = smem_u32(As), w_base = smem_u32(Ws);
      int it = 0;
      for (int chunk = 0; chunk < p.nchunks; ++chunk) {
        for (int tap = 0; tap < p.ntaps; ++tap, ++it) {
          const int stage = it % p.nstages;
          mbar_wait(&bar_full[stage], (it / p.nstages) & 1);
          tc_fence_after_sync();
          const int shift = (p.ntaps == 9) ? (tap / 3) * p.Wp + (tap % 3) : 0;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t db = umma_desc_kmajor(w_base + stage * L.w_stage_bytes + (kk * 2 * p.N) * 16, p.N * 16, 128);
            for (int tile = 0; tile < p.n_tiles; ++tile) {
              const uint32_t a_addr = a_base + (((chunk * 8 + kk * 2) * p.R) + tile * 128 + shift) * 16;
              umma_bf16_ss(tmem + tile * p.N, umma_desc_kmajor(a_addr, p.R * 16, 128), db, idesc, (it | kk) != 0);
            }
          }
          umma_commit(&bar_empty[stage]);  // slot reusable once these MMAs have read it
        }
      }
      umma_commit(&bar_acc);
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      for (int it = p.nstages; it < total_it; ++it) {
        const int stage = it % p.nstages;
        mbar_wait(&bar_empty[stage], ((it / p.nstages) - 1) & 1);
        mbar_arrive_expect_tx(&bar_full[stage], L.w_stage_bytes);
        bulk_g2s(Ws + stage * L.w_stage_bytes,
                 reinterpret_cast<const unsigned char*>(p.w) + static_cast<size_t>(it) * L.w_stage_bytes, L.w_stage_bytes,
                 &bar_full[stage]);
      }
    }
    __syncwarp();
  }

  // ------------------------------------------------------------------ epilogue (all 8 warps)
  mbar_wait(&bar_acc, 0);
  tc_fence_after_sync();
  {
    const int q = warp & 3;             // TMEM lane quarter this warp may read
    const int cblocks = p.N / 32;
    const int nblocks = p.n_tiles * cblocks;
    for (int blk = (warp >> 2); blk < nblocks; blk += 2) {
      const int tile = blk / cblocks, c0 = (blk % cblocks) * 32;
      uint32_t v[32];
      tmem_ld32(tmem + (static_cast<uint32_t>(q * 32) << 16) + tile * p.N + c0, v);
      tmem_ld_wait();
      const int row = tile * 128 + q * 32 + lane;
      const int s = row / p.rps, rem = row % p.rps;
      const int Y = rem / p.Wp, X = rem % p.Wp;
      bool valid = (s < S_act);
      int oy = Y, ox = X;
      if (p.stride == 2) {
        valid = valid && ((Y & 1) == 0) && ((X & 1) == 0);
        oy = Y >> 1; ox = X >> 1;
      }
      valid = valid && (oy < p.Ho) && (ox < p.Wo);
      if (valid) {
        const int sample = sample0 + s;
        const size_t o = ((static_cast<size_t>(sample) * p.Ho + oy) * p.Wo + ox) * p.N + c0;
        const float* tp = p.tproj ? p.tproj + static_cast<size_t>(sample) * p.tproj_stride + p.tproj_off + c0 : nullptr;
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float a = __uint_as_float(v[j]) + __ldg(p.bias + c0 + j);
          float b = __uint_as_float(v[j + 1]) + __ldg(p.bias + c0 + j + 1);
          if (tp) { a += __ldg(tp + j); b += __ldg(tp + j + 1); }
          if (p.residual) {
            const float2 r = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p.residual + o + j));
            a += r.x; b += r.y;
          }
          a *= p.out_scale; b *= p.out_scale;
          __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
          packed[j >> 1] = *reinterpret_cast<uint32_t*>(&h);
        }
        uint4* dst = reinterpret_cast<uint4*>(p.out + o);
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, p.tmem_cols);
}

// nearest-neighbour source index exactly as torch F.interpolate(mode='nearest'):
// src = min(floor(dst * (in/out as float)), in-1)
static void nearest_map(unsigned char* map, int out_size, int in_size) {
  const float scale = static_cast<float>(in_size) / static_cast<float>(out_size);
  for (int i = 0; i < out_size; ++i) {
    int s = static_cast<int>(floorf(static_cast<float>(i) * scale));
    map[i] = static_cast<unsigned char>(s < in_size - 1 ? s : in_size - 1);
  }
}

static int next_pow2_cols(int c) {
  int v = 32;
  while (v < c) v <<= 1;
  return v;
}

// Fills the launch geometry for an op; returns RD_OK or an error.
int conv_make_params(const rd_op_conv& op, ConvParams& p, int& smem_bytes, int& grid) {
  RD_REQUIRE(op.nsrc == 1 || op.nsrc == 2, "conv: nsrc must be 1 or 2");
  RD_REQUIRE(op.ntaps == 9 || op.ntaps == 1, "conv: ntaps must be 9 or 1");
  RD_REQUIRE(op.H_in >= 1 && op.W_in >= 1 && op.H_in <= MAX_HW && op.W_in <= MAX_HW, "conv: H,W must be in [1,%d]", MAX_HW);
  RD_REQUIRE(op.C_out % 32 == 0 && op.C_out >= 32 && op.C_out <= 256 && op.C_out % 16 == 0, "conv: C_out %d unsupported", op.C_out);
  RD_REQUIRE(op.stride == 1 || (op.stride == 2 && op.pad == 0 && op.ntaps == 9), "conv: stride 2 needs pad 0, 3x3");
  RD_REQUIRE(op.w && op.bias && op.out && op.B2 > 0, "conv: null pointer / empty batch");
  memset(&p, 0, sizeof(p));
  int cin = 0;
  for (int i = 0; i < op.nsrc; ++i) {
    RD_REQUIRE(op.src[i].ptr && op.src[i].C % 8 == 0 && op.src[i].C > 0, "conv: source %d channels must be a multiple of 8", i);
    RD_REQUIRE(op.src[i].Hs >= 1 && op.src[i].Ws >= 1 && op.src[i].Hs <= MAX_HW && op.src[i].Ws <= MAX_HW, "conv: bad source size");
    p.src[i] = static_cast<const __nv_bfloat16*>(op.src[i].ptr);
    p.C[i] = op.src[i].C; p.Hs[i] = op.src[i].Hs; p.Ws[i] = op.src[i].Ws;
    nearest_map(p.ymap[i], op.H_in, op.src[i].Hs);
    nearest_map(p.xmap[i], op.W_in, op.src[i].Ws);
    cin += op.src[i].C;
  }
  if (op.nsrc == 1) { p.C[1] = 0; }
  RD_REQUIRE(cin % 64 == 0, "conv: total input channels (%d) must be a multiple of 64", cin);
  p.nsrc = op.nsrc;
  p.H = op.H_in; p.W = op.W_in;
  p.ntaps = op.ntaps;
  if (op.ntaps == 9) {
    p.Wp = op.W_in + 1; p.rps = (op.H_in + 1) * p.Wp; p.pad = op.pad ? 1 : 0;
  } else {
    p.Wp = op.W_in; p.rps = op.H_in * op.W_in; p.pad = 0;
  }
  p.stride = op.stride;
  p.Ho = op.H_out; p.Wo = op.W_out;
  if (op.stride == 1) RD_REQUIRE(op.H_out == op.H_in && op.W_out == op.W_in, "conv: stride-1 output must match input size");
  else RD_REQUIRE(op.H_out == (op.H_in + 1 - 3) / 2 + 1 && op.W_out == (op.W_in + 1 - 3) / 2 + 1, "conv: bad downsample output size");
  p.Cin = cin; p.KC = cin / 8; p.nchunks = cin / 64; p.N = op.C_out;
  p.groups = op.gn_groups; p.silu = op.gn_silu; p.eps = op.gn_eps;
  if (p.groups > 0) {
    RD_REQUIRE(cin % p.groups == 0 && op.gn_gamma && op.gn_beta, "conv: bad GroupNorm arguments");
    p.cpg = cin / p.groups;  // groups may straddle the concat boundary (192 ch / 32 groups): stats are per channel
  }
  p.gamma = op.gn_gamma; p.beta = op.gn_beta;
  p.w = static_cast<const __nv_bfloat16*>(op.w);
  p.bias = op.bias; p.tproj = op.tproj; p.tproj_stride = op.tproj_stride; p.tproj_off = op.tproj_off;
  p.residual = static_cast<const __nv_bfloat16*>(op.residual);
  p.out_scale = op.out_scale; p.out = static_cast<__nv_bfloat16*>(op.out); p.B2 = op.B2;

  // tile geometry: as many 128-row tiles as TMEM (512 cols) and shared memory (227 KB) allow, capped at 4
  const int max_shift = (op.ntaps == 9) ? 2 * p.Wp + 2 : 0;
  const int smem_cap = 227 * 1024 - 1024;  // static __shared__ + alignment slack
  int best_tiles = 0;
  for (int nt = 4; nt >= 1; --nt) {
    if (nt * p.N > 512) continue;
    if (nt * 128 < p.rps) continue;  // must hold at least one sample
    int R = nt * 128 + max_shift; R |= 1;
    int S = (nt * 128) / p.rps;
    int stages = 3;
    ConvSmemLayout L = conv_smem_layout(cin, R, p.N, stages, S, p.groups);
    if (L.total > smem_cap) { stages = 2; L = conv_smem_layout(cin, R, p.N, stages, S, p.groups); }
    if (L.total > smem_cap) continue;
    // prefer a geometry that lets two CTAs share an SM (staging of one overlaps the MMAs of the other)
    if (best_tiles == 0) best_tiles = nt;
    if (L.total <= 110 * 1024 && nt * p.N <= 256) { best_tiles = nt; break; }
  }
  if (op.samples_per_cta > 0) {  // planner override: smallest tile count that holds that many samples
    for (int nt = 1; nt <= 4; ++nt)
      if (nt * 128 >= op.samples_per_cta * p.rps && nt * p.N <= 512) { best_tiles = nt; break; }
  }
  RD_REQUIRE(best_tiles > 0, "conv: no tile geometry fits (Cin=%d N=%d rps=%d)", cin, p.N, p.rps);
  p.n_tiles = best_tiles;
  p.R = (best_tiles * 128 + max_shift) | 1;
  p.S = (best_tiles * 128) / p.rps;
  if (op.samples_per_cta > 0 && op.samples_per_cta < p.S) p.S = op.samples_per_cta;
  p.nstages = 3;
  ConvSmemLayout L = conv_smem_layout(cin, p.R, p.N, p.nstages, p.S, p.groups);
  if (L.total > smem_cap) { p.nstages = 2; L = conv_smem_layout(cin, p.R, p.N, p.nstages, p.S, p.groups); }
  RD_REQUIRE(L.total <= smem_cap, "conv: shared memory budget exceeded (%d B)", L.total);
  // the statistics scratch ([PS][S*KC][16] floats <= max(256, S*KC)*64 B) is aliased onto the A buffer
  if (p.groups > 0) {
    const int pairs = p.S * p.KC;
    const int scratch = (pairs > CONV_THREADS ? pairs : CONV_THREADS) * 64;
    RD_REQUIRE(scratch <= L.a_bytes, "conv: GroupNorm scratch does not fit the staging buffer");
  }
  p.tmem_cols = next_pow2_cols(p.n_tiles * p.N);
  smem_bytes = L.total;
  grid = (op.B2 + p.S - 1) / p.S;
  return RD_OK;
}

int conv_launch(const rd_op_conv& op, cudaStream_t st) {
  ConvParams p;
  int smem = 0, grid = 0;
  int rc = conv_make_params(op, p, smem, grid);
  if (rc != RD_OK) return rc;
  static int configured_smem = 0;
  if (smem > configured_smem) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "conv: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured_smem = 227 * 1024;
  }
  conv_gemm_kernel<<<grid, CONV_THREADS, smem, st>>>(p);
  return check_launch("conv_gemm_kernel");
}

}  // namespace rd

extern "C" int rd_conv_launch_info(const rd_op_conv* op, int* smem_bytes, int* grid, int* rows_alloc) {
  RD_REQUIRE(op, "rd_conv_launch_info: null op");
  rd::ConvParams p;
  int smem = 0, g = 0;
  int rc = rd::conv_make_params(*op, p, smem, g);
  if (rc != RD_OK) return rc;
  if (smem_bytes) *smem_bytes = smem;
  if (grid) *grid = g;
  if (rows_alloc) *rows_alloc = p.R;
  return RD_OK;
}
