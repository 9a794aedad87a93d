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
//   * PERSISTENT kernel, one CTA per SM, looping over groups of S whole samples (GroupNorm statistics
//     never leave the CTA).  16 warps with fixed roles that overlap through mbarrier pipelines:
//       warp 0        MMA issuer (one elected thread, tcgen05.mma M=128 x N=C_out x K=16)
//       warp 1        weight producer (1-D bulk TMA copies from L2; whole filter resident in shared
//                     memory when it fits, else a ring recycled by tcgen05.commit)
//       warps 4-7     epilogue: TMEM -> registers (tcgen05.ld 32x32b) -> bias/temb/skip -> bf16 NHWC
//       warps 2,3,8-15 transform: global bf16 -> GroupNorm statistics -> normalise + SiLU -> operand ring
//     While the tensor core works on group g, the transform warps stage group g+1 and the epilogue
//     warps drain group g-1 (double-buffered TMEM accumulators).
//   * The operand image of a 64-channel chunk is staged ONCE as [8 k-chunks][R rows][8 ch] (K-major,
//     SWIZZLE_NONE, rows 16 B apart) with one shared zero column / zero row between image rows /
//     samples, so the A operand of tap (dy,dx) is the same buffer with the UMMA descriptor start
//     address advanced by (dy*Wp+dx)*16 B: nine taps, no im2col, no data movement (validated on B200 by
//     tools/probe_umma.cu).  Outputs at padding rows are computed and dropped.
#include <cstring>
#include "rd_common.h"
#include "rd_ptx.cuh"

namespace rd {

constexpr int CONV_THREADS = 512;
constexpr int XFORM_THREADS = 320;  // warps 2,3,8..15
constexpr int EPI_WARPS = 4;        // warps 4..7
constexpr int MAX_A_STAGES = 3;
constexpr int MAX_W_STAGES = 4;
constexpr int MAX_HW = 32;  // gather maps
constexpr int RC_MAX = 16;                           // pixels a transform thread may cache in registers
constexpr int STAT_PAIRS = 256;                      // (sample, 8-channel chunk) pairs per statistics batch
constexpr int STAT_SCRATCH_BYTES = XFORM_THREADS * 64;  // one 16-float partial record per transform thread

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
  int S, n_tiles, R; // samples per group, 128-row accumulator tiles, staged rows (odd)
  int n_groups;
  int a_stages, a_stage_bytes;
  int w_resident, w_stages, w_slab_bytes, n_slabs;
  int acc_bufs;
  int xmode, rc_PS;  // transform mode: 1 = register-cached single pass (rc_PS pixel slices), 0 = streaming
  int tmem_cols;     // power of two >= acc_bufs*n_tiles*N
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

// dynamic shared memory carve-up (bytes)
struct ConvSmemLayout {
  int a_off, w_off, tab_off, total;
};

__host__ __device__ inline ConvSmemLayout conv_smem_layout(const ConvParams& p) {
  ConvSmemLayout L;
  L.a_off = 0;
  L.w_off = p.a_stages * p.a_stage_bytes;
  L.tab_off = L.w_off + (p.w_resident ? p.n_slabs : p.w_stages) * p.w_slab_bytes;
  int tab = (p.groups > 0 ? p.Cin * 12 + p.S * p.groups * 8 + STAT_SCRATCH_BYTES : 0) + p.S * p.H * p.W * 12 + 64;
  L.total = L.tab_off + (tab + 127) / 128 * 128;
  return L;
}

__device__ __forceinline__ const __nv_bfloat16* src_pixel(const ConvParams& p, int which, int sample, int y, int x) {
  const int sy = p.ymap[which][y], sx = p.xmap[which][x];
  return p.src[which] + ((static_cast<size_t>(sample) * p.Hs[which] + sy) * p.Ws[which] + sx) * p.C[which];
}

__device__ __forceinline__ float silu_f(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

__device__ __forceinline__ void xform_bar() { asm volatile("bar.sync 1, %0;" ::"n"(XFORM_THREADS) : "memory"); }

// address of the 8-channel item (sample s of the group, pixel px, global k-chunk kc)
__device__ __forceinline__ const uint4* item_ptr(const ConvParams& p, int sample, int px, int kc) {
  const int y = px / p.W, x = px - y * p.W;
  const int which = (kc * 8 < p.C[0]) ? 0 : 1;
  const int coff = kc * 8 - (which ? p.C[0] : 0);
  return reinterpret_cast<const uint4*>(src_pixel(p, which, sample, y, x) + coff);
}

__device__ __forceinline__ void unpack8(const uint4& raw, float (&f)[8]) {
  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 v = __bfloat1622float2(h2[j]);
    f[2 * j] = v.x;
    f[2 * j + 1] = v.y;
  }
}

__global__ void __launch_bounds__(CONV_THREADS, 1) conv_gemm_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar_a_full[MAX_A_STAGES], bar_a_empty[MAX_A_STAGES];
  __shared__ uint64_t bar_w_full[MAX_W_STAGES], bar_w_empty[MAX_W_STAGES];
  __shared__ uint64_t bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint32_t tmem_slot;

  const ConvSmemLayout L = conv_smem_layout(p);
  unsigned char* As = smem + L.a_off;
  unsigned char* Ws = smem + L.w_off;
  // tables: source offsets [2][S*P] (int), staged rows [S*P] (u16), then the GroupNorm tables when used
  int* t_off = reinterpret_cast<int*>(smem + L.tab_off);
  unsigned short* t_row = reinterpret_cast<unsigned short*>(t_off + 2 * p.S * p.H * p.W);
  const int gn_cin = p.groups > 0 ? p.Cin : 0;
  float* s_gamma = reinterpret_cast<float*>(smem + L.tab_off + ((p.S * p.H * p.W * 10 + 63) / 64) * 64);
  float* s_beta = s_gamma + gn_cin;
  int* s_gidx = reinterpret_cast<int*>(s_beta + gn_cin);
  float* s_stat = reinterpret_cast<float*>(s_gidx + gn_cin);  // [S][G][2]: (mean, rstd)
  float* s_scr = s_stat + p.S * p.groups * 2;                 // statistics scratch, STAT_SCRATCH_BYTES (16-B aligned)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int P = p.H * p.W;
  const int my_groups = (p.n_groups > static_cast<int>(blockIdx.x))
                            ? (p.n_groups - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                            : 0;

  // ------------------------------------------------------------------ setup
  if (tid == 0) {
    for (int i = 0; i < MAX_A_STAGES; ++i) { mbar_init(&bar_a_full[i], 1); mbar_init(&bar_a_empty[i], 1); }
    for (int i = 0; i < MAX_W_STAGES; ++i) { mbar_init(&bar_w_full[i], 1); mbar_init(&bar_w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_acc_full[i], 1); mbar_init(&bar_acc_empty[i], EPI_WARPS); }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, p.tmem_cols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    // ================================================================ MMA issuer
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, p.N);
      const uint32_t a_base = smem_u32(As), w_base = smem_u32(Ws);
      int a_it = 0, w_it = 0;
      if (p.w_resident && my_groups > 0) { mbar_wait(&bar_w_full[0], 0); tc_fence_after_sync(); }
      for (int li = 0; li < my_groups; ++li) {
        const int buf = li % p.acc_bufs, useb = li / p.acc_bufs;
        if (useb > 0) { mbar_wait(&bar_acc_empty[buf], (useb - 1) & 1); tc_fence_after_sync(); }
        const uint32_t acc = tmem + buf * p.n_tiles * p.N;
        for (int chunk = 0; chunk < p.nchunks; ++chunk, ++a_it) {
          const int stage = a_it % p.a_stages;
          mbar_wait(&bar_a_full[stage], (a_it / p.a_stages) & 1);
          tc_fence_after_sync();
          const uint32_t a_stage = a_base + stage * p.a_stage_bytes;
          for (int tap = 0; tap < p.ntaps; ++tap) {
            uint32_t wslab;
            int ws = 0;
            if (p.w_resident) {
              wslab = w_base + (chunk * p.ntaps + tap) * p.w_slab_bytes;
            } else {
              ws = w_it % p.w_stages;
              mbar_wait(&bar_w_full[ws], (w_it / p.w_stages) & 1);
              tc_fence_after_sync();
              wslab = w_base + ws * p.w_slab_bytes;
            }
            const int shift = (p.ntaps == 9) ? (tap / 3) * p.Wp + (tap % 3) : 0;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t db = umma_desc_kmajor(wslab + (kk * 2 * p.N) * 16, p.N * 16, 128);
              for (int tile = 0; tile < p.n_tiles; ++tile) {
                const uint32_t a_addr = a_stage + ((kk * 2) * p.R + tile * 128 + shift) * 16;
                umma_bf16_ss(acc + tile * p.N, umma_desc_kmajor(a_addr, p.R * 16, 128), db, idesc, (chunk | tap | kk) != 0);
              }
            }
            if (!p.w_resident) { umma_commit(&bar_w_empty[ws]); ++w_it; }
          }
          umma_commit(&bar_a_empty[stage]);  // operand stage reusable once these MMAs have read it
        }
        umma_commit(&bar_acc_full[buf]);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================================================ weight producer
    if (lane == 0 && my_groups > 0) {
      const unsigned char* wg = reinterpret_cast<const unsigned char*>(p.w);
      if (p.w_resident) {
        mbar_arrive_expect_tx(&bar_w_full[0], p.n_slabs * p.w_slab_bytes);
        for (int sidx = 0; sidx < p.n_slabs; ++sidx)
          bulk_g2s(Ws + sidx * p.w_slab_bytes, wg + static_cast<size_t>(sidx) * p.w_slab_bytes, p.w_slab_bytes, &bar_w_full[0]);
      } else {
        const int total = my_groups * p.n_slabs;
        for (int it = 0; it < total; ++it) {
          const int ws = it % p.w_stages;
          if (it >= p.w_stages) mbar_wait(&bar_w_empty[ws], ((it / p.w_stages) - 1) & 1);
          mbar_arrive_expect_tx(&bar_w_full[ws], p.w_slab_bytes);
          bulk_g2s(Ws + ws * p.w_slab_bytes, wg + static_cast<size_t>(it % p.n_slabs) * p.w_slab_bytes, p.w_slab_bytes,
                   &bar_w_full[ws]);
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 8) {
    // ================================================================ epilogue
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int cblocks = p.N / 32;
    const int nblocks = p.n_tiles * cblocks;
    for (int li = 0; li < my_groups; ++li) {
      const int g = blockIdx.x + li * gridDim.x;
      const int sample0 = g * p.S;
      const int S_act = min(p.S, p.B2 - sample0);
      const int buf = li % p.acc_bufs;
      mbar_wait(&bar_acc_full[buf], (li / p.acc_bufs) & 1);
      tc_fence_after_sync();
      const uint32_t acc = tmem + buf * p.n_tiles * p.N + (static_cast<uint32_t>(q * 32) << 16);
      for (int tile = 0; tile < p.n_tiles; ++tile) {
        const int row = tile * 128 + q * 32 + lane;
        const int s = row / p.rps, rem = row - s * p.rps;
        const int Y = rem / p.Wp, X = rem - Y * p.Wp;
        bool valid = (s < S_act);
        int oy = Y, ox = X;
        if (p.stride == 2) {
          valid = valid && ((Y & 1) == 0) && ((X & 1) == 0);
          oy = Y >> 1; ox = X >> 1;
        }
        valid = valid && (oy < p.Ho) && (ox < p.Wo);
        const int sample = sample0 + s;
        const size_t o = valid ? ((static_cast<size_t>(sample) * p.Ho + oy) * p.Wo + ox) * p.N : 0;
        const float* tp = (p.tproj && valid) ? p.tproj + static_cast<size_t>(sample) * p.tproj_stride + p.tproj_off : nullptr;
        for (int cb = 0; cb < cblocks; ++cb) {
          const int c0 = cb * 32;
          uint32_t v[32];
          tmem_ld32(acc + tile * p.N + c0, v);
          // issue the global reads of this block's epilogue operands while the TMEM load is in flight
          uint4 res[4];
          float4 tv[8];
          if (valid) {
            if (p.residual) {
#pragma unroll
              for (int j = 0; j < 4; ++j) res[j] = *reinterpret_cast<const uint4*>(p.residual + o + c0 + 8 * j);
            }
            if (tp) {
#pragma unroll
              for (int j = 0; j < 8; ++j) tv[j] = __ldg(reinterpret_cast<const float4*>(tp + c0) + j);
            }
          }
          tmem_ld_wait();
          if (valid) {
            uint32_t packed[16];
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + c0) + j4);
              float a[4] = {__uint_as_float(v[4 * j4]) + bv.x, __uint_as_float(v[4 * j4 + 1]) + bv.y,
                            __uint_as_float(v[4 * j4 + 2]) + bv.z, __uint_as_float(v[4 * j4 + 3]) + bv.w};
              if (tp) { a[0] += tv[j4].x; a[1] += tv[j4].y; a[2] += tv[j4].z; a[3] += tv[j4].w; }
              if (p.residual) {
                const uint32_t* rw = reinterpret_cast<const uint32_t*>(&res[j4 >> 1]) + 2 * (j4 & 1);
                const float2 r0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(rw));
                const float2 r1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(rw + 1));
                a[0] += r0.x; a[1] += r0.y; a[2] += r1.x; a[3] += r1.y;
              }
              __nv_bfloat162 h0 = __floats2bfloat162_rn(a[0] * p.out_scale, a[1] * p.out_scale);
              __nv_bfloat162 h1 = __floats2bfloat162_rn(a[2] * p.out_scale, a[3] * p.out_scale);
              packed[2 * j4] = *reinterpret_cast<uint32_t*>(&h0);
              packed[2 * j4 + 1] = *reinterpret_cast<uint32_t*>(&h1);
            }
            uint4* dst = reinterpret_cast<uint4*>(p.out + o + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          }
        }
      }
      (void)nblocks;
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_acc_empty[buf]);
    }
  } else {
    // ================================================================ transform (warps 2,3,8..15)
    const int xt = (warp < 4 ? warp - 2 : warp - 6) * 32 + lane;  // 0..319
    // zero the operand ring once: pixel positions are identical for every group, so padding rows stay zero
    {
      uint4* a4 = reinterpret_cast<uint4*>(As);
      const int n16 = p.a_stages * p.a_stage_bytes / 16;
      const uint4 zero = make_uint4(0, 0, 0, 0);
      for (int i = xt; i < n16; i += XFORM_THREADS) a4[i] = zero;
    }
    // group-invariant addressing tables: staged row and source element offset of every (sample, pixel)
    for (int sp = xt; sp < p.S * P; sp += XFORM_THREADS) {
      const int s = sp / P, px = sp - s * P;
      const int y = px / p.W, x = px - y * p.W;
      t_row[sp] = static_cast<unsigned short>(s * p.rps + (y + p.pad) * p.Wp + (x + p.pad));
      t_off[sp] = ((s * p.Hs[0] + p.ymap[0][y]) * p.Ws[0] + p.xmap[0][x]) * p.C[0];
      if (p.nsrc > 1) t_off[p.S * P + sp] = ((s * p.Hs[1] + p.ymap[1][y]) * p.Ws[1] + p.xmap[1][x]) * p.C[1];
    }
    if (p.groups > 0) {
      for (int c = xt; c < p.Cin; c += XFORM_THREADS) {
        s_gamma[c] = p.gamma[c];
        s_beta[c] = p.beta[c];
        s_gidx[c] = c / p.cpg;
      }
    }
    xform_bar();
    const size_t gstride0 = static_cast<size_t>(p.S) * p.Hs[0] * p.Ws[0] * p.C[0];
    const size_t gstride1 = static_cast<size_t>(p.S) * p.Hs[1] * p.Ws[1] * p.C[1];
    const float inv_n = p.groups > 0 ? 1.0f / static_cast<float>(p.cpg * P) : 0.0f;
    int a_it = 0;

    // normalise + activate the 8 channels starting at c0 of sample s (of the group) in place
    auto gn_apply = [&](uint4& raw, int s, int c0) {
      float f[8];
      unpack8(raw, f);
      const float4 g0 = *reinterpret_cast<const float4*>(s_gamma + c0), g1 = *reinterpret_cast<const float4*>(s_gamma + c0 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(s_beta + c0), b1 = *reinterpret_cast<const float4*>(s_beta + c0 + 4);
      const float gam[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bet[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      const float* st = s_stat + s * p.groups * 2;
      if ((p.cpg & 7) == 0) {  // one group per 8-channel item
        const float2 mr = *reinterpret_cast<const float2*>(st + 2 * (c0 / p.cpg));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a = gam[j] * mr.y;
          f[j] = fmaf(f[j], a, fmaf(-mr.x, a, bet[j]));
        }
      } else if (p.cpg == 4) {  // two groups per item
        const float4 mr = *reinterpret_cast<const float4*>(st + 2 * (c0 >> 2));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a = gam[j] * (j < 4 ? mr.y : mr.w);
          f[j] = fmaf(f[j], a, fmaf(-(j < 4 ? mr.x : mr.z), a, bet[j]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 mr = *reinterpret_cast<const float2*>(st + 2 * s_gidx[c0 + j]);
          const float a = gam[j] * mr.y;
          f[j] = fmaf(f[j], a, fmaf(-mr.x, a, bet[j]));
        }
      }
      if (p.silu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = silu_f(f[j]);
      }
      __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
      for (int j = 0; j < 4; ++j) h2[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    };
    // fixed-order reduction of the per-thread partial records into per-(sample, group) mean / rstd
    auto stat_reduce = [&](int s_first, int ns, int pairs, int PS) {
      for (int sg = xt; sg < ns * p.groups; sg += XFORM_THREADS) {
        const int s = sg / p.groups, g = sg - s * p.groups;
        float sum = 0.0f, sq = 0.0f;
        for (int c = g * p.cpg; c < (g + 1) * p.cpg; ++c) {
          const int pair = s * p.KC + (c >> 3), j = c & 7;
          for (int slice = 0; slice < PS; ++slice) {
            const float* src = s_scr + (static_cast<size_t>(slice) * pairs + pair) * 16;
            sum += src[j];
            sq += src[8 + j];
          }
        }
        const float mean = sum * inv_n;
        const float var = fmaxf(sq * inv_n - mean * mean, 0.0f);
        s_stat[2 * ((s_first + s) * p.groups + g)] = mean;
        s_stat[2 * ((s_first + s) * p.groups + g) + 1] = 1.0f / sqrtf(var + p.eps);
      }
    };

    if (p.xmode == 1) {
      // ---------------- register-cached mode: every thread owns one (sample, 8-channel chunk, pixel slice);
      // its pixels are read from global memory ONCE, kept in registers across the statistics barrier,
      // then normalised and written to the operand ring.
      const int pairs = p.S * p.KC, PS = p.rc_PS;
      const int pair = xt % pairs, slice = xt / pairs;
      const int s = pair / p.KC, kc = pair - s * p.KC;
      const bool owner = xt < pairs * PS;
      const int which = (kc * 8 < p.C[0]) ? 0 : 1;
      const int coff = kc * 8 - (which ? p.C[0] : 0);
      const int* toff = t_off + (which ? p.S * P : 0) + s * P;
      const unsigned short* trow = t_row + s * P;
      const int my_chunk = kc >> 3, kcl = kc & 7;
      for (int li = 0; li < my_groups; ++li) {
        const int g = blockIdx.x + li * gridDim.x;
        const int S_act = min(p.S, p.B2 - g * p.S);
        const bool active = owner && s < S_act;
        const __nv_bfloat16* gbase = p.src[which] + static_cast<size_t>(g) * (which ? gstride1 : gstride0) + coff;
        uint4 raw[RC_MAX];
#pragma unroll
        for (int k = 0; k < RC_MAX; ++k) {
          const int px = slice + k * PS;
          if (active && px < P) raw[k] = *reinterpret_cast<const uint4*>(gbase + toff[px]);
        }
        float sum[8], sq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { sum[j] = 0.0f; sq[j] = 0.0f; }
#pragma unroll
        for (int k = 0; k < RC_MAX; ++k) {
          if (active && slice + k * PS < P) {
            float f[8];
            unpack8(raw[k], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) { sum[j] += f[j]; sq[j] = fmaf(f[j], f[j], sq[j]); }
          }
        }
        if (owner) {
          float4* dst = reinterpret_cast<float4*>(s_scr + (static_cast<size_t>(slice) * pairs + pair) * 16);
          dst[0] = make_float4(sum[0], sum[1], sum[2], sum[3]);
          dst[1] = make_float4(sum[4], sum[5], sum[6], sum[7]);
          dst[2] = make_float4(sq[0], sq[1], sq[2], sq[3]);
          dst[3] = make_float4(sq[4], sq[5], sq[6], sq[7]);
        }
        xform_bar();
        stat_reduce(0, S_act, pairs, PS);
        xform_bar();
        for (int chunk = 0; chunk < p.nchunks; ++chunk, ++a_it) {
          const int stage = a_it % p.a_stages;
          if (a_it >= p.a_stages) mbar_wait(&bar_a_empty[stage], ((a_it / p.a_stages) - 1) & 1);
          if (active && chunk == my_chunk) {
            uint4* a4 = reinterpret_cast<uint4*>(As + stage * p.a_stage_bytes) + kcl * p.R;
#pragma unroll
            for (int k = 0; k < RC_MAX; ++k) {
              const int px = slice + k * PS;
              if (px < P) {
                gn_apply(raw[k], s, kc * 8);
                a4[trow[px]] = raw[k];
              }
            }
          }
          fence_proxy_async_smem();  // st.shared above must be visible to the tensor core's async-proxy reads
          xform_bar();
          if (xt == 0) mbar_arrive(&bar_a_full[stage]);
        }
      }
    } else {
      // ---------------- streaming mode: optional statistics pass, then a copy/normalise pass per chunk
      for (int li = 0; li < my_groups; ++li) {
        const int g = blockIdx.x + li * gridDim.x;
        const int S_act = min(p.S, p.B2 - g * p.S);
        const __nv_bfloat16* gb0 = p.src[0] + static_cast<size_t>(g) * gstride0;
        const __nv_bfloat16* gb1 = p.nsrc > 1 ? p.src[1] + static_cast<size_t>(g) * gstride1 : gb0;
        if (p.groups > 0) {
          // deterministic statistics: per-thread partial records reduced in a fixed order, in batches of samples
          const int spb = max(1, STAT_PAIRS / p.KC);
          for (int sb = 0; sb < S_act; sb += spb) {
            const int ns = min(spb, S_act - sb);
            const int pairs = ns * p.KC;
            int PS = XFORM_THREADS / pairs;
            PS = max(1, min(PS, P));
            for (int item = xt; item < pairs * PS; item += XFORM_THREADS) {
              const int pair = item % pairs, slice = item / pairs;
              const int s = pair / p.KC, kc = pair - s * p.KC;
              const int which = (kc * 8 < p.C[0]) ? 0 : 1;
              const __nv_bfloat16* base = (which ? gb1 - p.C[0] : gb0) + kc * 8;
              const int* toff = t_off + (which ? p.S * P : 0) + (sb + s) * P;
              float sum[8], sq[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) { sum[j] = 0.0f; sq[j] = 0.0f; }
              for (int px = slice; px < P; px += 4 * PS) {
                uint4 raw[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  if (px + u * PS < P) raw[u] = *reinterpret_cast<const uint4*>(base + toff[px + u * PS]);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  if (px + u * PS < P) {
                    float f[8];
                    unpack8(raw[u], f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { sum[j] += f[j]; sq[j] = fmaf(f[j], f[j], sq[j]); }
                  }
              }
              float4* dst = reinterpret_cast<float4*>(s_scr + (static_cast<size_t>(slice) * pairs + pair) * 16);
              dst[0] = make_float4(sum[0], sum[1], sum[2], sum[3]);
              dst[1] = make_float4(sum[4], sum[5], sum[6], sum[7]);
              dst[2] = make_float4(sq[0], sq[1], sq[2], sq[3]);
              dst[3] = make_float4(sq[4], sq[5], sq[6], sq[7]);
            }
            xform_bar();
            stat_reduce(sb, ns, pairs, PS);
            xform_bar();
          }
        }
        const int items = S_act * P * 8;
        for (int chunk = 0; chunk < p.nchunks; ++chunk, ++a_it) {
          const int stage = a_it % p.a_stages;
          if (a_it >= p.a_stages) mbar_wait(&bar_a_empty[stage], ((a_it / p.a_stages) - 1) & 1);
          uint4* a4 = reinterpret_cast<uint4*>(As + stage * p.a_stage_bytes);
          const int which = (chunk * 64 < p.C[0]) ? 0 : 1;  // C[0] is a multiple of 64 whenever there are two sources
          const __nv_bfloat16* base = (which ? gb1 - p.C[0] : gb0) + chunk * 64;
          const int* toff = t_off + (which ? p.S * P : 0);
          for (int b0 = xt; b0 < items; b0 += 8 * XFORM_THREADS) {
            uint4 raw[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int item = b0 + u * XFORM_THREADS;
              if (item < items) raw[u] = *reinterpret_cast<const uint4*>(base + toff[item >> 3] + (item & 7) * 8);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int item = b0 + u * XFORM_THREADS;
              if (item < items) {
                const int sp = item >> 3, kcl = item & 7;
                if (p.groups > 0) gn_apply(raw[u], sp / P, chunk * 64 + kcl * 8);
                a4[kcl * p.R + t_row[sp]] = raw[u];
              }
            }
          }
          fence_proxy_async_smem();
          xform_bar();
          if (xt == 0) mbar_arrive(&bar_a_full[stage]);
        }
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

static int conv_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = kNumSMs;
  }
  return sms;
}

// Fills the launch geometry for an op; returns RD_OK or an error.
int conv_make_params(const rd_op_conv& op, ConvParams& p, int& smem_bytes, int& grid) {
  RD_REQUIRE(op.nsrc == 1 || op.nsrc == 2, "conv: nsrc must be 1 or 2");
  RD_REQUIRE(op.ntaps == 9 || op.ntaps == 1, "conv: ntaps must be 9 or 1");
  RD_REQUIRE(op.H_in >= 1 && op.W_in >= 1 && op.H_in <= MAX_HW && op.W_in <= MAX_HW, "conv: H,W must be in [1,%d]", MAX_HW);
  RD_REQUIRE(op.C_out % 32 == 0 && op.C_out >= 32 && op.C_out <= 256, "conv: C_out %d unsupported", op.C_out);
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
  else RD_REQUIRE(op.src[0].C % 64 == 0, "conv: with two sources the first must have a multiple of 64 channels");
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
  p.w_slab_bytes = p.N * 128;
  p.n_slabs = p.nchunks * p.ntaps;

  // Tile geometry.  Per candidate tile count: samples per group, row efficiency, whether the TMEM
  // accumulators can be double-buffered (2*nt*N <= 512 columns) and whether the whole filter stays
  // resident in shared memory.  Score = useful rows per MMA row, discounted when the epilogue cannot
  // overlap the next group's MMAs.
  const int max_shift = (op.ntaps == 9) ? 2 * p.Wp + 2 : 0;
  const int smem_cap = 227 * 1024 - 2048;  // static __shared__ barriers + alignment slack
  const int valid_px = op.H_out * op.W_out;
  double best_score = -1.0;
  ConvParams best = p;
  for (int nt = 1; nt <= 4; ++nt) {
    if (nt * p.N > 512 || nt * 128 < p.rps) continue;
    ConvParams c = p;
    c.n_tiles = nt;
    c.R = (nt * 128 + max_shift) | 1;
    c.S = (nt * 128) / p.rps;
    if (op.samples_per_cta > 0 && op.samples_per_cta < c.S) c.S = op.samples_per_cta;
    c.acc_bufs = (2 * nt * p.N <= 512) ? 2 : 1;
    c.a_stage_bytes = (8 * c.R * 16 + 127) / 128 * 128;
    c.a_stages = (p.nchunks == 1) ? 2 : 3;
    c.w_resident = 1;
    if (conv_smem_layout(c).total > smem_cap) {
      c.w_resident = 0;
      c.w_stages = MAX_W_STAGES;
      while (c.w_stages > 2 && conv_smem_layout(c).total > smem_cap) --c.w_stages;
      if (conv_smem_layout(c).total > smem_cap && c.a_stages == 3) c.a_stages = 2;
      if (conv_smem_layout(c).total > smem_cap) continue;
    }
    // transform mode: register-cached when every (sample, chunk) pair gets a thread and <= RC_MAX pixels
    c.xmode = 0; c.rc_PS = 1;
    if (p.groups > 0 && c.S * p.KC <= XFORM_THREADS) {
      int ps = XFORM_THREADS / (c.S * p.KC);
      if (ps > p.H * p.W) ps = p.H * p.W;
      if ((p.H * p.W + ps - 1) / ps <= RC_MAX) { c.xmode = 1; c.rc_PS = ps; }
    }
    double score = static_cast<double>(c.S * valid_px) / (nt * 128);
    if (c.acc_bufs == 1) score *= 0.75;
    if (!c.w_resident) score *= (nt >= 2 ? 0.97 : 0.85);  // streamed weights are re-read per group: favour larger groups
    if (score > best_score) { best_score = score; best = c; }
  }
  RD_REQUIRE(best_score > 0, "conv: no tile geometry fits (Cin=%d N=%d rps=%d)", cin, p.N, p.rps);
  p = best;
  p.n_groups = (op.B2 + p.S - 1) / p.S;
  p.tmem_cols = next_pow2_cols(p.acc_bufs * p.n_tiles * p.N);
  smem_bytes = conv_smem_layout(p).total;
  const int sms = conv_num_sms();
  grid = p.n_groups < sms ? p.n_groups : sms;
  return RD_OK;
}

int conv_launch(const rd_op_conv& op, cudaStream_t st) {
  ConvParams p;
  int smem = 0, grid = 0;
  int rc = conv_make_params(op, p, smem, grid);
  if (rc != RD_OK) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "conv: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
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
