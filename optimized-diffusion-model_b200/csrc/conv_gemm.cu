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
  int tab = p.groups > 0 ? p.Cin * 12 + p.S * p.groups * 8 + STAT_SCRATCH_BYTES : 0;
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
  float* s_gamma = reinterpret_cast<float*>(smem + L.tab_off);
  float* s_beta = s_gamma + p.Cin;
  int* s_gidx = reinterpret_cast<int*>(s_beta + p.Cin);
  float* s_stat = reinterpret_cast<float*>(s_gidx + p.Cin);  // [S][G][2]: (mean, rstd)
  float* s_scr = s_stat + p.S * p.groups * 2;                // statistics scratch, STAT_SCRATCH_BYTES

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
    if (p.groups > 0) {
      for (int c = xt; c < p.Cin; c += XFORM_THREADS) {
        s_gamma[c] = p.gamma[c];
        s_beta[c] = p.beta[c];
        s_gidx[c] = c / p.cpg;
      }
    }
    xform_bar();
    int a_it = 0;
    for (int li = 0; li < my_groups; ++li) {
      const int g = blockIdx.x + li * gridDim.x;
      const int sample0 = g * p.S;
      const int S_act = min(p.S, p.B2 - sample0);
      if (p.groups > 0) {
        // ---- GroupNorm statistics of this group's samples (fp32), one pass over the bf16 input.
        // Deterministic: per-thread partial sums go to a scratch table and are reduced in a fixed order
        // (no atomics), in batches of samples so the scratch stays small.
        const int spb = max(1, STAT_PAIRS / p.KC);
        const float inv_n = 1.0f / static_cast<float>(p.cpg * P);
        for (int sb = 0; sb < S_act; sb += spb) {
          const int ns = min(spb, S_act - sb);
          const int pairs = ns * p.KC;
          int PS = XFORM_THREADS / pairs;
          PS = max(1, min(PS, P));
          for (int item = xt; item < pairs * PS; item += XFORM_THREADS) {
            const int pair = item % pairs, slice = item / pairs;
            const int s = pair / p.KC, kc = pair - s * p.KC;
            float sum[8], sq[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { sum[j] = 0.0f; sq[j] = 0.0f; }
            for (int px = slice; px < P; px += 4 * PS) {
              uint4 raw[4];
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (px + u * PS < P) raw[u] = *item_ptr(p, sample0 + sb + s, px + u * PS, kc);
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
            s_stat[2 * ((sb + s) * p.groups + g)] = mean;
            s_stat[2 * ((sb + s) * p.groups + g) + 1] = 1.0f / sqrtf(var + p.eps);
          }
          xform_bar();
        }
      }
      // ---- stage the normalised / activated operand, one 64-channel chunk per ring stage
      const int items = S_act * P * 8;
      for (int chunk = 0; chunk < p.nchunks; ++chunk, ++a_it) {
        const int stage = a_it % p.a_stages;
        if (a_it >= p.a_stages) mbar_wait(&bar_a_empty[stage], ((a_it / p.a_stages) - 1) & 1);
        uint4* a4 = reinterpret_cast<uint4*>(As + stage * p.a_stage_bytes);
        for (int base = xt; base < items; base += 4 * XFORM_THREADS) {
          uint4 raw[4];
          int kcl[4], row[4], sidx[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int item = base + u * XFORM_THREADS;
            if (item < items) {
              kcl[u] = item & 7;
              const int sp = item >> 3;
              const int s = sp / P, px = sp - s * P;
              const int y = px / p.W, x = px - y * p.W;
              sidx[u] = s;
              row[u] = s * p.rps + (y + p.pad) * p.Wp + (x + p.pad);
              raw[u] = *item_ptr(p, sample0 + s, px, chunk * 8 + kcl[u]);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int item = base + u * XFORM_THREADS;
            if (item < items) {
              if (p.groups > 0) {
                float f[8];
                unpack8(raw[u], f);
                const int c0 = (chunk * 8 + kcl[u]) * 8;
                const float4 g0 = *reinterpret_cast<const float4*>(s_gamma + c0), g1 = *reinterpret_cast<const float4*>(s_gamma + c0 + 4);
                const float4 b0 = *reinterpret_cast<const float4*>(s_beta + c0), b1 = *reinterpret_cast<const float4*>(s_beta + c0 + 4);
                const float gam[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                const float bet[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                const float* st = s_stat + sidx[u] * p.groups * 2;
                if ((p.cpg & 7) == 0) {  // one group per 8-channel item
                  const float2 mr = *reinterpret_cast<const float2*>(st + 2 * (c0 / p.cpg));
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float a = gam[j] * mr.y;
                    f[j] = fmaf(f[j], a, fmaf(-mr.x, a, bet[j]));
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
                __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&raw[u]);
#pragma unroll
                for (int j = 0; j < 4; ++j) h2[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
              }
              a4[kcl[u] * p.R + row[u]] = raw[u];
            }
          }
        }
        fence_proxy_async_smem();  // st.shared above must be visible to the tensor core's async-proxy reads
        xform_bar();
        if (xt == 0) mbar_arrive(&bar_a_full[stage]);
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
