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
//       warp 0        MMA issuer (one elected thread, tcgen05.mma M=128 x N=C_out x K=16); the whole filter is
//                     fetched once per CTA by this thread when it fits in shared memory
//       warp 1        weight producer when the filter is streamed instead (1-D bulk TMA copies from L2 through
//                     a ring recycled by tcgen05.commit)
//       warps 4-7     epilogue: TMEM -> registers (tcgen05.ld 32x32b) -> bias/temb/skip -> bf16 NHWC
//       warps 2,3,8-15 transform: global bf16 -> GroupNorm statistics -> normalise + SiLU -> operand ring
//     (Role placement is measured, not arbitrary: issuing MMAs from several warps, or moving the issuer /
//     epilogue to other warp slots, was slower on B200 -- see DESIGN.md.)
//     While the tensor core works on group g, the transform warps stage group g+1 and the epilogue
//     warps drain group g-1 (double-buffered TMEM accumulators).
//   * The operand image of a 64-channel chunk is staged ONCE as [8 k-chunks][R rows][8 ch] (K-major,
//     SWIZZLE_NONE, rows 16 B apart) with one shared zero column / zero row between image rows /
//     samples, so the A operand of tap (dy,dx) is the same buffer with the UMMA descriptor start
//     address advanced by (dy*Wp+dx)*16 B: nine taps, no im2col, no data movement (validated on B200 by
//     tools/probe_umma.cu).  Outputs at padding rows are computed and dropped.
//   * Stride 2 (Downsample): the image is staged as its four parity planes (y&1, x&1) of (Ho+1) x (Wo+1) positions, one
//     after the other in the same buffer, so tap (dy,dx) is plane (dy&1, dx&1) advanced by ((dy>>1)*Wp + (dx>>1)) rows
//     and the accumulator rows are the OUTPUT positions only (a one-plane layout computes all H x W positions and keeps
//     a quarter: 0.155 -> 0.090 ms for the 8x9 -> 4x4 launch).  Still nine descriptor offsets, still no data movement.
//   * GroupNorm groups that do not align with the 8-channel operand items (192 channels / 32 groups = 6) and the fp32-class
//     plan take the streaming transform: a statistics pass (fixed-order shared-memory reduction), a per-(sample, channel)
//     affine table written over the idle statistics scratch, then a normalise pass per 64-channel chunk that reads its
//     coefficients with four 16-byte shared loads per item; the next group's pixels are requested into L2 meanwhile.
//   * All index arithmetic is group-invariant and tabulated once per CTA in shared memory.
//   * Two precisions (template parameter X3).  bf16: bf16 NHWC activations, one MMA per k-step.  fp32-class ("x3"):
//     fp32 NHWC activations; both operands are split into bf16 hi + bf16 lo (x = hi + lo to 2^-17) and every k-step
//     issues lo*hi + hi*lo + hi*hi into the same fp32 TMEM accumulator (tools/probe_split.cu: 5e-6 .. 1e-5 of the
//     output range against fp64 for K = 576 .. 2304, i.e. better than the TF32 convolutions the reference runs by
//     default on a GPU); GroupNorm / SiLU / bias / temb / skip stay in fp32 and SiLU uses libm tanh.
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include "rd_common.h"
#include "rd_ptx.cuh"

// Role-cost ablations for MEASUREMENT BUILDS ONLY (rdb200/build.py `defines`, tools/run_ablate.sh): each macro removes one
// role's work while keeping every barrier hand-off, so the per-launch times show which role a layer waits for.  The
// product library defines none of them; results of an ablated build are garbage.
//   RD_ABL_NO_MMA    no tcgen05.mma issued          RD_ABL_NO_EPI   epilogue does no TMEM load / math / global access
//   RD_ABL_NO_XFORM  transform does no global load / math / st.shared   RD_ABL_NO_WSTREAM  streamed filters not copied
//   register-cached transform only:  RD_ABL_XF_NOSTATS no statistics phase (records, reductions, their two barriers)
//   RD_ABL_XF_NOAPPLY no normalise / SiLU arithmetic (raw pixels are stored)   RD_ABL_XF_NOLOAD no global loads
namespace rd {

constexpr int CONV_THREADS = 512;
constexpr int XFORM_THREADS = 320;  // warps 2,3,8..15
constexpr int EPI_WARPS = 4;        // warps 4..7
constexpr int MAX_A_STAGES = 3;
constexpr int MAX_W_STAGES = 12;
constexpr int MAX_HW = 32;                              // gather maps
#ifndef RD_STREAM_SB
#define RD_STREAM_SB 6  // 16-byte loads in flight per thread in the streaming statistics pass (bf16 plan)
#endif
constexpr int STAT_PAIRS = 256;                         // (sample, 8-channel chunk) pairs per statistics batch (streaming mode)
constexpr int STAT_SCRATCH_BYTES = XFORM_THREADS * 64;  // one 16-float partial record per transform thread

// GroupNorm flavour of the operand transform (template parameter GNM)
constexpr int GNM_NONE = 0;     // plain copy (NIN shortcut, up/down-sampling convs, attention output projection)
constexpr int GNM_CPG8 = 1;     // channels-per-group multiple of 8: one group per 8-channel item
constexpr int GNM_CPG4 = 2;     // 4 channels per group: two groups per item
constexpr int GNM_GENERAL = 3;  // anything else (192 channels / 32 groups = 6)

struct ConvParams {
  const void* src[2];  // bf16 (or fp32 when x3) NHWC
  int C[2], Hs[2], Ws[2];
  int nsrc;
  // fused shortcut (ResBlock NIN_0 folded into Conv_1's launch): raw pixels of these sources are staged as sc_chunks extra
  // 64-channel chunks after the normalised ones and multiplied by the 1x1 filter slabs that follow the 3x3 ones in `w`,
  // into the SAME accumulator (centre tap only) -- the shortcut costs neither a launch nor a residual read
  const void* sc_src[2];
  int sc_C[2], sc_Hs[2], sc_Ws[2];
  int sc_nsrc, sc_chunks;
  int H, W;          // logical (gathered) input image
  int Wp, rps;       // padded row width, rows per sample in the staged image
  int pad;           // pixel (y,x) is staged at (y+pad, x+pad)
  int stride;        // output (oy,ox) = accumulator row (oy*stride, ox*stride)
  // stride 2, polyphase layout: the image is staged as four parity planes (y&1, x&1) of (Ho+1) x (Wo+1) positions, `plane_rows`
  // rows apart, so tap (dy,dx) reads plane (dy&1, dx&1) shifted by (dy>>1, dx>>1) and the accumulator rows are the OUTPUT
  // positions only (the one-plane layout computes every input position and keeps a quarter of them)
  int poly, plane_rows;
  int Ho, Wo;
  int ntaps;         // 9 or 1
  int Cin, KC;       // total input channels, Cin/8
  int nchunks;       // Cin/64
  int N;             // C_out of this launch (GEMM N)
  int out_stride;    // channels per pixel of `out` / `residual` (>= N: the launch may write a channel slice)
  int x3;            // fp32-class mode: fp32 activations, split-bf16 operands, three MMAs per k-step
  int S, n_tiles, R; // samples per group, 128-row accumulator tiles, staged rows (odd)
  int n_groups;
  int a_stages, a_stage_bytes;
  int w_resident, w_stages, w_slab_bytes, n_slabs;
  int acc_bufs;
  int xmode, rc_PS;  // transform mode: >0 = register-cached single pass (value = register slots, rc_PS pixel slices), 0 = streaming
  int gnm;
  int tmem_cols;     // power of two >= acc_bufs*n_tiles*N
  int groups, cpg, silu;
  float eps;
  const float* gamma;
  const float* beta;
  const __nv_bfloat16* w;  // [nchunks][ntaps][8][N][8]   (x3: [nchunks][ntaps][hi|lo][8][N][8])
  const float* bias;
  const float* tproj;
  int tproj_stride, tproj_off, tproj_wrap;
  const void* residual;
  float out_scale;
  void* out;
  int B2;
  unsigned char ymap[2][MAX_HW], xmap[2][MAX_HW];
  unsigned char sc_ymap[2][MAX_HW], sc_xmap[2][MAX_HW];
};

// dynamic shared memory carve-up (bytes, every region 128-B aligned)
struct ConvSmemLayout {
  int a_off, w_off, toff_off, trow_off, torow_off, tos_off, bt_off, gn_off, total;
};

__host__ __device__ inline int align128(int v) { return (v + 127) / 128 * 128; }

__host__ __device__ inline ConvSmemLayout conv_smem_layout(const ConvParams& p) {
  ConvSmemLayout L;
  const int SP = p.S * p.H * p.W, rows = p.n_tiles * 128;
  L.a_off = 0;
  L.w_off = p.a_stages * p.a_stage_bytes;
  L.toff_off = L.w_off + (p.w_resident ? p.n_slabs : p.w_stages) * p.w_slab_bytes;
  L.trow_off = L.toff_off + align128((p.sc_chunks > 0 ? 4 : 2) * SP * 4);  // source element offsets [2 (+2 shortcut)][S*P]
  L.torow_off = L.trow_off + align128(SP * 2);     // staged row of every (sample, pixel)
  L.tos_off = L.torow_off + align128(rows * 4);    // output element offset of every accumulator row (-1: dropped)
  L.bt_off = L.tos_off + align128(rows);           // sample index of every accumulator row
  L.gn_off = L.bt_off + align128(2 * p.S * p.N * 4);  // (bias + temb projection) * out_scale, double-buffered
  const int gn = p.groups > 0 ? p.Cin * 12 + p.S * p.groups * 8 + STAT_SCRATCH_BYTES : 0;
  L.total = L.gn_off + align128(gn);
  return L;
}

// SiLU(y) = y*sigmoid(y) = h + h*tanh(h) with h = y/2.  The GroupNorm affine is pre-halved when SiLU follows, so the
// transform costs FFMA + MUFU.TANH + FFMA per channel.  tanh.approx.f32 has 2^-11 relative error, a quarter of the
// bf16 rounding that follows.
__device__ __forceinline__ float tanh_approx(float x) {
#ifdef RD_EXACT_ACT  // error-budget builds only (tools/run_errbudget.sh): libm tanh instead of MUFU.TANH
  return tanhf(x);
#else
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}

// Programmatic dependent launch: the next conv launch of the stream may start its prologue (barriers, TMEM, index
// tables, operand-ring zeroing, filter fetch) on SMs this grid has already left; it must not touch activations
// before pdl_wait() (= the previous grid has completed and its writes are visible).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void xform_bar() { asm volatile("bar.sync 1, %0;" ::"n"(XFORM_THREADS) : "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 2, %0;" ::"n"(EPI_WARPS * 32) : "memory"); }

__device__ __forceinline__ void unpack8(const uint4& raw, float (&f)[8]) {
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __uint_as_float(w[j] << 16);
    f[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
  }
}

__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    w[j] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// All MMAs of one filter tap over one 64-channel chunk: 4 k-steps x NT accumulator tiles.  NT is a compile-time
// constant so that no MMA is predicated: with run-time tile counts ptxas re-materialises the (zero, idesc) and
// descriptor-high uniform-register pairs for every instruction, and the single issuing thread -- whose uniform-
// datapath instructions cost ~10 cycles each -- drops from the tensor core's 50 cycles per MMA to 70.
template <int NT, bool TILE_OUTER, bool X3>
__device__ __forceinline__ void issue_tap(uint32_t acc, uint32_t N, uint32_t a_lo, uint32_t w_lo, uint32_t kstep_a,
                                          uint32_t kstep_w, uint64_t desc_hi, uint32_t idesc, uint32_t accum,
                                          uint32_t a_half, uint32_t w_half) {
  // one k-step of one tile: a single MMA, or (x3) the three split-bf16 terms, small ones first
  auto kstep = [&](uint32_t d, uint32_t a, uint32_t w, uint32_t acc_flag) {
#ifdef RD_ABL_NO_MMA
    return;
#endif
    if (X3) {
      umma_bf16_ss(d, desc_hi | (a + a_half), desc_hi | w, idesc, acc_flag);  // lo * hi
      umma_bf16_ss(d, desc_hi | a, desc_hi | (w + w_half), idesc, 1u);        // hi * lo
      umma_bf16_ss(d, desc_hi | a, desc_hi | w, idesc, 1u);                   // hi * hi
    } else {
      umma_bf16_ss(d, desc_hi | a, desc_hi | w, idesc, acc_flag);
    }
  };
  if (!TILE_OUTER) {
    // narrow N: the tiles of one k-step together (they share the B descriptor)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t a_k = a_lo + kk * kstep_a, w_k = w_lo + kk * kstep_w;
#pragma unroll
      for (int t = 0; t < NT; ++t) kstep(acc + t * N, a_k + t * 128, w_k, kk == 0 ? accum : 1u);
    }
  } else {
    // N >= 128: the four k-steps of a tile back to back on the same accumulator (65 vs 86 cycles per MMA on
    // B200, tools/probe_umma2.cu)
#pragma unroll
    for (int t = 0; t < NT; ++t) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) kstep(acc + t * N, a_lo + t * 128 + kk * kstep_a, w_lo + kk * kstep_w, kk == 0 ? accum : 1u);
    }
  }
}

// Epilogue body for one block (32 accumulator columns of the row this thread owns): scale, add the staged
// (bias + temb) row, optionally add the residual, convert to bf16 and store 64 contiguous bytes.
template <bool RES>
__device__ __forceinline__ void epi_finish_block(const uint32_t (&v)[32], const float* __restrict__ btr, float oscale,
                                                 const u32x8 (&resv)[2], __nv_bfloat16* dst) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    u32x8 t;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int o = 16 * h + 8 * j;
      const float4 b0 = *reinterpret_cast<const float4*>(btr + o), b1 = *reinterpret_cast<const float4*>(btr + o + 4);
      float a[8] = {fmaf(__uint_as_float(v[o]), oscale, b0.x),     fmaf(__uint_as_float(v[o + 1]), oscale, b0.y),
                    fmaf(__uint_as_float(v[o + 2]), oscale, b0.z), fmaf(__uint_as_float(v[o + 3]), oscale, b0.w),
                    fmaf(__uint_as_float(v[o + 4]), oscale, b1.x), fmaf(__uint_as_float(v[o + 5]), oscale, b1.y),
                    fmaf(__uint_as_float(v[o + 6]), oscale, b1.z), fmaf(__uint_as_float(v[o + 7]), oscale, b1.w)};
      if (RES) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t w = resv[h].v[4 * j + e];
          a[2 * e] = fmaf(__uint_as_float(w << 16), oscale, a[2 * e]);
          a[2 * e + 1] = fmaf(__uint_as_float(w & 0xffff0000u), oscale, a[2 * e + 1]);
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        __nv_bfloat162 hh = __floats2bfloat162_rn(a[2 * e], a[2 * e + 1]);
        t.v[4 * j + e] = *reinterpret_cast<uint32_t*>(&hh);
      }
    }
    st_global_256(dst + 16 * h, t);
  }
}

// One accumulator row of one 128-row tile, walked in blocks of 32 columns.  With a residual, the 64 bytes of the next
// block are requested before the current block is converted (two register sets, no copies).
template <bool RES>
__device__ __forceinline__ void epi_row(uint32_t tacc, int N, const float* __restrict__ btr, float oscale,
                                        const __nv_bfloat16* __restrict__ rrow, __nv_bfloat16* orow, bool valid) {
  u32x8 ra[2], rb[2];
  if (RES && valid) { ra[0] = ld_global_256(rrow); ra[1] = ld_global_256(rrow + 16); }
  for (int c0 = 0; c0 < N; c0 += 64) {
    uint32_t v[32];
    tmem_ld32(tacc + c0, v);
    if (RES && valid && c0 + 32 < N) { rb[0] = ld_global_256(rrow + c0 + 32); rb[1] = ld_global_256(rrow + c0 + 48); }
    tmem_ld_wait32(v);
    if (valid) epi_finish_block<RES>(v, btr + c0, oscale, ra, orow + c0);
    if (c0 + 32 < N) {
      tmem_ld32(tacc + c0 + 32, v);
      if (RES && valid && c0 + 64 < N) { ra[0] = ld_global_256(rrow + c0 + 64); ra[1] = ld_global_256(rrow + c0 + 80); }
      tmem_ld_wait32(v);
      if (valid) epi_finish_block<RES>(v, btr + c0 + 32, oscale, rb, orow + c0 + 32);
    }
  }
}

// fp32-class epilogue: the same row walk with fp32 residual / output (128 bytes per 32-column block and thread).
template <bool RES>
__device__ __forceinline__ void epi_row_f32(uint32_t tacc, int N, const float* __restrict__ btr, float oscale,
                                            const float* __restrict__ rrow, float* orow, bool valid) {
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tacc + c0, v);
    u32x8 r[4];
    if (RES && valid) {
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = ld_global_256(rrow + c0 + 8 * j);
    }
    tmem_ld_wait32(v);
    if (valid) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 b0 = *reinterpret_cast<const float4*>(btr + c0 + 8 * j), b1 = *reinterpret_cast<const float4*>(btr + c0 + 8 * j + 4);
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        u32x8 t;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float a = fmaf(__uint_as_float(v[8 * j + e]), oscale, bb[e]);
          if (RES) a = fmaf(__uint_as_float(r[j].v[e]), oscale, a);
          t.v[e] = __float_as_uint(a);
        }
        st_global_256(orow + c0 + 8 * j, t);
      }
    }
  }
}

// x ~= hi + lo with both halves bf16 (round to nearest): |x - hi - lo| <= 2^-17 |x|
__device__ __forceinline__ void split8(const float (&f)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    const float2 hf = __bfloat1622float2(hh);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(f[2 * j] - hf.x, f[2 * j + 1] - hf.y);
    h[j] = *reinterpret_cast<const uint32_t*>(&hh);
    l[j] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void ldg8f(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

template <int GNM, int RC, bool X3>
__global__ void __launch_bounds__(CONV_THREADS, 1) conv_gemm_kernel(const __grid_constant__ ConvParams p) {
  static_assert(!X3 || RC == 0, "the fp32-class mode uses the streaming transform");
  using act_t = typename std::conditional<X3, float, __nv_bfloat16>::type;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar_a_full[MAX_A_STAGES], bar_a_empty[MAX_A_STAGES];
  __shared__ uint64_t bar_w_full[MAX_W_STAGES], bar_w_empty[MAX_W_STAGES];
  __shared__ uint64_t bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint32_t tmem_slot;

  pdl_launch_dependents();
  const ConvSmemLayout L = conv_smem_layout(p);
  unsigned char* As = smem + L.a_off;
  unsigned char* Ws = smem + L.w_off;
  int* t_off = reinterpret_cast<int*>(smem + L.toff_off);
  unsigned short* t_row = reinterpret_cast<unsigned short*>(smem + L.trow_off);
  int* t_orow = reinterpret_cast<int*>(smem + L.torow_off);
  unsigned char* t_os = smem + L.tos_off;
  float* s_bt = reinterpret_cast<float*>(smem + L.bt_off);
  float* s_gamma = reinterpret_cast<float*>(smem + L.gn_off);
  float* s_beta = s_gamma + p.Cin;
  int* s_gidx = reinterpret_cast<int*>(s_beta + p.Cin);
  float* s_stat = reinterpret_cast<float*>(s_gidx + p.Cin);  // [S][G][2]: (mean, rstd)
  float* s_scr = s_stat + p.S * p.groups * 2;                // statistics scratch (16-B aligned: S*G is even)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int P = p.H * p.W;
  const int my_groups = (p.n_groups > static_cast<int>(blockIdx.x))
                            ? (p.n_groups - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                            : 0;

  // ------------------------------------------------------------------ setup
  if (tid == 0) {
    for (int i = 0; i < MAX_A_STAGES; ++i) { mbar_init(&bar_a_full[i], XFORM_THREADS / 32); mbar_init(&bar_a_empty[i], 1); }
    for (int i = 0; i < MAX_W_STAGES; ++i) { mbar_init(&bar_w_full[i], 1); mbar_init(&bar_w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_acc_full[i], 1); mbar_init(&bar_acc_empty[i], EPI_WARPS); }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, p.tmem_cols);
  // group-invariant index tables (all threads help)
  for (int sp = tid; sp < p.S * P; sp += CONV_THREADS) {
    const int s = sp / P, px = sp - s * P;
    const int y = px / p.W, x = px - y * p.W;
    t_row[sp] = p.poly ? static_cast<unsigned short>(((y & 1) * 2 + (x & 1)) * p.plane_rows + s * p.rps + (y >> 1) * p.Wp + (x >> 1))
                       : static_cast<unsigned short>(s * p.rps + (y + p.pad) * p.Wp + (x + p.pad));
    t_off[sp] = ((s * p.Hs[0] + p.ymap[0][y]) * p.Ws[0] + p.xmap[0][x]) * p.C[0];
    if (p.nsrc > 1) t_off[p.S * P + sp] = ((s * p.Hs[1] + p.ymap[1][y]) * p.Ws[1] + p.xmap[1][x]) * p.C[1];
    if (p.sc_chunks > 0) {
      t_off[2 * p.S * P + sp] = ((s * p.sc_Hs[0] + p.sc_ymap[0][y]) * p.sc_Ws[0] + p.sc_xmap[0][x]) * p.sc_C[0];
      if (p.sc_nsrc > 1) t_off[3 * p.S * P + sp] = ((s * p.sc_Hs[1] + p.sc_ymap[1][y]) * p.sc_Ws[1] + p.sc_xmap[1][x]) * p.sc_C[1];
    }
  }
  for (int row = tid; row < p.n_tiles * 128; row += CONV_THREADS) {
    const int s = row / p.rps, rem = row - s * p.rps;
    const int Y = rem / p.Wp, X = rem - Y * p.Wp;
    bool valid = s < p.S;
    int oy = Y, ox = X;
    if (p.stride == 2 && !p.poly) {
      valid = valid && ((Y & 1) == 0) && ((X & 1) == 0);
      oy = Y >> 1; ox = X >> 1;
    }
    valid = valid && oy < p.Ho && ox < p.Wo;
    t_orow[row] = valid ? ((s * p.Ho + oy) * p.Wo + ox) * p.out_stride : -1;
    t_os[row] = static_cast<unsigned char>(valid ? s : 0);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    // ================================================================ MMA issuer
    // The whole warp runs the loop (warp-uniform control flow keeps the descriptor arithmetic on the uniform
    // datapath: inside a one-lane divergent region every tcgen05.mma operand needed its own R2UR move, and the tap
    // loop went through an indirect branch); only the elected lane issues MMAs, commits and copies.  The tile count
    // and the MMA order are compile-time constants of the loop body, chosen once per kernel.
    const bool leader = elect_one();
    const int N = p.N, ntaps = p.ntaps, nchunks = p.nchunks, Wp = p.Wp;
    const int a_stages = p.a_stages, w_stages = p.w_stages, acc_bufs = p.acc_bufs;
    const bool w_resident = p.w_resident != 0;
    const int sc_chunks = p.sc_chunks;
    const uint32_t sc_shift = static_cast<uint32_t>(p.pad * (p.Wp + 1));
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint64_t desc_hi = static_cast<uint64_t>((128u >> 4) | (1u << 14)) << 32;  // SBO = 128 B, version 1
    const uint32_t a_lo0 = (smem_u32(As) >> 4) | (static_cast<uint32_t>(p.R) << 16);  // LBO = R*16 B
    const uint32_t w_lo0 = (smem_u32(Ws) >> 4) | (static_cast<uint32_t>(N) << 16);    // LBO = N*16 B
    const uint32_t a_stage_u = p.a_stage_bytes >> 4, w_slab_u = p.w_slab_bytes >> 4;
    const uint32_t a_half = X3 ? (a_stage_u >> 1) : 0u, w_half = X3 ? (w_slab_u >> 1) : 0u;  // hi image, then lo image
    const uint32_t kstep_a = 2 * p.R, kstep_w = 2 * N;
    const uint32_t acc_stride = p.n_tiles * N;
    if (w_resident && my_groups > 0) {
      if (leader) {  // the resident filter is fetched once per CTA
        const unsigned char* wg = reinterpret_cast<const unsigned char*>(p.w);
        mbar_arrive_expect_tx(&bar_w_full[0], p.n_slabs * p.w_slab_bytes);
        for (int sidx = 0; sidx < p.n_slabs; ++sidx)
          bulk_g2s(Ws + sidx * p.w_slab_bytes, wg + static_cast<size_t>(sidx) * p.w_slab_bytes, p.w_slab_bytes, &bar_w_full[0]);
      }
      mbar_wait(&bar_w_full[0], 0);
      tc_fence_after_sync();
    }
    // The tap loop is unrolled at compile time (NTAPS = 9 or 1), the streamed and the resident filter get their own
    // copies of it, and nothing in it is looked up or divided: for N = 128 a tap is only eight MMAs (~540 cycles of
    // tensor-core work), and the ~130 scalar instructions per tap of a generic loop (ring index arithmetic, run-time
    // branches, register->uniform moves) kept the single issuing thread BEHIND the tensor core -- 102 cycles per MMA
    // issued against 67 executed (tools/probe_umma2.cu order 12).
    auto run = [&](auto nt_c, auto outer_c, auto taps_c) {
      constexpr int NT = decltype(nt_c)::value;
      constexpr bool TILE_OUTER = decltype(outer_c)::value;
      constexpr int NTAPS = decltype(taps_c)::value;
      int a_stage_i = 0, a_par = 0, ws = 0, w_par = 0;
      // (laundered through an opaque asm so that ptxas keeps the addresses in registers instead of re-deriving
      //  them from SR_CgaCtaId at every wait / commit)
      auto keep = [](uint32_t v) { uint32_t r; asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v)); return r; };
      const uint32_t wfull0 = keep(smem_u32(&bar_w_full[0])), wempty0 = keep(smem_u32(&bar_w_empty[0]));
      const uint32_t afull0 = keep(smem_u32(&bar_a_full[0])), aempty0 = keep(smem_u32(&bar_a_empty[0]));
      const uint32_t uN = static_cast<uint32_t>(N);
      // tap (dy,dx) -> row shift rs[dy] + cs[dx]: dy*Wp + dx, or (polyphase stride 2) plane (dy&1, dx&1) + (dy>>1)*Wp + (dx>>1)
      const uint32_t rs1 = p.poly ? 2u * p.plane_rows : static_cast<uint32_t>(Wp), rs2 = p.poly ? static_cast<uint32_t>(Wp) : 2u * Wp;
      const uint32_t cs1 = p.poly ? static_cast<uint32_t>(p.plane_rows) : 1u, cs2 = p.poly ? 1u : 2u;
      auto tap_shift = [&](int t) -> uint32_t {
        const int dy = t / 3, dx = t % 3;  // compile-time at every call site (the tap loops are fully unrolled)
        return (dy == 0 ? 0u : (dy == 1 ? rs1 : rs2)) + (dx == 0 ? 0u : (dx == 1 ? cs1 : cs2));
      };
      for (int li = 0; li < my_groups; ++li) {
        const int buf = li % acc_bufs, useb = li / acc_bufs;
        if (useb > 0) { mbar_wait(&bar_acc_empty[buf], (useb - 1) & 1); tc_fence_after_sync(); }
        const uint32_t acc = tmem + buf * acc_stride;
        for (int chunk = 0; chunk < nchunks; ++chunk) {
          const int stage = a_stage_i;
          mbar_wait_addr(afull0 + 8 * stage, a_par);
          tc_fence_after_sync();
          const uint32_t a_lo_stage = a_lo0 + stage * a_stage_u;
          if (w_resident) {
            const uint32_t w_chunk = w_lo0 + chunk * NTAPS * w_slab_u;
#pragma unroll
            for (int t = 0; t < NTAPS; ++t) {
              const uint32_t shift = NTAPS == 9 ? tap_shift(t) : 0u;
              if (leader)
                issue_tap<NT, TILE_OUTER, X3>(acc, uN, a_lo_stage + shift, w_chunk + t * w_slab_u, kstep_a, kstep_w, desc_hi,
                                              idesc, (chunk | t) != 0, a_half, w_half);
            }
          } else {
#pragma unroll
            for (int t = 0; t < NTAPS; ++t) {
              const uint32_t shift = NTAPS == 9 ? tap_shift(t) : 0u;
              mbar_wait_addr(wfull0 + 8 * ws, w_par);
              tc_fence_after_sync();
              if (leader) {
                issue_tap<NT, TILE_OUTER, X3>(acc, uN, a_lo_stage + shift, w_lo0 + ws * w_slab_u, kstep_a, kstep_w, desc_hi,
                                              idesc, (chunk | t) != 0, a_half, w_half);
                umma_commit_addr(wempty0 + 8 * ws);
              }
              if (++ws == w_stages) { ws = 0; w_par ^= 1; }
            }
          }
          if (leader) umma_commit_addr(aempty0 + 8 * stage);  // operand stage reusable once these MMAs have read it
          if (++a_stage_i == a_stages) { a_stage_i = 0; a_par ^= 1; }
        }
        // fused shortcut: raw chunks x 1x1 slabs, centre tap (row shift pad * (Wp + 1)), same accumulator
        for (int sc = 0; sc < sc_chunks; ++sc) {
          const int stage = a_stage_i;
          mbar_wait_addr(afull0 + 8 * stage, a_par);
          tc_fence_after_sync();
          const uint32_t a_lo_stage = a_lo0 + stage * a_stage_u + sc_shift;
          if (w_resident) {
            if (leader)
              issue_tap<NT, TILE_OUTER, X3>(acc, uN, a_lo_stage, w_lo0 + (nchunks * NTAPS + sc) * w_slab_u, kstep_a, kstep_w, desc_hi,
                                            idesc, 1u, a_half, w_half);
          } else {
            mbar_wait_addr(wfull0 + 8 * ws, w_par);
            tc_fence_after_sync();
            if (leader) {
              issue_tap<NT, TILE_OUTER, X3>(acc, uN, a_lo_stage, w_lo0 + ws * w_slab_u, kstep_a, kstep_w, desc_hi, idesc, 1u, a_half,
                                            w_half);
              umma_commit_addr(wempty0 + 8 * ws);
            }
            if (++ws == w_stages) { ws = 0; w_par ^= 1; }
          }
          if (leader) umma_commit_addr(aempty0 + 8 * stage);
          if (++a_stage_i == a_stages) { a_stage_i = 0; a_par ^= 1; }
        }
        if (leader) umma_commit(&bar_acc_full[buf]);
      }
    };
    using std::integral_constant;
    auto run_taps = [&](auto nt_c, auto outer_c) {
      if (ntaps == 9) run(nt_c, outer_c, integral_constant<int, 9>{});
      else run(nt_c, outer_c, integral_constant<int, 1>{});
    };
    const int sel = (N <= 64 ? 0 : 4) + p.n_tiles - 1;  // narrow N: tiles of a k-step together; N >= 128: tile-outer
    switch (sel) {
      case 0: run_taps(integral_constant<int, 1>{}, integral_constant<bool, false>{}); break;
      case 1: run_taps(integral_constant<int, 2>{}, integral_constant<bool, false>{}); break;
      case 2: run_taps(integral_constant<int, 3>{}, integral_constant<bool, false>{}); break;
      case 3: run_taps(integral_constant<int, 4>{}, integral_constant<bool, false>{}); break;
      case 4: run_taps(integral_constant<int, 1>{}, integral_constant<bool, true>{}); break;
      case 5: run_taps(integral_constant<int, 2>{}, integral_constant<bool, true>{}); break;
      case 6: run_taps(integral_constant<int, 3>{}, integral_constant<bool, true>{}); break;
      default: run_taps(integral_constant<int, 4>{}, integral_constant<bool, true>{}); break;
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================================================ weight producer (streamed filters only)
    if (lane == 0 && my_groups > 0 && !p.w_resident) {
      const unsigned char* wg = reinterpret_cast<const unsigned char*>(p.w);
      const int total = my_groups * p.n_slabs;
      const int w_stages = p.w_stages, n_slabs = p.n_slabs;
      // (slot, parity) and the slab index are counters, not `%` / `/` of the iteration index: this single thread has
      // to turn a slab around in well under the ~500 cycles the tensor core needs to consume one
      int ws = 0, e_par = 1, slab = 0;
      for (int it = 0; it < total; ++it) {
        if (it >= w_stages) mbar_wait(&bar_w_empty[ws], e_par);
#ifdef RD_ABL_NO_WSTREAM
        mbar_arrive(&bar_w_full[ws]);
#else
        mbar_arrive_expect_tx(&bar_w_full[ws], p.w_slab_bytes);
        bulk_g2s(Ws + ws * p.w_slab_bytes, wg + static_cast<size_t>(slab) * p.w_slab_bytes, p.w_slab_bytes, &bar_w_full[ws]);
#endif
        if (++ws == w_stages) { ws = 0; e_par ^= 1; }   // first wait on a slot (round 1) uses parity 0
        if (++slab == n_slabs) slab = 0;                  // natural (chunk, tap) order, restarted for every group
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 8) {
    // ================================================================ epilogue
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int et = q * 32 + lane;
    const int N = p.N, n_tiles = p.n_tiles, S = p.S;
    const float oscale = p.out_scale;
    const size_t out_gstride = static_cast<size_t>(S) * p.Ho * p.Wo * p.out_stride;
    // (bias + Dense_0(SiLU(temb))) * out_scale for every (sample, channel) of group g, written straight into one
    // half of the double-buffered table in shared memory (the other half is being read by the current group).
    auto bt_fill = [&](float* dst, int g) {
      const int n = min(S, p.B2 - g * S) * N;
      for (int i = et; i < n; i += EPI_WARPS * 32) {
        const int s = i / N, c = i - s * N;
        float v = __ldg(p.bias + c);
        if (p.tproj) {
          int row = g * S + s;
          if (p.tproj_wrap > 0 && row > p.tproj_wrap) row = p.tproj_wrap;  // shared unconditional-CFG row
          v += __ldg(p.tproj + static_cast<size_t>(row) * p.tproj_stride + p.tproj_off + c);
        }
        dst[i] = v * oscale;
      }
    };
    pdl_wait();  // tproj / residual come from earlier launches
    if (my_groups > 0) bt_fill(s_bt, blockIdx.x);
    act_t* const outp = static_cast<act_t*>(p.out);
    const act_t* const resp = static_cast<const act_t*>(p.residual);
    for (int li = 0; li < my_groups; ++li) {
      const int g = blockIdx.x + li * gridDim.x;
      const int S_act = max(0, min(S, p.B2 - g * S));
      const int buf = li % p.acc_bufs;
      const float* bt = s_bt + (li & 1) * S * N;
      epi_bar();  // this group's table is complete; the other half is no longer read by anyone
      // next group's table: its global-load latency hides behind the accumulator wait below.  (Requesting it into
      // registers before the rows and storing it after them was measured slower: +1.5 % conv time from the extra
      // register pressure in this 128-register role.)
      if (li + 1 < my_groups) bt_fill(s_bt + ((li + 1) & 1) * S * N, g + gridDim.x);
      mbar_wait(&bar_acc_full[buf], (li / p.acc_bufs) & 1);
      tc_fence_after_sync();
      const uint32_t acc = tmem + buf * n_tiles * N + (static_cast<uint32_t>(q * 32) << 16);
      act_t* og = outp + static_cast<size_t>(g) * out_gstride;
      const act_t* rg = resp ? resp + static_cast<size_t>(g) * out_gstride : nullptr;
      const int valid_limit = S_act * p.Ho * p.Wo * p.out_stride;
      // A thread owns one accumulator row per 128-row tile and walks it in blocks of 32 columns.  One warp per TMEM
      // lane quarter runs this loop, so it is bound by dependent-instruction latency: per-row state is computed once
      // per tile, nothing is divided, and the residual of the next block is requested before the current one is used.
      for (int tile = 0; tile < n_tiles; ++tile) {
        const int row = tile * 128 + et;
        const int orow = t_orow[row];
#ifdef RD_ABL_NO_EPI
        continue;
#endif
        const bool valid = orow >= 0 && orow < valid_limit;
        const float* btr = bt + t_os[row] * N;
        if constexpr (X3) {
          if (rg) epi_row_f32<true>(acc + tile * N, N, btr, oscale, rg + orow, og + orow, valid);
          else epi_row_f32<false>(acc + tile * N, N, btr, oscale, nullptr, og + orow, valid);
        } else {
          if (rg) epi_row<true>(acc + tile * N, N, btr, oscale, rg + orow, og + orow, valid);
          else epi_row<false>(acc + tile * N, N, btr, oscale, nullptr, og + orow, valid);
        }
      }
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
    if (GNM != GNM_NONE) {
      for (int c = xt; c < p.Cin; c += XFORM_THREADS) {
        s_gamma[c] = p.gamma[c];
        s_beta[c] = p.beta[c];
        s_gidx[c] = c / p.cpg;
      }
    }
    xform_bar();
    pdl_wait();  // the activations read below come from earlier launches
    const act_t* const src0 = static_cast<const act_t*>(p.src[0]);
    const act_t* const src1 = static_cast<const act_t*>(p.src[1]);
    const size_t gstride0 = static_cast<size_t>(p.S) * p.Hs[0] * p.Ws[0] * p.C[0];
    const size_t gstride1 = static_cast<size_t>(p.S) * p.Hs[1] * p.Ws[1] * p.C[1];
    const float inv_n = GNM != GNM_NONE ? 1.0f / static_cast<float>(p.cpg * P) : 0.0f;
    int a_it = 0;

    // per-channel affine (a, b) with y = x*a + b for the 8 channels starting at c0 of sample s
    auto gn_coeffs = [&](int s, int c0, float (&ca)[8], float (&cb)[8]) {
      const float4 g0 = *reinterpret_cast<const float4*>(s_gamma + c0), g1 = *reinterpret_cast<const float4*>(s_gamma + c0 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(s_beta + c0), b1 = *reinterpret_cast<const float4*>(s_beta + c0 + 4);
      const float gam[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bet[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      const float* st = s_stat + s * p.groups * 2;
      if (GNM == GNM_CPG8) {
        const float2 mr = *reinterpret_cast<const float2*>(st + 2 * (c0 / p.cpg));
#pragma unroll
        for (int j = 0; j < 8; ++j) { ca[j] = gam[j] * mr.y; cb[j] = fmaf(-mr.x, ca[j], bet[j]); }
      } else if (GNM == GNM_CPG4) {
        const float4 mr = *reinterpret_cast<const float4*>(st + 2 * (c0 >> 2));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          ca[j] = gam[j] * (j < 4 ? mr.y : mr.w);
          cb[j] = fmaf(-(j < 4 ? mr.x : mr.z), ca[j], bet[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 mr = *reinterpret_cast<const float2*>(st + 2 * s_gidx[c0 + j]);
          ca[j] = gam[j] * mr.y;
          cb[j] = fmaf(-mr.x, ca[j], bet[j]);
        }
      }
      if (p.silu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { ca[j] *= 0.5f; cb[j] *= 0.5f; }
      }
    };
    // (ca, cb) are pre-halved by gn_coeffs when SiLU follows: h = y/2, SiLU(y) = h + h*tanh(h)
    auto gn_apply_f = [&](float (&f)[8], const float (&ca)[8], const float (&cb)[8]) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], ca[j], cb[j]);
      if (p.silu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], X3 ? tanhf(f[j]) : tanh_approx(f[j]), f[j]);
      }
    };
    auto gn_apply = [&](uint4& raw, const float (&ca)[8], const float (&cb)[8]) {
      float f[8];
      unpack8(raw, f);
      gn_apply_f(f, ca, cb);
      raw = pack8(f);
    };
    // fixed-order two-step reduction of the per-thread partial records into per-(sample, group) mean / rstd:
    // (1) every (pair, value) column is summed over the pixel slices by its own thread, (2) one thread per
    // (sample, group) adds its cpg channels.  Contains one transform barrier.
    auto stat_reduce = [&](int s_first, int ns, int pairs, int PS) {
      if (PS > 1) {
        for (int col = xt; col < pairs * 16; col += XFORM_THREADS) {
          float acc = s_scr[col];
          for (int slice = 1; slice < PS; ++slice) acc += s_scr[static_cast<size_t>(slice) * pairs * 16 + col];
          s_scr[col] = acc;
        }
        xform_bar();
      }
      for (int sg = xt; sg < ns * p.groups; sg += XFORM_THREADS) {
        const int s = sg / p.groups, g = sg - s * p.groups;
        float sum = 0.0f, sq = 0.0f;
        for (int c = g * p.cpg; c < (g + 1) * p.cpg; ++c) {
          const float* src = s_scr + static_cast<size_t>(s * p.KC + (c >> 3)) * 16 + (c & 7);
          sum += src[0];
          sq += src[8];
        }
        const float mean = sum * inv_n;
        const float var = fmaxf(sq * inv_n - mean * mean, 0.0f);
        s_stat[2 * ((s_first + s) * p.groups + g)] = mean;
        s_stat[2 * ((s_first + s) * p.groups + g) + 1] = 1.0f / sqrtf(var + p.eps);
      }
    };
    auto put_record = [&](int idx, const float (&sum)[8], const float (&sq)[8]) {
      float4* dst = reinterpret_cast<float4*>(s_scr + static_cast<size_t>(idx) * 16);
      dst[0] = make_float4(sum[0], sum[1], sum[2], sum[3]);
      dst[1] = make_float4(sum[4], sum[5], sum[6], sum[7]);
      dst[2] = make_float4(sq[0], sq[1], sq[2], sq[3]);
      dst[3] = make_float4(sq[4], sq[5], sq[6], sq[7]);
    };

    if constexpr (RC > 0) {
      // ---------------- team mode (GroupNorm groups of 4 or 8 channels): a TEAM of PS adjacent lanes owns one
      // (sample, 8-channel item) of the current 64-channel chunk and every lane RC of its pixels.  A group's statistics
      // then live entirely inside one team, so they are a handful of warp shuffles -- no shared-memory records, no
      // CTA-wide barrier -- and each warp hands its rows to the tensor core on its own (the operand-full barrier counts
      // the ten transform warps).  Every pixel is read from global memory once, one (group, chunk) step ahead of use.
      static_assert(GNM == GNM_CPG8 || GNM == GNM_CPG4, "team mode needs whole GroupNorm groups inside an 8-channel item");
      const int PS = p.rc_PS, ps_log = 31 - __clz(PS);
      const int team = xt >> ps_log, slice = xt & (PS - 1);
      const int s = team >> 3, kcl = team & 7;
      const bool owner = s < p.S;
      const unsigned short* trow = t_row + s * P;
      const int steps_per_group = p.nchunks + p.sc_chunks;  // normalised chunks, then raw shortcut chunks
      const int total = my_groups * steps_per_group;
      const act_t* const sc0 = static_cast<const act_t*>(p.sc_src[0]);
      const act_t* const sc1 = static_cast<const act_t*>(p.sc_src[1]);
      const size_t sc_gstride0 = static_cast<size_t>(p.S) * p.sc_Hs[0] * p.sc_Ws[0] * p.sc_C[0];
      const size_t sc_gstride1 = static_cast<size_t>(p.S) * p.sc_Hs[1] * p.sc_Ws[1] * p.sc_C[1];
      // RC <= 4: the next step's pixels land in a second register set while this step is processed.  Larger RC: one set
      // (two do not fit the 128-register budget without spilling into the other roles' loops); the next step's loads are
      // issued into the same registers right after this step's stores, so their latency overlaps the operand-stage wait.
      constexpr bool PREFETCH = RC <= 4;
      uint4 raw[RC], nxt[PREFETCH ? RC : 1];
      // step st = (group li, chunk): the pixels of this lane's (sample, item) in that chunk
      // `l2_only`: warm the L2 with the step's lines instead of loading them (single-register-set variants issue this one
      // step ahead, so that the loads they can only issue after this step's stores find their data on chip)
      auto load_step = [&](uint4* dst, int li, int chunk, bool l2_only = false) {
        const int g = blockIdx.x + li * gridDim.x;
#ifdef RD_ABL_NO_XFORM
        const bool active = false;
#else
        const bool active = owner && s < min(p.S, p.B2 - g * p.S);
#endif
        const __nv_bfloat16* gbase;
        const int* toff;
        if (chunk < p.nchunks) {
          const int c0 = chunk * 64 + kcl * 8;
          const int which = (c0 < p.C[0]) ? 0 : 1;
          gbase = reinterpret_cast<const __nv_bfloat16*>(which ? src1 : src0) + static_cast<size_t>(g) * (which ? gstride1 : gstride0) +
                  (c0 - (which ? p.C[0] : 0));
          toff = t_off + (which ? p.S * P : 0) + s * P;
        } else {
          const int c0 = (chunk - p.nchunks) * 64 + kcl * 8;
          const int which = (c0 < p.sc_C[0]) ? 0 : 1;
          gbase = reinterpret_cast<const __nv_bfloat16*>(which ? sc1 : sc0) + static_cast<size_t>(g) * (which ? sc_gstride1 : sc_gstride0) +
                  (c0 - (which ? p.sc_C[0] : 0));
          toff = t_off + (2 + which) * p.S * P + s * P;
        }
#pragma unroll
        for (int k = 0; k < RC; ++k) {
          const int px = slice + (k << ps_log);
#ifdef RD_ABL_XF_NOLOAD
          dst[k] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
#else
          if (active && px < P) {
            if (l2_only) asm volatile("prefetch.global.L2 [%0];" ::"l"(gbase + toff[px]));
            else dst[k] = __ldg(reinterpret_cast<const uint4*>(gbase + toff[px]));
          }
#endif
        }
      };
      if (total > 0) load_step(raw, 0, 0);
      int li = 0, chunk = 0;
      for (int st = 0; st < total; ++st, ++a_it) {
        const int g = blockIdx.x + li * gridDim.x;
        const int S_act = max(0, min(p.S, p.B2 - g * p.S));
#ifdef RD_ABL_NO_XFORM
        const bool active = false;
#else
        const bool active = owner && s < S_act;
#endif
        int li_n = li, chunk_n = chunk + 1;
        if (chunk_n == steps_per_group) { chunk_n = 0; ++li_n; }
        const bool normalise = chunk < p.nchunks;  // shortcut chunks are staged raw
        if (st + 1 < total) load_step(nxt, li_n, chunk_n, !PREFETCH);
        // ---- statistics of this lane's pixels: (sum, sum of squares) of channels 0-3 and 4-7, then over the team
        float ca[8], cb[8];
#ifndef RD_ABL_XF_NOSTATS
        if (normalise) {
          float s_lo = 0.0f, q_lo = 0.0f, s_hi = 0.0f, q_hi = 0.0f;
#pragma unroll
          for (int k = 0; k < RC; ++k) {
            if (active && slice + (k << ps_log) < P) {
              float f[8];
              unpack8(raw[k], f);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                s_lo += f[j]; q_lo = fmaf(f[j], f[j], q_lo);
                s_hi += f[4 + j]; q_hi = fmaf(f[4 + j], f[4 + j], q_hi);
              }
            }
          }
          if (GNM == GNM_CPG8) { s_lo += s_hi; q_lo += q_hi; }
          for (int o = PS >> 1; o > 0; o >>= 1) {  // teams are aligned power-of-two lane groups: fixed-order butterfly
            s_lo += __shfl_xor_sync(0xffffffffu, s_lo, o);
            q_lo += __shfl_xor_sync(0xffffffffu, q_lo, o);
            if (GNM == GNM_CPG4) {
              s_hi += __shfl_xor_sync(0xffffffffu, s_hi, o);
              q_hi += __shfl_xor_sync(0xffffffffu, q_hi, o);
            }
          }
          const float m_lo = s_lo * inv_n, m_hi = GNM == GNM_CPG8 ? m_lo : s_hi * inv_n;
          const float r_lo = 1.0f / sqrtf(fmaxf(q_lo * inv_n - m_lo * m_lo, 0.0f) + p.eps);
          const float r_hi = GNM == GNM_CPG8 ? r_lo : 1.0f / sqrtf(fmaxf(q_hi * inv_n - m_hi * m_hi, 0.0f) + p.eps);
          const int c0 = chunk * 64 + kcl * 8;
          const float4 g0 = *reinterpret_cast<const float4*>(s_gamma + c0), g1 = *reinterpret_cast<const float4*>(s_gamma + c0 + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(s_beta + c0), b1 = *reinterpret_cast<const float4*>(s_beta + c0 + 4);
          const float gam[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float bet[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          const float hs = p.silu ? 0.5f : 1.0f;  // SiLU(y) = h + h tanh(h), h = y / 2: the halving is folded into the affine
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            ca[j] = gam[j] * (j < 4 ? r_lo : r_hi) * hs;
            cb[j] = fmaf(-(j < 4 ? m_lo : m_hi), ca[j], bet[j] * hs);
          }
        }
#else
#pragma unroll
        for (int j = 0; j < 8; ++j) { ca[j] = 0.5f; cb[j] = 0.0f; }
#endif
        const int stage = a_it % p.a_stages;
        if (a_it >= p.a_stages) mbar_wait_relaxed(&bar_a_empty[stage], ((a_it / p.a_stages) - 1) & 1);
        if (active) {
          uint4* a4 = reinterpret_cast<uint4*>(As + stage * p.a_stage_bytes) + kcl * p.R;
#pragma unroll
          for (int k = 0; k < RC; ++k) {
            const int px = slice + (k << ps_log);
            if (px < P) {
#ifndef RD_ABL_XF_NOAPPLY
              if (normalise) gn_apply(raw[k], ca, cb);
#endif
              a4[trow[px]] = raw[k];
            }
          }
        }
        if (!PREFETCH && st + 1 < total) load_step(raw, li_n, chunk_n);
        fence_proxy_async_smem();  // st.shared above must be visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_a_full[stage]);
        if (PREFETCH) {
#pragma unroll
          for (int k = 0; k < RC; ++k) raw[k] = nxt[k];
        }
        li = li_n; chunk = chunk_n;
      }
    } else if constexpr (GNM == GNM_NONE && !X3) {
      // ---------------- plain gather (NIN shortcuts, up/down-sampling convs, attention projections): there is nothing
      // to compute, so the pixels go global -> shared with 16-byte cp.async copies straight into their K-major slots,
      // one whole 64-channel chunk ahead of the chunk being handed to the tensor core (no registers, ~64 KB in
      // flight per SM instead of one batch of 8 loads per thread).
      const int total = my_groups * p.nchunks;
      auto issue_chunk = [&](int it) {
        const int li = it / p.nchunks, chunk = it - li * p.nchunks;
        const int g = blockIdx.x + li * gridDim.x;
        const int S_act = max(0, min(p.S, p.B2 - g * p.S));
        const int stage = it % p.a_stages;
        if (it >= p.a_stages) mbar_wait_relaxed(&bar_a_empty[stage], ((it / p.a_stages) - 1) & 1);
        uint4* a4 = reinterpret_cast<uint4*>(As + stage * p.a_stage_bytes);
        const int which = (chunk * 64 < p.C[0]) ? 0 : 1;
        const act_t* base = (which ? src1 + static_cast<size_t>(g) * gstride1 - p.C[0]
                                   : src0 + static_cast<size_t>(g) * gstride0) + chunk * 64;
        const int* toff = t_off + (which ? p.S * P : 0);
#ifdef RD_ABL_NO_XFORM
        const int items = 0;
#else
        const int items = S_act * P * 8;
#endif
        for (int item = xt; item < items; item += XFORM_THREADS) {
          const int sp = item >> 3, kcl = item & 7;
          cp_async16(a4 + kcl * p.R + t_row[sp], base + toff[sp] + kcl * 8);
        }
        cp_async_commit();
      };
      if (total > 0) issue_chunk(0);
      for (int it = 0; it < total; ++it) {
        if (it + 1 < total) { issue_chunk(it + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_a_full[it % p.a_stages]);  // every transform warp reports its own rows
      }
    } else {
      // ---------------- streaming mode: statistics pass, then a normalise pass per chunk (also the only transform of
      // the fp32-class mode, where a pixel's 8 channels are 32 bytes of fp32 and leave as a bf16 hi and a bf16 lo vector)
      const int a_half16 = p.a_stage_bytes / 32;  // uint4 index of the lo image inside a stage (x3)
#ifdef RD_STREAM_NOTABLE  // measurement build: per-item coefficient look-ups as before
      const bool coef_table = false;
#else
      const bool coef_table = GNM != GNM_NONE && 2 * p.S * p.Cin * 4 <= STAT_SCRATCH_BYTES;
#endif
      const float inv_P = 1.0f / static_cast<float>(P);
      auto load8 = [&](const act_t* ptr, uint4& raw, float (&f)[8]) {
        if constexpr (X3) ldg8f(reinterpret_cast<const float*>(ptr), f);
        else raw = __ldg(reinterpret_cast<const uint4*>(ptr));
      };
      for (int li = 0; li < my_groups; ++li) {
        const int g = blockIdx.x + li * gridDim.x;
#ifdef RD_ABL_NO_XFORM
        const int S_act = 0;
#else
        const int S_act = max(0, min(p.S, p.B2 - g * p.S));
#endif
        const act_t* gb0 = src0 + static_cast<size_t>(g) * gstride0;
        const act_t* gb1 = p.nsrc > 1 ? src1 + static_cast<size_t>(g) * gstride1 : gb0;
        if (GNM != GNM_NONE) {
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
              const act_t* base = (which ? gb1 - p.C[0] : gb0) + kc * 8;
              const int* toff = t_off + (which ? p.S * P : 0) + (sb + s) * P;
              float sum[8], sq[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) { sum[j] = 0.0f; sq[j] = 0.0f; }
              constexpr int SB = X3 ? 2 : RD_STREAM_SB;  // (12 in flight was measured slower: the streaming variants spill at the 128-register cap)
              for (int px = slice; px < P; px += SB * PS) {
                uint4 raw[SB];
                float f[X3 ? SB : 1][8];
#pragma unroll
                for (int u = 0; u < SB; ++u)
                  if (px + u * PS < P) load8(base + toff[px + u * PS], raw[u], f[X3 ? u : 0]);
#pragma unroll
                for (int u = 0; u < SB; ++u)
                  if (px + u * PS < P) {
                    float ff[8];
                    if constexpr (X3) {
#pragma unroll
                      for (int j = 0; j < 8; ++j) ff[j] = f[u][j];
                    } else {
                      unpack8(raw[u], ff);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) { sum[j] += ff[j]; sq[j] = fmaf(ff[j], ff[j], sq[j]); }
                  }
              }
              put_record(slice * pairs + pair, sum, sq);
            }
            xform_bar();
            stat_reduce(sb, ns, pairs, PS);
            xform_bar();
          }
          // per-(sample, channel) affine table over the (now idle) statistics scratch: the normalise pass then costs four
          // 16-byte shared loads per item instead of 20 scalar ones and 32 multiplies (same expressions, same values)
          if (coef_table) {
            for (int idx = xt; idx < S_act * p.Cin; idx += XFORM_THREADS) {
              const int s = idx / p.Cin, c = idx - s * p.Cin;
              const float2 mr = *reinterpret_cast<const float2*>(s_stat + 2 * (s * p.groups + s_gidx[c]));
              float ca = s_gamma[c] * mr.y, cb = fmaf(-mr.x, ca, s_beta[c]);
              if (p.silu) { ca *= 0.5f; cb *= 0.5f; }
              s_scr[idx] = ca;
              s_scr[p.S * p.Cin + idx] = cb;
            }
            xform_bar();
          }
        }
        // warm the L2 with the next group's pixels: its statistics pass otherwise waits for HBM, batch after batch
#ifndef RD_STREAM_NOPF  // (measurement build without the prefetch)
        if (li + 1 < my_groups) {
          const int gn = g + gridDim.x;
          const int S_n = max(0, min(p.S, p.B2 - gn * p.S));
          constexpr int EPL = 128 / static_cast<int>(sizeof(act_t));  // elements per 128-byte line
          for (int which = 0; which < p.nsrc; ++which) {
            const act_t* base = (which ? src1 + static_cast<size_t>(gn) * gstride1 : src0 + static_cast<size_t>(gn) * gstride0);
            const int lpp = p.C[which] / EPL;  // lines per pixel
            if (lpp * EPL != p.C[which]) continue;
            const int* toff = t_off + (which ? p.S * P : 0);
            for (int i = xt; i < S_n * P * lpp; i += XFORM_THREADS) {
              const int sp = i / lpp, l = i - sp * lpp;
              asm volatile("prefetch.global.L2 [%0];" ::"l"(base + toff[sp] + l * EPL));
            }
          }
        }
#endif
        const int items = S_act * P * 8;
        for (int chunk = 0; chunk < p.nchunks; ++chunk, ++a_it) {
          const int stage = a_it % p.a_stages;
          if (a_it >= p.a_stages) mbar_wait_relaxed(&bar_a_empty[stage], ((a_it / p.a_stages) - 1) & 1);
          uint4* a4 = reinterpret_cast<uint4*>(As + stage * p.a_stage_bytes);
          const int which = (chunk * 64 < p.C[0]) ? 0 : 1;  // C[0] is a multiple of 64 whenever there are two sources
          const act_t* base = (which ? gb1 - p.C[0] : gb0) + chunk * 64;
          const int* toff = t_off + (which ? p.S * P : 0);
          constexpr int NB = X3 ? 4 : 8;
          for (int b0 = xt; b0 < items; b0 += NB * XFORM_THREADS) {
            uint4 raw[NB];
            float f[X3 ? NB : 1][8];
#pragma unroll
            for (int u = 0; u < NB; ++u) {
              const int item = b0 + u * XFORM_THREADS;
              if (item < items) load8(base + toff[item >> 3] + (item & 7) * 8, raw[u], f[X3 ? u : 0]);
            }
#pragma unroll
            for (int u = 0; u < NB; ++u) {
              const int item = b0 + u * XFORM_THREADS;
              if (item < items) {
                const int sp = item >> 3, kcl = item & 7;
                float ca[8], cb[8];
                if (GNM != GNM_NONE) {
                  const int smp = __float2int_rz((static_cast<float>(sp) + 0.5f) * inv_P);  // sp / P, exact for these sizes
                  if (coef_table) {
                    const float* ta = s_scr + smp * p.Cin + chunk * 64 + kcl * 8;
                    const float* tb = ta + p.S * p.Cin;
                    const float4 a0 = *reinterpret_cast<const float4*>(ta), a1 = *reinterpret_cast<const float4*>(ta + 4);
                    const float4 b0 = *reinterpret_cast<const float4*>(tb), b1 = *reinterpret_cast<const float4*>(tb + 4);
                    ca[0] = a0.x; ca[1] = a0.y; ca[2] = a0.z; ca[3] = a0.w; ca[4] = a1.x; ca[5] = a1.y; ca[6] = a1.z; ca[7] = a1.w;
                    cb[0] = b0.x; cb[1] = b0.y; cb[2] = b0.z; cb[3] = b0.w; cb[4] = b1.x; cb[5] = b1.y; cb[6] = b1.z; cb[7] = b1.w;
                  } else {
                    gn_coeffs(smp, chunk * 64 + kcl * 8, ca, cb);
                  }
                }
                if constexpr (X3) {
                  if (GNM != GNM_NONE) gn_apply_f(f[u], ca, cb);
                  uint4 hi, lo;
                  split8(f[u], hi, lo);
                  a4[kcl * p.R + t_row[sp]] = hi;
                  a4[a_half16 + kcl * p.R + t_row[sp]] = lo;
                } else {
                  if (GNM != GNM_NONE) gn_apply(raw[u], ca, cb);
                  a4[kcl * p.R + t_row[sp]] = raw[u];
                }
              }
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_a_full[stage]);
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

constexpr int MAX_DEVICES = 64;

static int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) dev = 0;
  return dev;
}

// SM count of the CURRENT device (one process may drive several GPUs; launch state is kept per device ordinal)
static int conv_num_sms() {
  static int sms[MAX_DEVICES] = {};
  const int dev = current_device();
  if (sms[dev] == 0) {
    if (cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms[dev] <= 0) sms[dev] = kNumSMs;
  }
  return sms[dev];
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// Fills the launch geometry for an op; returns RD_OK or an error.
int conv_make_params(const rd_op_conv& op, ConvParams& p, int& smem_bytes, int& grid) {
  RD_REQUIRE(op.nsrc == 1 || op.nsrc == 2, "conv: nsrc must be 1 or 2");
  RD_REQUIRE(op.ntaps == 9 || op.ntaps == 1, "conv: ntaps must be 9 or 1");
  RD_REQUIRE(op.H_in >= 1 && op.W_in >= 1 && op.H_in <= MAX_HW && op.W_in <= MAX_HW, "conv: H,W must be in [1,%d]", MAX_HW);
  RD_REQUIRE(op.C_out % 32 == 0 && op.C_out >= 32 && op.C_out <= 256, "conv: C_out %d unsupported per launch (slice it, see out_stride)", op.C_out);
  RD_REQUIRE(op.out_stride == 0 || (op.out_stride >= op.C_out && op.out_stride % 8 == 0), "conv: out_stride %d < C_out %d", op.out_stride, op.C_out);
  RD_REQUIRE(op.stride == 1 || (op.stride == 2 && op.pad == 0 && op.ntaps == 9), "conv: stride 2 needs pad 0, 3x3");
  RD_REQUIRE(op.w && op.bias && op.out && op.B2 > 0, "conv: null pointer / empty batch");
  RD_REQUIRE(op.precision == RD_PREC_BF16 || op.precision == RD_PREC_F32X3, "conv: unknown precision %d", op.precision);
  memset(&p, 0, sizeof(p));
  p.x3 = op.precision == RD_PREC_F32X3;
  int cin = 0;
  for (int i = 0; i < op.nsrc; ++i) {
    RD_REQUIRE(op.src[i].ptr && op.src[i].C % 8 == 0 && op.src[i].C > 0, "conv: source %d channels must be a multiple of 8", i);
    RD_REQUIRE(op.src[i].Hs >= 1 && op.src[i].Ws >= 1 && op.src[i].Hs <= MAX_HW && op.src[i].Ws <= MAX_HW, "conv: bad source size");
    p.src[i] = op.src[i].ptr;
    p.C[i] = op.src[i].C; p.Hs[i] = op.src[i].Hs; p.Ws[i] = op.src[i].Ws;
    nearest_map(p.ymap[i], op.H_in, op.src[i].Hs);
    nearest_map(p.xmap[i], op.W_in, op.src[i].Ws);
    cin += op.src[i].C;
  }
  if (op.nsrc == 1) { p.C[1] = 0; }
  else RD_REQUIRE(op.src[0].C % 64 == 0, "conv: with two sources the first must have a multiple of 64 channels");
  RD_REQUIRE(cin % 64 == 0, "conv: total input channels (%d) must be a multiple of 64", cin);
  p.nsrc = op.nsrc;
  // fused 1x1 shortcut (see ConvParams): raw sources gathered to the same H_in x W_in grid
  RD_REQUIRE(op.sc_nsrc >= 0 && op.sc_nsrc <= 2, "conv: sc_nsrc must be 0, 1 or 2");
  int sc_cin = 0;
  for (int i = 0; i < op.sc_nsrc; ++i) {
    RD_REQUIRE(op.sc_src[i].ptr && op.sc_src[i].C % 8 == 0 && op.sc_src[i].C > 0, "conv: shortcut source %d channels must be a multiple of 8", i);
    RD_REQUIRE(op.sc_src[i].Hs >= 1 && op.sc_src[i].Ws >= 1 && op.sc_src[i].Hs <= MAX_HW && op.sc_src[i].Ws <= MAX_HW, "conv: bad shortcut source size");
    p.sc_src[i] = op.sc_src[i].ptr;
    p.sc_C[i] = op.sc_src[i].C; p.sc_Hs[i] = op.sc_src[i].Hs; p.sc_Ws[i] = op.sc_src[i].Ws;
    nearest_map(p.sc_ymap[i], op.H_in, op.sc_src[i].Hs);
    nearest_map(p.sc_xmap[i], op.W_in, op.sc_src[i].Ws);
    sc_cin += op.sc_src[i].C;
  }
  if (op.sc_nsrc > 0) {
    RD_REQUIRE(sc_cin % 64 == 0 && (op.sc_nsrc == 1 || op.sc_src[0].C % 64 == 0), "conv: shortcut channels must come in multiples of 64");
    RD_REQUIRE(op.stride == 1 && !p.x3 && !op.residual, "conv: a fused shortcut needs stride 1, the bf16 plan and no residual");
  }
  p.sc_nsrc = op.sc_nsrc;
  p.sc_chunks = sc_cin / 64;
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
  p.out_stride = op.out_stride ? op.out_stride : op.C_out;
  p.groups = op.gn_groups; p.silu = op.gn_silu; p.eps = op.gn_eps;
  p.gnm = GNM_NONE;
  if (p.groups > 0) {
    RD_REQUIRE(cin % p.groups == 0 && op.gn_gamma && op.gn_beta, "conv: bad GroupNorm arguments");
    p.cpg = cin / p.groups;  // groups may straddle the concat boundary (192 ch / 32 groups): stats are per channel
    p.gnm = (p.cpg % 8 == 0) ? GNM_CPG8 : (p.cpg == 4 ? GNM_CPG4 : GNM_GENERAL);
  }
  p.gamma = op.gn_gamma; p.beta = op.gn_beta;
  p.w = static_cast<const __nv_bfloat16*>(op.w);
  p.bias = op.bias; p.tproj = op.tproj; p.tproj_stride = op.tproj_stride; p.tproj_off = op.tproj_off; p.tproj_wrap = op.tproj_wrap;
  p.residual = op.residual;
  p.out_scale = op.out_scale; p.out = op.out; p.B2 = op.B2;
  const int planes = p.x3 ? 2 : 1;  // hi + lo images of every operand in the fp32-class mode
  p.w_slab_bytes = p.N * 128 * planes;
  p.n_slabs = p.nchunks * p.ntaps + p.sc_chunks;

  // Tile geometry.  Per candidate tile count: samples per group, row efficiency, whether the TMEM
  // accumulators can be double-buffered (2*nt*N <= 512 columns) and whether the whole filter stays
  // resident in shared memory.  Score = useful rows per MMA row, discounted when the epilogue cannot
  // overlap the next group's MMAs.
  const int smem_cap = 227 * 1024 - 2048;  // static __shared__ barriers + alignment slack
  const int valid_px = op.H_out * op.W_out;
  double best_score = -1.0;
  ConvParams best = p;
  // measurement switches (read once; defaults are the shipped configuration)
  static const int wmax = [] { int w = env_int("RD_CONV_WSTAGES", 4); return w < 2 ? 2 : (w > MAX_W_STAGES ? MAX_W_STAGES : w); }();
  static const int astream = env_int("RD_CONV_ASTAGES_STREAM", 3) <= 2 ? 2 : 3;
  // RD_CONV_FORCE_NT="cin,n,h,nt;..." pins the tile count of the launches with that (C_in, C_out, H_in) -- geometry experiments
  int force_nt = 0;
  {
    static const char* spec = getenv("RD_CONV_FORCE_NT");
    for (const char* q = spec; q && *q;) {
      int a = 0, b = 0, c = 0, d = 0;
      if (sscanf(q, "%d,%d,%d,%d", &a, &b, &c, &d) == 4 && a == cin && b == p.N && c == op.H_in) force_nt = d;
      q = strchr(q, ';');
      if (q) ++q;
    }
  }
  // stride 2 (bf16 plan): polyphase layout first, the one-plane layout if no polyphase geometry fits (RD_CONV_POLY=0: A/B runs)
  static const int poly_on = env_int("RD_CONV_POLY", 1);
  for (int poly = (op.stride == 2 && !p.x3 && poly_on) ? 1 : 0; poly >= 0 && best_score < 0; --poly) {
  if (poly) { p.poly = 1; p.Wp = op.W_out + 1; p.rps = (op.H_out + 1) * p.Wp; p.pad = 0; }
  else if (p.poly) { p.poly = 0; p.Wp = op.W_in + 1; p.rps = (op.H_in + 1) * p.Wp; p.pad = op.pad ? 1 : 0; }
  const int max_shift = p.poly ? p.Wp + 1 : ((op.ntaps == 9) ? 2 * p.Wp + 2 : 0);
  for (int nt = 1; nt <= 4; ++nt) {
    if (nt * p.N > 512 || nt * 128 < p.rps) continue;
    if (force_nt && nt != force_nt) continue;
    ConvParams c = p;
    c.n_tiles = nt;
    c.plane_rows = nt * 128 + max_shift;
    c.R = ((p.poly ? 4 : 1) * c.plane_rows) | 1;
    c.S = (nt * 128) / p.rps;
    if (op.samples_per_cta > 0 && op.samples_per_cta < c.S) c.S = op.samples_per_cta;
    // the per-group (bias + temb) table is S x N floats, double-buffered: keep it within 32 KB (only reached at 1x1 / 2x2
    // images with an un-padded 1x1 filter, where a 128-row tile could hold 32-128 samples)
    if (c.S > 4096 / p.N) c.S = 4096 / p.N > 0 ? 4096 / p.N : 1;
    c.acc_bufs = (2 * nt * p.N <= 512) ? 2 : 1;
    c.a_stage_bytes = (8 * c.R * 16 + 127) / 128 * 128 * planes;
    c.a_stages = (p.nchunks + p.sc_chunks == 1) ? 2 : 3;
    c.w_resident = 1;
    // launches with a fused shortcut have one short extra chunk per 64 shortcut channels: a resident filter with two
    // operand stages beats a streamed filter with three (RD_CONV_SC_RES2=0 restores the generic rule)
    static const int sc_res2 = env_int("RD_CONV_SC_RES2", 1);
    if (sc_res2 && p.sc_chunks > 0 && c.a_stages == 3 && conv_smem_layout(c).total > smem_cap) {
      c.a_stages = 2;
      if (conv_smem_layout(c).total > smem_cap) c.a_stages = 3;
    }
    if (conv_smem_layout(c).total > smem_cap) {
      // Streamed filter: the ring has to cover the L2 latency of a slab at the rate the tensor core consumes them
      // (8-16 KB per ~600 cycles), i.e. tens of KB in flight -- it gets whatever shared memory two operand stages
      // leave, up to `wmax` slabs.
      c.w_resident = 0;
      if (c.a_stages > astream) c.a_stages = astream;
      if (p.poly && c.a_stages > 2) c.a_stages = 2;  // a polyphase stage is four planes: two stages and a full filter ring
      c.w_stages = wmax < p.n_slabs ? wmax : p.n_slabs;
      if (c.w_stages < 2) c.w_stages = 2;
      while (c.w_stages > 2 && conv_smem_layout(c).total > smem_cap) --c.w_stages;
      if (conv_smem_layout(c).total > smem_cap && c.a_stages == 3) c.a_stages = 2;
      if (conv_smem_layout(c).total > smem_cap) c.a_stages = 1;  // last resort (fp32-class plan at 16x16): transform and MMAs alternate
      if (conv_smem_layout(c).total > smem_cap) continue;
      // the plain gather stages chunk i+1 before it reports chunk i: it needs two stages
      if (c.a_stages == 1 && !p.x3 && p.groups == 0) continue;
    }
    // transform mode: team mode (bf16 plan, GroupNorm groups of 4 / 8 channels) when the S x 8 (sample, item) teams of a
    // chunk fit the 320 transform threads with at most 16 pixels per lane; otherwise the two-pass streaming transform
    c.xmode = 0; c.rc_PS = 1;
    if (!p.x3 && p.groups > 0 && (p.cpg == 4 || p.cpg == 8) && c.S * 8 <= XFORM_THREADS) {
      int ps = 1;
      while (ps < 32 && c.S * 8 * (ps * 2) <= XFORM_THREADS) ps *= 2;
      const int slots = (p.H * p.W + ps - 1) / ps;
      if (slots <= 16) { c.xmode = slots <= 4 ? 4 : (slots <= 9 ? 9 : 16); c.rc_PS = ps; }
    }
    if (p.sc_chunks > 0 && c.xmode == 0) continue;  // the fused shortcut is staged by the team-mode transform only
    double score = static_cast<double>(c.S * valid_px) / (nt * 128);
    if (c.acc_bufs == 1) score *= 0.75;
    if (c.a_stages == 1) score *= 0.6;  // staging and MMAs alternate
    if (!c.w_resident) score *= (nt >= 2 ? 0.97 : 0.85);  // streamed weights are re-read per group: favour larger groups
    if (score > best_score) { best_score = score; best = c; }
  }
  }
  RD_REQUIRE(best_score > 0, "conv: no tile geometry fits (Cin=%d N=%d rps=%d x3=%d)", cin, p.N, p.rps, p.x3);
  p = best;
  p.n_groups = (op.B2 + p.S - 1) / p.S;
  p.tmem_cols = next_pow2_cols(p.acc_bufs * p.n_tiles * p.N);
  smem_bytes = conv_smem_layout(p).total;
  const int sms = conv_num_sms();
  grid = p.n_groups < sms ? p.n_groups : sms;
  return RD_OK;
}

typedef void (*conv_kernel_t)(const ConvParams);

static conv_kernel_t conv_pick(int gnm, int rc, int x3) {
#define RD_K(G, R) conv_gemm_kernel<G, R, false>
#define RD_K3(G) conv_gemm_kernel<G, 0, true>
  if (x3) {
    switch (gnm) {
      case GNM_NONE: return RD_K3(GNM_NONE);
      case GNM_CPG8: return RD_K3(GNM_CPG8);
      case GNM_CPG4: return RD_K3(GNM_CPG4);
      case GNM_GENERAL: return RD_K3(GNM_GENERAL);
      default: return nullptr;
    }
  }
  switch (gnm * 100 + rc) {
    case GNM_NONE * 100 + 0: return RD_K(GNM_NONE, 0);
    case GNM_CPG8 * 100 + 0: return RD_K(GNM_CPG8, 0);
    case GNM_CPG8 * 100 + 4: return RD_K(GNM_CPG8, 4);
    case GNM_CPG8 * 100 + 9: return RD_K(GNM_CPG8, 9);
    case GNM_CPG8 * 100 + 16: return RD_K(GNM_CPG8, 16);
    case GNM_CPG4 * 100 + 0: return RD_K(GNM_CPG4, 0);
    case GNM_CPG4 * 100 + 4: return RD_K(GNM_CPG4, 4);
    case GNM_CPG4 * 100 + 9: return RD_K(GNM_CPG4, 9);
    case GNM_CPG4 * 100 + 16: return RD_K(GNM_CPG4, 16);
    case GNM_GENERAL * 100 + 0: return RD_K(GNM_GENERAL, 0);
    default: return nullptr;
  }
#undef RD_K
#undef RD_K3
}

int conv_launch(const rd_op_conv& op, cudaStream_t st) {
  ConvParams p;
  int smem = 0, grid = 0;
  int rc = conv_make_params(op, p, smem, grid);
  if (rc != RD_OK) return rc;
  conv_kernel_t k = conv_pick(p.gnm, p.xmode, p.x3);
  if (!k) return fail(RD_E_STATE, "conv: no kernel for gnm=%d rc=%d x3=%d", p.gnm, p.xmode, p.x3);
  // cudaFuncSetAttribute is per (function, device): a process that drives several GPUs opts in on each of them
  static bool configured[MAX_DEVICES][2][4][17] = {};
  bool& done = configured[current_device()][p.x3][p.gnm][p.xmode];
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "conv: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    done = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(CONV_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = 0;
  static const int pdl = env_int("RD_CONV_PDL", 1);  // RD_CONV_PDL=0 restores fully serialised launches (A/B measurements)
  if (pdl) {
    attr[cfg.numAttrs].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[cfg.numAttrs].val.programmaticStreamSerializationAllowed = 1;
    ++cfg.numAttrs;
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, k, p);
  if (e != cudaSuccess) return fail(static_cast<int>(e), "conv_gemm_kernel launch: %s", cudaGetErrorString(e));
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

// planner / profiling feedback: the tile geometry the library picks for `op`
//   geom = {S, n_tiles, R, n_groups, a_stages, w_resident, w_stages, acc_bufs, xmode, smem_bytes, grid, tmem_cols}
extern "C" int rd_conv_geometry(const rd_op_conv* op, int* geom) {
  RD_REQUIRE(op && geom, "rd_conv_geometry: null argument");
  rd::ConvParams p;
  int smem = 0, g = 0;
  int rc = rd::conv_make_params(*op, p, smem, g);
  if (rc != RD_OK) return rc;
  const int v[12] = {p.S, p.n_tiles, p.R, p.n_groups, p.a_stages, p.w_resident, p.w_stages, p.acc_bufs, p.xmode, smem, g, p.tmem_cols};
  for (int i = 0; i < 12; ++i) geom[i] = v[i];
  return RD_OK;
}
