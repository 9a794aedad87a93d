// small_ops.cu -- the thin ends of the NCSN++ forward that are not GEMM-shaped enough for tcgen05:
//   temb_kernel     per-sample  Dense_0(SiLU(temb))  for all ResBlocks at once
//                   (reference ncsnpp.py:252-262, layerspp.py:201-202)
//   in_conv_kernel  input_conv: 3x3, channels(=1) -> nf, reads the fp32 sampler state, CFG-duplicated
//                   (ncsnpp.py:266, models/utils.py:121)
//   out_head_kernel out_norm + SiLU + out_conv (nf -> channels) + classifier-free-guidance combine
//                   (ncsnpp.py:343-347, models/utils.py:124-138), fp32 score out
#include "rd_common.h"
#include <cuda_bf16.h>
#include "rd_ptx.cuh"
#include <type_traits>
#include <cstdlib>
#include <cstdint>

namespace rd {

__device__ __forceinline__ float silu_acc(float v) { return v / (1.0f + expf(-v)); }

// ------------------------------------------------------------------------------------------------
// out[b][o] = dense_b[o] + sum_k dense_w[o][k] * SiLU(time_table[row(b)][k] + sum_c label_w[k][c]*labels[b][c])
// One GEMM for the Dense_0 of every ResBlock ([rows x K] x [K x 2176]).  Tensor cores with fp32-class accuracy: both
// operands are split into bf16 hi + bf16 lo (x = hi + lo to 2^-17 relative) and the product is accumulated in fp32
// as hi*hi + hi*lo + lo*hi with mma.sync.m16n8k16 -- the dropped lo*lo term is 2^-18 relative, so the temb path
// keeps the reference's fp32 semantics while running 5x faster than the fp32 SIMT tile it replaces.
constexpr int TE_BM = 128, TE_BN = 128, TE_BK = 16;
constexpr int TE_LD = TE_BK + 8;  // bf16 row stride (48 B): conflict-free 32-bit fragment loads

__device__ __forceinline__ void te_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// x -> (hi, lo) bf16 with x ~= hi + lo
__device__ __forceinline__ void te_split(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

__global__ void __launch_bounds__(256, 2) temb_kernel(const float* __restrict__ time_table, const float* __restrict__ label_w,
                                                   const float* __restrict__ labels, const float* __restrict__ dense_w,
                                                   const float* __restrict__ dense_b, float* __restrict__ out,
                                                   const int32_t* __restrict__ step_ctr,
                                                   const int32_t* __restrict__ row_idx, int B2, int K, int NC, int NO) {
  __shared__ __align__(16) __nv_bfloat16 As[2][TE_BM][TE_LD];  // [hi|lo][row][k]
  __shared__ __align__(16) __nv_bfloat16 Bs[2][TE_BN][TE_LD];  // [hi|lo][out][k]
  const int step = step_ctr ? *step_ctr : 0;
  const int m0 = blockIdx.x * TE_BM, n0 = blockIdx.y * TE_BN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q4 = lane & 3;
  const int wm = warp & 3, wn = warp >> 2;  // warp tile: rows wm*32..+31, columns wn*64..+63
  float acc[2][8][4] = {};
  // every thread stages 8 A and 8 B elements per k-tile (element = (row, k)); the next tile's global reads are
  // issued before the current tile's MMAs.  A thread's eight elements share one k (= k0 + tid % 16) and sit in eight
  // fixed rows, so everything per-row (time-table row, label) is looked up once and everything per-k once per tile.
  float pa[8], pb[8];
  const int kk = tid & (TE_BK - 1);
  int a_row[8];    // time-table row offset (elements) or -1: row outside the batch
  float a_lab[8];  // the sample's label (one class) -- more classes take the general loop below
  int b_off[8];    // filter row offset (elements) or -1
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int mm = (tid + u * 256) >> 4;
    const int b = m0 + mm, o = n0 + mm;
    a_row[u] = b < B2 ? (row_idx ? row_idx[b] : step) * K : -1;
    a_lab[u] = (b < B2 && NC == 1) ? labels[b] : 0.0f;
    b_off[u] = o < NO ? o * K : -1;
  }
  auto fast_silu = [](float v) { return __fdividef(v, 1.0f + __expf(-v)); };  // ~3e-7 relative: below the 2^-17 operand split
  auto fetch = [&](int k0) {
    const int k = k0 + kk;
    const bool k_ok = k < K;
    const float lw = (k_ok && NC == 1) ? label_w[k] : 0.0f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      float t = 0.0f;
      if (a_row[u] >= 0 && k_ok) {
        t = time_table[a_row[u] + k];
        if (NC == 1) {
          t = fmaf(lw, a_lab[u], t);
        } else {
          const int b = m0 + ((tid + u * 256) >> 4);
          for (int c = 0; c < NC; ++c) t = fmaf(label_w[k * NC + c], labels[static_cast<size_t>(b) * NC + c], t);
        }
        t = fast_silu(t);
      }
      pa[u] = t;
      pb[u] = (b_off[u] >= 0 && k_ok) ? dense_w[b_off[u] + k] : 0.0f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += TE_BK) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = tid + u * 256;
      const int kk = i & (TE_BK - 1), mm = i >> 4;
      te_split(pa[u], As[0][mm][kk], As[1][mm][kk]);
      te_split(pb[u], Bs[0][mm][kk], Bs[1][mm][kk]);
    }
    __syncthreads();
    if (k0 + TE_BK < K) fetch(k0 + TE_BK);
#pragma unroll
    for (int ks = 0; ks < TE_BK; ks += 16) {
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int r = wm * 32 + mt * 16 + g;
#pragma unroll
        for (int hl = 0; hl < 2; ++hl) {
          uint32_t (&dst)[4] = hl ? al[mt] : ah[mt];
          dst[0] = *reinterpret_cast<const uint32_t*>(&As[hl][r][ks + 2 * q4]);
          dst[1] = *reinterpret_cast<const uint32_t*>(&As[hl][r + 8][ks + 2 * q4]);
          dst[2] = *reinterpret_cast<const uint32_t*>(&As[hl][r][ks + 8 + 2 * q4]);
          dst[3] = *reinterpret_cast<const uint32_t*>(&As[hl][r + 8][ks + 8 + 2 * q4]);
        }
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int n = wn * 64 + nt * 8 + g;
        const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(&Bs[0][n][ks + 2 * q4]);
        const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(&Bs[0][n][ks + 8 + 2 * q4]);
        const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(&Bs[1][n][ks + 2 * q4]);
        const uint32_t bl1 = *reinterpret_cast<const uint32_t*>(&Bs[1][n][ks + 8 + 2 * q4]);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          te_mma(acc[mt][nt], al[mt], bh0, bh1);  // small terms first
          te_mma(acc[mt][nt], ah[mt], bl0, bl1);
          te_mma(acc[mt][nt], ah[mt], bh0, bh1);
        }
      }
    }
    __syncthreads();
  }
  // accumulator fragment: c0,c1 -> (row g, cols 2q,2q+1); c2,c3 -> (row g+8, same cols)
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int b = m0 + wm * 32 + mt * 16 + g + hr * 8;
      if (b >= B2) continue;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int o = n0 + wn * 64 + nt * 8 + 2 * q4;
        if (o + 1 < NO) {
          const float2 bb = *reinterpret_cast<const float2*>(dense_b + o);
          *reinterpret_cast<float2*>(out + static_cast<size_t>(b) * NO + o) =
              make_float2(acc[mt][nt][2 * hr] + bb.x, acc[mt][nt][2 * hr + 1] + bb.y);
        } else if (o < NO) {
          out[static_cast<size_t>(b) * NO + o] = acc[mt][nt][2 * hr] + dense_b[o];
        }
      }
    }
}

int temb_launch(const rd_op_temb& op, cudaStream_t st) {
  RD_REQUIRE(op.time_table && op.dense_w && op.dense_b && op.out && op.B2 > 0, "temb: null pointer / empty batch");
  RD_REQUIRE(op.num_classes == 0 || (op.label_w && op.labels), "temb: labels missing");
  RD_REQUIRE(op.n_out_total % 4 == 0, "temb: n_out_total must be a multiple of 4");
  dim3 grid((op.B2 + TE_BM - 1) / TE_BM, (op.n_out_total + TE_BN - 1) / TE_BN);
  temb_kernel<<<grid, 256, 0, st>>>(op.time_table, op.label_w, op.labels, static_cast<const float*>(op.dense_w), op.dense_b,
                                    op.out, op.step_ctr, op.row_idx, op.B2, op.temb_dim, op.num_classes, op.n_out_total);
  return check_launch("temb_kernel");
}

// ------------------------------------------------------------------------------------------------
// input_conv: one thread per (sample, pixel) computes all C_out channels (weights broadcast from shared memory),
// writes its 2*C_out bytes with 256-bit stores; NHWC bf16 out.  C_out <= 64 per pass (template-free: register tile of 64).
template <typename TO>
__global__ void __launch_bounds__(256) in_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, TO* __restrict__ out, int B,
                                                      int B2, int Cin, int Cout, int H, int W) {
  extern __shared__ __align__(16) float sw[];  // [Cin*9][Cout] (tap-major) + bias[Cout]
  // the conv launch that follows may start its prologue as SMs free up (it waits for this grid before reading `out`)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int wn = Cout * Cin * 9;
  for (int i = threadIdx.x; i < wn; i += blockDim.x) {
    const int co = i / (Cin * 9), t = i - co * (Cin * 9);
    sw[t * Cout + co] = w[i];
  }
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[wn + i] = bias[i];
  __syncthreads();
  const int P = H * W;
  // B2 is B or a multiple of it (x.repeat(2,1,1,1) under classifier-free guidance): every distinct pixel is computed
  // once and stored B2/B times
  const int total = B * P, copies = B2 / B;
  const size_t copy_stride = static_cast<size_t>(B) * P * Cout;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += gridDim.x * blockDim.x) {
    const int b2 = pix / P, px = pix - b2 * P;
    const int yh = px / W, xw = px - yh * W;
    const float* xb = x + static_cast<size_t>(b2) * Cin * P;
    for (int c0 = 0; c0 < Cout; c0 += 64) {
      float acc[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) acc[j] = sw[wn + c0 + j];
      for (int ci = 0; ci < Cin; ++ci)
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const int yy = yh + dy - 1;
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const int xx = xw + dx - 1;
            const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
            const float v = in ? __ldg(xb + ci * P + yy * W + xx) : 0.0f;
            const float4* wp = reinterpret_cast<const float4*>(sw + (ci * 9 + dy * 3 + dx) * Cout + c0);
#pragma unroll
            for (int j4 = 0; j4 < 16; ++j4) {
              const float4 wv = wp[j4];
              acc[4 * j4] = fmaf(v, wv.x, acc[4 * j4]);
              acc[4 * j4 + 1] = fmaf(v, wv.y, acc[4 * j4 + 1]);
              acc[4 * j4 + 2] = fmaf(v, wv.z, acc[4 * j4 + 2]);
              acc[4 * j4 + 3] = fmaf(v, wv.w, acc[4 * j4 + 3]);
            }
          }
        }
      TO* dst = out + static_cast<size_t>(pix) * Cout + c0;
      if constexpr (std::is_same<TO, float>::value) {
#pragma unroll
        for (int j8 = 0; j8 < 8; ++j8) {
          u32x8 pk;
#pragma unroll
          for (int j = 0; j < 8; ++j) pk.v[j] = __float_as_uint(acc[8 * j8 + j]);
          for (int cp = 0; cp < copies; ++cp) st_global_256(dst + cp * copy_stride + 8 * j8, pk);
        }
      } else {
#pragma unroll
        for (int j16 = 0; j16 < 4; ++j16) {   // whole 32-byte sectors per store
          u32x8 pk;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            __nv_bfloat162 h = __floats2bfloat162_rn(acc[16 * j16 + 2 * j], acc[16 * j16 + 2 * j + 1]);
            pk.v[j] = *reinterpret_cast<uint32_t*>(&h);
          }
          for (int cp = 0; cp < copies; ++cp) st_global_256(dst + cp * copy_stride + 16 * j16, pk);
        }
      }
    }
  }
}

int inconv_launch(const rd_op_inconv& op, cudaStream_t st) {
  RD_REQUIRE(op.x && op.w && op.bias && op.out && op.B2 > 0 && op.B > 0, "in_conv: null pointer / empty batch");
  RD_REQUIRE(op.C_out % 64 == 0, "in_conv: C_out must be a multiple of 64");
  const int smem = (op.C_out * op.C_in * 9 + op.C_out) * 4;
  RD_REQUIRE(smem <= 48 * 1024, "in_conv: weights do not fit shared memory");
  RD_REQUIRE(op.B2 % op.B == 0, "in_conv: B2 must be a multiple of B");
  const size_t total = static_cast<size_t>(op.B) * op.H * op.W;
  RD_REQUIRE(static_cast<size_t>(op.B2) * op.H * op.W < (1u << 31), "in_conv: batch too large");
  size_t blocks = (total + 255) / 256;
  if (blocks > static_cast<size_t>(kNumSMs) * 16) blocks = static_cast<size_t>(kNumSMs) * 16;
  if (op.precision == RD_PREC_F32X3)
    in_conv_kernel<float><<<static_cast<unsigned>(blocks), 256, smem, st>>>(op.x, op.w, op.bias, static_cast<float*>(op.out),
                                                                            op.B, op.B2, op.C_in, op.C_out, op.H, op.W);
  else
    in_conv_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), 256, smem, st>>>(op.x, op.w, op.bias, static_cast<__nv_bfloat16*>(op.out),
                                                                                    op.B, op.B2, op.C_in, op.C_out, op.H, op.W);
  return check_launch("in_conv_kernel");
}

// ------------------------------------------------------------------------------------------------
// out head: out_norm + SiLU + 3x3 out_conv (C -> C_img) + CFG combine, one CTA per guided sample.
// A warp owns whole pixels and a lane two channels of every 64-channel slab: pass 1 accumulates the GroupNorm sums,
// pass 2 re-reads the same bf16 pairs (L1/L2 hits), normalises, applies SiLU and forms
// the nine per-tap dot products q[px][tap] = sum_c w[tap][c] * act[px][c] (a vector warp reduction), and the conv
// output is the 9-term gather y[px] = bias + sum_tap q[px + tap][tap].  Nothing is staged as fp32 in shared memory.  With CFG, warps 0-3 take the conditional copy and warps 4-7 the
// unconditional one; the combine is the reference's (1+w)*cond - w*uncond in fp32.
constexpr int OH_WARPS = 4;        // per copy

// Sum each of the 8 values of v[] over the 32 lanes with 9 shuffles (recursive halving): on return lane l holds the
// total of value ((l >> 2) & 7) in v[0].
__device__ __forceinline__ void warp_reduce8(float (&v)[8], int lane) {
#pragma unroll
  for (int step = 0; step < 3; ++step) {
    const int half = 4 >> step;           // values kept after this step
    const int mask = 16 >> step;          // partner distance
    const bool upper = (lane & mask) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i < half) {
        const float send = upper ? v[i] : v[i + half];
        const float keep = upper ? v[i + half] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
      }
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// two adjacent channels of one pixel as fp32
template <typename TI>
__device__ __forceinline__ float2 oh_load2(const TI* p) {
  if constexpr (std::is_same<TI, float>::value) {
    return __ldg(reinterpret_cast<const float2*>(p));
  } else {
    const uint32_t xv = __ldg(reinterpret_cast<const uint32_t*>(p));
    return make_float2(__uint_as_float(xv << 16), __uint_as_float(xv & 0xffff0000u));
  }
}

template <typename TI>
__global__ void __launch_bounds__(2 * OH_WARPS * 32, 4) out_head_kernel(const TI* __restrict__ h, const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta, const float* __restrict__ w,
                                                                     const float* __restrict__ bias, const float* __restrict__ cfg_w,
                                                                     float cfg_w_scalar, float* __restrict__ score, int B, int C,
                                                                     int Cimg, int H, int W, int groups, int cfg, float eps,
                                                                     const float* __restrict__ sigma_table,
                                                                     const int32_t* __restrict__ step_ctr) {
  extern __shared__ __align__(16) float sm[];
  const int P = H * W;
  const int npass = cfg ? 2 : 1;
  float* q_all = sm;                                   // [2][Cimg][P][9]  per-tap dot products
  float* csum = q_all + 2 * Cimg * P * 9;              // [2][OH_WARPS][C][2] per-warp channel sums
  float* gstat = csum + 2 * OH_WARPS * C * 2;          // [2][groups][2] mean, rstd
  const int warp_all = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pass = warp_all / OH_WARPS, wq = warp_all % OH_WARPS;
  const int b = blockIdx.x;
  const int cpg = C / groups;
  const int nslab = C / 64;
  const int ppw = (P + OH_WARPS - 1) / OH_WARPS;       // pixels per warp
  const int px0 = wq * ppw;
  const TI* hb = h + static_cast<size_t>(b + pass * B) * P * C;
  float* qh = q_all + pass * Cimg * P * 9;

  for (int slab = 0; slab < nslab; ++slab) {
    const int c = slab * 64 + 2 * lane;
    // ---- pass 1: this lane's two channels of this warp's pixels -> registers, channel sums
    float s0 = 0.0f, s1 = 0.0f, q0 = 0.0f, q1 = 0.0f;
#pragma unroll 6
    for (int i = 0; i < ppw; ++i) {
      const int px = px0 + i;
      if (px < P) {
        const float2 xv = oh_load2(hb + static_cast<size_t>(px) * C + c);
        const float x0 = xv.x, x1 = xv.y;
        s0 += x0; q0 = fmaf(x0, x0, q0);
        s1 += x1; q1 = fmaf(x1, x1, q1);
      }
    }
    float* cs = csum + ((pass * OH_WARPS + wq) * C + c) * 2;
    cs[0] = s0; cs[1] = q0; cs[2] = s1; cs[3] = q1;
    __syncthreads();
    // group statistics of this slab's groups (fixed summation order: warps, then channels)
    for (int g = threadIdx.x; g < 2 * (64 / cpg); g += blockDim.x) {
      const int ps = g / (64 / cpg), gl = g % (64 / cpg);
      if (ps < npass) {
        float sa = 0.0f, sq = 0.0f;
        for (int wv = 0; wv < OH_WARPS; ++wv)
          for (int cc = 0; cc < cpg; ++cc) {
            const float* src = csum + ((ps * OH_WARPS + wv) * C + slab * 64 + gl * cpg + cc) * 2;
            sa += src[0]; sq += src[1];
          }
        const float inv = 1.0f / static_cast<float>(P * cpg);
        const float mean = sa * inv;
        const float var = fmaxf(sq * inv - mean * mean, 0.0f);
        gstat[(ps * groups + slab * (64 / cpg) + gl) * 2] = mean;
        gstat[(ps * groups + slab * (64 / cpg) + gl) * 2 + 1] = 1.0f / sqrtf(var + eps);
      }
    }
    __syncthreads();
    // ---- pass 2: normalise + SiLU from registers, nine per-tap dot products per pixel and output channel
    const float* gs0 = gstat + (pass * groups + c / cpg) * 2;
    const float* gs1 = gstat + (pass * groups + (c + 1) / cpg) * 2;
    const float a0 = gs0[1] * gamma[c], b0 = fmaf(-gs0[0], a0, beta[c]);
    const float a1 = gs1[1] * gamma[c + 1], b1 = fmaf(-gs1[0], a1, beta[c + 1]);
    for (int co = 0; co < Cimg; ++co) {
      float w0[9], w1[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        w0[t] = w[(static_cast<size_t>(co) * C + c) * 9 + t];
        w1[t] = w[(static_cast<size_t>(co) * C + c + 1) * 9 + t];
      }
#pragma unroll 2
      for (int i = 0; i < ppw; ++i) {
        const int px = px0 + i;
        if (px < P) {  // warp-uniform; the second read of the row hits L1/L2
          const float2 xv = oh_load2(hb + static_cast<size_t>(px) * C + c);
          const float y0 = fmaf(xv.x, a0, b0), y1 = fmaf(xv.y, a1, b1);
          const float act0 = silu_acc(y0), act1 = silu_acc(y1);
          float v[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) v[t] = fmaf(act0, w0[t], act1 * w1[t]);
          float v8 = fmaf(act0, w0[8], act1 * w1[8]);
          warp_reduce8(v, lane);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v8 += __shfl_xor_sync(0xffffffffu, v8, o);
          float* dst = qh + (co * P + px) * 9;
          if ((lane & 3) == 0) {
            const int t = (lane >> 2) & 7;
            if (slab == 0) dst[t] = v[0]; else dst[t] += v[0];
          }
          if (lane == 1) { if (slab == 0) dst[8] = v8; else dst[8] += v8; }
        }
      }
    }
    __syncthreads();
  }
  // ---- 9-term gather per output pixel, CFG combine, NCHW fp32 store
  for (int i = threadIdx.x; i < Cimg * P; i += blockDim.x) {
    const int co = i / P, px = i % P, y = px / W, x = px % W;
    float r[2] = {0.0f, 0.0f};
    // model.scale_by_sigma (ncsnpp.py:350-351): each network output is divided by its sample's sigma before the
    // guidance combine; inside the sampler every sample shares sigma_table[*step_ctr]
    const float sig0 = sigma_table ? (step_ctr ? sigma_table[*step_ctr] : sigma_table[b]) : 1.0f;
    const float sig1 = sigma_table ? (step_ctr ? sig0 : sigma_table[b + B]) : 1.0f;
    for (int ps = 0; ps < npass; ++ps) {
      float acc = bias[co];
      const float* qq = q_all + (ps * Cimg + co) * P * 9;
      for (int dy = 0; dy < 3; ++dy) {
        const int yy = y + dy - 1;
        if (yy < 0 || yy >= H) continue;
        for (int dx = 0; dx < 3; ++dx) {
          const int xx = x + dx - 1;
          if (xx < 0 || xx >= W) continue;
          acc += qq[(yy * W + xx) * 9 + dy * 3 + dx];
        }
      }
      r[ps] = sigma_table ? __fdiv_rn(acc, ps ? sig1 : sig0) : acc;
    }
    float v = r[0];
    if (npass == 2) {
      const float wv = cfg_w ? cfg_w[b] : cfg_w_scalar;
      v = __fsub_rn(__fmul_rn(__fadd_rn(1.0f, wv), r[0]), __fmul_rn(wv, r[1]));
    }
    score[static_cast<size_t>(b) * Cimg * P + i] = v;  // NCHW [B, Cimg, H, W]
  }
}

// ------------------------------------------------------------------------------------------------
// out head, bf16 plan at the GTO-Halo shape (C = 64, 16 groups of 4 channels, one image channel, <= 88 pixels): one warp
// per guided sample, everything in registers.  The warp reads the sample's [P][64] bf16 image ONCE, laid out as the B
// fragments of mma.sync.m16n8k16 (N = 8 pixels per block, K = 16 channels per k-block): lane (g, q) holds pixel 8 nb + g,
// channels 16 kb + 4 q .. + 3.  A GroupNorm group (four adjacent channels) is then one register pair, its statistics a
// per-lane sum over the pixel blocks plus three shuffles over g.  The nine per-tap dot products q[tap][px] = sum_c
// w[tap][c] act[c][px] are ONE M = 16 (taps, 9 used) x N = 8 (pixels) x K = 64 product per pixel block with fp32-class
// operands: act and w are split into bf16 hi + lo and accumulated as lo*hi + hi*lo + hi*hi in fp32 (2^-17 relative, the
// scheme of temb_kernel).  This replaces the 14 shuffles + ~25 selects per pixel of the SIMT warp reduction: ~25 warp
// instructions per pixel instead of ~100.  The conv output is the 9-term gather of q over the 3x3 neighbourhood, then the
// reference's guidance combine.  The k index of the product is a permutation of the channels (fragment positions 2q, 2q+1,
// 2q+8, 2q+9 <-> channels 4q .. 4q+3) so that a lane's four channels are one 8-byte load.
template <int NB>
__global__ void __launch_bounds__(128, 3) out_head_mma_kernel(const __nv_bfloat16* __restrict__ h, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const float* __restrict__ w,
                                                             const float* __restrict__ bias, const float* __restrict__ cfg_w,
                                                             float cfg_w_scalar, float* __restrict__ score, int B, int H, int W,
                                                             int cfg, float eps, const float* __restrict__ sigma_table,
                                                             const int32_t* __restrict__ step_ctr) {
  constexpr int C = 64, QP = 8 * NB;
  __shared__ __align__(16) uint32_t s_wf[4][2][32][4];  // [k-block][hi|lo][lane][a0..a3]
  __shared__ __align__(16) float s_q[4][9][QP];         // [warp][tap][pixel]
  __shared__ __align__(16) float s_gb[2][C];            // gamma, beta
  __shared__ __align__(16) float s_ab[4][2][C];         // [warp][a|b][channel]: y = x a + b of the current (sample, pass)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;
  const int P = H * W;
  for (int i = tid; i < 4 * 32 * 4; i += 128) {
    const int r = i & 3, ln = (i >> 2) & 31, kb = i >> 7;
    const int tap = (ln >> 2) + ((r & 1) ? 8 : 0), ch = 16 * kb + 4 * (ln & 3) + ((r & 2) ? 2 : 0);
    const float w0 = tap < 9 ? w[ch * 9 + tap] : 0.0f, w1 = tap < 9 ? w[(ch + 1) * 9 + tap] : 0.0f;
    __nv_bfloat16 h0, l0, h1, l1;
    te_split(w0, h0, l0);
    te_split(w1, h1, l1);
    s_wf[kb][0][ln][r] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
    s_wf[kb][1][ln][r] = static_cast<uint32_t>(__bfloat16_as_ushort(l0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l1)) << 16);
  }
  if (tid < C) { s_gb[0][tid] = gamma[tid]; s_gb[1][tid] = beta[tid]; }
  __syncthreads();
  const int b = blockIdx.x * 4 + warp;
  if (b >= B) return;
  const int npass = cfg ? 2 : 1;
  float* qs = &s_q[warp][0][0];
  const float inv_n = 1.0f / static_cast<float>(4 * P);
  float r0[3] = {0.0f, 0.0f, 0.0f};
  const float bias0 = bias[0];
  const float sig0 = sigma_table ? (step_ctr ? sigma_table[*step_ctr] : sigma_table[b]) : 1.0f;
  const float sig1 = sigma_table ? (step_ctr ? sig0 : sigma_table[b + B]) : 1.0f;

  for (int ps = 0; ps < npass; ++ps) {
    const __nv_bfloat16* hb = h + static_cast<size_t>(b + ps * B) * P * C + 4 * q;
    uint2 x[NB][4];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      const int px = 8 * nb + g;
#pragma unroll
      for (int kb = 0; kb < 4; ++kb)
        x[nb][kb] = px < P ? __ldg(reinterpret_cast<const uint2*>(hb + static_cast<size_t>(px) * C + 16 * kb)) : make_uint2(0u, 0u);
    }
    // ---- GroupNorm statistics: the lane's four channels 16kb + 4q .. + 3 are one group; pixels beyond P contribute zeros
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
      float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        const float a0 = __uint_as_float(x[nb][kb].x << 16), a1 = __uint_as_float(x[nb][kb].x & 0xffff0000u);
        const float b0 = __uint_as_float(x[nb][kb].y << 16), b1 = __uint_as_float(x[nb][kb].y & 0xffff0000u);
        s1 += (a0 + a1) + (b0 + b1);
        s2 = fmaf(a0, a0, fmaf(a1, a1, fmaf(b0, b0, fmaf(b1, b1, s2))));
      }
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      const float mean = s1 * inv_n;
      const float rstd = 1.0f / sqrtf(fmaxf(s2 * inv_n - mean * mean, 0.0f) + eps);
      const float4 ga = *reinterpret_cast<const float4*>(&s_gb[0][16 * kb + 4 * q]);
      const float4 be = *reinterpret_cast<const float4*>(&s_gb[1][16 * kb + 4 * q]);
      if (g == 0) {  // (the eight lanes that share q hold the same values)
        const float4 a4 = make_float4(ga.x * rstd, ga.y * rstd, ga.z * rstd, ga.w * rstd);
        *reinterpret_cast<float4*>(&s_ab[warp][0][16 * kb + 4 * q]) = a4;
        *reinterpret_cast<float4*>(&s_ab[warp][1][16 * kb + 4 * q]) =
            make_float4(fmaf(-mean, a4.x, be.x), fmaf(-mean, a4.y, be.y), fmaf(-mean, a4.z, be.z), fmaf(-mean, a4.w, be.w));
      }
    }
    __syncwarp();
    // ---- per pixel block: normalise + SiLU + split, q[tap][px] on the tensor core
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      float d[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        float v[4] = {__uint_as_float(x[nb][kb].x << 16), __uint_as_float(x[nb][kb].x & 0xffff0000u),
                      __uint_as_float(x[nb][kb].y << 16), __uint_as_float(x[nb][kb].y & 0xffff0000u)};
        const float4 ca4 = *reinterpret_cast<const float4*>(&s_ab[warp][0][16 * kb + 4 * q]);
        const float4 cb4 = *reinterpret_cast<const float4*>(&s_ab[warp][1][16 * kb + 4 * q]);
        const float ca[4] = {ca4.x, ca4.y, ca4.z, ca4.w}, cb[4] = {cb4.x, cb4.y, cb4.z, cb4.w};
        uint32_t bh[2], bl[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float a[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float y = fmaf(v[2 * j + e], ca[2 * j + e], cb[2 * j + e]);
            a[e] = __fdividef(y, 1.0f + __expf(-y));
          }
          const __nv_bfloat162 hi = __floats2bfloat162_rn(a[0], a[1]);
          const float2 hf = __bfloat1622float2(hi);
          const __nv_bfloat162 lo = __floats2bfloat162_rn(a[0] - hf.x, a[1] - hf.y);
          bh[j] = *reinterpret_cast<const uint32_t*>(&hi);
          bl[j] = *reinterpret_cast<const uint32_t*>(&lo);
        }
        const uint4 wh = *reinterpret_cast<const uint4*>(&s_wf[kb][0][lane][0]);
        const uint4 wl = *reinterpret_cast<const uint4*>(&s_wf[kb][1][lane][0]);
        const uint32_t ah[4] = {wh.x, wh.y, wh.z, wh.w}, al[4] = {wl.x, wl.y, wl.z, wl.w};
        te_mma(d, al, bh[0], bh[1]);  // small terms first
        te_mma(d, ah, bl[0], bl[1]);
        te_mma(d, ah, bh[0], bh[1]);
      }
      *reinterpret_cast<float2*>(qs + g * QP + 8 * nb + 2 * q) = make_float2(d[0], d[1]);
      if (g == 0) *reinterpret_cast<float2*>(qs + 8 * QP + 8 * nb + 2 * q) = make_float2(d[2], d[3]);
    }
    __syncwarp();
    // ---- 9-term gather per output pixel (+ scale_by_sigma), guidance combine after the second pass
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int px = lane + 32 * k;
      if (px < P) {
        const int y = px / W, xx0 = px - y * W;
        float acc = bias0;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const int yy = y + dy - 1;
          if (yy < 0 || yy >= H) continue;
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const int xx = xx0 + dx - 1;
            if (xx < 0 || xx >= W) continue;
            acc += qs[(dy * 3 + dx) * QP + yy * W + xx];
          }
        }
        const float r = sigma_table ? __fdiv_rn(acc, ps ? sig1 : sig0) : acc;
        if (ps == 0 && npass == 2) {
          r0[k] = r;
        } else {
          float v = r;
          if (npass == 2) {
            const float wv = cfg_w ? cfg_w[b] : cfg_w_scalar;
            v = __fsub_rn(__fmul_rn(__fadd_rn(1.0f, wv), r0[k]), __fmul_rn(wv, r));
          }
          score[static_cast<size_t>(b) * P + px] = v;
        }
      }
    }
    __syncwarp();  // the next pass overwrites q
  }
}

int outhead_launch(const rd_op_outhead& op, cudaStream_t st) {
  RD_REQUIRE(op.h && op.gamma && op.beta && op.w && op.bias && op.score, "out_head: null pointer");
  RD_REQUIRE(op.groups > 0 && op.C % op.groups == 0 && op.C % 64 == 0, "out_head: C must be a multiple of 64 and of the group count");
  RD_REQUIRE(64 % (op.C / op.groups) == 0, "out_head: channels per group must divide 64");
  RD_REQUIRE(op.cfg ? (op.B2 == 2 * op.B) : (op.B2 == op.B), "out_head: B2 must be 2B with cfg, B otherwise");
  const int P = op.H * op.W;
  const int smem = (2 * op.C_img * P * 9 + 2 * OH_WARPS * op.C * 2 + 2 * op.groups * 2) * 4;
  RD_REQUIRE(smem <= 48 * 1024, "out_head: image too large for shared memory");
  const int threads = (op.cfg ? 2 : 1) * OH_WARPS * 32;
  // bf16 plan at the GTO-Halo shape: one warp per sample on mma.sync (RD_OUTHEAD_MMA=0 keeps the general kernel: A/B runs)
  static const bool use_mma = !(getenv("RD_OUTHEAD_MMA") && atoi(getenv("RD_OUTHEAD_MMA")) == 0);
  if (use_mma && op.precision != RD_PREC_F32X3 && op.C == 64 && op.groups == 16 && op.C_img == 1 && P <= 88 &&
      (reinterpret_cast<uintptr_t>(op.h) & 7) == 0) {
    const __nv_bfloat16* hp = static_cast<const __nv_bfloat16*>(op.h);
    const int grid = (op.B + 3) / 4;
    if (P <= 72)
      out_head_mma_kernel<9><<<grid, 128, 0, st>>>(hp, op.gamma, op.beta, op.w, op.bias, op.cfg_w, op.cfg_w_scalar, op.score, op.B, op.H,
                                                   op.W, op.cfg, op.eps, op.sigma_table, op.step_ctr);
    else
      out_head_mma_kernel<11><<<grid, 128, 0, st>>>(hp, op.gamma, op.beta, op.w, op.bias, op.cfg_w, op.cfg_w_scalar, op.score, op.B, op.H,
                                                    op.W, op.cfg, op.eps, op.sigma_table, op.step_ctr);
    return check_launch("out_head_mma_kernel");
  }
  if (op.precision == RD_PREC_F32X3)
    out_head_kernel<float><<<op.B, threads, smem, st>>>(static_cast<const float*>(op.h), op.gamma, op.beta, op.w, op.bias, op.cfg_w,
                                                        op.cfg_w_scalar, op.score, op.B, op.C, op.C_img, op.H, op.W, op.groups, op.cfg,
                                                        op.eps, op.sigma_table, op.step_ctr);
  else
    out_head_kernel<__nv_bfloat16><<<op.B, threads, smem, st>>>(static_cast<const __nv_bfloat16*>(op.h), op.gamma, op.beta, op.w, op.bias,
                                                                op.cfg_w, op.cfg_w_scalar, op.score, op.B, op.C, op.C_img, op.H, op.W,
                                                                op.groups, op.cfg, op.eps, op.sigma_table, op.step_ctr);
  return check_launch("out_head_kernel");
}

}  // namespace rd
