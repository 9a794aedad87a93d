// small_ops.cu -- the thin ends of the NCSN++ forward that are not GEMM-shaped enough for tcgen05:
//   temb_kernel     per-sample  Dense_0(SiLU(temb))  for all ResBlocks at once
//                   (reference ncsnpp.py:252-262, layerspp.py:201-202)
//   in_conv_kernel  input_conv: 3x3, channels(=1) -> nf, reads the fp32 sampler state, CFG-duplicated
//                   (ncsnpp.py:266, models/utils.py:121)
//   out_head_kernel out_norm + SiLU + out_conv (nf -> channels) + classifier-free-guidance combine
//                   (ncsnpp.py:343-347, models/utils.py:124-138), fp32 score out
#include "rd_common.h"
#include <cuda_bf16.h>

namespace rd {

__device__ __forceinline__ float silu_acc(float v) { return v / (1.0f + expf(-v)); }

// ------------------------------------------------------------------------------------------------
// out[b][o] = dense_b[o] + sum_k dense_w[o][k] * SiLU(time_table[row(b)][k] + sum_c label_w[k][c]*labels[b][c])
// fp32 SIMT GEMM (the temb path stays fp32): 128x128 block tile, 8x8 outputs per thread, K tile 16.
constexpr int TE_BM = 128, TE_BN = 128, TE_BK = 16;

__global__ void __launch_bounds__(256) temb_kernel(const float* __restrict__ time_table, const float* __restrict__ label_w,
                                                   const float* __restrict__ labels, const float* __restrict__ dense_w,
                                                   const float* __restrict__ dense_b, float* __restrict__ out,
                                                   const int32_t* __restrict__ step_ctr,
                                                   const int32_t* __restrict__ row_idx, int B2, int K, int NC, int NO) {
  __shared__ __align__(16) float As[TE_BK][TE_BM];
  __shared__ __align__(16) float Bs[TE_BK][TE_BN];
  const int step = step_ctr ? *step_ctr : 0;
  const int m0 = blockIdx.x * TE_BM, n0 = blockIdx.y * TE_BN;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[8][8] = {};
  // every thread stages 8 A and 8 B elements per k-tile; the next tile's global reads are issued before the
  // current tile's 1024 FMAs so their latency is hidden
  float pa[8], pb[8];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = tid + u * 256;
      const int kk = i & (TE_BK - 1), mm = i >> 4;
      const int b = m0 + mm, k = k0 + kk;
      float t = 0.0f;
      if (b < B2 && k < K) {
        const int row = row_idx ? row_idx[b] : step;
        t = time_table[static_cast<size_t>(row) * K + k];
        for (int c = 0; c < NC; ++c) t = fmaf(label_w[k * NC + c], labels[static_cast<size_t>(b) * NC + c], t);
      }
      pa[u] = t;
      const int o = n0 + mm;
      pb[u] = (o < NO && k < K) ? dense_w[static_cast<size_t>(o) * K + k] : 0.0f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += TE_BK) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = tid + u * 256;
      const int kk = i & (TE_BK - 1), mm = i >> 4;
      const bool live = (m0 + mm < B2) && (k0 + kk < K);
      As[kk][mm] = live ? silu_acc(pa[u]) : 0.0f;
      Bs[kk][mm] = pb[u];
    }
    __syncthreads();
    if (k0 + TE_BK < K) fetch(k0 + TE_BK);
#pragma unroll
    for (int kk = 0; kk < TE_BK; ++kk) {
      // rows ty*4..+3 and 64+ty*4..+3, columns tx*4..+3 and 64+tx*4..+3: conflict-free float4 reads
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]), a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]), b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int b = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (b >= B2) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int o = n0 + jh * 64 + tx * 4;
      if (o + 3 < NO) {
        const float4 bb = *reinterpret_cast<const float4*>(dense_b + o);
        *reinterpret_cast<float4*>(out + static_cast<size_t>(b) * NO + o) =
            make_float4(acc[i][jh * 4] + bb.x, acc[i][jh * 4 + 1] + bb.y, acc[i][jh * 4 + 2] + bb.z, acc[i][jh * 4 + 3] + bb.w);
      } else {
        for (int j = 0; j < 4; ++j)
          if (o + j < NO) out[static_cast<size_t>(b) * NO + o + j] = acc[i][jh * 4 + j] + dense_b[o + j];
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
__global__ void __launch_bounds__(256) in_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int B,
                                                      int B2, int Cin, int Cout, int H, int W) {
  extern __shared__ __align__(16) float sw[];  // [Cin*9][Cout] (tap-major) + bias[Cout]
  const int wn = Cout * Cin * 9;
  for (int i = threadIdx.x; i < wn; i += blockDim.x) {
    const int co = i / (Cin * 9), t = i - co * (Cin * 9);
    sw[t * Cout + co] = w[i];
  }
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[wn + i] = bias[i];
  __syncthreads();
  const int P = H * W;
  const int total = B2 * P;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += gridDim.x * blockDim.x) {
    const int b2 = pix / P, px = pix - b2 * P;
    const int yh = px / W, xw = px - yh * W;
    const float* xb = x + static_cast<size_t>(b2 % B) * Cin * P;  // x.repeat(2,1,1,1)
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
      __nv_bfloat16* dst = out + static_cast<size_t>(pix) * Cout + c0;
#pragma unroll
      for (int j8 = 0; j8 < 8; ++j8) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          __nv_bfloat162 h = __floats2bfloat162_rn(acc[8 * j8 + 2 * j], acc[8 * j8 + 2 * j + 1]);
          pk[j] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(dst + 8 * j8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
  }
}

int inconv_launch(const rd_op_inconv& op, cudaStream_t st) {
  RD_REQUIRE(op.x && op.w && op.bias && op.out && op.B2 > 0 && op.B > 0, "in_conv: null pointer / empty batch");
  RD_REQUIRE(op.C_out % 64 == 0, "in_conv: C_out must be a multiple of 64");
  const int smem = (op.C_out * op.C_in * 9 + op.C_out) * 4;
  RD_REQUIRE(smem <= 48 * 1024, "in_conv: weights do not fit shared memory");
  const size_t total = static_cast<size_t>(op.B2) * op.H * op.W;
  RD_REQUIRE(total < (1u << 31), "in_conv: batch too large");
  size_t blocks = (total + 255) / 256;
  if (blocks > static_cast<size_t>(kNumSMs) * 16) blocks = static_cast<size_t>(kNumSMs) * 16;
  in_conv_kernel<<<static_cast<unsigned>(blocks), 256, smem, st>>>(op.x, op.w, op.bias, static_cast<__nv_bfloat16*>(op.out),
                                                                   op.B, op.B2, op.C_in, op.C_out, op.H, op.W);
  return check_launch("in_conv_kernel");
}

// ------------------------------------------------------------------------------------------------
// out head: one CTA per guided sample; with CFG the conditional and unconditional copies are processed
// concurrently by the two 128-thread halves of the CTA and combined at the end.
constexpr int OH_THREADS = 128;  // per half

__global__ void __launch_bounds__(2 * OH_THREADS) out_head_kernel(const __nv_bfloat16* __restrict__ h, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, const float* __restrict__ w,
                                                                  const float* __restrict__ bias, const float* __restrict__ cfg_w,
                                                                  float cfg_w_scalar, float* __restrict__ score, int B, int C,
                                                                  int Cimg, int H, int W, int groups, int cfg, float eps) {
  extern __shared__ __align__(16) float sm[];
  const int P = H * W;
  const int LD = C + 4;                // row stride: 16-B aligned, rows 4 banks apart -> conflict-free float4 reads
  const int npass = cfg ? 2 : 1;
  float* act_all = sm;                             // [2][P][LD]   (16-B aligned: LD % 4 == 0)
  float* sw = act_all + 2 * P * LD;                // [Cimg][9][C] (tap-major; C % 4 == 0 keeps float4 alignment)
  float* gsum_all = sw + Cimg * C * 9;             // [2][groups][2]
  float* res = gsum_all + 2 * groups * 2;          // [2][Cimg][P]
  const int tid = threadIdx.x & (OH_THREADS - 1), pass = threadIdx.x / OH_THREADS;
  const int b = blockIdx.x;
  const int cpg = C / groups;
  float* act = act_all + pass * P * LD;
  float* gsum = gsum_all + pass * groups * 2;
  for (int i = threadIdx.x; i < Cimg * C * 9; i += blockDim.x) {
    const int co = i / (C * 9), r = i - co * C * 9, c = r / 9, t = r - c * 9;
    sw[(co * 9 + t) * C + c] = w[i];
  }
  const __nv_bfloat16* hb = h + static_cast<size_t>(b + pass * B) * P * C;
  // load (coalesced 4-byte reads)
  for (int i = tid; i < P * C / 2; i += OH_THREADS) {
    const float2 f = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(hb)[i]);
    const int px = (2 * i) / C, c = (2 * i) % C;
    act[px * LD + c] = f.x;
    act[px * LD + c + 1] = f.y;
  }
  __syncthreads();
  for (int g = tid >> 5; g < groups; g += OH_THREADS / 32) {  // one warp per group
    float sacc = 0.0f, q = 0.0f;
    for (int i = tid & 31; i < P * cpg; i += 32) {
      const float v = act[(i / cpg) * LD + g * cpg + (i % cpg)];
      sacc += v; q += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sacc += __shfl_xor_sync(0xffffffffu, sacc, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if ((tid & 31) == 0) {
      const float mean = sacc / static_cast<float>(P * cpg);
      const float var = fmaxf(q / static_cast<float>(P * cpg) - mean * mean, 0.0f);
      gsum[2 * g] = mean;
      gsum[2 * g + 1] = 1.0f / sqrtf(var + eps);
    }
  }
  __syncthreads();
  for (int i = tid; i < P * C; i += OH_THREADS) {
    const int px = i / C, c = i % C, g = c / cpg;
    const float v = (act[px * LD + c] - gsum[2 * g]) * gsum[2 * g + 1] * gamma[c] + beta[c];
    act[px * LD + c] = silu_acc(v);
  }
  __syncthreads();
  for (int i = tid; i < Cimg * P; i += OH_THREADS) {
    const int co = i / P, px = i % P, y = px / W, x = px % W;
    float acc = bias[co];
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = y + dy - 1;
      if (yy < 0 || yy >= H) continue;
      for (int dx = 0; dx < 3; ++dx) {
        const int xx = x + dx - 1;
        if (xx < 0 || xx >= W) continue;
        const float4* a = reinterpret_cast<const float4*>(act + (yy * W + xx) * LD);
        const float4* ww = reinterpret_cast<const float4*>(sw + (co * 9 + dy * 3 + dx) * C);
        float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f, p3 = 0.0f;
        for (int c4 = 0; c4 < C / 4; ++c4) {
          const float4 av = a[c4], wv = ww[c4];
          p0 = fmaf(av.x, wv.x, p0); p1 = fmaf(av.y, wv.y, p1); p2 = fmaf(av.z, wv.z, p2); p3 = fmaf(av.w, wv.w, p3);
        }
        acc += (p0 + p1) + (p2 + p3);
      }
    }
    res[(pass * Cimg + co) * P + px] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cimg * P; i += blockDim.x) {
    float v = res[i];
    if (npass == 2) {
      const float wv = cfg_w ? cfg_w[b] : cfg_w_scalar;
      v = __fsub_rn(__fmul_rn(__fadd_rn(1.0f, wv), res[i]), __fmul_rn(wv, res[Cimg * P + i]));
    }
    score[static_cast<size_t>(b) * Cimg * P + i] = v;  // NCHW [B, Cimg, H, W]
  }
}

int outhead_launch(const rd_op_outhead& op, cudaStream_t st) {
  RD_REQUIRE(op.h && op.gamma && op.beta && op.w && op.bias && op.score, "out_head: null pointer");
  RD_REQUIRE(op.groups > 0 && op.C % op.groups == 0 && op.C % 4 == 0, "out_head: bad GroupNorm geometry");
  RD_REQUIRE(op.cfg ? (op.B2 == 2 * op.B) : (op.B2 == op.B), "out_head: B2 must be 2B with cfg, B otherwise");
  const int P = op.H * op.W;
  const int smem = (2 * P * (op.C + 4) + op.C_img * op.C * 9 + 2 * op.groups * 2 + 2 * op.C_img * P) * 4;
  static int configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    RD_REQUIRE(smem <= 227 * 1024, "out_head: image too large for shared memory");
    cudaError_t e = cudaFuncSetAttribute(out_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "out_head: %s", cudaGetErrorString(e));
    configured = smem;
  }
  out_head_kernel<<<op.B, (op.cfg ? 2 : 1) * OH_THREADS, smem, st>>>(static_cast<const __nv_bfloat16*>(op.h), op.gamma, op.beta, op.w, op.bias,
                                                  op.cfg_w, op.cfg_w_scalar, op.score, op.B, op.C, op.C_img, op.H, op.W,
                                                  op.groups, op.cfg, op.eps);
  return check_launch("out_head_kernel");
}

}  // namespace rd
