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
// out[b][o] = dense_b[o] + sum_k dense_w[o][k] * SiLU(time_table[step][k] + sum_c label_w[k][c]*labels[b][c])
// fp32 SIMT GEMM, 64x64 block tile, 4x4 per thread, K tile 32.
constexpr int TE_BM = 64, TE_BN = 64, TE_BK = 32;

__global__ void __launch_bounds__(256) temb_kernel(const float* __restrict__ time_table, const float* __restrict__ label_w,
                                                   const float* __restrict__ labels, const float* __restrict__ dense_w,
                                                   const float* __restrict__ dense_b, float* __restrict__ out,
                                                   const int32_t* __restrict__ step_ctr,
                                                   const int32_t* __restrict__ row_idx, int B2, int K, int NC, int NO) {
  __shared__ float As[TE_BK][TE_BM + 1];
  __shared__ float Bs[TE_BK][TE_BN + 1];
  const int step = step_ctr ? *step_ctr : 0;
  const int m0 = blockIdx.x * TE_BM, n0 = blockIdx.y * TE_BN;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += TE_BK) {
    for (int i = tid; i < TE_BM * TE_BK; i += 256) {
      const int kk = i % TE_BK, mm = i / TE_BK;
      const int b = m0 + mm, k = k0 + kk;
      float v = 0.0f;
      if (b < B2 && k < K) {
        const int row = row_idx ? row_idx[b] : step;
        float t = time_table[static_cast<size_t>(row) * K + k];
        for (int c = 0; c < NC; ++c) t += label_w[k * NC + c] * labels[static_cast<size_t>(b) * NC + c];
        v = silu_acc(t);
      }
      As[kk][mm] = v;
    }
    for (int i = tid; i < TE_BN * TE_BK; i += 256) {
      const int kk = i % TE_BK, nn = i / TE_BK;
      const int o = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (o < NO && k < K) ? dense_w[static_cast<size_t>(o) * K + k] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TE_BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int b = m0 + ty * 4 + i;
    if (b >= B2) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int o = n0 + tx * 4 + j;
      if (o < NO) out[static_cast<size_t>(b) * NO + o] = acc[i][j] + dense_b[o];
    }
  }
}

int temb_launch(const rd_op_temb& op, cudaStream_t st) {
  RD_REQUIRE(op.time_table && op.dense_w && op.dense_b && op.out && op.B2 > 0, "temb: null pointer / empty batch");
  RD_REQUIRE(op.num_classes == 0 || (op.label_w && op.labels), "temb: labels missing");
  dim3 grid((op.B2 + TE_BM - 1) / TE_BM, (op.n_out_total + TE_BN - 1) / TE_BN);
  temb_kernel<<<grid, 256, 0, st>>>(op.time_table, op.label_w, op.labels, static_cast<const float*>(op.dense_w), op.dense_b,
                                    op.out, op.step_ctr, op.row_idx, op.B2, op.temb_dim, op.num_classes, op.n_out_total);
  return check_launch("temb_kernel");
}

// ------------------------------------------------------------------------------------------------
// input_conv: one thread per (sample, pixel, 8 output channels); NHWC bf16 out
__global__ void __launch_bounds__(256) in_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int B,
                                                      int B2, int Cin, int Cout, int H, int W) {
  extern __shared__ float sw[];  // [Cin*9][Cout] (tap-major: lanes read consecutive output channels) + bias[Cout]
  const int wn = Cout * Cin * 9;
  for (int i = threadIdx.x; i < wn; i += blockDim.x) {
    const int co = i / (Cin * 9), t = i - co * (Cin * 9);
    sw[t * Cout + co] = w[i];
  }
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[wn + i] = bias[i];
  __syncthreads();
  const int KC = Cout / 8;
  const size_t total = static_cast<size_t>(B2) * H * W * KC;
  for (size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int kc = static_cast<int>(idx % KC);
    const size_t pix = idx / KC;
    const int xw = static_cast<int>(pix % W), yh = static_cast<int>((pix / W) % H);
    const int b2 = static_cast<int>(pix / (static_cast<size_t>(W) * H));
    const float* xb = x + static_cast<size_t>(b2 % B) * Cin * H * W;  // x.repeat(2,1,1,1)
    float acc[8];
    {
      const float4 b0 = *reinterpret_cast<const float4*>(sw + wn + kc * 8), b1 = *reinterpret_cast<const float4*>(sw + wn + kc * 8 + 4);
      acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w; acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
    }
    for (int ci = 0; ci < Cin; ++ci)
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int yy = yh + dy - 1;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int xx = xw + dx - 1;
          if (xx < 0 || xx >= W) continue;
          const float v = __ldg(xb + (static_cast<size_t>(ci) * H + yy) * W + xx);
          const float* wp = sw + (ci * 9 + dy * 3 + dx) * Cout + kc * 8;
          const float4 w0 = *reinterpret_cast<const float4*>(wp), w1 = *reinterpret_cast<const float4*>(wp + 4);
          acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
          acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]); acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
        }
      }
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 h = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
      pk[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(out + pix * Cout + kc * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

int inconv_launch(const rd_op_inconv& op, cudaStream_t st) {
  RD_REQUIRE(op.x && op.w && op.bias && op.out && op.B2 > 0 && op.B > 0, "in_conv: null pointer / empty batch");
  RD_REQUIRE(op.C_out % 8 == 0, "in_conv: C_out must be a multiple of 8");
  const int smem = (op.C_out * op.C_in * 9 + op.C_out) * 4;
  RD_REQUIRE(smem <= 48 * 1024, "in_conv: weights do not fit shared memory");
  const size_t total = static_cast<size_t>(op.B2) * op.H * op.W * (op.C_out / 8);
  size_t blocks = (total + 255) / 256;
  if (blocks > static_cast<size_t>(kNumSMs) * 8) blocks = static_cast<size_t>(kNumSMs) * 8;
  in_conv_kernel<<<static_cast<unsigned>(blocks), 256, smem, st>>>(op.x, op.w, op.bias, static_cast<__nv_bfloat16*>(op.out),
                                                                   op.B, op.B2, op.C_in, op.C_out, op.H, op.W);
  return check_launch("in_conv_kernel");
}

// ------------------------------------------------------------------------------------------------
// out head: one CTA per guided sample; processes the conditional and (if cfg) unconditional copy.
constexpr int OH_THREADS = 128;

__global__ void __launch_bounds__(OH_THREADS) out_head_kernel(const __nv_bfloat16* __restrict__ h, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, const float* __restrict__ w,
                                                              const float* __restrict__ bias, const float* __restrict__ cfg_w,
                                                              float cfg_w_scalar, float* __restrict__ score, int B, int C,
                                                              int Cimg, int H, int W, int groups, int cfg, float eps) {
  extern __shared__ float sm[];
  const int P = H * W;
  const int LD = C + 4;                // row stride: 16-B aligned, rows 4 banks apart -> conflict-free float4 reads
  float* act = sm;                     // [P][LD]
  float* sw = act + P * LD;            // [Cimg][9][C] (tap-major)
  float* gsum = sw + Cimg * C * 9;     // [groups][2]
  float* res = gsum + groups * 2;      // [2][Cimg][P]
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const int cpg = C / groups;
  for (int i = tid; i < Cimg * C * 9; i += OH_THREADS) {
    const int co = i / (C * 9), r = i - co * C * 9, c = r / 9, t = r - c * 9;
    sw[(co * 9 + t) * C + c] = w[i];
  }
  const int npass = cfg ? 2 : 1;
  for (int pass = 0; pass < npass; ++pass) {
    const __nv_bfloat16* hb = h + static_cast<size_t>(b + pass * B) * P * C;
    __syncthreads();
    for (int i = tid; i < groups * 2; i += OH_THREADS) gsum[i] = 0.0f;
    __syncthreads();
    // load + statistics: thread handles channel pair columns; coalesced 4-byte loads
    for (int i = tid; i < P * C / 2; i += OH_THREADS) {
      const float2 f = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(hb)[i]);
      const int px = (2 * i) / C, c = (2 * i) % C;
      act[px * LD + c] = f.x;
      act[px * LD + c + 1] = f.y;
    }
    __syncthreads();
    for (int g = tid >> 5; g < groups; g += OH_THREADS / 32) {  // one warp per group
      float s = 0.0f, q = 0.0f;
      for (int i = tid & 31; i < P * cpg; i += 32) {
        const float v = act[(i / cpg) * LD + g * cpg + (i % cpg)];
        s += v; q += v * v;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
      if ((tid & 31) == 0) {
        const float mean = s / static_cast<float>(P * cpg);
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
  }
  __syncthreads();
  for (int i = tid; i < Cimg * P; i += OH_THREADS) {
    float v = res[i];
    if (cfg) {
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
  const int smem = (P * (op.C + 4) + op.C_img * op.C * 9 + op.groups * 2 + 2 * op.C_img * P) * 4;
  static int configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    RD_REQUIRE(smem <= 227 * 1024, "out_head: image too large for shared memory");
    cudaError_t e = cudaFuncSetAttribute(out_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "out_head: %s", cudaGetErrorString(e));
    configured = smem;
  }
  out_head_kernel<<<op.B, OH_THREADS, smem, st>>>(static_cast<const __nv_bfloat16*>(op.h), op.gamma, op.beta, op.w, op.bias,
                                                  op.cfg_w, op.cfg_w_scalar, op.score, op.B, op.C, op.C_img, op.H, op.W,
                                                  op.groups, op.cfg, op.eps);
  return check_launch("out_head_kernel");
}

}  // namespace rd
