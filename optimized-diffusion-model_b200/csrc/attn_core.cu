// attn_core.cu -- softmax(q k^T / sqrt(C)) v for the single-head pixel attention of AttnBlockpp
// (reference models/layerspp.py:87-91), one CTA per sample, T = H*W tokens (72 / 81 for GTO-Halo),
// C = 64 channels.  q/k/v come from the fused qkv projection (conv_gemm.cu, N = 3C) as bf16
// [B2, T, 3C]; the output feeds the NIN_3 projection kernel.
//
// Round-1 implementation: register-resident flash-style kernel on mma.sync.m16n8k16 (bf16 in,
// fp32 accumulate, fp32 softmax).  It is ~3 % of the network's FLOPs; the projections around it
// (11.8 of the 18.4 MFLOP per attention-bearing sample) already run on tcgen05.  Moving QK^T / PV
// to tcgen05 (M=128 tile per sample, S in TMEM) is the planned follow-up.
#include "rd_common.h"
#include <cuda_bf16.h>

namespace rd {

constexpr int ATT_C = 64;
constexpr int ATT_LD = 72;  // smem row stride in bf16 (144 B): conflict-free fragment loads

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// T16 = number of 16-key blocks (keys padded to 16*T16)
template <int T16>
__global__ void __launch_bounds__(32 * T16) attn_core_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                             __nv_bfloat16* __restrict__ out, int T, float scale) {
  constexpr int TP = 16 * T16;
  __shared__ __align__(16) __nv_bfloat16 Ks[TP * ATT_LD];
  __shared__ __align__(16) __nv_bfloat16 Vs[TP * ATT_LD];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const __nv_bfloat16* base = qkv + static_cast<size_t>(b) * T * (3 * ATT_C);

  // stage K and V (zero rows beyond T)
  for (int i = tid; i < TP * 16; i += blockDim.x) {
    const int row = i >> 4, seg = i & 15;  // seg 0-7: K chunks, 8-15: V chunks
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row < T) v = *reinterpret_cast<const uint4*>(base + static_cast<size_t>(row) * (3 * ATT_C) + ATT_C + seg * 8);
    __nv_bfloat16* dst = (seg < 8 ? Ks : Vs) + row * ATT_LD + (seg & 7) * 8;
    *reinterpret_cast<uint4*>(dst) = v;
  }
  __syncthreads();

  const int g = lane >> 2, q4 = lane & 3;
  const int r0 = warp * 16 + g, r1 = r0 + 8;
  const int r0c = min(r0, T - 1), r1c = min(r1, T - 1);
  // Q fragments: 4 k-steps of 16 channels
  uint32_t qa[4][4];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const __nv_bfloat16* q0 = base + static_cast<size_t>(r0c) * (3 * ATT_C) + kk * 16 + 2 * q4;
    const __nv_bfloat16* q1 = base + static_cast<size_t>(r1c) * (3 * ATT_C) + kk * 16 + 2 * q4;
    qa[kk][0] = *reinterpret_cast<const uint32_t*>(q0);
    qa[kk][1] = *reinterpret_cast<const uint32_t*>(q1);
    qa[kk][2] = *reinterpret_cast<const uint32_t*>(q0 + 8);
    qa[kk][3] = *reinterpret_cast<const uint32_t*>(q1 + 8);
  }
  // S = Q K^T
  float s[2 * T16][4];
#pragma unroll
  for (int nb = 0; nb < 2 * T16; ++nb) {
    s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.0f;
    const __nv_bfloat16* kr = Ks + (nb * 8 + g) * ATT_LD + 2 * q4;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kr + kk * 16);
      const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kr + kk * 16 + 8);
      mma_bf16_16816(s[nb], qa[kk], b0, b1);
    }
  }
  // softmax over the T valid keys (rows r0: elements [0],[1]; r1: [2],[3]; key = nb*8 + 2*q4 + {0,1})
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int nb = 0; nb < 2 * T16; ++nb) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const bool ok = (nb * 8 + 2 * q4 + e) < T;
      s[nb][e] = ok ? s[nb][e] * scale : -INFINITY;
      s[nb][2 + e] = ok ? s[nb][2 + e] * scale : -INFINITY;
      m0 = fmaxf(m0, s[nb][e]);
      m1 = fmaxf(m1, s[nb][2 + e]);
    }
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  float l0 = 0.0f, l1 = 0.0f;
#pragma unroll
  for (int nb = 0; nb < 2 * T16; ++nb) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      s[nb][e] = __expf(s[nb][e] - m0);
      s[nb][2 + e] = __expf(s[nb][2 + e] - m1);
      l0 += s[nb][e];
      l1 += s[nb][2 + e];
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);

  // O = P V : k-steps of 16 keys, 8 output n-blocks of 8 channels
  float o[ATT_C / 8][4];
#pragma unroll
  for (int nb = 0; nb < ATT_C / 8; ++nb) o[nb][0] = o[nb][1] = o[nb][2] = o[nb][3] = 0.0f;
#pragma unroll
  for (int kk = 0; kk < T16; ++kk) {
    uint32_t pa[4];
    pa[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
    pa[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
    pa[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    pa[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
    for (int nb = 0; nb < ATT_C / 8; ++nb) {
      // B fragment (16 keys x 8 channels) from row-major V via ldmatrix.trans
      uint32_t b0, b1;
      const uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(Vs + (kk * 16 + (lane & 15)) * ATT_LD + nb * 8));
      asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(addr));
      mma_bf16_16816(o[nb], pa, b0, b1);
    }
  }
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  __nv_bfloat16* ob = out + static_cast<size_t>(b) * T * ATT_C;
#pragma unroll
  for (int nb = 0; nb < ATT_C / 8; ++nb) {
    if (r0 < T) *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(r0) * ATT_C + nb * 8 + 2 * q4) = pack_bf16(o[nb][0] * i0, o[nb][1] * i0);
    if (r1 < T) *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(r1) * ATT_C + nb * 8 + 2 * q4) = pack_bf16(o[nb][2] * i1, o[nb][3] * i1);
  }
}

int attn_launch(const rd_op_attn& op, cudaStream_t st) {
  RD_REQUIRE(op.qkv && op.out && op.B2 > 0, "attn: null pointer / empty batch");
  RD_REQUIRE(op.C == ATT_C, "attn: only C == %d is supported in this round (got %d)", ATT_C, op.C);
  RD_REQUIRE(op.T >= 1 && op.T <= 128, "attn: T must be in [1,128] (got %d)", op.T);
  const float scale = 1.0f / sqrtf(static_cast<float>(op.C));  // int(C) ** (-0.5)
  const __nv_bfloat16* qkv = static_cast<const __nv_bfloat16*>(op.qkv);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(op.out);
  const int t16 = (op.T + 15) / 16;
  switch (t16) {
#define RD_ATT_CASE(n) \
  case n: attn_core_kernel<n><<<op.B2, 32 * n, 0, st>>>(qkv, out, op.T, scale); break;
    RD_ATT_CASE(1) RD_ATT_CASE(2) RD_ATT_CASE(3) RD_ATT_CASE(4) RD_ATT_CASE(5) RD_ATT_CASE(6) RD_ATT_CASE(7) RD_ATT_CASE(8)
#undef RD_ATT_CASE
    default: return fail(RD_E_UNSUPPORTED, "attn: T=%d unsupported", op.T);
  }
  return check_launch("attn_core_kernel");
}

}  // namespace rd
