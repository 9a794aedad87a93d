// probe_split.cu -- how accurate is a split-bf16 ("x3") GEMM on tcgen05?
//   D[128 x N] = A[128 x K] * B[N x K]^T with fp32 inputs, each split into bf16 hi + bf16 lo (x ~= hi + lo to 2^-17),
//   accumulated in TMEM (fp32) as  lo*hi + hi*lo + hi*hi  (mode 3), hi*hi only (mode 1), or all four terms (mode 4).
// The result is compared on the host with (a) the fp64 product of the ORIGINAL fp32 operands and (b) the fp64 product of
// the hi+lo representations (isolates the tensor core's accumulation error from the 16-bit representation error).
// This decides whether the fp32-class mode of conv_gemm.cu can rely on TMEM accumulation.
// usage: probe_split K N mode [seed]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "../optimized-diffusion-model_b200/csrc/rd_ptx.cuh"

using namespace rd;

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                       \
    }                                                                                \
  } while (0)

// operands in global memory: [K/64 chunks][hi|lo][8 k-chunks][rows][8] bf16 (the conv kernel's K-major image)
__global__ void __launch_bounds__(128) split_kernel(const __nv_bfloat16* __restrict__ Ag, const __nv_bfloat16* __restrict__ Bg,
                                                    float* __restrict__ D, int nchunks, int N, int mode) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int a_half = 8 * 128 * 16, b_half = 8 * N * 16;
  unsigned char* As = smem;               // hi then lo
  unsigned char* Bs = smem + 2 * a_half;  // hi then lo
  if (tid == 0) { mbar_init(&bar_done, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 64 > N ? 64 : (N <= 128 ? 128 : 256));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = umma_idesc_bf16(128, N);
  for (int c = 0; c < nchunks; ++c) {
    const uint4* ag = reinterpret_cast<const uint4*>(Ag) + static_cast<size_t>(c) * 2 * a_half / 16;
    const uint4* bg = reinterpret_cast<const uint4*>(Bg) + static_cast<size_t>(c) * 2 * b_half / 16;
    for (int i = tid; i < 2 * a_half / 16; i += 128) reinterpret_cast<uint4*>(As)[i] = ag[i];
    for (int i = tid; i < 2 * b_half / 16; i += 128) reinterpret_cast<uint4*>(Bs)[i] = bg[i];
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after_sync();
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t a_hi = smem_u32(As) + kk * 2 * 128 * 16, a_lo = a_hi + a_half;
        const uint32_t b_hi = smem_u32(Bs) + kk * 2 * N * 16, b_lo = b_hi + b_half;
        const uint64_t dah = umma_desc_kmajor(a_hi, 128 * 16, 128), dal = umma_desc_kmajor(a_lo, 128 * 16, 128);
        const uint64_t dbh = umma_desc_kmajor(b_hi, N * 16, 128), dbl = umma_desc_kmajor(b_lo, N * 16, 128);
        uint32_t acc = (c | kk) != 0;
        if (mode >= 4) { umma_bf16_ss(tmem, dal, dbl, idesc, acc); acc = 1; }
        if (mode >= 3) { umma_bf16_ss(tmem, dal, dbh, idesc, acc); umma_bf16_ss(tmem, dah, dbl, idesc, 1); acc = 1; }
        umma_bf16_ss(tmem, dah, dbh, idesc, acc);
      }
      umma_commit(&bar_done);
    }
    mbar_wait(&bar_done, c & 1);
    tc_fence_after_sync();
    __syncthreads();
  }
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + ((warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64 > N ? 64 : (N <= 128 ? 128 : 256));
}

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main(int argc, char** argv) {
  const int K = argc > 1 ? atoi(argv[1]) : 576;
  const int N = argc > 2 ? atoi(argv[2]) : 64;
  const int mode = argc > 3 ? atoi(argv[3]) : 3;
  const int seed = argc > 4 ? atoi(argv[4]) : 1;
  if (K % 64 || N % 16 || N > 256) { printf("bad K/N\n"); return 1; }
  const int nch = K / 64;
  std::mt19937 rng(seed);
  std::normal_distribution<float> nd(0.0f, 1.0f);
  std::uniform_real_distribution<float> ud(-0.04f, 0.04f);
  std::vector<float> A(128 * K), B(N * K);
  for (auto& v : A) { float y = nd(rng); v = y / (1.0f + expf(-y)); }  // SiLU of a normalised activation
  for (auto& v : B) v = ud(rng);
  const size_t a_half = 8 * 128 * 8, b_half = static_cast<size_t>(8) * N * 8;  // elements
  std::vector<__nv_bfloat16> Ap(nch * 2 * a_half), Bp(nch * 2 * b_half);
  std::vector<double> Ar(128 * K), Br(N * K);  // the hi+lo representations
  for (int r = 0; r < 128; ++r)
    for (int k = 0; k < K; ++k) {
      const float x = A[r * K + k], hi = bf16_round(x), lo = bf16_round(x - hi);
      const int c = k / 64, kc = (k % 64) / 8, j = k % 8;
      Ap[c * 2 * a_half + (kc * 128 + r) * 8 + j] = __float2bfloat16(hi);
      Ap[c * 2 * a_half + a_half + (kc * 128 + r) * 8 + j] = __float2bfloat16(lo);
      Ar[r * K + k] = static_cast<double>(hi) + static_cast<double>(lo);
    }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      const float x = B[n * K + k], hi = bf16_round(x), lo = bf16_round(x - hi);
      const int c = k / 64, kc = (k % 64) / 8, j = k % 8;
      Bp[c * 2 * b_half + (kc * N + n) * 8 + j] = __float2bfloat16(hi);
      Bp[c * 2 * b_half + b_half + (kc * N + n) * 8 + j] = __float2bfloat16(lo);
      Br[n * K + k] = static_cast<double>(hi) + static_cast<double>(lo);
    }
  __nv_bfloat16 *dA, *dB;
  float* dD;
  CK(cudaMalloc(&dA, Ap.size() * 2));
  CK(cudaMalloc(&dB, Bp.size() * 2));
  CK(cudaMalloc(&dD, 128 * N * 4));
  CK(cudaMemcpy(dA, Ap.data(), Ap.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bp.data(), Bp.size() * 2, cudaMemcpyHostToDevice));
  const int smem = 2 * 8 * 128 * 16 + 2 * 8 * N * 16;
  CK(cudaFuncSetAttribute(split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  split_kernel<<<1, 128, smem>>>(dA, dB, dD, nch, N, mode);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> D(128 * N);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxref = 0, e_true = 0, e_repr = 0, e_f32 = 0, s_true = 0, s_ref = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double t = 0, r = 0;
      float f = 0.0f;
      for (int k = 0; k < K; ++k) {
        t += static_cast<double>(A[m * K + k]) * B[n * K + k];
        r += Ar[m * K + k] * Br[n * K + k];
        f = fmaf(A[m * K + k], B[n * K + k], f);
      }
      const double d = D[m * N + n];
      maxref = fmax(maxref, fabs(t));
      e_true = fmax(e_true, fabs(d - t));
      e_repr = fmax(e_repr, fabs(d - r));
      e_f32 = fmax(e_f32, fabs(static_cast<double>(f) - t));
      s_true += (d - t) * (d - t);
      s_ref += t * t;
    }
  printf("SPLIT K=%d N=%d mode=%d: max|ref| %.4g | tcgen05 vs fp64(fp32 operands): max %.3e (rel-to-max %.3e, rms rel %.3e) | "
         "vs fp64(hi+lo operands): rel-to-max %.3e | sequential fp32 FMA vs fp64: rel-to-max %.3e\n",
         K, N, mode, maxref, e_true, e_true / maxref, sqrt(s_true / s_ref), e_repr / maxref, e_f32 / maxref);
  return 0;
}
