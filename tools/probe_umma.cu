// probe_umma.cu -- stand-alone check of the tcgen05 operand conventions the conv kernel relies on:
//   * K-major SWIZZLE_NONE descriptor with SBO=128 B, LBO = R*16 B (R = rows in the staged image, odd or even)
//   * start-address row shifts (descriptor advanced by s*16 B == matrix shifted by s rows)
//   * weights delivered by a 1-D bulk TMA copy + mbarrier transaction count
//   * accumulator read-back with tcgen05.ld.32x32b
// usage: probe_umma R N shift use_bulk     (prints "PROBE ... mismatches=<n>")
#include <cstdio>
#include <cstdlib>
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

constexpr int KC = 8;  // K = 64 -> 8 chunks of 8 bf16 (16 B)

__global__ void __launch_bounds__(128) probe_kernel(const __nv_bfloat16* __restrict__ Ag,  // [KC][R][8]
                                                    const __nv_bfloat16* __restrict__ Bg,  // [KC][N][8]
                                                    float* __restrict__ D,                 // [2][128][N]
                                                    int R, int N, int shift, int use_bulk) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar_w, bar_acc;
  __shared__ uint32_t tmem_slot;
  __nv_bfloat16* As = reinterpret_cast<__nv_bfloat16*>(smem);
  const int a_bytes = KC * R * 16;
  __nv_bfloat16* Bs = reinterpret_cast<__nv_bfloat16*>(smem + ((a_bytes + 127) / 128) * 128);
  const int b_bytes = KC * N * 16;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    mbar_init(&bar_w, 1);
    mbar_init(&bar_acc, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  // stage A with ordinary stores
  for (int i = tid; i < KC * R; i += blockDim.x)
    reinterpret_cast<uint4*>(As)[i] = reinterpret_cast<const uint4*>(Ag)[i];
  if (!use_bulk)
    for (int i = tid; i < KC * N; i += blockDim.x)
      reinterpret_cast<uint4*>(Bs)[i] = reinterpret_cast<const uint4*>(Bg)[i];
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp == 1) {
    if (elect_one()) {
      if (use_bulk) {
        mbar_arrive_expect_tx(&bar_w, b_bytes);
        bulk_g2s(Bs, Bg, b_bytes, &bar_w);
        mbar_wait(&bar_w, 0);
      }
      tc_fence_after_sync();
      const uint32_t idesc = umma_idesc_bf16(128, N);
      for (int tile = 0; tile < 2; ++tile) {
        for (int kk = 0; kk < 4; ++kk) {
          uint32_t a_addr = smem_u32(As) + ((kk * 2) * R + tile * 128 + shift) * 16;
          uint32_t b_addr = smem_u32(Bs) + ((kk * 2) * N) * 16;
          uint64_t da = umma_desc_kmajor(a_addr, R * 16, 128);
          uint64_t db = umma_desc_kmajor(b_addr, N * 16, 128);
          umma_bf16_ss(tmem + tile * N, da, db, idesc, kk > 0);
        }
      }
      umma_commit(&bar_acc);
    }
    __syncwarp();
  }
  mbar_wait(&bar_acc, 0);
  tc_fence_after_sync();
  for (int tile = 0; tile < 2; ++tile) {
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((warp * 32) << 16) + tile * N + c0, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) D[(tile * 128 + tid) * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main(int argc, char** argv) {
  int R = argc > 1 ? atoi(argv[1]) : 288;
  int N = argc > 2 ? atoi(argv[2]) : 64;
  int shift = argc > 3 ? atoi(argv[3]) : 0;
  int use_bulk = argc > 4 ? atoi(argv[4]) : 0;
  if (R < 256 + shift) { printf("R too small\n"); return 1; }
  std::vector<int> A(R * 64), B(N * 64);
  for (int r = 0; r < R; ++r)
    for (int k = 0; k < 64; ++k) A[r * 64 + k] = ((r * 7 + k * 3) % 13) - 6;
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < 64; ++k) B[n * 64 + k] = ((n * 5 + k) % 11) - 5;
  std::vector<__nv_bfloat16> Ap(KC * R * 8), Bp(KC * N * 8);
  for (int r = 0; r < R; ++r)
    for (int k = 0; k < 64; ++k) Ap[((k / 8) * R + r) * 8 + (k % 8)] = __float2bfloat16((float)A[r * 64 + k]);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < 64; ++k) Bp[((k / 8) * N + n) * 8 + (k % 8)] = __float2bfloat16((float)B[n * 64 + k]);
  __nv_bfloat16 *dA, *dB;
  float* dD;
  CK(cudaMalloc(&dA, Ap.size() * 2));
  CK(cudaMalloc(&dB, Bp.size() * 2));
  CK(cudaMalloc(&dD, 2 * 128 * N * 4));
  CK(cudaMemcpy(dA, Ap.data(), Ap.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bp.data(), Bp.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, 2 * 128 * N * 4));
  int smem = ((KC * R * 16 + 127) / 128) * 128 + KC * N * 16;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_kernel<<<1, 128, smem>>>(dA, dB, dD, R, N, shift, use_bulk);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> D(2 * 128 * N);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < N; ++n) {
      int acc = 0;
      for (int k = 0; k < 64; ++k) acc += A[(m + shift) * 64 + k] * B[n * 64 + k];
      if (D[m * N + n] != (float)acc) {
        if (bad < 5) printf("  mismatch m=%d n=%d got %f want %d\n", m, n, D[m * N + n], acc);
        ++bad;
      }
    }
  printf("PROBE R=%d N=%d shift=%d bulk=%d mismatches=%d\n", R, N, shift, use_bulk, bad);
  return bad ? 1 : 0;
}
