// next_rows.cu -- the callers either side of the sampling hot path:
//   * evaluation DSM loss pieces: perturbation x_t = reflect(x_0 + sigma z) and the weighted squared
//     error against the heat-kernel score (reference Reflected-Diffusion/losses.py:77-92);
//   * latent -> physical-unit codec of the GTO-Halo benchmark
//     (reference Benchmark/gto_halo_benchmarking.py:255-328 and :335-363).
// All are HBM-bound streaming kernels (12-16 B per element), one warp per sample where a
// per-sample reduction is needed (fixed shuffle order => deterministic).
#include "rd_common.h"
#include "rd_math.cuh"

namespace rd {

// x_t = reflect(mean + std[b] * z): product and sum rounded separately like the eager reference.
__global__ void __launch_bounds__(256) perturb_reflect_kernel(const float* __restrict__ x0, const float* __restrict__ z,
                                                              const float* __restrict__ std, float* __restrict__ out,
                                                              size_t B, int D) {
  const size_t n = B * static_cast<size_t>(D);
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) {
    const float s = std[i / D];
    out[i] = reflect1(__fadd_rn(x0[i], __fmul_rn(s, z[i])));
  }
}

// losses[b] = scale * sum_d w[b] * (score - target)^2   (losses.py:86-92; scale = 1/D or 0.5)
__global__ void __launch_bounds__(256) dsm_reduce_kernel(const float* __restrict__ score, const float* __restrict__ target,
                                                         const float* __restrict__ w, float* __restrict__ out, size_t B,
                                                         int D, float scale) {
  const size_t warp = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B) return;
  const float wb = w[warp];
  float acc = 0.0f;
  for (int j = lane; j < D; j += 32) {
    const float d = __fsub_rn(score[warp * D + j], target[warp * D + j]);
    acc = __fadd_rn(acc, __fmul_rn(wb, __fmul_rn(d, d)));
  }
  acc = warp_sum(acc);
  if (lane == 0) out[warp] = acc * scale;
}

// Probability-flow drift of the reflected VE SDE times the boundary mollifier (sampling.py:345-383,
// sde_lib.py:93-101 with probability_flow=True): out = (0 - g^2 * score * 0.5) * bump(x),
// bump(x) = exp((-1 / (0.25 - (0.5 - x)^2) + 4) / moll) for moll > 0, else x itself.
__global__ void __launch_bounds__(256) pf_drift_kernel(const float* __restrict__ x, const float* __restrict__ score,
                                                       const float* __restrict__ g, float g_scalar, float moll,
                                                       float* __restrict__ out, size_t B, int D) {
  const size_t n = B * static_cast<size_t>(D);
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) {
    const float gb = g ? g[i / D] : g_scalar;
    const float drift = __fsub_rn(0.0f, __fmul_rn(__fmul_rn(__fmul_rn(gb, gb), score[i]), 0.5f));
    const float xv = x[i];
    float b = xv;
    if (moll > 0.0f) {
      const float c = __fsub_rn(0.5f, xv);
      const float q = __fsub_rn(0.25f, __fmul_rn(c, c));
      b = expf(__fdiv_rn(__fadd_rn(__fdiv_rn(-1.0f, q), 4.0f), moll));
    }
    out[i] = __fmul_rn(drift, b);
  }
}

// One thread per sample: 1 label + (3 times, n_triplets controls, 3 tail values).
__global__ void __launch_bounds__(128) gto_halo_decode_kernel(const float* __restrict__ lat, float* __restrict__ out,
                                                              size_t n, int row_stride, rd_gto_halo_codec c) {
  const size_t s = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const float* in = lat + s * row_stride;
  const int n_out = 1 + 3 + 3 * c.n_triplets + 3;
  float* o = out + s * n_out;
  const float TWO_PI = 6.283185307179586f;
  auto unnorm = [&](float v) { return __fadd_rn(__fmul_rn(v, c.data_std), c.data_mean); };  // latents -> [0,1] units
  o[0] = __fadd_rn(__fmul_rn(in[0], c.halo_energy_span), c.halo_energy_min);
  o[1] = __fadd_rn(__fmul_rn(unnorm(in[1]), c.shooting_time_span), c.shooting_time_min);
  o[2] = __fadd_rn(__fmul_rn(unnorm(in[2]), c.coast_time_span), c.coast_time_min);
  o[3] = __fadd_rn(__fmul_rn(unnorm(in[3]), c.coast_time_span), c.coast_time_min);
  for (int k = 0; k < c.n_triplets; ++k) {
    float u3[3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
      u3[j] = __fsub_rn(__fmul_rn(__fmul_rn(unnorm(in[4 + 3 * k + j]), 2.0f), c.thrust), c.thrust);
    float u = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(u3[0], u3[0]), __fmul_rn(u3[1], u3[1])), __fmul_rn(u3[2], u3[2])));
    float theta = (u != 0.0f) ? asinf(__fdiv_rn(u3[2], u)) : 0.0f;
    float alpha = atan2f(u3[1], u3[0]);
    if (!(alpha >= 0.0f)) alpha = __fadd_rn(TWO_PI, alpha);
    if (!(theta >= 0.0f)) theta = __fadd_rn(TWO_PI, theta);
    if (u > 1.0f) u = 1.0f;
    o[4 + 3 * k + 0] = alpha;
    o[4 + 3 * k + 1] = theta;
    o[4 + 3 * k + 2] = u;
  }
  const int t = 4 + 3 * c.n_triplets;
  o[t + 0] = __fadd_rn(__fmul_rn(unnorm(in[t + 0]), c.fuel_mass_span), c.fuel_mass_min);
  o[t + 1] = unnorm(in[t + 1]);  // halo period stays normalised (gto_halo_benchmarking.py:319)
  o[t + 2] = __fadd_rn(__fmul_rn(unnorm(in[t + 2]), c.manifold_length_span), c.manifold_length_min);
}

// dataset row -> latent (reference Reflected-Diffusion/datasets.py:82-98, GTOHaloImageDataset.__getitem__): the row's
// n_in values are zero-padded to the latent size, then EVERY entry (padding included) is z-scored, fp32 like numpy's.
__global__ void __launch_bounds__(256) gto_halo_encode_kernel(const float* __restrict__ raw, float* __restrict__ lat,
                                                              float* __restrict__ label, size_t n, int n_in, int n_lat,
                                                              float mean, float std) {
  const size_t total = n * static_cast<size_t>(n_lat);
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < total; i += stride) {
    const size_t s = i / n_lat;
    const int j = static_cast<int>(i - s * n_lat);
    const float v = j < n_in ? raw[s * n_in + j] : 0.0f;
    lat[i] = __fdiv_rn(__fsub_rn(v, mean), std);
    if (j == 0 && label) label[s] = v;  // class label = the un-normalised first value (datasets.py:92)
  }
}

}  // namespace rd

using namespace rd;

extern "C" {

int rd_gto_halo_encode_f32(const float* raw, float* latents, float* labels, size_t n, size_t n_in, size_t n_latent,
                           float data_mean, float data_std, void* stream) {
  if (n == 0) return RD_OK;
  RD_REQUIRE(raw && latents, "rd_gto_halo_encode_f32: null pointer");
  RD_REQUIRE(n_in >= 1 && n_in <= n_latent && n_latent <= (1u << 20), "rd_gto_halo_encode_f32: %zu values do not fit a latent of %zu",
             n_in, n_latent);
  size_t blocks = (n * n_latent + 255) / 256;
  if (blocks > static_cast<size_t>(kNumSMs) * 16) blocks = static_cast<size_t>(kNumSMs) * 16;
  gto_halo_encode_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      raw, latents, labels, n, static_cast<int>(n_in), static_cast<int>(n_latent), data_mean, data_std);
  return check_launch("gto_halo_encode_kernel");
}

int rd_perturb_reflect_f32(const float* x0, const float* z, const float* std, float* out, size_t B, size_t D,
                           void* stream) {
  if (B == 0 || D == 0) return RD_OK;
  RD_REQUIRE(x0 && z && std && out, "rd_perturb_reflect_f32: null pointer");
  RD_REQUIRE(D <= (1u << 24), "rd_perturb_reflect_f32: D too large");
  size_t blocks = (B * D + 255) / 256;
  if (blocks > static_cast<size_t>(kNumSMs) * 16) blocks = static_cast<size_t>(kNumSMs) * 16;
  perturb_reflect_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x0, z, std, out, B, static_cast<int>(D));
  return check_launch("perturb_reflect_kernel");
}

int rd_dsm_reduce_f32(const float* score, const float* target, const float* weight, float* out, size_t B, size_t D,
                      int reduce_mean, void* stream) {
  if (B == 0) return RD_OK;
  RD_REQUIRE(score && target && weight && out, "rd_dsm_reduce_f32: null pointer");
  RD_REQUIRE(D >= 1 && D <= (1u << 24), "rd_dsm_reduce_f32: D out of range");
  const size_t blocks = (B + 7) / 8;
  RD_REQUIRE(blocks <= 0x7fffffffu, "rd_dsm_reduce_f32: batch too large");
  const float scale = reduce_mean ? 1.0f / static_cast<float>(D) : 0.5f;
  dsm_reduce_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      score, target, weight, out, B, static_cast<int>(D), scale);
  return check_launch("dsm_reduce_kernel");
}

int rd_pf_drift_f32(const float* x, const float* score, const float* g, float g_scalar, float moll, float* out,
                    size_t B, size_t D, void* stream) {
  if (B == 0 || D == 0) return RD_OK;
  RD_REQUIRE(x && score && out, "rd_pf_drift_f32: null pointer");
  RD_REQUIRE(D <= (1u << 24), "rd_pf_drift_f32: D too large");
  size_t blocks = (B * D + 255) / 256;
  if (blocks > static_cast<size_t>(kNumSMs) * 16) blocks = static_cast<size_t>(kNumSMs) * 16;
  pf_drift_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, score, g, g_scalar, moll, out, B, static_cast<int>(D));
  return check_launch("pf_drift_kernel");
}

int rd_gto_halo_decode_f32(const float* latents, float* out, size_t n, size_t row_stride, const rd_gto_halo_codec* codec,
                           void* stream) {
  if (n == 0) return RD_OK;
  RD_REQUIRE(latents && out && codec, "rd_gto_halo_decode_f32: null pointer");
  RD_REQUIRE(codec->n_triplets >= 0 && codec->n_triplets <= 4096, "rd_gto_halo_decode_f32: bad n_triplets");
  const size_t need = 1 + 3 + 3 * static_cast<size_t>(codec->n_triplets) + 3;
  RD_REQUIRE(row_stride >= need && row_stride <= 0x7fffffffu,
             "rd_gto_halo_decode_f32: rows of %zu values cannot hold %zu variables", row_stride, need);
  const size_t blocks = (n + 127) / 128;
  RD_REQUIRE(blocks <= 0x7fffffffu, "rd_gto_halo_decode_f32: too many samples");
  gto_halo_decode_kernel<<<static_cast<unsigned>(blocks), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      latents, out, n, static_cast<int>(row_stride), *codec);
  return check_launch("gto_halo_decode_kernel");
}

}  // extern "C"
