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

// ---------------------------------------------------------------------------------------------------------------------
// Device-side Dormand-Prince RK45 for the probability-flow ODE sampler (reference sampling.py:342-392 hands the state to
// scipy.integrate.solve_ivp, which keeps it as a float64 numpy vector on the host and round-trips host <-> GPU for every
// right-hand side).  Here the float64 state and the seven stage derivatives stay on the GPU; per step the host reads ONE
// scalar (the error norm) to decide accept / reject, exactly scipy's control law.
//   stage:  y_s = y + (sum_j a[j] K_j) * h  (fp64, scipy's order: dot first, then * h, then + y), also cast to fp32 for the network
__global__ void __launch_bounds__(256) rk45_stage_kernel(const double* __restrict__ y, const float* __restrict__ K, size_t n, int s,
                                                         double a0, double a1, double a2, double a3, double a4, double a5, double h,
                                                         double* __restrict__ y_out, float* __restrict__ x_out) {
  const double a[6] = {a0, a1, a2, a3, a4, a5};
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) {
    double acc = 0.0;
    for (int j = 0; j < s; ++j) acc += static_cast<double>(K[static_cast<size_t>(j) * n + i]) * a[j];
    const double v = y[i] + acc * h;
    if (y_out) y_out[i] = v;
    x_out[i] = static_cast<float>(v);
  }
}

//   error:  err = (sum_j E[j] K_j) * h ; scale = atol + max(|y|, |y_new|) * rtol ; partial[blk] = sum (err / scale)^2
__global__ void __launch_bounds__(256) rk45_error_kernel(const double* __restrict__ y, const double* __restrict__ y_new,
                                                         const float* __restrict__ K, size_t n, double e0, double e2, double e3,
                                                         double e4, double e5, double e6, double h, double atol, double rtol,
                                                         double* __restrict__ partial) {
  __shared__ double red[256];
  double acc = 0.0;
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) {
    const double err = (e0 * K[i] + e2 * K[2 * n + i] + e3 * K[3 * n + i] + e4 * K[4 * n + i] + e5 * K[5 * n + i] + e6 * K[6 * n + i]) * h;
    const double sc = atol + fmax(fabs(y[i]), fabs(y_new[i])) * rtol;
    const double r = err / sc;
    acc += r * r;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {  // fixed-order tree: run-to-run deterministic
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

__global__ void rk45_sum_kernel(const double* __restrict__ partial, int nblk, double* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < nblk; ++i) s += partial[i];
    out[0] = s;
  }
}

}  // namespace rd

using namespace rd;

extern "C" {

int rd_rk45_stage_f64(const double* y, const float* K, size_t n, int s, const double* a, double h, double* y_out, float* x_out,
                      void* stream) {
  if (n == 0) return RD_OK;
  RD_REQUIRE(y && x_out && (s == 0 || (K && a)) && s >= 0 && s <= 6, "rd_rk45_stage_f64: bad arguments");
  double c[6] = {0, 0, 0, 0, 0, 0};
  for (int j = 0; j < s; ++j) c[j] = a[j];
  size_t blocks = (n + 255) / 256;
  if (blocks > static_cast<size_t>(kNumSMs) * 8) blocks = static_cast<size_t>(kNumSMs) * 8;
  rk45_stage_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(y, K, n, s, c[0], c[1], c[2], c[3], c[4],
                                                                                               c[5], h, y_out, x_out);
  return check_launch("rk45_stage_kernel");
}

int rd_rk45_error_f64(const double* y, const double* y_new, const float* K, size_t n, const double* E, double h, double atol,
                      double rtol, double* partial, int partial_len, double* sumsq_out, void* stream) {
  RD_REQUIRE(y && y_new && K && E && partial && sumsq_out && n > 0 && partial_len >= 1, "rd_rk45_error_f64: bad arguments");
  size_t blocks = (n + 255) / 256;
  if (blocks > static_cast<size_t>(partial_len)) blocks = partial_len;
  if (blocks > static_cast<size_t>(kNumSMs) * 8) blocks = static_cast<size_t>(kNumSMs) * 8;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  rk45_error_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(y, y_new, K, n, E[0], E[2], E[3], E[4], E[5], E[6], h, atol, rtol, partial);
  int rc = check_launch("rk45_error_kernel");
  if (rc != RD_OK) return rc;
  rk45_sum_kernel<<<1, 32, 0, st>>>(partial, static_cast<int>(blocks), sumsq_out);
  return check_launch("rk45_sum_kernel");
}

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
