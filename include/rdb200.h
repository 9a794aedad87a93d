/* rdb200.h -- C ABI of the B200-native Reflected-Diffusion sampling hot path.
 *
 * Every entry point takes plain device pointers (caller-owned, e.g. torch allocations),
 * sizes and a cudaStream_t passed as void*; nothing here allocates device memory except
 * rd_plan_* (host-side bookkeeping only) and nothing synchronises the stream, so all
 * calls are CUDA-graph capturable.  Return value: 0 on success, otherwise a cudaError_t
 * (positive) or an RD_E_* code (negative); rd_last_error() gives the message.
 *
 * Reference interfaces replaced (all under /root/reference/Reflected-Diffusion):
 *   rd_reflect_f32            cube.reflect                         cube.py:34-49
 *   rd_inside_f32             cube.inside                          cube.py:17-31
 *   rd_score_hk_f32           cube.score_hk                        cube.py:149-193
 *   rd_philox_normal_f32      torch.randn_like call sites          sampling.py:200,224
 *   rd_pc_norms / rd_pc_corrector_apply
 *                             ReflectedLangevinCorrector.update_fn sampling.py:215-233
 *   rd_pc_predictor_step      ReflectedEulerMaruyamaPredictor.update_fn + RSDE.sde
 *                                                                  sampling.py:198-207, sde_lib.py:93-101
 *   rd_cfg_combine_f32        get_cf_score_fn                      models/utils.py:120-138
 *   rd_plan_* / rd_op         NCSNpp.forward and its layers        models/ncsnpp.py:226-354,
 *                                                                  models/layerspp.py:67-214, models/layers.py:531-540
 *   rd_sampler_*              get_pc_sampler / pc_sampler loop     sampling.py:292-339
 *   rd_checksum_f32           (no counterpart: the reference re-reads nn.Parameters on every call; packed kernel
 *                             weights must notice ema.copy_to / restore, models/ema.py:60-88)
 * "next" rows (callers either side of the path):
 *   rd_perturb_reflect_f32    x_t = cube.reflect(mean + std z)      losses.py:80-82
 *   rd_dsm_reduce_f32         weighted squared error + reduce_op    losses.py:86-92
 *   rd_pf_drift_f32           probability-flow drift * mollifier    sampling.py:345-383, sde_lib.py:93-101
 *   rd_rk45_stage/error_f64   scipy.integrate.solve_ivp(method='RK45') stage / error-norm arithmetic   sampling.py:383
 *   rd_gto_halo_decode_f32    latent -> physical units              ../Benchmark/gto_halo_benchmarking.py:255-328, 335-363
 *   rd_gto_halo_encode_f32    dataset row -> latent                 datasets.py:82-98 (GTOHaloImageDataset.__getitem__)
 */
#ifndef RDB200_H
#define RDB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RD_OK 0
#define RD_E_INVALID (-1)   /* bad argument */
#define RD_E_UNSUPPORTED (-2)
#define RD_E_STATE (-3)

const char* rd_last_error(void);
int rd_version(void);
/* compute capability major*10+minor of the current device, or negative error */
int rd_device_cc(void);

/* ---------------------------------------------------------------- cube.* */
/* out[i] = reflect(x[i]);  n elements, out may alias x. */
int rd_reflect_f32(const float* x, float* out, size_t n, void* stream);
/* ok[b] = all(0 <= x[b,:] <= 1) as uint8;  x is [B, D]. */
int rd_inside_f32(const float* x, uint8_t* ok, size_t B, size_t D, void* stream);
/* Reflected heat-kernel score.  x, x_orig, out: [B, D];  sigma: [B] device floats, or NULL and
 * sigma_scalar is used for every sample.  efs / refls / min_cutoff as in the reference; the sums
 * are pruned to the terms that can change the fp32 result (see csrc/elementwise.cu). */
int rd_score_hk_f32(const float* x, const float* x_orig, const float* sigma, float sigma_scalar, float* out,
                    size_t B, size_t D, int efs, int refls, float min_cutoff, void* stream);

/* Same function through the streaming path: a pre-pass classifies the samples into `workspace` (caller-owned device
 * memory, >= rd_score_hk_workspace_bytes(B), 16-byte aligned), then one barrier-free grid-stride kernel evaluates the
 * elements.  workspace == NULL (or too small) falls back to the single-kernel path above. */
size_t rd_score_hk_workspace_bytes(size_t B);
int rd_score_hk_ws_f32(const float* x, const float* x_orig, const float* sigma, float sigma_scalar, float* out,
                       size_t B, size_t D, int efs, int refls, float min_cutoff, void* workspace, size_t workspace_bytes,
                       void* stream);

/* ---------------------------------------------------------------- noise */
/* out[0..n) = the N(0,1) stream the fused step kernels draw for (seed, draw): element i is lane i%4 of Philox quad i/4
 * (any n; quads are indexed over the flat tensor). */
int rd_philox_normal_f32(float* out, size_t n, uint64_t seed, uint32_t draw, void* stream);

/* ---------------------------------------------------------------- predictor / corrector */
/* Per-step scalars live in device tables indexed by *step_ctr (device int32) so that one captured
 * graph serves every iteration.  Any table pointer may instead be used with step_ctr == NULL,
 * in which case index 0 is read.  A noise pointer is advanced by (*step_ctr) * noise_step_stride
 * elements before use, so a pre-generated noise tape can be replayed from inside a captured graph. */

/* partial[blk*2+{0,1}] = sum over the block's samples of ||grad_b||, ||noise_b||.
 * noise == NULL -> Philox(seed, draw = draw_base + 2*step + 0).  returns number of blocks via *nblk. */
int rd_pc_norms(const float* grad, const float* noise, float* partial, int* nblk, size_t B, size_t D,
                uint64_t seed, uint32_t draw_base, const int32_t* step_ctr, size_t noise_step_stride,
                void* stream);
/* x_mean = reflect(x + eps*grad); x_out = reflect(x + eps*grad + sqrt(2 eps)*noise),
 * eps = 2*(snr*nbar/gbar)^2 with nbar,gbar the batch means rebuilt from `partial`.
 * x_mean_out may be NULL.  stats_out (optional, 3 floats): gbar, nbar, eps. */
int rd_pc_corrector_apply(const float* x, const float* grad, const float* noise, const float* partial,
                          int nblk, float snr, float* x_out, float* x_mean_out, float* stats_out, size_t B,
                          size_t D, uint64_t seed, uint32_t draw_base, const int32_t* step_ctr,
                          size_t noise_step_stride, void* stream);
/* x_mean = reflect(x + g^2*score/N) ; x_out = reflect(x_mean_raw + g*sqrt(1/N)*z).
 * g_table[step] = diffusion coefficient, dt = -1/N (negative), sqrt_dt = sqrt(1/N).
 * z == NULL -> Philox(seed, draw = draw_base + 2*step + 1).  advance_ctr != 0 -> (*step_ctr)++ after use.
 * g_per_sample != 0 -> g_table is [B], one coefficient per sample (the update_fn API with arbitrary t). */
int rd_pc_predictor_step(const float* x, const float* score, const float* z, const float* g_table, float dt,
                         float sqrt_dt, float* x_out, float* x_mean_out, size_t B, size_t D, uint64_t seed,
                         uint32_t draw_base, int32_t* step_ctr, size_t noise_step_stride, int advance_ctr,
                         int g_per_sample, void* stream);
/* out = (1+w)*s_cond - w*s_uncond ; s is [2B, D] (cond first); w: [B] device or NULL -> w_scalar. */
int rd_cfg_combine_f32(const float* s, const float* w, float w_scalar, float* out, size_t B, size_t D,
                       void* stream);

/* ---------------------------------------------------------------- evaluation loss + codec ("next" rows) */
/* out[b,:] = reflect(x0[b,:] + std[b] * z[b,:])  (product and sum rounded separately). */
int rd_perturb_reflect_f32(const float* x0, const float* z, const float* std, float* out, size_t B, size_t D,
                           void* stream);
/* out[b] = scale * sum_d weight[b] * (score[b,d] - target[b,d])^2 ; scale = 1/D if reduce_mean else 0.5. */
int rd_dsm_reduce_f32(const float* score, const float* target, const float* weight, float* out, size_t B, size_t D,
                      int reduce_mean, void* stream);
/* out = (0 - g^2 * score * 0.5) * bump(x); g: [B] device or NULL -> g_scalar;
 * bump(x) = exp((-1/(0.25 - (0.5-x)^2) + 4) / moll) when moll > 0, else x. */
int rd_pf_drift_f32(const float* x, const float* score, const float* g, float g_scalar, float moll, float* out,
                    size_t B, size_t D, void* stream);
/* Device-side Dormand-Prince RK45 building blocks for get_ode_sampler (sampling.py:342-392; the reference round-trips the
 * float64 state through scipy on the host for every right-hand side).  K: [7][n] fp32 stage derivatives.
 *   stage:  y_out (optional) = y + (sum_{j<s} a[j] K_j) * h in fp64 ; x_out = fp32 cast of the same (the network's input)
 *   error:  sumsq_out[0] = sum_i ((sum_j E[j] K_j[i]) * h / (atol + max(|y_i|, |y_new_i|) * rtol))^2, fixed summation order
 *           (E has 7 entries, E[1] is zero in Dormand-Prince and not read); partial: scratch of partial_len doubles */
int rd_rk45_stage_f64(const double* y, const float* K, size_t n, int s, const double* a, double h, double* y_out, float* x_out,
                      void* stream);
int rd_rk45_error_f64(const double* y, const double* y_new, const float* K, size_t n, const double* E, double h, double atol,
                      double rtol, double* partial, int partial_len, double* sumsq_out, void* stream);

/* Constants of the GTO-Halo un-normalisation; spans are (max - min). Row layout of a decoded sample:
 * [halo energy | shooting time, 2 coast times | n_triplets x (alpha, beta, r) | fuel mass, halo period, manifold length]. */
typedef struct rd_gto_halo_codec {
  float data_mean, data_std;
  float shooting_time_min, shooting_time_span;
  float coast_time_min, coast_time_span;
  float halo_energy_min, halo_energy_span;
  float fuel_mass_min, fuel_mass_span;
  float manifold_length_min, manifold_length_span;
  float thrust;
  int32_t n_triplets;
} rd_gto_halo_codec;
/* latents: [n, row_stride] fp32 (row 0 = class label, then the model variables); out: [n, 7 + 3 n_triplets]. */
int rd_gto_halo_decode_f32(const float* latents, float* out, size_t n, size_t row_stride,
                           const rd_gto_halo_codec* codec, void* stream);

/* dataset rows -> latents (datasets.py:82-98): raw [n, n_in] fp32 is zero-padded to n_latent values per sample and
 * every entry z-scored ((v - mean) / std, padding included); labels (optional) [n] receives the un-normalised first value. */
int rd_gto_halo_encode_f32(const float* raw, float* latents, float* labels, size_t n, size_t n_in, size_t n_latent,
                           float data_mean, float data_std, void* stream);

/* ---------------------------------------------------------------- NCSN++ forward as an op plan */
/* The host (python, mirroring NCSNpp.__init__) lowers the network to a flat list of ops over
 * device buffers; the library owns only the kernels.  Activations are NHWC bf16. */

enum rd_op_kind {
  RD_OP_CONV = 1,      /* implicit-GEMM conv / 1x1 on tcgen05 with fused GN+SiLU prologue and epilogue */
  RD_OP_ATTN_CORE = 2, /* softmax(q k^T / sqrt(C)) v per sample */
  RD_OP_TEMB = 3,      /* per-sample Dense_0(SiLU(temb)) for every ResBlock */
  RD_OP_IN_CONV = 4,   /* 3x3 conv C_in(=channels) -> nf from the fp32 state x */
  RD_OP_OUT_HEAD = 5,  /* GN + SiLU + 3x3 conv nf -> channels, CFG combine, fp32 score */
  RD_OP_ATTN_BLOCK = 6 /* whole AttnBlockpp fused: GN, q/k/v, softmax(qk^T)v, output projection, skip */
};

/* Precision of the network plan (every op of a plan uses the same one):
 *   RD_PREC_BF16   bf16 NHWC activations, bf16 MMA operands, fp32 accumulation / statistics / epilogues
 *   RD_PREC_F32X3  "fp32-class": fp32 NHWC activations, operands split into bf16 hi + lo, three tcgen05 MMAs per k-step
 *                  (the reference is fp32 end to end, layers.py:103-109) */
#define RD_PREC_BF16 0
#define RD_PREC_F32X3 1

typedef struct rd_conv_src {
  const void* ptr; /* NHWC [B2, Hs, Ws, C], bf16 (fp32 when precision == RD_PREC_F32X3) */
  int32_t C;       /* channels of this source (multiple of 8; the total over sources a multiple of 64) */
  int32_t Hs, Ws;  /* stored spatial size; gathered to (H_in, W_in) by nearest mapping */
} rd_conv_src;

typedef struct rd_op_conv {
  rd_conv_src src[2];
  int32_t nsrc;
  int32_t H_in, W_in;         /* logical input image (after nearest gather)            */
  int32_t pad;                /* 1: 3x3 pad 1 (output H_in x W_in); 0: pad (0,1,0,1)    */
  int32_t stride;             /* 1 or 2 (2 only with pad == 0: Downsample)              */
  int32_t H_out, W_out;
  int32_t ntaps;              /* 9 (3x3 conv) or 1 (1x1: NIN / attention projections)   */
  int32_t C_out;              /* N of the GEMM (this launch): multiple of 32, <= 256    */
  int32_t gn_groups;          /* 0: no GroupNorm/SiLU prologue                          */
  int32_t gn_silu;            /* 1: SiLU after GN (ResBlock), 0: GN only (attention)    */
  float gn_eps;
  const float* gn_gamma;      /* [C_in_total] */
  const float* gn_beta;
  const void* w;              /* bf16 packed [C_in/64][ntaps][8][C_out][8] (rdb200/pack.py); F32X3: [C_in/64][ntaps][hi|lo][8][C_out][8] */
  const float* bias;          /* [C_out] (sum of all fused biases) */
  const float* tproj;         /* [B2, tproj_stride] per-sample additive term or NULL */
  int32_t tproj_stride, tproj_off;
  int32_t tproj_wrap;         /* >0: samples with index > tproj_wrap read row tproj_wrap (shared unconditional CFG row) */
  const void* residual;       /* NHWC [B2,H_out,W_out,out_stride] identity skip or NULL (same dtype as out) */
  float out_scale;            /* 1/sqrt(2) when skip_rescale, else 1 */
  void* out;                  /* NHWC [B2,H_out,W_out,out_stride], bf16 or fp32 by `precision` */
  int32_t B2;                 /* samples (2B under CFG) */
  int32_t samples_per_cta;    /* 0: library picks the tile geometry; >0: planner override */
  int32_t precision;          /* RD_PREC_* */
  int32_t out_stride;         /* 0 = C_out; otherwise channels per pixel of out / residual: a layer wider than one launch
                                 (C_out > 256, or a tile geometry that does not fit) is issued as channel slices, each
                                 with out / residual / bias / tproj_off / w advanced to its first channel */
  /* Fused 1x1 shortcut (ResnetBlockDDPMpp's NIN_0 folded into Conv_1, layerspp.py:211-214): when sc_nsrc > 0 the RAW
   * pixels of sc_src (gathered to H_in x W_in like src) are multiplied by the 1x1 filter slabs that follow the 3x3 slabs
   * in `w` ([sum(sc C)/64][1][8][C_out][8]) and accumulated into the same output tile; `bias` then carries both biases
   * and `residual` must be NULL.  bf16 plan, stride 1, GroupNorm groups of 4 or 8 channels. */
  rd_conv_src sc_src[2];
  int32_t sc_nsrc;
  int32_t _pad0;
} rd_op_conv;

typedef struct rd_op_attn {
  const void* qkv; /* [B2, T, 3C] (q | k | v), bf16 or fp32 by `precision` */
  void* out;       /* [B2, T, C] */
  int32_t B2, T, C;
  int32_t precision; /* RD_PREC_* */
} rd_op_attn;

typedef struct rd_op_attn_block {
  const void* x;        /* bf16 NHWC [B2, T, C] */
  void* out;            /* bf16 NHWC [B2, T, C] = (x + attn(x)) * out_scale */
  const void* wqkv_t;   /* bf16 [3C][72]: rows = output channel of q|k|v, cols = input channel (padded to 72) */
  const void* wproj_t;  /* bf16 [C][72]: NIN_3 transposed the same way */
  const float* bqkv;    /* [3C] */
  const float* bproj;   /* [C] */
  const float* gamma;   /* GroupNorm_0 */
  const float* beta;
  int32_t B2, T, C, groups;
  float eps, out_scale;
} rd_op_attn_block;

typedef struct rd_op_temb {
  const float* time_table; /* [n_steps, temb_dim] = time_mlp(fourier(log sigma_i)) + label_emb.bias */
  const float* label_w;    /* [temb_dim, num_classes] label_emb.weight */
  const float* labels;     /* [B2, num_classes] */
  const void* dense_w;     /* fp32 [n_out_total, temb_dim]: all Dense_0 weights stacked */
  const float* dense_b;    /* [n_out_total] */
  float* out;              /* [B2, n_out_total] */
  const int32_t* step_ctr; /* device step index: every sample uses row *step_ctr (sampler) ... */
  const int32_t* row_idx;  /* ... unless row_idx != NULL: per-sample table row [B2] (generic forward) */
  int32_t B2, temb_dim, num_classes, n_out_total;
} rd_op_temb;

typedef struct rd_op_inconv {
  const float* x;    /* fp32 [B, C_in, H, W] (NCHW state, C_in small) ; sample b2 reads x[b2 % B] */
  const float* w;    /* fp32 [C_out, C_in, 3, 3] */
  const float* bias; /* [C_out] */
  void* out;         /* NHWC [B2,H,W,C_out], bf16 or fp32 by `precision` */
  int32_t B, B2, C_in, C_out, H, W;
  int32_t precision; /* RD_PREC_* */
} rd_op_inconv;

typedef struct rd_op_outhead {
  const void* h;       /* NHWC [B2,H,W,C], bf16 or fp32 by `precision` */
  const float* gamma;  /* out_norm */
  const float* beta;
  const float* w;      /* fp32 [C_img, C, 3, 3] out_conv.weight */
  const float* bias;   /* [C_img] */
  const float* cfg_w;  /* [B] guidance weights or NULL */
  float cfg_w_scalar;
  float* score;        /* fp32 [B, C_img, H, W] guided score (or [B2,...] raw if cfg == 0) */
  int32_t B, B2, C, C_img, H, W, groups, cfg;
  float eps;
  int32_t precision;   /* RD_PREC_* */
  const float* sigma_table; /* NULL, or (model.scale_by_sigma, ncsnpp.py:350-351) the score is divided by sigma: */
  const int32_t* step_ctr;  /*   sigma_table[*step_ctr] when step_ctr != NULL (sampler), else sigma_table[b2] per sample */
} rd_op_outhead;

typedef struct rd_op {
  int32_t kind;
  int32_t _pad;
  union {
    rd_op_conv conv;
    rd_op_attn attn;
    rd_op_attn_block attn_block;
    rd_op_temb temb;
    rd_op_inconv inconv;
    rd_op_outhead outhead;
  } u;
} rd_op;

/* Unit entry points -- one reference layer per call, on caller-owned NHWC buffers:
 *   rd_resblock   ResnetBlockDDPMpp.forward (layerspp.py:198-214): [temb: Dense_0(act(temb)) rows] -> [shortcut: NIN_0] ->
 *                 conv0 (GN0 + SiLU + Conv_0 + temb rows) -> conv1 (GN1 + SiLU + Conv_1 + skip, * out_scale); temb /
 *                 shortcut may be NULL; shortcut / conv0 / conv1 are arrays of n_slices channel-slice launches each
 *                 (1 for C_out <= 128, see rd_op_conv.out_stride)
 *   rd_attn_block AttnBlockpp.forward (layerspp.py:80-96), the fused kernel (C = 64, T <= 128, bf16 plan) */
int rd_resblock(const rd_op_temb* temb, const rd_op_conv* shortcut, const rd_op_conv* conv0, const rd_op_conv* conv1, int n_slices,
                void* stream);
int rd_attn_block(const rd_op_attn_block* op, void* stream);

typedef struct rd_plan rd_plan;
int rd_plan_create(rd_plan** out);
int rd_plan_add(rd_plan* p, const rd_op* op);
int rd_plan_size(const rd_plan* p);
int rd_plan_run(rd_plan* p, void* stream);
/* run ops [first, first+count) -- unit-test entry point for single layers */
int rd_plan_run_range(rd_plan* p, int first, int count, void* stream);
int rd_plan_destroy(rd_plan* p);
/* shared-memory bytes / CTAs a conv op will launch with (planner feedback, also validates the op) */
int rd_conv_launch_info(const rd_op_conv* op, int* smem_bytes, int* grid, int* rows_alloc);
/* the tile geometry picked for a conv op: geom[12] = {samples per group, 128-row tiles, staged rows, groups, operand stages,
 * filter resident?, filter ring slabs, accumulator buffers, register-cached transform slots (0 = streaming), shared-memory
 * bytes, grid, TMEM columns} (profiling / planner feedback) */
int rd_conv_geometry(const rd_op_conv* op, int* geom);

/* ---------------------------------------------------------------- parameter change detection */
/* 64-bit content checksum of a set of fp32 device tensors (value AND position of every element enter the sum), so the
 * host can tell with one launch + 8 bytes of D2H whether parameters were rewritten in place (EMA copy_to / restore,
 * load_state_dict; reference models/ema.py:60-88).  segs: device [n_segs][3] int64 = (address, index of the first
 * element in the concatenation of all tensors, element count); out: device uint64. */
int rd_checksum_f32(const int64_t* segs, int n_segs, uint64_t* out, void* stream);

/* ---------------------------------------------------------------- whole sampler (sampling.py:292-339) */
typedef struct rd_sampler_desc {
  rd_plan* forward;        /* guided score: x -> score (reads x, step_ctr via the ops' pointers) */
  float* x;                /* [B, D] state, updated in place */
  float* score;            /* [B, D] written by `forward` */
  float* partial;          /* scratch >= 2*ceil(B/8) floats */
  const float* g_table;    /* [N] */
  int32_t* step_ctr;       /* device */
  const float* noise_tape; /* NULL (Philox) or [(N-1)*2 or (N-1), B, D]: corrector noise then predictor z per step */
  uint64_t seed;
  float snr, dt, sqrt_dt;
  int32_t B, D, n_corrector_steps; /* 0 -> corrector 'none' */
} rd_sampler_desc;
typedef struct rd_sampler rd_sampler;
int rd_sampler_create(const rd_sampler_desc* d, rd_sampler** out);
/* enqueue iterations [*step_ctr, *step_ctr + n_iter): corrector(s) then predictor, each with one
 * guided-score evaluation.  use_graph != 0 captures one iteration into a CUDA graph on first use. */
int rd_sampler_run(rd_sampler* s, int n_iter, int use_graph, void* stream);
/* number of kernel launches one iteration issues (for bench.py's gpu_launches) */
int rd_sampler_launches_per_iter(const rd_sampler* s);
int rd_sampler_destroy(rd_sampler* s);

#ifdef __cplusplus
}
#endif
#endif /* RDB200_H */
