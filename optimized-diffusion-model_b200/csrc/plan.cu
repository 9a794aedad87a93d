// plan.cu -- host-side execution of an op plan (the lowered NCSN++ forward) and of the
// predictor-corrector loop (reference sampling.py:292-339), with one iteration captured in a CUDA
// graph and replayed; every per-step scalar is read on the device through the step counter.
#include <vector>
#include <new>
#include "rd_common.h"

namespace rd {
int conv_launch(const rd_op_conv& op, cudaStream_t st);
int attn_launch(const rd_op_attn& op, cudaStream_t st);
int attn_block_launch(const rd_op_attn_block& op, cudaStream_t st);
int temb_launch(const rd_op_temb& op, cudaStream_t st);
int inconv_launch(const rd_op_inconv& op, cudaStream_t st);
int outhead_launch(const rd_op_outhead& op, cudaStream_t st);
}  // namespace rd

struct rd_plan {
  std::vector<rd_op> ops;
};

struct rd_sampler {
  rd_sampler_desc d;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int launches_per_iter = 0;
};

using namespace rd;

static int run_op(const rd_op& op, cudaStream_t st) {
  switch (op.kind) {
    case RD_OP_CONV: return conv_launch(op.u.conv, st);
    case RD_OP_ATTN_CORE: return attn_launch(op.u.attn, st);
    case RD_OP_ATTN_BLOCK: return attn_block_launch(op.u.attn_block, st);
    case RD_OP_TEMB: return temb_launch(op.u.temb, st);
    case RD_OP_IN_CONV: return inconv_launch(op.u.inconv, st);
    case RD_OP_OUT_HEAD: return outhead_launch(op.u.outhead, st);
    default: return fail(RD_E_INVALID, "plan: unknown op kind %d", op.kind);
  }
}

// one predictor-corrector iteration: [corrector: score, norms, apply] x n_corrector_steps, then [score, predictor]
// (sampling.py:327-332 with ReflectedLangevinCorrector.update_fn :221-231 looping n_steps times).
// Noise: tape slots per iteration are [corrector 0 .. n-1, predictor]; in Philox mode corrector move c draws from
// stream (c * kDrawStride + 2*step), the predictor from (2*step + 1) -- disjoint for any grid below kDrawStride/2 points.
constexpr uint32_t kDrawStride = 1u << 22;

static int enqueue_iteration(rd_sampler* s, cudaStream_t st, int* launches) {
  const rd_sampler_desc& d = s->d;
  const size_t n = static_cast<size_t>(d.B) * d.D;
  const size_t draws_per_step = static_cast<size_t>(d.n_corrector_steps) + 1;
  // The time-embedding projections depend on (step, labels) only, not on x: the first forward of an iteration fills
  // them, every later forward of the same iteration skips the temb launch (the step counter advances after the predictor).
  const int has_temb = (rd_plan_size(d.forward) > 1 && d.forward->ops[0].kind == RD_OP_TEMB) ? 1 : 0;
  int count = 0;
  int rc;
  for (int c = 0; c <= d.n_corrector_steps; ++c) {
    const int skip = c > 0 ? has_temb : 0;
    if ((rc = rd_plan_run_range(d.forward, skip, rd_plan_size(d.forward) - skip, st)) != RD_OK) return rc;
    count += rd_plan_size(d.forward) - skip;
    const float* noise = d.noise_tape ? d.noise_tape + static_cast<size_t>(c) * n : nullptr;
    if (c < d.n_corrector_steps) {
      int nblk = 0;
      const uint32_t base = static_cast<uint32_t>(c) * kDrawStride;
      if ((rc = rd_pc_norms(d.score, noise, d.partial, &nblk, d.B, d.D, d.seed, base, d.step_ctr, draws_per_step * n, st)) != RD_OK) return rc;
      if ((rc = rd_pc_corrector_apply(d.x, d.score, noise, d.partial, nblk, d.snr, d.x, nullptr, nullptr, d.B, d.D, d.seed, base,
                                      d.step_ctr, draws_per_step * n, st)) != RD_OK)
        return rc;
      count += 2;
    } else {
      if ((rc = rd_pc_predictor_step(d.x, d.score, noise, d.g_table, d.dt, d.sqrt_dt, d.x, nullptr, d.B, d.D, d.seed, 0, d.step_ctr,
                                     draws_per_step * n, 1, 0, st)) != RD_OK)
        return rc;
      count += 2;  // the update + the step-counter advance it carries
    }
  }
  if (launches) *launches = count;
  return RD_OK;
}

extern "C" {

// ---- unit entry points: one reference layer per call on caller-owned buffers (layer-level parity tests, and the
// `forward` of the drop-in's ResnetBlockDDPMpp / AttnBlockpp modules)
int rd_resblock(const rd_op_temb* temb, const rd_op_conv* shortcut, const rd_op_conv* conv0, const rd_op_conv* conv1, int n_slices,
                void* stream) {
  RD_REQUIRE(conv0 && conv1 && n_slices >= 1, "rd_resblock: conv0 / conv1 are required");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  if (temb && (rc = temb_launch(*temb, st)) != RD_OK) return rc;                    // Dense_0(act(temb)) -> conv0's additive rows
  for (int i = 0; shortcut && i < n_slices; ++i)
    if ((rc = conv_launch(shortcut[i], st)) != RD_OK) return rc;                     // NIN_0 when the channel count changes
  for (int i = 0; i < n_slices; ++i)
    if ((rc = conv_launch(conv0[i], st)) != RD_OK) return rc;                        // GN0 + SiLU + Conv_0 + temb rows
  for (int i = 0; i < n_slices; ++i)
    if ((rc = conv_launch(conv1[i], st)) != RD_OK) return rc;                        // GN1 + SiLU + Conv_1 + skip, * out_scale
  return RD_OK;
}

int rd_attn_block(const rd_op_attn_block* op, void* stream) {
  RD_REQUIRE(op, "rd_attn_block: null op");
  return attn_block_launch(*op, static_cast<cudaStream_t>(stream));
}

int rd_plan_create(rd_plan** out) {
  RD_REQUIRE(out, "rd_plan_create: null out");
  *out = new (std::nothrow) rd_plan();
  RD_REQUIRE(*out, "rd_plan_create: out of memory");
  return RD_OK;
}
int rd_plan_add(rd_plan* p, const rd_op* op) {
  RD_REQUIRE(p && op, "rd_plan_add: null argument");
  if (op->kind == RD_OP_CONV) {  // validate geometry eagerly so planner bugs surface at build time
    int rc = rd_conv_launch_info(&op->u.conv, nullptr, nullptr, nullptr);
    if (rc != RD_OK) return rc;
  }
  p->ops.push_back(*op);
  return RD_OK;
}
int rd_plan_size(const rd_plan* p) { return p ? static_cast<int>(p->ops.size()) : 0; }
int rd_plan_run_range(rd_plan* p, int first, int count, void* stream) {
  RD_REQUIRE(p, "rd_plan_run: null plan");
  RD_REQUIRE(first >= 0 && count >= 0 && first + count <= static_cast<int>(p->ops.size()), "rd_plan_run_range: bad range");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int i = first; i < first + count; ++i) {
    int rc = run_op(p->ops[i], st);
    if (rc != RD_OK) return rc;
  }
  return RD_OK;
}
int rd_plan_run(rd_plan* p, void* stream) { return rd_plan_run_range(p, 0, rd_plan_size(p), stream); }
int rd_plan_destroy(rd_plan* p) {
  delete p;
  return RD_OK;
}

int rd_sampler_create(const rd_sampler_desc* d, rd_sampler** out) {
  RD_REQUIRE(d && out, "rd_sampler_create: null argument");
  RD_REQUIRE(d->forward && d->x && d->score && d->partial && d->g_table && d->step_ctr, "rd_sampler_create: null pointer");
  RD_REQUIRE(d->B > 0 && d->D > 0 && d->n_corrector_steps >= 0, "rd_sampler_create: bad sizes");
  // (any D: Philox quads are indexed over the flat tensor, ragged ends are masked in the kernels)
  rd_sampler* s = new (std::nothrow) rd_sampler();
  RD_REQUIRE(s, "rd_sampler_create: out of memory");
  s->d = *d;
  s->launches_per_iter = (d->n_corrector_steps + 1) * rd_plan_size(d->forward) + 2 * d->n_corrector_steps + 2;
  if (rd_plan_size(d->forward) > 1 && d->forward->ops[0].kind == RD_OP_TEMB) s->launches_per_iter -= d->n_corrector_steps;
  *out = s;
  return RD_OK;
}

int rd_sampler_run(rd_sampler* s, int n_iter, int use_graph, void* stream) {
  RD_REQUIRE(s && n_iter >= 0, "rd_sampler_run: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!use_graph) {
    for (int i = 0; i < n_iter; ++i) {
      int rc = enqueue_iteration(s, st, nullptr);
      if (rc != RD_OK) return rc;
    }
    return RD_OK;
  }
  if (!s->exec) {
    // warm every kernel once outside capture is the caller's job (first-use cudaFuncSetAttribute calls
    // are not capturable); here we only capture.
    cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "sampler: begin capture: %s", cudaGetErrorString(e));
    int rc = enqueue_iteration(s, st, nullptr);
    cudaGraph_t g = nullptr;
    e = cudaStreamEndCapture(st, &g);
    if (rc != RD_OK) {
      if (g) cudaGraphDestroy(g);
      return rc;
    }
    if (e != cudaSuccess) return fail(static_cast<int>(e), "sampler: end capture: %s", cudaGetErrorString(e));
    s->graph = g;
    e = cudaGraphInstantiate(&s->exec, g, 0);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "sampler: graph instantiate: %s", cudaGetErrorString(e));
  }
  for (int i = 0; i < n_iter; ++i) {
    cudaError_t e = cudaGraphLaunch(s->exec, st);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "sampler: graph launch: %s", cudaGetErrorString(e));
  }
  return RD_OK;
}

int rd_sampler_launches_per_iter(const rd_sampler* s) { return s ? s->launches_per_iter : 0; }

int rd_sampler_destroy(rd_sampler* s) {
  if (s) {
    if (s->exec) cudaGraphExecDestroy(s->exec);
    if (s->graph) cudaGraphDestroy(s->graph);
    delete s;
  }
  return RD_OK;
}

}  // extern "C"
