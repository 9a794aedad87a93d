"""Evaluation loss of the reflected SDE -- drop-in for the inference half of the reference's `losses`
module (/root/reference/Reflected-Diffusion/losses.py:51-160).  `get_sde_loss_fn(sde, train=False)`
and `get_step_fn(sde, train=False)` keep the reference's signatures and value semantics; the
perturbation, the heat-kernel score and the weighted reduction run as CUDA kernels
(csrc/next_rows.cu, csrc/elementwise.cu) and the network runs on the B200 forward plan.
Training (`train=True`, optimizers) needs a backward pass, which this inference path does not
have: those entry points raise NotImplementedError instead of silently running something else.
"""
import torch

import cube
from models import utils as mutils
from rdb200 import ops as _ops


def get_optimizer(config, params):
    raise NotImplementedError('training is outside the B200 inference path (SURVEY.md section 8)')


def optimization_manager(config):
    raise NotImplementedError('training is outside the B200 inference path (SURVEY.md section 8)')


def get_sde_loss_fn(sde, train, reduce_mean=True, likelihood_weighting=True, eps=1e-5):
    """Returns `loss_fn(model, batch, class_labels=None)` (losses.py:51-107, eval branch).

    Two keyword-only additions make runs replayable: `rd_t` ([B] uniform draws in [0,1)) and `rd_z`
    (noise shaped like `batch`) replace the two RNG calls; otherwise torch's device RNG is used exactly
    where the reference uses it."""
    if train:
        raise NotImplementedError('train=True needs gradients; the B200 path implements the evaluation loss only')

    def loss_fn(model, batch, class_labels=None, *, rd_t=None, rd_z=None):
        score_fn = mutils.get_score_fn(sde, model, train=False)
        u = torch.rand(batch.shape[0], device=batch.device) if rd_t is None else rd_t.to(batch.device)
        t = u * (sde.T - eps) + eps
        z = torch.randn_like(batch) if rd_z is None else rd_z.to(batch.device)
        mean, std = sde.marginal_prob(batch, t)
        perturbed = _ops.perturb_reflect(mean, z, std)
        score = score_fn(perturbed, t, class_labels=class_labels)
        target = cube.score_hk(perturbed, mean, std)
        if likelihood_weighting:
            weight = sde.sde(torch.zeros_like(batch), t)[1] ** 2
        else:
            weight = std ** 2
        per_sample = _ops.dsm_reduce(score, target, weight, reduce_mean)
        loss = torch.mean(per_sample)
        if torch.isnan(loss):
            print('WARNING: NaN detected in loss!')
        return loss

    return loss_fn


def get_step_fn(sde, train, optimize_fn=None, reduce_mean=False, likelihood_weighting=False):
    """Returns `step_fn(state, batch, class_labels=None)` (losses.py:110-160, eval branch): evaluates the
    loss with the EMA weights swapped in and the live weights restored afterwards."""
    if train:
        raise NotImplementedError('train=True needs gradients; the B200 path implements the evaluation step only')
    loss_fn = get_sde_loss_fn(sde, train, reduce_mean=reduce_mean, likelihood_weighting=likelihood_weighting)

    def step_fn(state, batch, class_labels=None, **rd_replay):
        model = state['model']
        with torch.no_grad():
            ema = state['ema']
            ema.store(model.parameters())
            ema.copy_to(model.parameters())
            try:
                loss = loss_fn(model, batch, class_labels=class_labels, **rd_replay)
            finally:
                ema.restore(model.parameters())
        return loss

    return step_fn
