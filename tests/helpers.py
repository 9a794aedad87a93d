"""Shared test helpers: reference-style configs, synthetic conditioned weights, tape patching."""
import contextlib
import os
import types

import numpy as np
import torch

from oracle import rd_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def make_config(image_size=8, attn=8, corrector="langevin", nf=64, ch_mult=(1, 2, 2), W=9, scale_by_sigma=False,
                n_steps_each=1, precision=None):
    """A types.SimpleNamespace tree with the fields of configs/model/ncsnpp.yaml + sampling.* (SURVEY.md 8c)."""
    model = types.SimpleNamespace(
        name="ncsnpp", channels=1, image_size=image_size, image_width=W, num_classes=1, cond_drop_prob=0.5,
        conditional=True, init_scale=0.0, ema_rate=0.999, nf=nf, ch_mult=list(ch_mult), num_res_blocks=2,
        attn_resolutions=[attn], resamp_with_conv=True, embedding_type="fourier", fourier_scale=16,
        skip_rescale=True, nonlinearity="swish", fir=False, fir_kernel=[1, 3, 3, 1], dropout=0.2,
        scale_by_sigma=scale_by_sigma)
    if precision is not None:
        model.rd_precision = precision   # B200 extension: "bf16" (default) | "fp32" (fp32-class plan)
    samp = types.SimpleNamespace(method="pc", predictor="euler_maruyama", corrector=corrector, denoiser="none",
                                 snr=0.01, n_steps_each=n_steps_each)
    return types.SimpleNamespace(model=model, sampling=samp)


def oracle_cfg(image_size=8, attn=8, **kw):
    return O.NetConfig(image_size=image_size, attn_resolutions=(attn,), **kw)


@contextlib.contextmanager
def patched(mod, name, fn):
    old = getattr(mod, name)
    setattr(mod, name, fn)
    try:
        yield
    finally:
        setattr(mod, name, old)


def rel_to_max(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-300))


def same_bits(a: torch.Tensor, b: torch.Tensor) -> bool:
    """Bitwise equality of fp32 tensors, treating every NaN as equal to every NaN."""
    a, b = a.cpu(), b.cpu()
    nan = torch.isnan(a) & torch.isnan(b)
    ai, bi = a.contiguous().view(torch.int32), b.contiguous().view(torch.int32)
    return bool(((ai == bi) | nan).all())
