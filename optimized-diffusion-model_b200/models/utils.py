"""Model registry and score-function wrappers -- drop-in for the reference's `models.utils`
(/root/reference/Reflected-Diffusion/models/utils.py:8-150).

The wrappers keep the reference's call conventions (`score_fn(x, t)`, `model(x, sigma, class_labels)`);
when the model is the B200 `NCSNpp`, the classifier-free-guidance wrapper runs both passes and
the combine inside one native plan (no `x.repeat`, no concat, no separate combine kernels).
"""
import numpy as np
import torch

_MODELS = {}


def register_model(cls=None, *, name=None):
    """Class decorator: register under `name` (default: the class name); duplicates raise ValueError."""

    def _do(c):
        key = c.__name__ if name is None else name
        if key in _MODELS:
            raise ValueError(f'Already registered model with name: {key}')
        _MODELS[key] = c
        return c

    return _do if cls is None else _do(cls)


def get_model(name):
    return _MODELS[name]


def get_sigmas(config):
    """Geometric noise ladder sigma_max -> sigma_min (models/utils.py:34-45)."""
    s = config.sde
    return np.exp(np.linspace(np.log(s.sigma_max), np.log(s.sigma_min), s.num_scales))


def create_model(config):
    return get_model(config.model.name)(config)


def get_model_fn(model, train=False):
    """`model_fn(x, time_cond, class_labels=None)`: puts the model in train/eval mode, then calls it."""

    def model_fn(x, time_cond, class_labels=None):
        model.train() if train else model.eval()
        return model(x, time_cond, class_labels=class_labels)

    return model_fn


def get_score_fn(sde, model, train=False):
    """`score_fn(x, t, class_labels=None)`: conditions the network on sigma(t), not on t
    (models/utils.py:87-105); the network output is the score itself."""
    model_fn = get_model_fn(model, train=train)

    def score_fn(x, t, class_labels=None):
        sigma = sde.marginal_prob(torch.zeros_like(x), t)[1]
        return model_fn(x, sigma, class_labels=class_labels)

    score_fn.sde, score_fn.model = sde, model
    return score_fn


def get_cf_score_fn(sde, model, class_labels, weight):
    """Classifier-free-guided score `(1 + w) s(x, c) - w s(x, 0)` (models/utils.py:108-140).
    `weight`: None (-> 0), a python number, or a [B] tensor."""
    plain = get_score_fn(sde, model, train=False)
    native = getattr(model, "rd_guided_score", None)

    def weighted_score_fn(x, t):
        if native is not None and x.is_cuda:
            sigma = sde.marginal_prob(torch.zeros_like(x), t)[1]
            model.eval()
            return native(x, sigma, class_labels, weight)
        both = plain(x.repeat(2, 1, 1, 1), t.repeat(2), torch.cat([class_labels, torch.zeros_like(class_labels)], dim=0))
        if both.is_cuda:
            from rdb200 import ops
            return ops.cfg_combine(both, weight)
        raise RuntimeError("the B200 drop-in has no CPU path: scores must be CUDA tensors")

    weighted_score_fn.sde, weighted_score_fn.model = sde, model
    weighted_score_fn.class_labels, weighted_score_fn.weight = class_labels, weight
    return weighted_score_fn


def to_flattened_numpy(x):
    return x.detach().cpu().numpy().reshape((-1,))


def from_flattened_numpy(x, shape):
    return torch.from_numpy(x.reshape(shape))
