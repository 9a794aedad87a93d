"""Reflected predictor-corrector sampling -- drop-in for the reference's `sampling` module
(/root/reference/Reflected-Diffusion/sampling.py:13-339).

Kept from the reference: the registries and their decorators, `get_sampling_fn(config, sde, shape,
eps, device)`, `get_pc_sampler(...)`, the `Predictor` / `Corrector` / `Denoiser` classes with
`update_fn`, and the sampler's observable behaviour (corrector before predictor, last grid point
skipped, noisy `x` returned, denoiser output discarded, NFE reported as N*(n_steps+1)).

B200 execution:
  * `update_fn` of the reflected Euler-Maruyama predictor / Langevin corrector = one score call +
    one fused kernel (update + reflection; warp-shuffle norms for the Langevin step size).
  * `pc_sampler` with the B200 `NCSNpp` runs the WHOLE loop natively: guided score (both CFG passes
    in one plan) -> fused update, one CUDA graph per iteration replayed N-1 times, per-step scalars
    read on the device, noise from an in-kernel Philox stream (or an injected tape for parity).
"""
import abc
import contextlib

import numpy as np
import torch

import cube  # noqa: F401  (re-exported like the reference module does)
from models import utils as mutils
from models.utils import from_flattened_numpy, to_flattened_numpy, get_score_fn  # noqa: F401
from rdb200 import ops as _ops

_CORRECTORS = {}
_PREDICTORS = {}
_DENOISERS = {}


def _registrar(table):
    def register(cls=None, *, name=None):
        def _do(c):
            key = c.__name__ if name is None else name
            if key in table:
                raise ValueError(f'Already registered model with name: {key}')
            table[key] = c
            return c
        return _do if cls is None else _do(cls)
    return register


register_predictor = _registrar(_PREDICTORS)
register_corrector = _registrar(_CORRECTORS)
register_denoiser = _registrar(_DENOISERS)


def get_predictor(name):
    return _PREDICTORS[name]


def get_corrector(name):
    return _CORRECTORS[name]


def get_denoiser(name):
    return _DENOISERS[name]


def get_sampling_fn(config, sde, shape, eps, device):
    """Build the sampling function named by `config.sampling.method` (sampling.py:87-130)."""
    s = config.sampling
    method = s.method.lower()
    if method == 'ode':
        get_denoiser(s.denoiser.lower())  # looked up (and validated) before the sampler is built, like the reference
        return get_ode_sampler(sde=sde, shape=shape, eps=eps, moll=s.moll, side_eps=s.side_eps, device=device)
    if method == 'pc':
        return get_pc_sampler(sde=sde, shape=shape, predictor=get_predictor(s.predictor.lower()),
                              corrector=get_corrector(s.corrector.lower()), denoiser=get_denoiser(s.denoiser.lower()),
                              snr=s.snr, n_steps=s.n_steps_each, eps=eps, device=device)
    raise ValueError(f"Sampler name {s.method} unknown.")


class Predictor(abc.ABC):
    """A predictor advances x one step along the reverse SDE built from `score_fn`."""

    def __init__(self, sde, score_fn, probability_flow=False):
        super().__init__()
        self.sde = sde
        self.rsde = sde.reverse(score_fn, probability_flow)
        self.score_fn = score_fn

    @abc.abstractmethod
    def update_fn(self, x, t):
        """-> (x_next, x_next_mean)"""


class Corrector(abc.ABC):
    """A corrector refines x at a fixed noise level with `n_steps` MCMC moves."""

    def __init__(self, sde, score_fn, snr, n_steps):
        super().__init__()
        self.sde, self.score_fn, self.snr, self.n_steps = sde, score_fn, snr, n_steps

    @abc.abstractmethod
    def update_fn(self, x, t):
        """-> (x_next, x_next_mean)"""


class Denoiser(abc.ABC):
    def __init__(self, denoiser):
        super().__init__()
        self.denoiser = denoiser

    @abc.abstractmethod
    def update_fn(self, x, x_mean, t):
        pass


@register_predictor(name='euler_maruyama')
class ReflectedEulerMaruyamaPredictor(Predictor):
    """x <- reflect(x + g^2 s / N + g sqrt(1/N) z) (sampling.py:193-207); one fused kernel after the score."""

    def update_fn(self, x, t):
        z = torch.randn_like(x)  # drawn before the score call, like the reference
        score = self.score_fn(x, t)
        _, g = self.sde.sde(torch.zeros(1, 1, 1, 1, device=t.device), t)
        if self.rsde.probability_flow:  # ODE variant: half drift, no noise (sde_lib.py:97-100)
            score, z = score * 0.5, torch.zeros_like(z)
        return _ops.predictor_step(x, score, z, g, self.rsde.N)


@register_corrector(name='langevin')
class ReflectedLangevinCorrector(Corrector):
    """Langevin move with batch-mean signal-to-noise step size (sampling.py:210-233)."""

    def update_fn(self, x, t):
        x_mean = x
        for _ in range(self.n_steps):
            grad = self.score_fn(x, t)
            noise = torch.randn_like(x)
            x, x_mean, _ = _ops.corrector_step(x, grad, noise, self.snr)
        return x, x_mean


@register_corrector(name='none')
class NoneCorrector(Corrector):
    def update_fn(self, x, t):
        return x, x


@register_denoiser(name='network')
class TrainedDenoiser(Denoiser):
    def update_fn(self, x, x_mean, t):
        return (x - self.denoiser(x, t)).clamp(min=0, max=1)


@register_denoiser(name='mean')
class MeanDenoiser(Denoiser):
    def update_fn(self, x, x_mean, t):
        return x_mean


@register_denoiser(name='none')
class NoneDenoiser(Denoiser):
    def update_fn(self, x, x_mean, t):
        return x


def _native_engine(model, sde, shape, predictor, corrector, snr, n_steps, eps, device, class_labels):
    """The fully native loop applies when the model is the B200 NCSNpp, the SDE tabulates its steps,
    and the step algorithms are the registered reflected EM predictor / Langevin-or-none corrector."""
    if not hasattr(model, 'rd_sampler_engine') or not hasattr(sde, 'step_tables'):
        return None
    if torch.device(device).type != 'cuda' or predictor is not ReflectedEulerMaruyamaPredictor:
        return None
    if corrector is ReflectedLangevinCorrector:
        n_corr = n_steps
    elif corrector is NoneCorrector:
        n_corr = 0
    else:
        return None
    cfg = class_labels is not None
    if not cfg and getattr(model, 'conditional', False):
        return None  # (the reference crashes here: label_emb(None), ncsnpp.py:262)
    B, _, H, W = shape
    return model.rd_sampler_engine(B, H, W, device, sde, eps, snr, n_corr, cfg=cfg)


def get_pc_sampler(sde, shape, predictor, corrector, denoiser, snr, n_steps=1, eps=1e-3, device='cuda'):
    """Create the predictor-corrector sampler closure (sampling.py:292-339)."""

    def pc_sampler(model, z=None, noise_removal_model=None, weight=0, class_labels=None, *, rd_tape=None,
                   rd_seed=None, rd_native=True, rd_graph=True):
        """-> (samples [B,C,H,W] on `device`, reported number of function evaluations).

        `z` is accepted and ignored exactly like the reference (the prior is re-drawn, sampling.py:324).
        Extras (keyword-only, not in the reference): `rd_tape` [(N-1)*k, B, C, H, W] replays injected
        N(0,1) noise (k = n_steps + 1 with the Langevin corrector, 1 without) for parity runs; `rd_seed` fixes the
        Philox stream; `rd_native=False` forces the generic python loop over `update_fn`.
        """
        if z is None:
            torch.rand(shape)  # the reference draws (and discards) a first prior here (sampling.py:308)
        with torch.no_grad():
            x = torch.rand(shape).to(device)
            engine = _native_engine(model, sde, shape, predictor, corrector, snr, n_steps, eps, device,
                                    class_labels) if rd_native else None
            if engine is not None:
                model.eval()
                seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if rd_seed is None else int(rd_seed)
                x = engine.sample(x, class_labels, weight, tape=rd_tape, seed=seed, use_graph=rd_graph)
                return x, sde.N * (n_steps + 1)

            # generic loop: any score model, fused step kernels, torch RNG (tape = patched randn_like)
            if class_labels is None:
                score_fn = mutils.get_score_fn(sde, model, train=False)
            else:
                score_fn = mutils.get_cf_score_fn(sde, model, class_labels, weight)
            pred, corr = predictor(sde, score_fn), corrector(sde, score_fn, snr, n_steps)
            tape = None if rd_tape is None else list(rd_tape)
            timesteps = torch.linspace(sde.T, eps, sde.N, device=device)
            # nothing can rewrite the parameters inside this loop: check / pack the weights once, not per call
            frozen = model.rd_freeze_weights() if hasattr(model, 'rd_freeze_weights') else contextlib.nullcontext()
            with frozen:
                for i in range(sde.N - 1):  # the last grid point performs no update (sampling.py:330)
                    vec_t = torch.ones(shape[0], device=device) * timesteps[i]
                    if tape is not None:
                        x = _tape_iteration(pred, corr, x, vec_t, tape, device)
                    else:
                        x, _ = corr.update_fn(x, vec_t)
                        x, _ = pred.update_fn(x, vec_t)
            return x, sde.N * (n_steps + 1)

    return pc_sampler


def _tape_iteration(pred, corr, x, vec_t, tape, device):
    """One generic iteration with torch.randn_like served from `tape` (parity runs)."""
    real = torch.randn_like
    torch.randn_like = lambda t, *a, **k: tape.pop(0).to(device).reshape(t.shape)
    try:
        x, _ = corr.update_fn(x, vec_t)
        x, _ = pred.update_fn(x, vec_t)
    finally:
        torch.randn_like = real
    return x


def get_ode_sampler(sde, shape, rtol=1e-5, atol=1e-5, method='RK45', eps=1e-3, moll=200, side_eps=1e-2, device='cuda'):
    """Probability-flow ODE sampler (sampling.py:342-392).

    The reference hands the flattened float64 state to scipy.integrate.solve_ivp on the HOST and crosses host <-> GPU for
    every right-hand side.  For method='RK45' (the default, and the only one any shipped config uses) the B200 path runs
    the same Dormand-Prince scheme with scipy's step-size control law on the DEVICE (rdb200/ode.py): float64 state and
    stage arithmetic in CUDA kernels, each right-hand side = one guided network plan + one fused drift-times-mollifier
    kernel, one 8-byte read per step for the accept / reject decision.  Other `method`s (or rd_host_solver=True) take
    the reference's scipy loop with the same device right-hand side."""

    def ode_sampler(model, z=None, noise_removal_model=None, weight=0, class_labels=None, *, rd_host_solver=False):
        """-> (samples [B,C,H,W] on `device`, number of function evaluations)."""
        with torch.no_grad():
            if z is None:
                x = (1 - 2 * side_eps) * torch.rand(shape).to(device) + side_eps
            else:
                x = z
            if class_labels is None:
                score_fn = mutils.get_score_fn(sde, model, train=False)
            else:
                score_fn = mutils.get_cf_score_fn(sde, model, class_labels, weight)

            def rhs(t, xt):
                vec_t = torch.ones(shape[0], device=xt.device) * t
                g = sde.sde(xt, vec_t)[1]
                return _ops.pf_drift(xt, score_fn(xt, vec_t), g, moll)

            frozen = model.rd_freeze_weights() if hasattr(model, 'rd_freeze_weights') else contextlib.nullcontext()
            with frozen:
                if method == 'RK45' and not rd_host_solver and x.is_cuda:
                    from rdb200.ode import DeviceRK45
                    solver = DeviceRK45(rhs, float(sde.T), x.to(torch.float32).reshape(shape), float(eps), rtol=rtol, atol=atol)
                    y = solver.solve()
                    return y.to(torch.float32), solver.nfev

                from scipy import integrate

                def ode_func(t, flat):
                    xt = from_flattened_numpy(flat, shape).to(device).type(torch.float32)
                    return to_flattened_numpy(rhs(t, xt))

                solution = integrate.solve_ivp(ode_func, (sde.T, eps), to_flattened_numpy(x), rtol=rtol, atol=atol,
                                               method=method)
            x = torch.tensor(solution.y[:, -1]).reshape(shape).to(device).type(torch.float32)
            return x, solution.nfev

    return ode_sampler
