"""CPU/GPU-agnostic restatement of the Reflected-Diffusion sampling hot path in plain PyTorch.

TEST INFRASTRUCTURE ONLY.  This module is the parity checker for the CUDA kernels: it may be
imported from ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` and from nowhere else.  The product path (``optimized-diffusion-model_b200/``)
never imports it and has no CPU fallback.

Every function restates one piece of the reference (file:line under
``/root/reference/Reflected-Diffusion``).  The reference ships no tests or golden vectors
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference itself, generated in
the build container by ``oracle/make_golden.py`` and committed under ``tests/golden/``;
``tests/test_oracle_golden.py`` re-checks the restatement against them on every run.

The network restatement is functional (driven by a ``state_dict`` with the reference's key names,
SURVEY.md App. C) so that the same weights can be fed to the reference, to this oracle and to the
weight packer of the CUDA path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# cube.py
def reflect(x: Tensor) -> Tensor:
    """cube.reflect (cube.py:34-49): fold x into [0,1] by repeated reflection (period 2)."""
    m = torch.remainder(x, 2)
    return torch.where(m > 1, 2 - m, m)


def inside(x: Tensor) -> Tensor:
    """cube.inside (cube.py:17-31)."""
    f = x.flatten(1)
    return ((f >= 0) & (f <= 1)).all(dim=-1)


def _score_hk_ef(x: Tensor, x0: Tensor, t: Tensor, efs: int) -> Tensor:
    """cube._score_hk_ef (cube.py:73-107): cosine eigenfunction series of the Neumann heat kernel."""
    shape1 = (-1,) + (1,) * (x.dim())
    k = torch.arange(1, efs + 1).to(x)
    kx = (math.pi * x.unsqueeze(0)) * k.view(shape1)
    k0 = (math.pi * x0.unsqueeze(0)) * k.view(shape1)
    e_den = ((-t.unsqueeze(0)) * k.unsqueeze(-1).pow(2) * (math.pi ** 2)).exp()  # [efs, B]
    e_num = e_den * k.unsqueeze(-1)
    bshape = e_den.shape + (1,) * (x.dim() - 1)
    c0 = k0.cos()
    num = -2 * math.pi * (e_num.view(bshape) * (kx.sin() * c0)).sum(0)
    den = 1 + 2 * (e_den.view(bshape) * (kx.cos() * c0)).sum(0)
    return num / (den + 1e-12)


def _score_hk_refl(x: Tensor, x0: Tensor, t: Tensor, refls: int) -> Tensor:
    """cube._score_hk_refl (cube.py:110-146): method of images, 2*(2*refls+1) images."""
    m2 = torch.arange(-2 * refls, 2 * refls + 1, 2).to(x)
    shape1 = (-1,) + (1,) * (x.dim())
    images = torch.cat((m2.view(shape1) + x.unsqueeze(0), m2.view(shape1) - x.unsqueeze(0)), dim=0)
    sign = torch.cat((torch.ones_like(m2), -torch.ones_like(m2)), dim=0).view(shape1)
    d = images - x0.unsqueeze(0)
    fourt = 4 * t.view((1, -1) + (1,) * (x.dim() - 1))
    coeff = -2 * d / fourt
    e = (-d.pow(2) / fourt).exp()
    return (coeff * e * sign).sum(0) / (e.sum(0) + 1e-12)


def score_hk(x: Tensor, x_orig: Tensor, sigma, efs: int = 20, refls: int = 10, min_cutoff: float = 1e-2) -> Tensor:
    """cube.score_hk (cube.py:149-193): per-sample branch on t = sigma^2/2 > min_cutoff."""
    t = sigma ** 2 / 2
    if not torch.is_tensor(t):
        t = t * torch.ones(x.shape[0]).to(x)
    ef = t > min_cutoff
    out = torch.zeros_like(x)
    if ef.any():
        out[ef] = _score_hk_ef(x[ef], x_orig[ef], t[ef], efs)
    if (~ef).any():
        out[~ef] = _score_hk_refl(x[~ef], x_orig[~ef], t[~ef], refls)
    return out


def sample_hk(x: Tensor, sigma, noise: Optional[Tensor] = None) -> Tensor:
    """cube.sample_hk (cube.py:52-70) with injectable noise."""
    if not torch.is_tensor(sigma):
        sigma = sigma * torch.ones(x.shape[0]).to(x)
    if noise is None:
        noise = torch.randn_like(x)
    return reflect(noise * sigma.view((-1,) + (1,) * (x.dim() - 1)) + x)


# ----------------------------------------------------------------------------------------------
# sde_lib.py
@dataclass
class VESchedule:
    """RVESDE (sde_lib.py:114-161) evaluated on the sampler's time grid (sampling.py:325)."""
    sigma_min: float = 0.01
    sigma_max: float = 5.0
    N: int = 1000
    T: float = 1.0
    eps: float = 1e-5

    def timesteps(self) -> Tensor:
        return torch.linspace(self.T, self.eps, self.N)

    def sigma(self, t: Tensor) -> Tensor:
        # sde_lib.py:136 / :143  sigma_min * (sigma_max / sigma_min) ** t
        return self.sigma_min * (self.sigma_max / self.sigma_min) ** t

    def diffusion(self, t: Tensor) -> Tensor:
        # sde_lib.py:138-139
        c = torch.sqrt(torch.tensor(2 * (np.log(self.sigma_max) - np.log(self.sigma_min)), dtype=torch.float32))
        return self.sigma(t) * c.to(t.device)


def predictor_step(x: Tensor, score: Tensor, g: Tensor, N: int, z: Tensor):
    """ReflectedEulerMaruyamaPredictor.update_fn (sampling.py:198-207) given the score.

    g: [B] diffusion coefficients.  Reverse drift = 0 - g^2 * score (sde_lib.py:95-98).
    """
    dt = -1.0 / N
    drift = torch.zeros_like(x) - g[:, None, None, None] ** 2 * score * 1.0
    x_mean = x + drift * dt
    x_new = x_mean + g[:, None, None, None] * np.sqrt(-dt) * z
    return reflect(x_new), reflect(x_mean)


def corrector_step(x: Tensor, grad: Tensor, noise: Tensor, snr: float):
    """One inner iteration of ReflectedLangevinCorrector.update_fn (sampling.py:222-231)."""
    B = x.shape[0]
    grad_norm = torch.norm(grad.reshape(B, -1), dim=-1).mean()
    noise_norm = torch.norm(noise.reshape(B, -1), dim=-1).mean()
    step = (snr * noise_norm / grad_norm) ** 2 * 2 * torch.ones(B, device=x.device, dtype=x.dtype)
    x_mean = x + step[:, None, None, None] * grad
    x_new = x_mean + torch.sqrt(step * 2)[:, None, None, None] * noise
    return reflect(x_new), reflect(x_mean), (grad_norm, noise_norm, step[0])


def cfg_combine(s2: Tensor, weight) -> Tensor:
    """get_cf_score_fn tail (models/utils.py:124-138): (1 + w) s_cond - w s_uncond."""
    B = s2.shape[0] // 2
    if weight is None:
        w = torch.zeros(B, device=s2.device)
    elif isinstance(weight, (float, int)):
        w = torch.full((B,), float(weight), device=s2.device)
    else:
        w = weight
    w = w.view(-1, 1, 1, 1)
    return (1 + w) * s2[:B] - w * s2[B:]


# ----------------------------------------------------------------------------------------------
# models/ncsnpp.py + layerspp.py, functional
@dataclass
class NetConfig:
    """The subset of configs/model/ncsnpp.yaml that shapes the network."""
    channels: int = 1
    image_size: int = 8
    nf: int = 64
    ch_mult: Sequence[int] = (1, 2, 2)
    num_res_blocks: int = 2
    attn_resolutions: Sequence[int] = (8,)
    num_classes: int = 1
    conditional: bool = True
    skip_rescale: bool = True
    scale_by_sigma: bool = False
    fourier_scale: float = 16.0


def _groups(c: int) -> int:
    return min(c // 4, 32)


def _gn(h: Tensor, sd: Dict[str, Tensor], key: str) -> Tensor:
    return F.group_norm(h, _groups(h.shape[1]), sd[key + ".weight"], sd[key + ".bias"], eps=1e-6)


def _conv3(h: Tensor, sd, key: str, stride: int = 1, padding: int = 1) -> Tensor:
    return F.conv2d(h, sd[key + ".weight"], sd[key + ".bias"], stride=stride, padding=padding)


def _nin(h: Tensor, sd, key: str) -> Tensor:
    # layers.NIN (layers.py:531-540): channel matmul with W stored [in, out]
    return torch.einsum("bchw,cd->bdhw", h, sd[key + ".W"]) + sd[key + ".b"][None, :, None, None]


def resblock(h: Tensor, temb: Tensor, sd, p: str, skip_rescale: bool = True) -> Tensor:
    """ResnetBlockDDPMpp.forward (layerspp.py:198-214), eval mode (dropout = identity)."""
    x = h
    h = F.silu(_gn(h, sd, p + ".GroupNorm_0"))
    h = _conv3(h, sd, p + ".Conv_0")
    h = h + F.linear(F.silu(temb), sd[p + ".Dense_0.weight"], sd[p + ".Dense_0.bias"])[:, :, None, None]
    h = F.silu(_gn(h, sd, p + ".GroupNorm_1"))
    h = _conv3(h, sd, p + ".Conv_1")
    if (p + ".NIN_0.W") in sd:
        x = _nin(x, sd, p + ".NIN_0")
    return (x + h) / np.sqrt(2.0) if skip_rescale else x + h


def attnblock(x: Tensor, sd, p: str, skip_rescale: bool = True) -> Tensor:
    """AttnBlockpp.forward (layerspp.py:80-96): single-head attention over the H*W pixels."""
    B, C, H, W = x.shape
    h = _gn(x, sd, p + ".GroupNorm_0")
    q = _nin(h, sd, p + ".NIN_0").flatten(2)  # [B, C, T]
    k = _nin(h, sd, p + ".NIN_1").flatten(2)
    v = _nin(h, sd, p + ".NIN_2").flatten(2)
    w = torch.einsum("bct,bcs->bts", q, k) * (int(C) ** (-0.5))
    w = F.softmax(w, dim=-1)
    h = torch.einsum("bts,bcs->bct", w, v).reshape(B, C, H, W)
    h = _nin(h, sd, p + ".NIN_3")
    return (x + h) / np.sqrt(2.0) if skip_rescale else x + h


def temb_trunk(sigma: Tensor, labels: Optional[Tensor], sd, cfg: NetConfig) -> Tensor:
    """ncsnpp.py:252-262 + GaussianFourierProjection (layerspp.py:26-28)."""
    proj = torch.log(sigma)[:, None] * sd["time_embed.W"][None, :] * 2 * np.pi
    emb = torch.cat([torch.sin(proj), torch.cos(proj)], dim=-1)
    t = F.linear(emb, sd["time_mlp.0.weight"], sd["time_mlp.0.bias"])
    t = F.linear(F.silu(t), sd["time_mlp.2.weight"], sd["time_mlp.2.bias"])
    if cfg.conditional:
        t = t + F.linear(labels, sd["label_emb.weight"], sd["label_emb.bias"])
    return t


def ncsnpp_forward(x: Tensor, sigma: Tensor, labels: Optional[Tensor], sd: Dict[str, Tensor], cfg: NetConfig,
                   taps: Optional[dict] = None) -> Tensor:
    """NCSNpp.forward (ncsnpp.py:226-354), eval mode.  `taps` (optional dict) receives named
    intermediate activations for layer-level parity tests."""
    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    L = len(cfg.ch_mult)
    has_attn = [(cfg.image_size // (2 ** i)) in cfg.attn_resolutions for i in range(L)]
    temb = tap("temb", temb_trunk(sigma, labels, sd, cfg))
    h = tap("input_conv", _conv3(x, sd, "input_conv"))
    hs: List[Tensor] = [h]
    d = 0
    for i in range(L):
        for _ in range(cfg.num_res_blocks):
            h = tap(f"down_blocks.{d}", resblock(h, temb, sd, f"down_blocks.{d}", cfg.skip_rescale))
            if has_attn[i]:
                h = tap(f"down_attn.{d}", attnblock(h, sd, f"down_attn.{d}", cfg.skip_rescale))
            hs.append(h)
            d += 1
        hs.append(h)  # extra skip for the (num_res_blocks+1)-th up block (ncsnpp.py:288)
        if i != L - 1:
            # Downsample (layerspp.py:157-159): pad right/bottom by one, 3x3 stride 2, no padding
            h = tap(f"downsample.{i}", _conv3(F.pad(h, (0, 1, 0, 1)), sd, f"downsample.{i}.Conv_0", stride=2, padding=0))
    h = tap("mid_block1", resblock(h, temb, sd, "mid_block1", cfg.skip_rescale))
    if (cfg.image_size // (2 ** (L - 1))) in cfg.attn_resolutions:
        h = tap("mid_attn", attnblock(h, sd, "mid_attn", cfg.skip_rescale))
    h = tap("mid_block2", resblock(h, temb, sd, "mid_block2", cfg.skip_rescale))
    u = 0
    # NOTE: the reference walks the *module lists* in construction order (coarsest level first):
    # upsample[j] is applied after the j-th group of up blocks, and is None for the last group.
    for j, i in enumerate(reversed(range(L))):
        for _ in range(cfg.num_res_blocks + 1):
            skip = hs.pop()
            if h.shape[2:] != skip.shape[2:]:
                h = F.interpolate(h, size=skip.shape[2:], mode="nearest")  # ncsnpp.py:319-320
            h = torch.cat([h, skip], dim=1)
            h = tap(f"up_blocks.{u}", resblock(h, temb, sd, f"up_blocks.{u}", cfg.skip_rescale))
            if has_attn[i]:
                h = tap(f"up_attn.{u}", attnblock(h, sd, f"up_attn.{u}", cfg.skip_rescale))
            u += 1
        if i != 0:
            Bn, Cn, Hn, Wn = h.shape
            h = F.interpolate(h, size=(Hn * 2, Wn * 2), mode="nearest")  # Upsample (layerspp.py:122-124)
            h = tap(f"upsample.{j}", _conv3(h, sd, f"upsample.{j}.Conv_0"))
    h = F.silu(_gn(h, sd, "out_norm"))
    h = _conv3(h, sd, "out_conv")
    if cfg.scale_by_sigma:
        h = h / sigma.view(-1, 1, 1, 1)
    return tap("out", h)


def guided_score(x: Tensor, sigma: Tensor, labels: Tensor, weight, sd, cfg: NetConfig) -> Tensor:
    """get_cf_score_fn (models/utils.py:108-140): one forward at 2B, unconditional half = label 0."""
    x2 = x.repeat(2, 1, 1, 1)
    s2 = sigma.repeat(2)
    l2 = torch.cat([labels, torch.zeros_like(labels)], dim=0)
    return cfg_combine(ncsnpp_forward(x2, s2, l2, sd, cfg), weight)


# ----------------------------------------------------------------------------------------------
# sampling.py
def state_dict_shapes(cfg: NetConfig) -> Dict[str, tuple]:
    """Key -> shape of NCSNpp.state_dict(), derived from the constructor logic
    (ncsnpp.py:42-224)."""
    nf, L = cfg.nf, len(cfg.ch_mult)
    td = nf * 4
    shapes: Dict[str, tuple] = {}

    def lin(p, o, i):
        shapes[p + ".weight"] = (o, i)
        shapes[p + ".bias"] = (o,)

    def conv(p, o, i):
        shapes[p + ".weight"] = (o, i, 3, 3)
        shapes[p + ".bias"] = (o,)

    def gn(p, c):
        shapes[p + ".weight"] = (c,)
        shapes[p + ".bias"] = (c,)

    def nin(p, i, o):
        shapes[p + ".W"] = (i, o)
        shapes[p + ".b"] = (o,)

    def res(p, i, o):
        gn(p + ".GroupNorm_0", i)
        conv(p + ".Conv_0", o, i)
        lin(p + ".Dense_0", o, td)
        gn(p + ".GroupNorm_1", o)
        conv(p + ".Conv_1", o, o)
        if i != o:
            nin(p + ".NIN_0", i, o)

    def attn(p, c):
        gn(p + ".GroupNorm_0", c)
        for j in range(4):
            nin(p + f".NIN_{j}", c, c)

    shapes["time_embed.W"] = (nf,)
    lin("time_mlp.0", td, 2 * nf)
    lin("time_mlp.2", td, td)
    if cfg.conditional:
        lin("label_emb", td, cfg.num_classes)
    conv("input_conv", nf, cfg.channels)
    has_attn = [(cfg.image_size // (2 ** i)) in cfg.attn_resolutions for i in range(L)]
    in_ch = nf
    skips = []
    d = 0
    for i, m in enumerate(cfg.ch_mult):
        out_ch = nf * m
        for _ in range(cfg.num_res_blocks):
            res(f"down_blocks.{d}", in_ch, out_ch)
            in_ch = out_ch
            if has_attn[i]:
                attn(f"down_attn.{d}", in_ch)
            skips.append(in_ch)
            d += 1
        skips.append(in_ch)
        if i != L - 1:
            conv(f"downsample.{i}.Conv_0", in_ch, in_ch)
    res("mid_block1", in_ch, in_ch)
    if (cfg.image_size // (2 ** (L - 1))) in cfg.attn_resolutions:
        attn("mid_attn", in_ch)
    res("mid_block2", in_ch, in_ch)
    u = 0
    for j, i in enumerate(reversed(range(L))):
        out_ch = nf * cfg.ch_mult[i]
        for _ in range(cfg.num_res_blocks + 1):
            res(f"up_blocks.{u}", in_ch + skips.pop(), out_ch)
            in_ch = out_ch
            if has_attn[i]:
                attn(f"up_attn.{u}", in_ch)
            u += 1
        if i != 0:
            conv(f"upsample.{j}.Conv_0", in_ch, in_ch)
    gn("out_norm", in_ch)
    conv("out_conv", cfg.channels, in_ch)
    return shapes


def synth_state_dict(cfg: NetConfig, seed: int = 0, out_scale: float = 0.1, degenerate: bool = False,
                     dtype=torch.float32) -> Dict[str, Tensor]:
    """Deterministic synthetic weights keyed like the reference state_dict.

    `degenerate=False` is the *conditioned* policy of SURVEY.md section 4 / App. E: every conv / NIN /
    linear weight ~ U(+-sqrt(3/fan_avg)) at scale 1 (including the layers the reference initialises
    at 1e-10), `out_conv` scaled by `out_scale`, biases 0.1 N(0,1), GroupNorm affine 1 + 0.1 N(0,1)
    / 0.1 N(0,1).  `degenerate=True` reproduces the magnitude pattern of the untouched reference
    init (Conv_1 / NIN_3 / out_conv at variance scale 1e-10, zero biases, unit GroupNorm).
    The values only depend on (cfg, seed, flags) and torch's CPU generator.
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for key, shape in state_dict_shapes(cfg).items():
        leaf = key.rsplit(".", 1)[-1]
        parent = key.rsplit(".", 1)[0]
        if key == "time_embed.W":
            v = torch.randn(shape, generator=g) * cfg.fourier_scale
        elif "GroupNorm" in key or key.startswith("out_norm"):
            if degenerate:
                v = torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
            else:
                v = (1.0 if leaf == "weight" else 0.0) + 0.1 * torch.randn(shape, generator=g)
        elif leaf in ("bias", "b"):
            v = torch.zeros(shape) if degenerate else 0.1 * torch.randn(shape, generator=g)
        else:
            if leaf == "W":  # NIN [in, out]
                fan_in, fan_out = shape[0], shape[1]
            else:
                rf = int(np.prod(shape[2:])) if len(shape) > 2 else 1
                fan_in, fan_out = shape[1] * rf, shape[0] * rf
            scale = 1.0
            small = parent.endswith("Conv_1") or parent.endswith("NIN_3") or parent == "out_conv"
            if degenerate and small:
                scale = 1e-10
            elif parent == "out_conv":
                scale = out_scale ** 2  # variance scale -> amplitude out_scale
            elif parent.split(".")[-1].startswith("NIN") and not small:
                scale = 0.1 if degenerate else 1.0  # layers.NIN default init_scale=0.1 (layers.py:532)
            bound = math.sqrt(3.0 * scale / ((fan_in + fan_out) / 2.0))
            v = (torch.rand(shape, generator=g) * 2.0 - 1.0) * bound
        sd[key] = v.to(dtype)
    return sd


@dataclass
class SamplerConfig:
    """configs/train.yaml:31-39 `sampling.*` + the sampler eps (run_train.py:105)."""
    predictor: str = "euler_maruyama"
    corrector: str = "langevin"
    snr: float = 0.01
    n_steps_each: int = 1
    eps: float = 1e-5


def make_tape(B: int, D_shape: Sequence[int], n_draws: int, seed: int):
    """x0 ~ U[0,1] and `n_draws` N(0,1) tensors, in the order the sampler consumes them
    (per iteration corrector noise, then predictor z)."""
    g = torch.Generator().manual_seed(seed)
    shape = (B,) + tuple(D_shape)
    x0 = torch.rand(shape, generator=g)
    noise = torch.randn((n_draws,) + shape, generator=g)
    return x0, noise


def pc_sampler(score_fn, sched: VESchedule, scfg: SamplerConfig, x0: Tensor, noise: Tensor,
               trace: Optional[list] = None) -> Tensor:
    """pc_sampler (sampling.py:295-337) with an injected noise tape.

    score_fn(x, sigma[B]) -> guided score.  Corrector first, then predictor; the last grid point is
    skipped; the (discarded) denoiser is omitted; the noisy x is returned (sampling.py:327-337).
    `trace` (optional list) receives (i, x_after_corrector, x_after_predictor) for teacher forcing.
    """
    x = x0.clone()
    B = x.shape[0]
    ts = sched.timesteps().to(x.device)
    n_corr = scfg.n_steps_each if scfg.corrector == "langevin" else 0
    draw = 0
    for i in range(sched.N - 1):
        vec_t = torch.ones(B, device=x.device) * ts[i]
        sig = sched.sigma(vec_t)
        for _ in range(n_corr):
            grad = score_fn(x, sig)
            x, _, _ = corrector_step(x, grad, noise[draw].to(x.device), scfg.snr)
            draw += 1
        x_c = x
        z = noise[draw].to(x.device)  # drawn before the score call (sampling.py:200)
        draw += 1
        score = score_fn(x, sig)
        x, _ = predictor_step(x, score, sched.diffusion(vec_t), sched.N, z)
        if trace is not None:
            trace.append((i, x_c.clone(), x.clone()))
    return x


# ----------------------------------------------------------------------------------------------
# "next" rows: evaluation DSM loss, EMA, latent -> physical codec
def dsm_loss(score_fn, sched: VESchedule, batch: Tensor, u: Tensor, z: Tensor, reduce_mean: bool = True,
             likelihood_weighting: bool = True, eps: float = 1e-5):
    """Evaluation loss (losses.py:77-92) with the two RNG draws passed in: `u` uniform [B] and `z`
    like `batch`.  `score_fn(x, sigma)` is the (unguided) network call.  Returns (loss, per-sample
    losses, perturbed data, heat-kernel target)."""
    t = u * (sched.T - eps) + eps  # losses.py:78
    std = sched.sigma(t)  # marginal_prob: mean = batch (sde_lib.py:141-143)
    perturbed = reflect(batch + std[:, None, None, None] * z)
    score = score_fn(perturbed, std)
    target = score_hk(perturbed, batch, std)
    w = sched.diffusion(t) ** 2 if likelihood_weighting else std ** 2
    losses = w[:, None, None, None] * (score - target).pow(2)
    flat = losses.reshape(losses.shape[0], -1)
    per_sample = torch.mean(flat, dim=-1) if reduce_mean else 0.5 * torch.sum(flat, dim=-1)
    return torch.mean(per_sample), per_sample, perturbed, target


def ode_sampler(score_fn, sched: VESchedule, x0: Tensor, rtol: float = 1e-5, atol: float = 1e-5, method: str = "RK45",
                moll: float = 200, eps: Optional[float] = None):
    """Probability-flow ODE sampler (sampling.py:342-392): scipy solve_ivp over
    drift(x,t) = -(g(t)^2 * score * 0.5) * bump(x).  `score_fn(x, sigma)`; returns (x, nfev)."""
    from scipy import integrate
    shape = tuple(x0.shape)
    eps = sched.eps if eps is None else eps

    def bump(x):
        return ((-1 / (0.5 ** 2 - (0.5 - x).pow(2)) + 4) / moll).exp() if moll > 0 else x

    def f(t, flat):
        x = torch.from_numpy(flat.reshape(shape)).type(torch.float32)
        vec_t = torch.ones(shape[0]) * t
        g = sched.diffusion(vec_t)
        drift = torch.zeros_like(x) - g[:, None, None, None] ** 2 * score_fn(x, sched.sigma(vec_t)) * 0.5
        return (drift * bump(x)).detach().cpu().numpy().reshape((-1,))

    sol = integrate.solve_ivp(f, (sched.T, eps), x0.detach().cpu().numpy().reshape((-1,)), rtol=rtol, atol=atol,
                              method=method)
    return torch.tensor(sol.y[:, -1]).reshape(shape).type(torch.float32), sol.nfev


def ema_update(shadow: List[Tensor], params: List[Tensor], decay: float, num_updates: Optional[int]):
    """One ExponentialMovingAverage.update (models/ema.py:32-52); returns the new num_updates."""
    d = decay
    if num_updates is not None:
        num_updates += 1
        d = min(d, (1 + num_updates) / (10 + num_updates))
    for s, p in zip(shadow, params):
        s.sub_((1.0 - d) * (s - p))
    return num_updates


def gto_halo_decode(samples: np.ndarray, n_variables: int = 67) -> np.ndarray:
    """Latents -> physical units (Benchmark/gto_halo_benchmarking.py:255-328, :335-363), numpy fp32.
    Pinned: oracle/make_golden_r2.py imports the benchmark module (absent plotting / optimiser dependencies stubbed),
    drives the real GTOHaloBenchmarker.generate_samples tail on seeded latents and found this restatement
    bit-identical (tests/golden/codec.npz, REPORT_r2.txt); tests/test_oracle_golden.py re-checks it."""
    s = np.asarray(samples, dtype=np.float32).reshape(samples.shape[0], -1)[:, :n_variables]
    label = s[:, 0]
    m = s[:, 1:] * np.float32(0.1811) + np.float32(0.4652)  # global mean/std un-normalisation (:266-268)
    m[:, 0] = m[:, 0] * np.float32(40 - 0) + np.float32(0)
    m[:, 1] = m[:, 1] * np.float32(15 - 0) + np.float32(0)
    m[:, 2] = m[:, 2] * np.float32(15 - 0) + np.float32(0)
    m[:, 3:-3] = m[:, 3:-3] * np.float32(2) * np.float32(1.0) - np.float32(1.0)
    n_ctrl = (m.shape[1] - 6) // 3 * 3
    c = m[:, 3:3 + n_ctrl].reshape(m.shape[0], -1, 3)
    ux, uy, uz = c[:, :, 0].copy(), c[:, :, 1].copy(), c[:, :, 2].copy()
    u = np.sqrt(ux ** 2 + uy ** 2 + uz ** 2)
    theta = np.zeros_like(u)
    nz = u != 0
    theta[nz] = np.arcsin(uz[nz] / u[nz])
    alpha = np.arctan2(uy, ux)
    two_pi = np.float32(2 * np.pi)
    alpha = np.where(alpha >= 0, alpha, two_pi + alpha)
    theta = np.where(theta >= 0, theta, two_pi + theta)
    u = np.minimum(u, np.float32(1))
    c[:, :, 0], c[:, :, 1], c[:, :, 2] = alpha, theta, u
    m[:, 3:3 + n_ctrl] = c.reshape(m.shape[0], n_ctrl)
    m[:, -3] = m[:, -3] * np.float32(470 - 408) + np.float32(408)
    m[:, -1] = m[:, -1] * np.float32(11 - 5) + np.float32(5)
    halo = label * np.float32(0.095 - 0.008) + np.float32(0.008)
    return np.column_stack((halo, m)).astype(np.float32)


def gto_halo_encode(raw: np.ndarray, image_size: int = 9, image_width: Optional[int] = None):
    """Dataset rows -> latents (datasets.GTOHaloImageDataset.__getitem__, datasets.py:88-98), numpy fp32:
    zero-pad to H*W, z-score every entry, reshape to [N,1,H,W]; label = un-normalised first value."""
    W = image_size if image_width is None else image_width
    raw = np.asarray(raw, dtype=np.float32)
    padded = np.pad(raw, ((0, 0), (0, image_size * W - raw.shape[1])), "constant")
    padded = (padded - 0.4652) / 0.1811
    return padded.reshape(raw.shape[0], 1, image_size, W).astype(np.float32), raw[:, :1].copy()
