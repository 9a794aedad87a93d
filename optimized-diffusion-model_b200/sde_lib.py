"""Forward / reverse SDE objects of the reflected VE diffusion -- drop-in for the reference's
`sde_lib` module (/root/reference/Reflected-Diffusion/sde_lib.py:7-161).

These classes are host-side bookkeeping (a handful of [B]-sized tensor ops); the sampler's hot
loop does not call them per step -- it reads the same quantities from device tables built once
by `RVESDE.step_tables` and indexed by a device-side step counter inside the captured graph.
"""
import abc

import numpy as np
import torch


class SDE(abc.ABC):
    """Interface of a forward SDE on mini-batches (sde_lib.py:7-69)."""

    def __init__(self, N):
        super().__init__()
        self.N = N  # number of discretisation steps

    @property
    @abc.abstractmethod
    def T(self):
        """Terminal time."""

    @abc.abstractmethod
    def sde(self, x, t):
        """(drift, diffusion) at (x, t)."""

    @abc.abstractmethod
    def marginal_prob(self, x, t):
        """(mean, std) of the perturbation kernel p_t(x_t | x_0 = x)."""

    @abc.abstractmethod
    def prior_sampling(self, shape):
        """One draw from p_T."""

    @abc.abstractmethod
    def prior_logp(self, z):
        """log p_T(z)."""

    def discretize(self, x, t):
        """Euler-Maruyama coefficients: x_{i+1} = x_i + f + G z."""
        h = 1 / self.N
        drift, diffusion = self.sde(x, t)
        return drift * h, diffusion * torch.sqrt(torch.tensor(h, device=t.device))

    def reverse(self, score_fn, probability_flow=False):
        """Reverse-time SDE (or probability-flow ODE) driven by `score_fn(x, t)` (sde_lib.py:71-111)."""
        return ReverseSDE(self, score_fn, probability_flow)


class ReverseSDE:
    """What `SDE.reverse` returns: same `.sde / .discretize / .N / .T` surface as the reference's
    dynamically built RSDE class (sde_lib.py:83-109)."""

    def __init__(self, forward, score_fn, probability_flow):
        self.forward = forward
        self.score_fn = score_fn
        self.probability_flow = probability_flow
        self.N = forward.N

    @property
    def T(self):
        return self.forward.T

    def _weight(self):
        return 0.5 if self.probability_flow else 1.0

    def sde(self, x, t):
        drift, g = self.forward.sde(x, t)
        drift = drift - g[:, None, None, None] ** 2 * self.score_fn(x, t) * self._weight()
        return drift, (torch.zeros_like(g) if self.probability_flow else g)

    def discretize(self, x, t):
        f, G = self.forward.discretize(x, t)
        f = f - G[:, None, None, None] ** 2 * self.score_fn(x, t) * self._weight()
        return f, (torch.zeros_like(G) if self.probability_flow else G)

    def marginal_prob(self, x, t):
        return self.forward.marginal_prob(x, t)

    def prior_sampling(self, shape):
        return self.forward.prior_sampling(shape)

    def prior_logp(self, z):
        return self.forward.prior_logp(z)


class RVESDE(SDE):
    """Reflected variance-exploding SDE, sigma(t) = sigma_min (sigma_max/sigma_min)^t (sde_lib.py:114-161)."""

    def __init__(self, sigma_min=0.01, sigma_max=50, N=1000, T=1):
        super().__init__(N)
        self.sigma_min = sigma_min
        self.sigma_max = sigma_max
        self.T_val = T
        self.discrete_sigmas = torch.exp(torch.linspace(np.log(sigma_min), np.log(sigma_max), N))

    @property
    def T(self):
        return self.T_val

    def _sigma(self, t):
        return self.sigma_min * (self.sigma_max / self.sigma_min) ** t

    def sde(self, x, t):
        scale = torch.sqrt(torch.tensor(2 * (np.log(self.sigma_max) - np.log(self.sigma_min)),
                                        device=t.device, dtype=torch.float32))
        return torch.zeros_like(x), self._sigma(t) * scale

    def marginal_prob(self, x, t):
        return x, self._sigma(t)

    def prior_sampling(self, shape):
        return torch.rand(*shape)

    def prior_logp(self, z):
        return torch.zeros_like(z)

    def discretize(self, x, t):
        """SMLD-style discretisation on the geometric sigma ladder (sde_lib.py:153-161)."""
        idx = (t * (self.N - 1) / self.T).long()
        ladder = self.discrete_sigmas.to(t.device)
        below = torch.where(idx == 0, torch.zeros_like(t), ladder[idx - 1])
        return torch.zeros_like(x), torch.sqrt(ladder[idx] ** 2 - below ** 2)

    # ---- B200 path: per-step tables for the graph-captured sampler --------------------------
    def step_tables(self, eps):
        """(t_i, sigma_i, g_i) on the sampler grid linspace(T, eps, N) (sampling.py:325), evaluated on
        the CPU with exactly the tensor ops of `sde` / `marginal_prob` so the values are bit-identical
        to what the reference's per-step code would compute."""
        t = torch.linspace(self.T, eps, self.N)
        _, g = self.sde(torch.zeros(self.N, 1, 1, 1), t)
        return t, self._sigma(t), g
