"""Helper functions for the unit hypercube [0, 1]^D -- drop-in for the reference's `cube` module
(/root/reference/Reflected-Diffusion/cube.py), backed by the sm_100a kernels in
csrc/elementwise.cu through the C ABI (include/rdb200.h).  Same names, arguments and return
conventions; tensors must be float32 CUDA tensors (there is no CPU path).
"""
import torch

from rdb200 import ops as _ops


def unsqueeze_as(x, y, back=True):
    """View x with trailing (or leading) singleton dims so that it broadcasts against y (cube.py:5-14)."""
    extra = (1,) * (len(y.shape) - len(x.shape))
    return x.view(*x.shape, *extra) if back else x.view(*extra, *x.shape)


def inside(x):
    """[B] bool: is every coordinate of x[b] inside [0, 1] (cube.py:17-31)."""
    return _ops.inside(x)


def reflect(x):
    """Reflect x into the unit cube; returns a new tensor (cube.py:34-49)."""
    return _ops.reflect(x)


def sample_hk(x, sigma):
    """Sample the reflected heat kernel started at x (cube.py:52-70)."""
    if not torch.is_tensor(sigma):
        sigma = sigma * torch.ones(x.shape[0]).to(x)
    return reflect(torch.randn_like(x) * unsqueeze_as(sigma, x) + x)


def score_hk(x, x_orig, sigma, efs=20, refls=10, min_cutoff=1e-2):
    """Score of the reflected heat kernel (cube.py:149-193).  The reference evaluates sigma^2/2 > min_cutoff with
    `efs` cosine modes and the rest with images |m| <= `refls`; both series are the same function, and the fused
    kernel sums whichever converged form is cheaper per sample -- truncated series
    (efs / refls below what convergence needs) and the reference's 1e-12 denominator epsilon are reproduced as
    written.  One launch, no temporaries."""
    return _ops.score_hk(x, x_orig, sigma, efs=efs, refls=refls, min_cutoff=min_cutoff)
