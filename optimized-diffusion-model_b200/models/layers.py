"""Base layers of the NCSN++ drop-in: activation lookup, DDPM-style initialisers and the parameter
containers whose names/shapes define the reference state_dict (reference models/layers.py:14-26,
39-109, 522-540).  The modules hold parameters only; the arithmetic runs in csrc/conv_gemm.cu via
the plan built in rdb200/engine.py, so `forward` on an individual layer is not provided.
"""
import numpy as np
import torch
import torch.nn as nn


def get_act(config):
    """Activation named by config.model.nonlinearity; the CUDA path implements swish/SiLU."""
    name = config.model.nonlinearity.lower()
    table = {'elu': nn.ELU, 'relu': nn.ReLU, 'swish': nn.SiLU}
    if name == 'lrelu':
        return nn.LeakyReLU(negative_slope=0.2)
    if name not in table:
        raise NotImplementedError('activation function does not exist!')
    return table[name]()


def variance_scaling(scale, mode, distribution, in_axis=1, out_axis=0, dtype=torch.float32, device='cpu'):
    """JAX-style variance-scaling initialiser factory (layers.py:39-70)."""

    def init(shape, dtype=dtype, device=device):
        receptive = np.prod(shape) / shape[in_axis] / shape[out_axis]
        fan_in, fan_out = shape[in_axis] * receptive, shape[out_axis] * receptive
        denom = {'fan_in': fan_in, 'fan_out': fan_out, 'fan_avg': (fan_in + fan_out) / 2}.get(mode)
        if denom is None:
            raise ValueError("invalid mode for variance scaling initializer: {}".format(mode))
        var = scale / denom
        if distribution == 'normal':
            return torch.randn(*shape, dtype=dtype, device=device) * np.sqrt(var)
        if distribution == 'uniform':
            return (torch.rand(*shape, dtype=dtype, device=device) * 2. - 1.) * np.sqrt(3 * var)
        raise ValueError("invalid distribution for variance scaling initializer")

    return init


def default_init(scale=1.):
    """DDPM default: uniform, fan_avg; scale 0 means 1e-10 (layers.py:73-76)."""
    return variance_scaling(1e-10 if scale == 0 else scale, 'fan_avg', 'uniform')


def _ddpm_conv(in_planes, out_planes, k, stride, padding, bias, init_scale, dilation=1):
    conv = nn.Conv2d(in_planes, out_planes, kernel_size=k, stride=stride, padding=padding, dilation=dilation, bias=bias)
    conv.weight.data = default_init(init_scale)(conv.weight.data.shape)
    if bias:
        nn.init.zeros_(conv.bias)
    return conv


def ddpm_conv1x1(in_planes, out_planes, stride=1, bias=True, init_scale=1., padding=0):
    return _ddpm_conv(in_planes, out_planes, 1, stride, padding, bias, init_scale)


def ddpm_conv3x3(in_planes, out_planes, stride=1, bias=True, dilation=1, init_scale=1., padding=1):
    return _ddpm_conv(in_planes, out_planes, 3, stride, padding, bias, init_scale, dilation)


class NIN(nn.Module):
    """Per-pixel channel mixing y = x W + b with W stored [in, out] (layers.py:531-540)."""

    def __init__(self, in_dim, num_units, init_scale=0.1):
        super().__init__()
        self.W = nn.Parameter(default_init(scale=init_scale)((in_dim, num_units)), requires_grad=True)
        self.b = nn.Parameter(torch.zeros(num_units), requires_grad=True)
