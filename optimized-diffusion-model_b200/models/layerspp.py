"""NCSN++ building blocks of the drop-in (reference models/layerspp.py:19-28, 67-214): the reference's
attribute names (so state_dicts match).  Inside NCSNpp their arithmetic is executed by the fused CUDA
plan (GN+SiLU -> tcgen05 conv -> temb/skip epilogue), not module by module; called on their own,
`ResnetBlockDDPMpp.forward` / `AttnBlockpp.forward` run the same kernels through the unit entry
points `rd_resblock` / `rd_attn_block` (rdb200/unit.py).  CUDA only, inference only.
"""
import torch
import torch.nn as nn

from . import layers

conv1x1 = layers.ddpm_conv1x1
conv3x3 = layers.ddpm_conv3x3
NIN = layers.NIN
default_init = layers.default_init


def _gn(ch):
    return nn.GroupNorm(num_groups=min(ch // 4, 32), num_channels=ch, eps=1e-6)


class GaussianFourierProjection(nn.Module):
    """Frozen random Fourier features of log(sigma) (layerspp.py:19-28)."""

    def __init__(self, embedding_size=256, scale=1.0):
        super().__init__()
        self.W = nn.Parameter(torch.randn(embedding_size) * scale, requires_grad=False)


class AttnBlockpp(nn.Module):
    """Single-head self-attention over pixels: GroupNorm_0, NIN_0..2 (q,k,v), NIN_3 (out)."""

    def __init__(self, channels, skip_rescale=False, init_scale=0.):
        super().__init__()
        self.GroupNorm_0 = _gn(channels)
        self.NIN_0, self.NIN_1, self.NIN_2 = NIN(channels, channels), NIN(channels, channels), NIN(channels, channels)
        self.NIN_3 = NIN(channels, channels, init_scale=init_scale)
        self.skip_rescale = skip_rescale

    def forward(self, x):
        """x [B,C,H,W] fp32 CUDA -> (x + attn(x)) [/ sqrt 2] (layerspp.py:80-96), one fused kernel."""
        from rdb200 import unit
        return unit.attnblock_forward(self, x)


class _Resample(nn.Module):
    def __init__(self, in_ch, out_ch, with_conv, fir, fir_kernel, stride, padding):
        super().__init__()
        if fir:
            raise NotImplementedError('fir resampling (up_or_down_sampling.py) is outside the B200 hot path; '
                                      'the GTO-Halo configuration uses fir: false')
        out_ch = out_ch if out_ch else in_ch
        if not with_conv:
            raise NotImplementedError('resamp_with_conv: false is not supported by the B200 path')
        self.Conv_0 = conv3x3(in_ch, out_ch, stride=stride, padding=padding)
        self.fir, self.with_conv, self.fir_kernel, self.out_ch = fir, with_conv, fir_kernel, out_ch


class Upsample(_Resample):
    """Nearest x2 followed by Conv_0 (layerspp.py:99-131)."""

    def __init__(self, in_ch=None, out_ch=None, with_conv=False, fir=False, fir_kernel=(1, 3, 3, 1)):
        super().__init__(in_ch, out_ch, with_conv, fir, fir_kernel, stride=1, padding=1)


class Downsample(_Resample):
    """Pad (0,1,0,1) then stride-2 Conv_0 without padding (layerspp.py:134-168)."""

    def __init__(self, in_ch=None, out_ch=None, with_conv=False, fir=False, fir_kernel=(1, 3, 3, 1)):
        super().__init__(in_ch, out_ch, with_conv, fir, fir_kernel, stride=2, padding=0)


class ResnetBlockDDPMpp(nn.Module):
    """GroupNorm_0 -> act -> Conv_0 -> + Dense_0(act(temb)) -> GroupNorm_1 -> act -> Conv_1, NIN_0 shortcut
    when the channel count changes, optional 1/sqrt(2) rescale (layerspp.py:171-214)."""

    def __init__(self, act, in_ch, out_ch=None, temb_dim=None, conv_shortcut=False, dropout=0.1, skip_rescale=False,
                 init_scale=0.):
        super().__init__()
        out_ch = out_ch if out_ch else in_ch
        if conv_shortcut:
            raise NotImplementedError('conv_shortcut is never used by NCSNpp (ncsnpp.py:140)')
        self.GroupNorm_0 = _gn(in_ch)
        self.Conv_0 = conv3x3(in_ch, out_ch)
        if temb_dim is not None:
            self.Dense_0 = nn.Linear(temb_dim, out_ch)
            self.Dense_0.weight.data = default_init()(self.Dense_0.weight.data.shape)
            nn.init.zeros_(self.Dense_0.bias)
        self.GroupNorm_1 = _gn(out_ch)
        self.Dropout_0 = nn.Dropout(dropout)
        self.Conv_1 = conv3x3(out_ch, out_ch, init_scale=init_scale)
        if in_ch != out_ch:
            self.NIN_0 = NIN(in_ch, out_ch)
        self.skip_rescale, self.act, self.out_ch, self.conv_shortcut = skip_rescale, act, out_ch, conv_shortcut
        self.rd_precision = "bf16"  # "fp32": the fp32-class kernels (split-bf16 tensor-core operands)

    def forward(self, x, temb=None):
        """x [B,C_in,H,W] fp32 CUDA, temb [B,temb_dim] or None -> [B,C_out,H,W] (layerspp.py:198-214, eval mode)."""
        if not isinstance(self.act, nn.SiLU):
            raise NotImplementedError('the B200 kernels fuse swish/SiLU only')
        from rdb200 import unit
        return unit.resblock_forward(self, x, temb, precision=self.rd_precision)
