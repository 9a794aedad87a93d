"""NCSN++ score network -- drop-in for the reference's `models.ncsnpp.NCSNpp`
(/root/reference/Reflected-Diffusion/models/ncsnpp.py:22-354).

Same constructor (a Hydra-style `config` object), same parameter names and shapes (so
`load_state_dict` of reference checkpoints and `ExponentialMovingAverage.copy_to/restore` work),
same call `model(x, time_cond, class_labels=None)` where `time_cond` is the noise level sigma.
The forward pass is NOT a stack of torch modules: the parameters are packed into kernel layouts
(rdb200/pack.py) and the network is executed as one plan of hand-written sm_100a kernels
(rdb200/engine.py -> librdb200.so).  CUDA only -- a CPU tensor raises.
"""
import torch
import torch.nn as nn

from . import layers, layerspp, utils
from rdb200.engine import PRECISIONS, ForwardEngine, SamplerEngine, spec_from_config
from rdb200.pack import PackedWeights
from rdb200 import ops as _rd_ops

ResnetBlockDDPM = layerspp.ResnetBlockDDPMpp
conv3x3 = layerspp.conv3x3
get_act = layers.get_act
default_initializer = layers.default_init


@utils.register_model(name='ncsnpp')
class NCSNpp(nn.Module):
    """U-Net over [B, C, H, W] latents with Fourier noise-level embedding, scalar-label conditioning,
    skip concatenation and optional pixel attention."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        m = config.model
        self.nf = nf = m.nf
        self.ch_mult = ch_mult = m.ch_mult
        self.num_res_blocks = num_res_blocks = m.num_res_blocks
        self.attn_resolutions = attn_resolutions = m.attn_resolutions
        self.dropout = dropout = m.dropout
        self.resamp_with_conv = m.resamp_with_conv
        self.embedding_type = m.embedding_type
        self.conditional = m.conditional
        self.cond_drop_prob = m.cond_drop_prob if hasattr(m, 'cond_drop_prob') else 0.0
        self.num_classes = getattr(m, 'num_classes', 1)
        self.init_scale = m.init_scale
        self.skip_rescale = m.skip_rescale
        self.fir = m.fir
        self.fir_kernel = tuple(m.fir_kernel)
        self.image_size = m.image_size
        self.image_width = getattr(m, 'image_width', m.image_size)  # read but unused, like the reference
        self.channels = m.channels
        self.scale_by_sigma = getattr(m, 'scale_by_sigma', False)
        self.act = get_act(config)
        if not isinstance(self.act, nn.SiLU):
            raise NotImplementedError('the B200 kernels fuse swish/SiLU only (configs/model/ncsnpp.yaml: swish)')
        if self.embedding_type != 'fourier':
            raise NotImplementedError('Only fourier embedding supported')

        block = dict(temb_dim=nf * 4, dropout=dropout, skip_rescale=self.skip_rescale, init_scale=self.init_scale)
        attn = dict(skip_rescale=self.skip_rescale, init_scale=self.init_scale)
        resample = dict(with_conv=self.resamp_with_conv, fir=self.fir, fir_kernel=self.fir_kernel)

        self.time_embed = layerspp.GaussianFourierProjection(embedding_size=nf, scale=m.fourier_scale)
        self.time_mlp = nn.Sequential(nn.Linear(2 * nf, nf * 4), self.act, nn.Linear(nf * 4, nf * 4))
        if self.conditional:
            self.label_emb = nn.Linear(self.num_classes, nf * 4)
        self.input_conv = conv3x3(self.channels, nf)

        # encoder: num_res_blocks blocks per level; one extra skip per level for the decoder's extra block
        self.down_blocks, self.down_attn, self.downsample = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        self.skip_channels = []
        width = nf
        levels = len(ch_mult)
        for lvl, mult in enumerate(ch_mult):
            for _ in range(num_res_blocks):
                self.down_blocks.append(ResnetBlockDDPM(self.act, width, nf * mult, **block))
                width = nf * mult
                use_attn = (self.image_size // (2 ** lvl)) in attn_resolutions
                self.down_attn.append(layerspp.AttnBlockpp(width, **attn) if use_attn else None)
                self.skip_channels.append(width)
            self.skip_channels.append(width)
            self.downsample.append(layerspp.Downsample(width, **resample) if lvl != levels - 1 else None)
        assert len(self.skip_channels) == levels * (num_res_blocks + 1)

        self.mid_block1 = ResnetBlockDDPM(self.act, width, width, **block)
        self.mid_attn = (layerspp.AttnBlockpp(width, **attn)
                         if (self.image_size // (2 ** (levels - 1))) in attn_resolutions else None)
        self.mid_block2 = ResnetBlockDDPM(self.act, width, width, **block)

        # decoder: consumes the skips last-in first-out
        self.up_blocks, self.up_attn, self.upsample = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        pending = list(self.skip_channels)
        for lvl in reversed(range(levels)):
            for _ in range(num_res_blocks + 1):
                self.up_blocks.append(ResnetBlockDDPM(self.act, width + pending.pop(), nf * ch_mult[lvl], **block))
                width = nf * ch_mult[lvl]
                use_attn = (self.image_size // (2 ** lvl)) in attn_resolutions
                self.up_attn.append(layerspp.AttnBlockpp(width, **attn) if use_attn else None)
            self.upsample.append(layerspp.Upsample(width, **resample) if lvl != 0 else None)

        self.out_norm = nn.GroupNorm(num_groups=min(width // 4, 32), num_channels=width, eps=1e-6)
        self.out_act = self.act
        self.out_conv = conv3x3(width, self.channels, init_scale=self.init_scale)

        # ---- B200 execution state (not part of the state_dict)
        self._rd_spec = spec_from_config(config)
        self._rd_packed = None
        self._rd_token = None          # content checksum of the parameters the packed weights were made from
        self._rd_segments = None       # (data_ptr key, device segment table, device result) of the checksum kernel
        self._rd_frozen = 0            # > 0 inside a sampling loop: parameters cannot change, skip the check
        self._rd_forward_engines = {}
        self._rd_sampler_engines = {}

    # ------------------------------------------------------------------ weight packing
    def _rd_device(self):
        return self.input_conv.weight.device

    @property
    def rd_precision(self):
        return self._rd_spec.precision

    def rd_set_precision(self, precision):
        """'bf16' (default): bf16 activations / operands.  'fp32': the fp32-class plan (fp32 activations, split-bf16
        tensor-core operands, fp32 attention core).  Also settable as config.model.rd_precision / RDB200_PRECISION."""
        if precision not in PRECISIONS:
            raise ValueError(f"unknown precision {precision!r}: expected one of {sorted(PRECISIONS)}")
        if precision != self._rd_spec.precision:
            self._rd_spec.precision = precision
            self._rd_packed, self._rd_token = None, None
            self._rd_forward_engines.clear()
            self._rd_sampler_engines.clear()
        return self

    def _rd_weights_token(self):
        """64-bit checksum of every parameter's values and positions: one kernel launch + an 8-byte read (EMA
        copy_to / restore and load_state_dict write through `.data` / copy_ in place, which no version counter of
        the Parameter objects records)."""
        ps = [p.detach() for p in self.parameters()]
        key = tuple((p.data_ptr(), p.numel()) for p in ps)
        if self._rd_segments is None or self._rd_segments[0] != key:
            for p in ps:
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError('NCSNpp (B200) expects contiguous float32 parameters')
            self._rd_segments = (key,) + _rd_ops.checksum_segments(ps)
        return _rd_ops.checksum(self._rd_segments[1], self._rd_segments[2])

    @torch.no_grad()
    def rd_sync_weights(self, force=False):
        """(Re)pack the parameters into kernel layouts if their content changed; returns True when repacked."""
        dev = self._rd_device()
        if dev.type != 'cuda':
            raise RuntimeError('NCSNpp (B200) has no CPU path: move the model to a CUDA device')
        if self._rd_frozen and self._rd_packed is not None and not force:
            return False
        token = self._rd_weights_token()
        if not force and self._rd_packed is not None and self._rd_packed.device == dev and token == self._rd_token:
            return False
        x3 = self._rd_spec.precision == 'fp32'
        if self._rd_packed is None or self._rd_packed.device != dev or self._rd_packed.x3 != x3:
            self._rd_packed = PackedWeights(dev, x3=x3)
            self._rd_forward_engines.clear()
            self._rd_sampler_engines.clear()
        self._rd_packed.update(self.state_dict(), self._rd_spec.res_blocks(), self._rd_spec.attn_blocks())
        self._rd_token = token
        for eng in self._rd_sampler_engines.values():
            eng.refresh_tables()
        return True

    def rd_freeze_weights(self):
        """Context manager for loops that call the model many times while nothing can touch its parameters (the
        generic update_fn sampler loop): the content check runs once on entry instead of once per call."""
        model = self

        class _Frozen:
            def __enter__(self_inner):
                model.rd_sync_weights()
                model._rd_frozen += 1

            def __exit__(self_inner, *exc):
                model._rd_frozen -= 1
                return False
        return _Frozen()

    # ------------------------------------------------------------------ reference call surface
    def forward(self, x, time_cond, class_labels=None):
        """x [B,C,H,W] fp32 CUDA, time_cond [B] noise levels sigma, class_labels [B,num_classes]."""
        if not x.is_cuda:
            raise RuntimeError('NCSNpp (B200) has no CPU path: inputs must be CUDA tensors')
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and self.training:
            raise NotImplementedError('the B200 path implements inference only (training is out of scope)')
        if self.conditional and self.training and self.cond_drop_prob > 0:
            keep = (torch.rand(x.shape[0], device=x.device) >= self.cond_drop_prob).float().unsqueeze(1)
            class_labels = class_labels * keep
        self.rd_sync_weights()
        B, _, H, W = x.shape
        key = (B, H, W, x.device.index)
        eng = self._rd_forward_engines.get(key)
        if eng is None:
            eng = self._rd_forward_engines[key] = ForwardEngine(self._rd_spec, self._rd_packed, B, H, W, x.device)
        return eng(x.float(), time_cond.float(), class_labels)

    # ------------------------------------------------------------------ native fast paths
    def rd_guided_score(self, x, sigma, class_labels, weight):
        """(1+w) net(x,sigma,c) - w net(x,sigma,0) in one plan (both passes batched, combine fused)."""
        B = x.shape[0]
        s2 = self.forward(x.repeat(2, 1, 1, 1), sigma.repeat(2),
                          torch.cat([class_labels, torch.zeros_like(class_labels)], dim=0))
        from rdb200 import ops
        return ops.cfg_combine(s2, weight)

    def rd_sampler_engine(self, B, H, W, device, sde, eps, snr, n_corrector_steps, cfg=True):
        self.rd_sync_weights()
        key = (B, H, W, torch.device(device).index, sde.N, float(sde.sigma_min), float(sde.sigma_max), float(sde.T),
               float(eps), float(snr), int(n_corrector_steps), bool(cfg))
        eng = self._rd_sampler_engines.get(key)
        if eng is None:
            eng = SamplerEngine(self._rd_spec, self._rd_packed, B, H, W, device, sde, eps, snr, n_corrector_steps, cfg=cfg)
            eng.refresh_tables()
            self._rd_sampler_engines[key] = eng
        return eng
