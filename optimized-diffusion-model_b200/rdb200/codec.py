"""Latent -> physical-unit codec of the GTO-Halo benchmark, the consumer directly downstream of the
sampler (reference Benchmark/gto_halo_benchmarking.py:255-328 `generate_samples` tail and :335-363
`_convert_to_spherical`).  One CUDA kernel, one thread per sample (csrc/next_rows.cu)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from ._lib import check, lib, ptr, require_cuda_f32, stream_ptr
from .cdefs import GtoHaloCodec


@dataclass
class GtoHaloConstants:
    """The constants hard-coded at gto_halo_benchmarking.py:266-280."""
    data_mean: float = 0.4652
    data_std: float = 0.1811
    min_shooting_time: float = 0
    max_shooting_time: float = 40
    min_coast_time: float = 0
    max_coast_time: float = 15
    min_halo_energy: float = 0.008
    max_halo_energy: float = 0.095
    min_final_fuel_mass: float = 408
    max_final_fuel_mass: float = 470
    min_manifold_length: float = 5
    max_manifold_length: float = 11
    thrust: float = 1.0
    n_variables: int = 67  # values kept per sample (label + 66 model outputs)

    def c_struct(self) -> GtoHaloCodec:
        n_ctrl = self.n_variables - 1 - 3 - 3
        if n_ctrl < 0:
            raise ValueError("n_variables must be at least 7")
        if n_ctrl % 3 != 0:
            # the reference converts whole (ux, uy, uz) triplets and leaves stragglers in cartesian form in the row
            # (gto_halo_benchmarking.py:295-298, :314); the kernel's row layout has no slot for them
            raise ValueError(f"n_variables - 7 = {n_ctrl} control values do not form whole (ux, uy, uz) triplets")
        return GtoHaloCodec(
            self.data_mean, self.data_std,
            self.min_shooting_time, self.max_shooting_time - self.min_shooting_time,
            self.min_coast_time, self.max_coast_time - self.min_coast_time,
            self.min_halo_energy, self.max_halo_energy - self.min_halo_energy,
            self.min_final_fuel_mass, self.max_final_fuel_mass - self.min_final_fuel_mass,
            self.min_manifold_length, self.max_manifold_length - self.min_manifold_length,
            self.thrust, n_ctrl // 3)


def gto_halo_decode(samples: torch.Tensor, consts: GtoHaloConstants = GtoHaloConstants()) -> torch.Tensor:
    """samples: [N, ...] fp32 CUDA latents straight from the sampler (e.g. [N,1,9,9] or [N,1,8,9]);
    returns [N, 7 + 3*n_triplets] physical variables on the same device (n_variables - 7 must be a multiple of 3)."""
    x = require_cuda_f32(samples, "samples")
    n = x.shape[0]
    stride = x.numel() // max(n, 1)
    cs = consts.c_struct()
    width = 7 + 3 * cs.n_triplets
    if stride < consts.n_variables:
        raise ValueError(f"samples carry {stride} values each; the codec needs {consts.n_variables}")
    out = torch.empty((n, width), dtype=torch.float32, device=x.device)
    check(lib().rd_gto_halo_decode_f32(ptr(x), ptr(out), n, stride, C.byref(cs), stream_ptr(x.device)),
          "rd_gto_halo_decode_f32")
    return out


def gto_halo_encode(raw: torch.Tensor, image_size: int = 9, image_width: int = None,
                    consts: GtoHaloConstants = GtoHaloConstants()):
    """Dataset rows -> latents, the inverse direction (reference datasets.GTOHaloImageDataset.__getitem__,
    datasets.py:88-98): raw [N, n_values <= H*W] fp32 CUDA -> (latents [N, 1, H, W], labels [N, 1]); rows are zero-padded
    to H*W and z-scored with the data mean / std, the label is the un-normalised first value."""
    x = require_cuda_f32(raw, "raw")
    if x.dim() != 2:
        raise ValueError("raw must be [N, n_values]")
    W = image_size if image_width is None else image_width
    n, n_in, n_lat = x.shape[0], x.shape[1], image_size * W
    if n_in > n_lat:
        raise ValueError(f"{n_in} values per row do not fit a {image_size}x{W} latent")
    lat = torch.empty((n, 1, image_size, W), dtype=torch.float32, device=x.device)
    lab = torch.empty((n, 1), dtype=torch.float32, device=x.device)
    check(lib().rd_gto_halo_encode_f32(ptr(x), ptr(lat), ptr(lab), n, n_in, n_lat, consts.data_mean, consts.data_std,
                                       stream_ptr(x.device)), "rd_gto_halo_encode_f32")
    return lat, lab
