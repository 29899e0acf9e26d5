"""Latent bottleneck of the sigmaVAE path (reference stable_audio_tools/models/bottleneck.py:10-107).

The reference's ``VAEBottleneck`` was edited into an identity pass-through and ``vae_sample`` into
``randn_like(mean) * scale + mean`` (raw scale); callers split mean/scale and sample
themselves (twj_dataset.py:251-252).  Both are mirrored here.  The noise is drawn with torch on the
caller's device/generator and handed to the kernel, so the sampled latents are bit-identical to the
reference given the same RNG state.  (The reference's ``vae_sample`` also prints tensor statistics;
that debugging print is not reproduced.)
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch
from torch import nn

from . import _lib


class Bottleneck(nn.Module):
    def __init__(self, is_discrete: bool = False):
        super().__init__()
        self.is_discrete = is_discrete

    def encode(self, x, return_info=False, **kwargs):
        raise NotImplementedError

    def decode(self, x):
        raise NotImplementedError


def vae_sample(mean: torch.Tensor, scale: torch.Tensor, noise: Optional[torch.Tensor] = None):
    """latents = noise*scale + mean ; kl = (mean^2 + var - log var - 1).sum(1).mean() with
    stdev = softplus(scale) + 1e-4 (bottleneck.py:51-62).  ``noise`` defaults to ``torch.randn_like(mean)``."""
    _lib.require_cuda(mean, "vae_sample")
    if mean.shape != scale.shape or mean.dim() != 3:
        raise ValueError("mean and scale must both be [B, D, T]")
    if noise is None:
        noise = torch.randn_like(mean)
    dt = mean.dtype if mean.dtype in (torch.float32, torch.bfloat16) else torch.float32
    m, s, n = (t.to(dt).contiguous() for t in (mean, scale, noise))
    out = torch.empty_like(m)
    kl = torch.empty((), dtype=torch.float32, device=mean.device)
    scratch = torch.empty(8 * 1024, dtype=torch.uint8, device=mean.device)
    B, D, T = m.shape
    _lib.check(_lib.lib().kvae_vae_sample(m.data_ptr(), s.data_ptr(), n.data_ptr(), out.data_ptr(), kl.data_ptr(), B, D,
                                          T, _lib.dtype_code(dt), scratch.data_ptr(), _lib.stream_ptr(mean.device)))
    return out.to(mean.dtype), kl.to(mean.dtype)


class VAEBottleneck(Bottleneck):
    """Identity at inference time, exactly as the reference's edited class (bottleneck.py:85-107)."""

    def __init__(self):
        super().__init__(is_discrete=False)

    def encode(self, x, return_info=False, **kwargs):
        info: Dict[str, Any] = {}
        if return_info:
            return x, info
        return x

    def decode(self, x):
        return x


def create_bottleneck_from_config(bottleneck_config):
    """factory.py:112-153 restricted to the bottleneck the sigmaVAE path uses."""
    bottleneck_type = bottleneck_config.get("type", None)
    assert bottleneck_type is not None, "type must be specified in bottleneck config"
    if bottleneck_type == "vae":
        bottleneck = VAEBottleneck()
    else:
        raise NotImplementedError(f"bottleneck type {bottleneck_type!r} is outside the sigmaVAE hot path")
    if not bottleneck_config.get("requires_grad", True):
        for p in bottleneck.parameters():
            p.requires_grad = False
    return bottleneck
