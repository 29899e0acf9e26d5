"""sigma-VAE fixed-sigma latent sampler (reference model_sigmaVAE.py:150-178 method, :187-213 function).

``sample(mean, 'fix')`` = ``mean + 0.5 * randn_like(mean)``; ``'gaussian'`` draws one sigma per batch item,
``randn(B) * (0.5 / 0.8)``; anything else returns ``mean``.  Noise is drawn by torch in the same order as the
reference (for 'gaussian': the per-item sigmas first, then the element noise), and the kernel evaluates
mul-then-add with torch's two roundings, so results are bit-identical given the RNG state.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

_STD = 0.5


def _run(mean, noise, std_noise):
    dt = mean.dtype if mean.dtype in (torch.float32, torch.bfloat16) else None
    if dt is None:
        raise _lib.KvaeError(f"sample(): dtype {mean.dtype} unsupported (fp32 / bf16)")
    m = mean.contiguous()
    n = noise.to(dt).contiguous()
    out = torch.empty_like(m)
    if m.numel() == 0:
        return out
    value = float(torch.tensor(_STD) / 0.8)     # tensor(0.5) / 0.8, as the reference computes it
    sn = None if std_noise is None else std_noise.to(dt).contiguous()
    per_batch = m.numel() // m.shape[0]
    _lib.check(_lib.lib().kvae_sigma_sample(m.data_ptr(), n.data_ptr(), out.data_ptr(), m.numel(), _lib.dtype_code(dt),
                                            _STD, _lib.ptr(sn), value, per_batch, _lib.stream_ptr(mean.device)))
    return out


def sample(mean: torch.Tensor, dist_type: str = "fix", noise: Optional[torch.Tensor] = None,
           std_noise: Optional[torch.Tensor] = None) -> torch.Tensor:
    if dist_type == "fix":
        _lib.require_cuda(mean, "sample")
        if noise is None:
            noise = torch.randn_like(mean)
        return _run(mean, noise, None)
    if dist_type == "gaussian":
        _lib.require_cuda(mean, "sample")
        if std_noise is None:
            std_noise = torch.randn(mean.size(0), device=mean.device, dtype=mean.dtype)
        if noise is None:
            noise = torch.randn_like(mean)
        return _run(mean, noise, std_noise)
    return mean


class SigmaVAESampler:
    """The ``init_sigmaVAE`` / ``sample`` pair the reference mixes into its LM class (model_sigmaVAE.py:150-178)."""

    def init_sigmaVAE(self):
        self.std = torch.tensor(_STD)

    def sample(self, mean, dist_type="fix"):
        return sample(mean, dist_type)
