"""LM <-> VAE glue of the reference's autoregressive loop (``Llasa``, /root/reference/model_sigmaVAE.py).

The reference's ``Llasa`` owns ``audio_linear`` (latent -> LM embedding, :33-35) and ``distribution_linear`` (LM hidden
-> latent mean, an ``nn.Sequential(Linear, GELU, Linear)``, :42-50), and every generated frame of ``infer`` (:123-145)
runs ``distribution_linear -> sample('fix') -> KL stop test -> audio_linear`` as ~20 eager launches around three tiny
GEMVs.  ``LatentGlue`` carries the same two sub-modules under the same names (so the corresponding ``state_dict``
entries of a ``Llasa`` checkpoint load with ``strict=False``) and ``step`` does the whole chain in ONE launch
(``kvae_lm_glue_step``: a thread-block cluster per batch row, csrc/glue.cuh).  The Llama backbone itself is not part
of this package: the caller feeds its last hidden state in and takes the next input embedding out.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import nn

from . import _lib
from .sampling import _STD, sample


class LatentGlue(nn.Module):
    def __init__(self, latent_dim: int, audio_proj_dim: int, dtype: Optional[torch.dtype] = None):
        super().__init__()
        kw = {} if dtype is None else {"dtype": dtype}
        self.audio_linear = nn.Linear(latent_dim, audio_proj_dim, **kw)
        self.distribution_linear = nn.Sequential(nn.Linear(audio_proj_dim, latent_dim, **kw), nn.GELU(),
                                                 nn.Linear(latent_dim, latent_dim, **kw))
        self.latent_dim, self.audio_proj_dim = latent_dim, audio_proj_dim
        self.init_sigmaVAE()
        self._packed = None

    # -- the reference's sampler mix-in (model_sigmaVAE.py:150-178)
    def init_sigmaVAE(self):
        self.std = torch.tensor(_STD)

    def sample(self, mean, dist_type="fix"):
        return sample(mean, dist_type)

    def _weights(self, device):
        ps = [self.distribution_linear[0].weight, self.distribution_linear[0].bias, self.distribution_linear[2].weight,
              self.distribution_linear[2].bias, self.audio_linear.weight, self.audio_linear.bias]
        fp = tuple((p.data_ptr(), p._version) for p in ps)
        if self._packed is None or self._packed[0] != fp:
            for p in ps:
                _lib.require_cuda(p, "LatentGlue")
            self._packed = (fp, [p.detach().float().contiguous() for p in ps])
        return self._packed[1]

    @torch.no_grad()
    def step(self, last_hidden: torch.Tensor, noise: Optional[torch.Tensor] = None
             ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """One generated frame.  ``last_hidden`` [B, 1, H] (or [B, H]) -> (mean, audio_latent, audio_embed, kl_end):
        ``mean`` = distribution_linear(last_hidden), ``audio_latent`` = sample(mean) with ``noise`` (default
        ``torch.randn_like(mean)``, drawn exactly where the reference draws it), ``audio_embed`` = audio_linear(
        audio_latent) and ``kl_end`` [B, 1] = KL(N(mean, std) || N(1, e)).sum(-1) / latent_dim, the quantity
        ``infer`` compares with ``end_disp_kl_thres``."""
        _lib.require_cuda(last_hidden, "LatentGlue.step")
        shape = last_hidden.shape
        h = last_hidden.reshape(-1, shape[-1])
        if h.shape[1] != self.audio_proj_dim:
            raise ValueError(f"hidden size {h.shape[1]} != audio_proj_dim {self.audio_proj_dim}")
        h = h if h.dtype in (torch.float32, torch.bfloat16) else h.float()
        h = h.contiguous()
        out_dt = self.audio_linear.weight.dtype if self.audio_linear.weight.dtype in (torch.float32, torch.bfloat16) \
            else torch.float32
        B, D, H = h.shape[0], self.latent_dim, self.audio_proj_dim
        mean = torch.empty(B, D, dtype=out_dt, device=h.device)
        if noise is None:
            noise = torch.randn_like(mean.view(*shape[:-1], D))
        n = noise.reshape(B, D).to(out_dt).contiguous()
        latent = torch.empty_like(mean)
        embed = torch.empty(B, H, dtype=out_dt, device=h.device)
        kl = torch.empty(B, dtype=torch.float32, device=h.device)
        w1, b1, w2, b2, wa, ba = self._weights(h.device)
        _lib.check(_lib.lib().kvae_lm_glue_step(h.data_ptr(), _lib.dtype_code(h.dtype), w1.data_ptr(), b1.data_ptr(),
                                                w2.data_ptr(), b2.data_ptr(), wa.data_ptr(), ba.data_ptr(), n.data_ptr(),
                                                mean.data_ptr(), latent.data_ptr(), embed.data_ptr(), kl.data_ptr(),
                                                _lib.dtype_code(out_dt), B, H, D, float(self.std),
                                                _lib.stream_ptr(h.device)))
        lead = shape[:-1]
        return mean.view(*lead, D), latent.view(*lead, D), embed.view(*lead, H), kl.view(*lead)
