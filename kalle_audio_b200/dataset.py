"""GPU drop-in for the dataset-side VAE latent extraction of the reference (``twj_dataset.py:231-256``).

There every ``__getitem__`` runs, on a CPU DataLoader worker and one clip at a time:

    wav      = librosa.load(..., sr=44100, mono=True)
    norm_wav = librosa.util.normalize(wav) * 0.95                 # peak normalisation
    dual     = norm_wav.reshape(1, -1).repeat(2, 1).unsqueeze(0)  # stereo duplication, [1, 2, L]
    ms       = generator.pretransform.encode(dual)                # [1, 128, T], T = floor(L / 2048)
    mean, scale = ms.chunk(2, dim=1)
    latents, kl = vae_sample(mean, scale)                         # randn_like(mean) * scale + mean
    latents  = latents.squeeze(0).transpose(0, 1)                 # [T, 64]

``LatentExtractor.extract`` does the same for a whole LIST of clips of different lengths in a few launches: one
normalise / duplicate / pad kernel pair over the concatenated clips (``kvae_prep_mono_clips``), ONE ragged encoder pass
(``kvae_encode_ragged``: every clip sees its own end-of-clip zero padding in every layer, and lengths that are not
multiples of the downsampling ratio floor exactly as the reference's strided convolutions do), one ``vae_sample``
launch, and per-clip views of the result.  Decoding of the audio container and resampling stay with the caller.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import _lib
from .bottleneck import vae_sample


class LatentExtractor:
    def __init__(self, autoencoder, gain: float = 0.95, scale: float = 1.0, max_batch_samples: int = 64 * 442368):
        """``autoencoder``: an ``AudioAutoencoder`` (or an ``AutoencoderPretransform``, whose ``.model`` and ``.scale``
        are used) on a CUDA device.  ``max_batch_samples`` bounds clips x padded length per encoder call."""
        if hasattr(autoencoder, "model") and hasattr(autoencoder, "scale"):
            scale = float(autoencoder.scale)
            autoencoder = autoencoder.model
        self.ae = autoencoder
        self.gain, self.scale = float(gain), float(scale)
        self.ratio = int(autoencoder.downsampling_ratio)
        self.channels = int(autoencoder.in_channels)
        self.max_batch_samples = int(max_batch_samples)

    def prepare(self, wavs: Sequence[torch.Tensor]):
        """normalise * gain, duplicate to the encoder's channels, zero-pad to a common multiple of the ratio:
        ([B, channels, L_pad] fp32, lengths)."""
        dev = wavs[0].device
        lens = [int(w.numel()) for w in wavs]
        if min(lens) <= 0:
            raise ValueError("empty clip")
        flat = torch.cat([w.reshape(-1).float() for w in wavs])
        _lib.require_cuda(flat, "LatentExtractor")
        L_pad = -(-max(lens) // self.ratio) * self.ratio
        offs = torch.tensor([0] + lens[:-1], dtype=torch.int64).cumsum(0).to(dev)
        lens_d = torch.tensor(lens, dtype=torch.int32, device=dev)
        B = len(lens)
        out = torch.empty(B, self.channels, L_pad, dtype=torch.float32, device=dev)
        scratch = torch.empty(B, dtype=torch.int32, device=dev)
        _lib.check(_lib.lib().kvae_prep_mono_clips(flat.data_ptr(), offs.data_ptr(), lens_d.data_ptr(), B, L_pad,
                                                   self.channels, self.gain, out.data_ptr(), scratch.data_ptr(),
                                                   _lib.stream_ptr(dev)))
        return out, lens

    @torch.no_grad()
    def extract(self, wavs: Sequence[torch.Tensor], noise: Optional[Sequence[torch.Tensor]] = None,
                return_mean_scale: bool = False) -> List[torch.Tensor]:
        """``wavs``: mono clips (1-D CUDA tensors at the model's sample rate, any lengths).  Returns one ``[T_i, D]``
        latent tensor per clip (``T_i`` = the reference's own output length for ``len_i`` samples, about len_i / ratio).  ``noise[i]`` ([1, D, T_i] or [D, T_i]) replaces the
        ``torch.randn_like(mean)`` draw of clip i (bit-exact sampling given the reference's noise)."""
        order = sorted(range(len(wavs)), key=lambda i: -int(wavs[i].numel()))
        results: List = [None] * len(wavs)
        i = 0
        while i < len(order):
            L_pad = -(-int(wavs[order[i]].numel()) // self.ratio) * self.ratio
            n = max(1, min(len(order) - i, self.max_batch_samples // L_pad))
            idx = order[i:i + n]
            i += n
            x, lens = self.prepare([wavs[j] for j in idx])
            ms = self.ae.encoder(x, valid_len=lens)                       # [n, 2D, L_pad / ratio]
            if self.scale != 1.0:
                ms = ms / self.scale
            mean, scale = ms.chunk(2, dim=1)
            r = self.ae.encoder.runner(x.device)
            T = [r.valid_out_length(l) for l in lens]
            if noise is None:
                nz = torch.randn_like(mean)
            else:
                nz = torch.zeros_like(mean)
                for k, j in enumerate(idx):
                    nz[k, :, :T[k]] = noise[j].reshape(mean.shape[1], -1)[:, :T[k]]
            lat, _ = vae_sample(mean.contiguous(), scale.contiguous(), nz)
            for k, j in enumerate(idx):
                z = lat[k, :, :T[k]].transpose(0, 1)
                results[j] = (z, mean[k, :, :T[k]], scale[k, :, :T[k]]) if return_mean_scale else z
        return results
