"""CPU oracle for the Oobleck / sigmaVAE autoencoder hot path.  TEST INFRASTRUCTURE ONLY.

This is a plain restatement of the reference's algorithm, written against a *state_dict with the
reference's key names*, with no nn.Module machinery.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the product package
(``kalle_audio_b200``) never does and fails loudly when its CUDA library is missing.

Parity pin: the reference ships no tests or golden vectors, so the oracle is
pinned against outputs of the reference's own modules, imported from /root/reference in the build
container by ``tests/golden/make_golden.py`` and committed as fixtures under ``tests/golden/``
(see ``tests/test_oracle_golden.py``).

Reference lines each function follows (relative to /root/reference):
  weight_norm_fold      dac.nn.layers.WNConv1d == torch.nn.utils.weight_norm(nn.Conv1d), dim=0
                        (un-vendored third party; call sites stable_audio_tools/models/autoencoders.py:9,49,52,76,98)
  snake_beta            stable_audio_tools/models/blocks.py:301-302, 331-339
  residual_unit         stable_audio_tools/models/autoencoders.py:39-62
  encoder_block         autoencoders.py:64-81
  decoder_block         autoencoders.py:83-114
  oobleck_encoder       autoencoders.py:116-147
  oobleck_decoder       autoencoders.py:150-191
  decode_audio/encode_audio (chunked)   autoencoders.py:429-560
  vae_sample            stable_audio_tools/models/bottleneck.py:51-62
  sigma_sample          model_sigmaVAE.py:153-178, 187-213
  training_loss         the generator branch of AutoencoderTrainingWrapper.training_step
                        (stable_audio_tools/training/autoencoders.py:221-352: encode -> bottleneck -> decode ->
                        losses), KL wired as in :446-456; gradients come from torch autograd through this
                        restatement, exactly as the reference gets them from autograd through its modules
  gaussian_nll          sigma-VAE reconstruction term of BASELINE config 5; NOT defined in the reference tree
                        (dangling sigma-vae-pytorch symlink): standard form with a fixed
                        scalar sigma, summed per clip, averaged over the batch
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def weight_norm_fold(v: Tensor, g: Tensor) -> Tensor:
    """w = v * (g / ||v||_2), the norm taken over every dim but 0 (old-style weight_norm, dim=0).

    For Conv1d ``v`` is [Cout, Cin, K] (norm per out-channel); for ConvTranspose1d ``v`` is
    [Cin, Cout, K], so the norm is per *in*-channel (autoencoders.py:98).
    """
    norm = v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
    return v * (g / norm)


def snake_beta(x: Tensor, alpha: Tensor, beta: Tensor, logscale: bool = True) -> Tensor:
    a = alpha.view(1, -1, 1)
    b = beta.view(1, -1, 1)
    if logscale:
        a = torch.exp(a)
        b = torch.exp(b)
    return x + (1.0 / (b + 0.000000001)) * torch.sin(x * a) ** 2


def _wn_conv1d(sd: Dict[str, Tensor], key: str, x: Tensor, stride=1, padding=0, dilation=1) -> Tensor:
    w = weight_norm_fold(sd[key + ".weight_v"], sd[key + ".weight_g"])
    return F.conv1d(x, w, sd.get(key + ".bias"), stride=stride, padding=padding, dilation=dilation)


def _wn_conv_transpose1d(sd: Dict[str, Tensor], key: str, x: Tensor, stride: int, padding: int) -> Tensor:
    w = weight_norm_fold(sd[key + ".weight_v"], sd[key + ".weight_g"])
    return F.conv_transpose1d(x, w, sd.get(key + ".bias"), stride=stride, padding=padding)


def _snake(sd: Dict[str, Tensor], key: str, x: Tensor) -> Tensor:
    return snake_beta(x, sd[key + ".alpha"], sd[key + ".beta"])


def residual_unit(sd: Dict[str, Tensor], key: str, x: Tensor, dilation: int) -> Tensor:
    h = _snake(sd, f"{key}.layers.0", x)
    h = _wn_conv1d(sd, f"{key}.layers.1", h, padding=(dilation * 6) // 2, dilation=dilation)
    h = _snake(sd, f"{key}.layers.2", h)
    h = _wn_conv1d(sd, f"{key}.layers.3", h)
    return h + x


def encoder_block(sd: Dict[str, Tensor], key: str, x: Tensor, stride: int) -> Tensor:
    for j, d in enumerate((1, 3, 9)):
        x = residual_unit(sd, f"{key}.layers.{j}", x, d)
    x = _snake(sd, f"{key}.layers.3", x)
    return _wn_conv1d(sd, f"{key}.layers.4", x, stride=stride, padding=math.ceil(stride / 2))


def decoder_block(sd: Dict[str, Tensor], key: str, x: Tensor, stride: int, use_nearest_upsample: bool = False) -> Tensor:
    x = _snake(sd, f"{key}.layers.0", x)
    if use_nearest_upsample:     # autoencoders.py:87-96: Upsample(nearest) -> WNConv1d(k = 2 stride, 'same', no bias)
        x = F.interpolate(x, scale_factor=stride, mode="nearest")
        x = _wn_conv1d(sd, f"{key}.layers.1.1", x, padding="same")
    else:
        x = _wn_conv_transpose1d(sd, f"{key}.layers.1", x, stride=stride, padding=math.ceil(stride / 2))
    for j, d in enumerate((1, 3, 9)):
        x = residual_unit(sd, f"{key}.layers.{2 + j}", x, d)
    return x


def oobleck_encoder(sd: Dict[str, Tensor], x: Tensor, strides: Sequence[int], prefix: str = "") -> Tensor:
    """x [B, C_io, L] -> [B, latent_dim, L / prod(strides)]; ``sd`` keys are ``{prefix}layers.N...``."""
    p = prefix
    x = _wn_conv1d(sd, f"{p}layers.0", x, padding=3)
    for i, s in enumerate(strides):
        x = encoder_block(sd, f"{p}layers.{1 + i}", x, s)
    n = len(strides)
    x = _snake(sd, f"{p}layers.{1 + n}", x)
    return _wn_conv1d(sd, f"{p}layers.{2 + n}", x, padding=1)


def oobleck_decoder(sd: Dict[str, Tensor], z: Tensor, strides: Sequence[int], prefix: str = "",
                    final_tanh: bool = False, use_nearest_upsample: bool = False) -> Tensor:
    """z [B, latent_dim, T] -> [B, C_io, T * prod(strides)].  ``strides`` in *encoder* order
    (the decoder walks them reversed, autoencoders.py:171-180)."""
    p = prefix
    x = _wn_conv1d(sd, f"{p}layers.0", z, padding=3)
    n = len(strides)
    for i in range(n):
        x = decoder_block(sd, f"{p}layers.{1 + i}", x, strides[n - 1 - i], use_nearest_upsample)
    x = _snake(sd, f"{p}layers.{1 + n}", x)
    x = _wn_conv1d(sd, f"{p}layers.{2 + n}", x, padding=3)   # bias=False in the reference: no bias key
    return torch.tanh(x) if final_tanh else x


# ----------------------------------------------------------------------------- chunked paths
def decode_audio_chunked(decode_fn, latents: Tensor, downsampling_ratio: int, out_channels: int,
                         overlap: int = 32, chunk_size: int = 128) -> Tensor:
    """autoencoders.py:514-560 (chunked=True branch), incl. its fp32 y_final and edge trimming."""
    hop = chunk_size - overlap
    total, bsz = latents.shape[2], latents.shape[0]
    if total < chunk_size:
        raise UnboundLocalError("latent length shorter than chunk_size (reference raises here too)")
    starts = list(range(0, total - chunk_size + 1, hop))
    chunks = [latents[:, :, i:i + chunk_size] for i in starts]
    if starts[-1] + chunk_size != total:
        chunks.append(latents[:, :, -chunk_size:])
    n = len(chunks)
    spl = downsampling_ratio
    y_size = total * spl
    y = torch.zeros((bsz, out_channels, y_size), device=latents.device)
    for i, ch in enumerate(chunks):
        yc = decode_fn(ch)
        if i == n - 1:
            t_end = y_size
            t_start = t_end - yc.shape[2]
        else:
            t_start = i * hop * spl
            t_end = t_start + chunk_size * spl
        ol = (overlap // 2) * spl
        c0, c1 = 0, yc.shape[2]
        if i > 0:
            t_start += ol
            c0 += ol
        if i < n - 1:
            t_end -= ol
            c1 -= ol
        y[:, :, t_start:t_end] = yc[:, :, c0:c1]
    return y


def encode_audio_chunked(encode_fn, audio: Tensor, downsampling_ratio: int, latent_dim: int,
                         overlap: int = 32, chunk_size: int = 128) -> Tensor:
    """autoencoders.py:446-497 (chunked=True branch)."""
    spl = downsampling_ratio
    total, bsz = audio.shape[2], audio.shape[0]
    cs, ov = chunk_size * spl, overlap * spl
    hop = cs - ov
    if total < cs:
        raise UnboundLocalError("audio shorter than chunk_size (reference raises here too)")
    starts = list(range(0, total - cs + 1, hop))
    chunks = [audio[:, :, i:i + cs] for i in starts]
    if starts[-1] + cs != total:
        chunks.append(audio[:, :, -cs:])
    n = len(chunks)
    y_size = total // spl
    y = torch.zeros((bsz, latent_dim, y_size), device=audio.device)
    for i, ch in enumerate(chunks):
        yc = encode_fn(ch)
        if i == n - 1:
            t_end = y_size
            t_start = t_end - yc.shape[2]
        else:
            t_start = i * hop // spl
            t_end = t_start + cs // spl
        ol = ov // spl // 2
        c0, c1 = 0, yc.shape[2]
        if i > 0:
            t_start += ol
            c0 += ol
        if i < n - 1:
            t_end -= ol
            c1 -= ol
        y[:, :, t_start:t_end] = yc[:, :, c0:c1]
    return y


# ----------------------------------------------------------------------------- latent sampling
def vae_sample(mean: Tensor, scale: Tensor, noise: Tensor):
    """bottleneck.py:51-62 as edited in the reference: latents = noise*scale + mean (raw scale,
    two roundings: mul then add); kl uses stdev = softplus(scale) + 1e-4."""
    stdev = F.softplus(scale) + 1e-4
    var = stdev * stdev
    logvar = torch.log(var)
    latents = noise * scale + mean
    kl = (mean * mean + var - logvar - 1).sum(1).mean()
    return latents, kl


def sigma_sample(mean: Tensor, noise: Tensor, dist_type: str = "fix", std_noise: Optional[Tensor] = None) -> Tensor:
    """model_sigmaVAE.py:153-178 / 187-213.  ``noise`` replaces randn_like(mean); for 'gaussian'
    ``std_noise`` replaces randn(batch)."""
    std = torch.tensor(0.5)
    if dist_type == "fix":
        return mean + std.to(mean.device) * noise
    if dist_type == "gaussian":
        value = std / 0.8
        s = std_noise.to(mean.dtype) * value.to(mean.device)
        while s.dim() < mean.dim():
            s = s.unsqueeze(-1)
        return mean + s * noise
    return mean


# ----------------------------------------------------------------------------- training step
def gaussian_nll(x: Tensor, xhat: Tensor, log_sigma: float = 0.0) -> Tensor:
    """sum_{c,t} [0.5 ((x - xhat)/sigma)^2 + log sigma + 0.5 log 2 pi], mean over the batch."""
    sigma = math.exp(log_sigma)
    nll = 0.5 * ((x - xhat) / sigma) ** 2 + log_sigma + 0.5 * math.log(2.0 * math.pi)
    return nll.flatten(1).sum(1).mean()


def training_loss(sd: Dict[str, Tensor], x: Tensor, noise: Tensor, strides: Sequence[int], kl_weight: float = 1e-6,
                  log_sigma: float = 0.0, enc_prefix: str = "encoder.", dec_prefix: str = "decoder."):
    """encode -> (mean, scale) -> vae_sample -> decode -> nll + kl_weight * kl.  Returns (loss, nll, kl, decoded)."""
    e = oobleck_encoder(sd, x, strides, prefix=enc_prefix)
    mean, scale = e.chunk(2, dim=1)
    z, kl = vae_sample(mean, scale, noise)
    y = oobleck_decoder(sd, z, strides, prefix=dec_prefix)
    nll = gaussian_nll(x, y, log_sigma)
    return nll + kl_weight * kl, nll, kl, y


# ----------------------------------------------------------------------------- bookkeeping
def conv_flops_decoder(latent_dim: int, channels: int, c_mults: Sequence[int], strides: Sequence[int],
                       out_channels: int, B: int, T: int) -> float:
    """Algorithmic FLOPs of one decoder pass: 2*B*T_out*Cout*Cin*K per conv,
    2*B*T_in*Cin*Cout*K per transposed conv, all taps counted."""
    cm = [1] + list(c_mults)
    f = 2.0 * B * T * latent_dim * cm[-1] * channels * 7
    t = T
    for i in range(len(cm) - 1, 0, -1):
        cin, cout, s = cm[i] * channels, cm[i - 1] * channels, strides[i - 1]
        k = 2 * s + s % 2
        f += 2.0 * B * t * cin * cout * k
        t *= s
        f += 3 * (2.0 * B * t * cout * cout * 7 + 2.0 * B * t * cout * cout)
    f += 2.0 * B * t * channels * out_channels * 7
    return f


def conv_flops_encoder(latent_dim: int, channels: int, c_mults: Sequence[int], strides: Sequence[int],
                       in_channels: int, B: int, L: int) -> float:
    cm = [1] + list(c_mults)
    f = 2.0 * B * L * in_channels * channels * 7
    t = L
    for i in range(len(cm) - 1):
        cin, cout, s = cm[i] * channels, cm[i + 1] * channels, strides[i]
        f += 3 * (2.0 * B * t * cin * cin * 7 + 2.0 * B * t * cin * cin)
        t //= s
        f += 2.0 * B * t * cin * cout * 2 * s
    f += 2.0 * B * t * cm[-1] * channels * latent_dim * 3
    return f


# ------------------------------------------------------------------------------------------------ LM <-> VAE glue
def lm_glue_step(sd: Dict[str, Tensor], last_hidden: Tensor, noise: Tensor, std: float = 0.5):
    """One generated frame of ``Llasa.infer`` (model_sigmaVAE.py:123-145): distribution_linear (Linear, GELU, Linear;
    :42-50) -> sample 'fix' (:153-157) -> KL against N(1, e), the stop criterion (:134-139) -> audio_linear (:143).
    ``sd`` holds the reference's keys ``distribution_linear.{0,2}.{weight,bias}`` / ``audio_linear.{weight,bias}``."""
    h = F.linear(last_hidden, sd["distribution_linear.0.weight"], sd["distribution_linear.0.bias"])
    mean = F.linear(F.gelu(h), sd["distribution_linear.2.weight"], sd["distribution_linear.2.bias"])
    latent = mean + torch.tensor(std) * noise
    end = torch.distributions.Normal(torch.ones_like(mean), torch.exp(torch.ones_like(mean)))
    cur = torch.distributions.Normal(mean, torch.tensor(std))
    kl = torch.distributions.kl_divergence(cur, end).sum(2) / mean.shape[2]
    embed = F.linear(latent, sd["audio_linear.weight"], sd["audio_linear.bias"])
    return mean, latent, embed, kl


def llasa_infer(sd: Dict[str, Tensor], backbone, text_embed: Tensor, audio_latents: Optional[Tensor], noises,
                end_disp_kl_thres: float = 0.5, max_length: int = 200) -> Tensor:
    """``Llasa.infer`` (model_sigmaVAE.py:105-148) around a caller-supplied backbone ``inputs_embeds -> hidden``;
    ``noises[i]`` replaces the i-th ``torch.randn_like`` draw.  Returns [1, D, n] like the reference."""
    audio_embed = None if audio_latents is None else F.linear(audio_latents, sd["audio_linear.weight"], sd["audio_linear.bias"])
    input_embed = text_embed if audio_embed is None else torch.cat((text_embed, audio_embed), dim=1)
    outs = []
    for i in range(max_length):
        last_hidden = backbone(input_embed)[:, -1:, :]
        _, latent, embed, kl = lm_glue_step(sd, last_hidden, noises[i])
        outs.append(latent)
        if kl < end_disp_kl_thres and i > 3:
            break
        input_embed = torch.cat((input_embed, embed), dim=1)
    return torch.stack(outs[:-1], dim=1).squeeze(1).squeeze(2).transpose(1, 2)


# ------------------------------------------------------------------------------------------------ dataset side
def peak_normalize(wav: Tensor, gain: float = 0.95) -> Tensor:
    """``librosa.util.normalize(wav) * 0.95`` (twj_dataset.py:232; librosa is an un-vendored dependency without a pin
    in the reference -- its default is the infinity norm, ``x / max|x|``, inputs whose norm is below float tiny are
    left as they are)."""
    peak = wav.abs().max()
    return (wav / peak if float(peak) > 1.1754943508222875e-38 else wav) * gain


def dataset_latents(enc_sd: Dict[str, Tensor], strides: Sequence[int], wav: Tensor, noise: Tensor, gain: float = 0.95,
                    scale: float = 1.0):
    """twj_dataset.py:231-256 for one mono clip: peak-normalise * 0.95 -> stereo duplicate -> pretransform.encode
    (= the Oobleck encoder, / scale) -> chunk into mean | scale -> vae_sample -> [T, D].  Returns (latents, mean_scale)."""
    dual = peak_normalize(wav.float(), gain).reshape(1, -1).repeat(2, 1).unsqueeze(0)
    mean_scale = oobleck_encoder(enc_sd, dual, strides) / scale
    mean, sc = mean_scale.chunk(2, dim=1)
    latents, _ = vae_sample(mean, sc, noise)
    return latents.squeeze(0).transpose(0, 1), mean_scale
