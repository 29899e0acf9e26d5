"""CPU oracle for the BigVGANFlowVAE inference path.  TEST INFRASTRUCTURE ONLY (tests/ and make_golden.py use it).

Functional restatement, over a state_dict with the reference's key names, of
/root/reference/backup/flows.py:
  conv (weight-normed, optionally causal)   Conv1d :566-620 (causal: left padding d (K - 1)), Conv1d_S :139-172
  upsampler                                 ConvTranspose1d :336-391 (causal: kernel 2 stride, drop the last `stride` samples)
  ResStack / Encoder (extract_latents)      :174-241, :494-496
  AMPBlock1 / AMPBlock2                     :243-330
  inference_from_latents                    :498-529
  the anti-aliased activation               alias_free_torch.Activation1d, un-vendored: oracle/alias_free_restated.py
Pinned by tests/golden/bigvgan.npz, recorded from the reference's own flows.py (with the restated alias_free_torch
installed under its name, the only way the file imports here)."""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

from . import alias_free_restated as AF

Tensor = torch.Tensor


def _fold(sd, key):
    if key + ".weight_g" in sd:
        v, g = sd[key + ".weight_v"], sd[key + ".weight_g"]
        norm = v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
        return v * (g / norm)
    return sd[key + ".weight"]


def conv(sd, key, x, k, stride=1, dilation=1, causal=False, padding=None):
    w = _fold(sd, key)
    if padding is None:
        padding = 0 if causal else int((k * dilation - dilation) / 2)
    if causal:
        x = F.pad(x, (dilation * (k - 1), 0))
    return F.conv1d(x, w, sd.get(key + ".bias"), stride=stride, padding=padding, dilation=dilation)


def conv_transpose(sd, key, x, k, stride, causal):
    w = _fold(sd, key)
    padding = 0 if causal else (k - stride) // 2
    y = F.conv_transpose1d(x, w, sd.get(key + ".bias"), stride=stride, padding=padding)
    return y[:, :, :-stride] if causal else y


def snake(x, alpha, beta, logscale):
    a = alpha.view(1, -1, 1)
    b = a if beta is None else beta.view(1, -1, 1)
    if logscale:
        a, b = torch.exp(a), torch.exp(b)
    return x + (1.0 / (b + 0.000000001)) * torch.sin(x * a) ** 2


def aa_activation(sd, key, x, logscale):
    """Activation1d(Snake | SnakeBeta): upsample x2 -> activation -> downsample x2."""
    up, down = AF.UpSample1d(2, 12), AF.DownSample1d(2, 12)
    u = up(x)
    v = snake(u, sd[key + ".act.alpha"], sd.get(key + ".act.beta"), logscale)
    return down(v)


def amp_block1(sd, key, x, k, dilations, causal, logscale):
    for i, d in enumerate(dilations):
        xt = aa_activation(sd, f"{key}.activations.{2 * i}", x, logscale)
        xt = conv(sd, f"{key}.convs1.{i}", xt, k, dilation=d, causal=causal)
        xt = aa_activation(sd, f"{key}.activations.{2 * i + 1}", xt, logscale)
        xt = conv(sd, f"{key}.convs2.{i}", xt, k, dilation=1, causal=causal)
        x = xt + x
    return x


def amp_block2(sd, key, x, k, dilations, causal, logscale):
    for i, d in enumerate(dilations):
        xt = aa_activation(sd, f"{key}.activations.{i}", x, logscale)
        xt = conv(sd, f"{key}.convs.{i}", xt, k, dilation=d, causal=causal)
        x = xt + x
    return x


def inference_from_latents(sd: Dict[str, Tensor], h, x: Tensor, noise: Tensor = None) -> Tensor:
    """flows.py:498-529.  ``noise`` given: the do_sample branch on [B, 2 D, T] (mean | log-scale)."""
    if noise is not None:
        m_q, logs_q = torch.split(x, h["latent_dim"], dim=1)
        x = m_q + noise * torch.exp(logs_q)
    causal, logscale = h["causal"], h["snake_logscale"]
    x = conv(sd, "conv_pre", x, 7, causal=False)
    nk = len(h["resblock_kernel_sizes"])
    block = amp_block1 if h["resblock"] == "1" else amp_block2
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        x = conv_transpose(sd, f"ups.{i}.0", x, k, u, causal)
        xs = None
        for j, (rk, rd) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            y = block(sd, f"resblocks.{i * nk + j}", x, rk, rd, causal, logscale)
            xs = y if xs is None else xs + y
        x = xs / nk
    x = aa_activation(sd, "activation_post", x, logscale)
    x = conv(sd, "conv_post", x, 7, causal=causal)
    return torch.tanh(x)


def extract_latents(sd: Dict[str, Tensor], h, x: Tensor) -> Tensor:
    """flows.py:494-496 -> Encoder :191-241: Conv1d_S k3 -> LeakyReLU(0.2) -> per stage [Conv1d_S(k = 2 f, stride f) ->
    ResStack(k3, dilation base 2, 6 layers) -> LeakyReLU(0.2)] -> Conv1d_S k3."""
    p = "audio_encoder.generator"
    x = F.leaky_relu(conv(sd, f"{p}.0.layer", x, 3, padding=1), 0.2)
    idx = 2
    for f in h["downsample_rates"]:
        x = conv(sd, f"{p}.{idx}.layer", x, 2 * f, stride=f, padding=(2 * f - 1) // 2)
        for i in range(6):                                   # ResStack :174-189 (LeakyReLU default slope 0.01)
            y = F.leaky_relu(x, 0.01)
            y = conv(sd, f"{p}.{idx + 1}.layers.{i}.1", y, 3, dilation=2 ** i, padding=2 ** i)
            y = F.leaky_relu(y, 0.01)
            y = conv(sd, f"{p}.{idx + 1}.layers.{i}.3", y, 3, padding=1)
            x = x + y
        x = F.leaky_relu(x, 0.2)
        idx += 3
    return conv(sd, f"{p}.{idx}.layer", x, 3, padding=1)
