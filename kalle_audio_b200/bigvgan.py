"""Drop-in for the inference path of the reference's 12.5 Hz VAE, ``BigVGANFlowVAE`` (/root/reference/backup/flows.py).

Module tree, constructor argument (the hyper-parameter object ``h``) and ``state_dict`` keys mirror the reference:
``audio_encoder`` (Encoder :191-241: Conv1d_S / LeakyReLU / ResStack), ``flow`` (ResidualCouplingBlock -- training
only; kept as a parameter container so checkpoints load), ``conv_pre``, ``ups``, ``resblocks`` (AMPBlock1 / AMPBlock2
:243-330 with their anti-aliased ``Activation1d``), ``activation_post``, ``conv_post``.  Implemented:

    extract_latents(x)                      :494-496   waveform [B, 1, L] -> [B, 2 D, L / prod(downsample_rates)]
    inference_from_latents(x, do_sample)    :498-529   latents -> waveform [B, 1, T * prod(upsample_rates)]

Every step runs through libkvae: convolutions and transposed convolutions (causal or not) through the generic conv
kernels (``kvae_conv1d_fwd``; weight norm folded on the device), the anti-aliased Snake / SnakeBeta activation as ONE
fused kernel (``kvae_aa_act_fwd``: x2 Kaiser-sinc upsampling, activation, low-pass + x2 decimation -- six eager
kernels and four double-rate intermediates in the reference), LeakyReLU / tanh / residual adds / the AMP average / the
Gaussian sample as small elementwise kernels.  ``forward`` (the training pass through the flow) is not built.
This is the parity-first form of SURVEY section 8(f) item 2: CUDA end to end, not yet on the tensor cores.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
from torch import nn

from . import _lib
from .layers import WNConv1d, WNConvTranspose1d


def _kaiser_sinc_filter1d(cutoff: float, half_width: float, kernel_size: int) -> torch.Tensor:
    """Filter design of alias_free_torch (filter.py): Kaiser window x sinc, normalised to unit sum.  Load-time only."""
    half_size = kernel_size // 2
    delta_f = 4 * half_width
    A = 2.285 * (half_size - 1) * math.pi * delta_f + 7.95
    beta = 0.1102 * (A - 8.7) if A > 50.0 else (0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21.0) if A >= 21.0 else 0.0)
    window = torch.kaiser_window(kernel_size, beta=beta, periodic=False)
    time = (torch.arange(-half_size, half_size) + 0.5) if kernel_size % 2 == 0 else torch.arange(kernel_size) - half_size
    arg = 2 * cutoff * time
    sinc = torch.where(arg == 0, torch.ones_like(arg), torch.sin(math.pi * arg) / math.pi / arg)
    f = 2 * cutoff * window * sinc
    return (f / f.sum()).view(1, 1, kernel_size)


# ------------------------------------------------------------------------------------------------- kernel wrappers
def _k(t: torch.Tensor) -> torch.Tensor:
    t = t if t.dtype in (torch.float32, torch.bfloat16) else t.float()
    return t.contiguous()


def _unary(x: torch.Tensor, op: int, param: float = 0.0) -> torch.Tensor:
    _lib.require_cuda(x, "bigvgan")
    xin = _k(x)
    y = torch.empty_like(xin)
    _lib.check(_lib.lib().kvae_unary_fwd(xin.data_ptr(), y.data_ptr(), xin.numel(), op, float(param),
                                         _lib.dtype_code(xin.dtype), _lib.stream_ptr(x.device)))
    return y if y.dtype == x.dtype else y.to(x.dtype)


def _axpby(a: torch.Tensor, b: torch.Tensor, alpha: float, beta: float) -> torch.Tensor:
    ain, bin_ = _k(a), _k(b).to(_k(a).dtype)
    out = torch.empty_like(ain)
    _lib.check(_lib.lib().kvae_axpby(ain.data_ptr(), bin_.data_ptr(), out.data_ptr(), ain.numel(), alpha, beta,
                                     _lib.dtype_code(ain.dtype), _lib.stream_ptr(a.device)))
    return out if out.dtype == a.dtype else out.to(a.dtype)


class LeakyReLU(nn.Module):
    def __init__(self, negative_slope: float = 0.01, inplace: bool = False):
        super().__init__()
        self.negative_slope = negative_slope

    def forward(self, x):
        return _unary(x, 0, self.negative_slope)


class Snake(nn.Module):
    """Parameter holder of flows.py:9-61 (alpha); applied inside Activation1d's fused kernel."""

    def __init__(self, in_features, alpha=1.0, alpha_trainable=True, alpha_logscale=False):
        super().__init__()
        self.in_features, self.alpha_logscale = in_features, alpha_logscale
        init = torch.zeros if alpha_logscale else torch.ones
        self.alpha = nn.Parameter(init(in_features) * alpha, requires_grad=alpha_trainable)


class SnakeBeta(nn.Module):
    """Parameter holder of flows.py:64-125 (alpha, beta); applied inside Activation1d's fused kernel."""

    def __init__(self, in_features, alpha=1.0, alpha_trainable=True, alpha_logscale=False):
        super().__init__()
        self.in_features, self.alpha_logscale = in_features, alpha_logscale
        init = torch.zeros if alpha_logscale else torch.ones
        self.alpha = nn.Parameter(init(in_features) * alpha, requires_grad=alpha_trainable)
        self.beta = nn.Parameter(init(in_features) * alpha, requires_grad=alpha_trainable)


class _Filter(nn.Module):
    def __init__(self, ratio: int, kernel_size: int):
        super().__init__()
        self.register_buffer("filter", _kaiser_sinc_filter1d(0.5 / ratio, 0.6 / ratio, kernel_size))


class _Down(nn.Module):
    def __init__(self, ratio: int, kernel_size: int):
        super().__init__()
        self.lowpass = _Filter(ratio, kernel_size)


class Activation1d(nn.Module):
    """alias_free_torch.Activation1d (used at flows.py:266, 312, 443) with the same buffers (``upsample.filter``,
    ``downsample.lowpass.filter``) and ``act`` sub-module; forward is ONE kernel."""

    def __init__(self, activation, up_ratio: int = 2, down_ratio: int = 2, up_kernel_size: int = 12,
                 down_kernel_size: int = 12):
        super().__init__()
        if (up_ratio, down_ratio, up_kernel_size, down_kernel_size) != (2, 2, 12, 12):
            raise NotImplementedError("Activation1d: ratio 2 / 12 taps only (the values flows.py uses)")
        self.act = activation
        self.upsample = _Filter(up_ratio, up_kernel_size)
        self.downsample = _Down(down_ratio, down_kernel_size)

    def forward(self, x):
        _lib.require_cuda(x, "Activation1d")
        xin = _k(x)
        B, C, T = xin.shape
        y = torch.empty_like(xin)
        beta = getattr(self.act, "beta", None)
        fu = self.upsample.filter.detach().float().reshape(-1).contiguous()
        fd = self.downsample.lowpass.filter.detach().float().reshape(-1).contiguous()
        _lib.check(_lib.lib().kvae_aa_act_fwd(xin.data_ptr(), y.data_ptr(), self.act.alpha.detach().float().contiguous().data_ptr(),
                                              None if beta is None else beta.detach().float().contiguous().data_ptr(),
                                              int(self.act.alpha_logscale), fu.data_ptr(), fd.data_ptr(), B, C, T,
                                              _lib.dtype_code(xin.dtype), _lib.stream_ptr(x.device)))
        return y if y.dtype == x.dtype else y.to(x.dtype)


class Conv1d(WNConv1d):
    """flows.py:566-620: weight-normed Conv1d, 'same' padding or causal (left padding d (K - 1), :607-608).  A causal conv
    is the symmetric-padded conv's first T outputs, which is how it runs here."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, dilation=1, causal=False, bias=True):
        self.causal = bool(causal)
        pad = dilation * (kernel_size - 1) if causal else int((kernel_size * dilation - dilation) / 2)
        super().__init__(in_channels, out_channels, kernel_size, stride=stride, padding=pad, dilation=dilation, bias=bias)

    def forward(self, x):
        if self.causal and self.kernel_size[0] > 1:
            y = self._forward_tc(x, keep=x.shape[2])         # tensor-core layers truncate inside the kernel
            if y is not None:
                return y
            return super().forward(x)[:, :, :x.shape[2]].contiguous()
        return super().forward(x)


class ConvTranspose1d(WNConvTranspose1d):
    """flows.py:336-391: padding (k - stride) // 2, or causal: k == 2 stride, no padding, last `stride` samples dropped."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, causal=False):
        self.causal = bool(causal)
        if causal and kernel_size != 2 * stride:
            raise AssertionError("kernel_size must be equal to 2*stride in Causal ConvTranspose1d.")
        super().__init__(in_channels, out_channels, kernel_size, stride=stride,
                         padding=0 if causal else (kernel_size - stride) // 2)
        self._drop = stride if causal else 0

    def forward(self, x):
        if self._drop:
            y = self._forward_tc(x, keep=x.shape[2] * self.stride[0])
            if y is not None:
                return y
            return super().forward(x)[:, :, :-self._drop].contiguous()
        return super().forward(x)


class Conv1d_S(nn.Module):
    """flows.py:139-172 (weight_norm form): ``self.layer`` = weight-normed Conv1d with padding d (k - 1) // 2."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, dilation=1):
        super().__init__()
        self.layer = WNConv1d(in_channels, out_channels, kernel_size, stride=stride, padding=dilation * (kernel_size - 1) // 2,
                              dilation=dilation)

    def forward(self, x):
        return self.layer(x)


class ResStack(nn.Module):
    """flows.py:174-189."""

    def __init__(self, channel, kernel_size=3, base=3, nums=4):
        super().__init__()
        self.layers = nn.ModuleList([
            nn.Sequential(LeakyReLU(), WNConv1d(channel, channel, kernel_size, dilation=base ** i, padding=base ** i),
                          LeakyReLU(), WNConv1d(channel, channel, kernel_size, dilation=1, padding=1))
            for i in range(nums)])

    def forward(self, x):
        for layer in self.layers:
            x = _axpby(x, layer(x), 1.0, 1.0)
        return x


class Encoder(nn.Module):
    """flows.py:191-241."""

    def __init__(self, in_channels=1, out_channels=100, base_channels=12, proj_kernel_size=3, stack_kernel_size=3,
                 stack_dilation_base=2, stacks=6, channels=(12, 24, 48, 96, 192, 384, 768),
                 down_sample_factors=(2, 2, 2, 2, 4, 4), use_vae=False):
        super().__init__()
        if use_vae:
            out_channels = out_channels * 2
        layers: List[nn.Module] = [Conv1d_S(in_channels, base_channels, kernel_size=proj_kernel_size), LeakyReLU(0.2)]
        for (in_c, out_c), f in zip(zip(channels[:-1], channels[1:]), down_sample_factors):
            layers += [Conv1d_S(in_c, out_c, kernel_size=f * 2, stride=f),
                       ResStack(out_c, stack_kernel_size, stack_dilation_base, stacks), LeakyReLU(0.2)]
        layers += [Conv1d_S(channels[-1], out_channels, proj_kernel_size)]
        self.generator = nn.Sequential(*layers)

    def forward(self, conditions, z_inputs=None):
        return self.generator(conditions)


def _make_act(h, channels):
    if h.activation == "snake":
        return Activation1d(Snake(channels, alpha_logscale=h.snake_logscale))
    if h.activation == "snakebeta":
        return Activation1d(SnakeBeta(channels, alpha_logscale=h.snake_logscale))
    raise NotImplementedError("activation incorrectly specified. check the config file and look for 'activation'.")


class AMPBlock1(nn.Module):
    """flows.py:243-291."""

    def __init__(self, h, channels, kernel_size=3, dilation=(1, 3, 5), activation=None, causal=True):
        super().__init__()
        self.convs1 = nn.ModuleList([Conv1d(channels, channels, kernel_size, dilation=d, causal=causal) for d in dilation])
        self.convs2 = nn.ModuleList([Conv1d(channels, channels, kernel_size, dilation=1, causal=causal) for _ in dilation])
        self.num_layers = len(self.convs1) + len(self.convs2)
        self.activations = nn.ModuleList([_make_act(h, channels) for _ in range(self.num_layers)])

    def forward(self, x):
        acts1, acts2 = self.activations[::2], self.activations[1::2]
        for c1, c2, a1, a2 in zip(self.convs1, self.convs2, acts1, acts2):
            x = _axpby(c2(a2(c1(a1(x)))), x, 1.0, 1.0)
        return x


class AMPBlock2(nn.Module):
    """flows.py:294-335."""

    def __init__(self, h, channels, kernel_size=3, dilation=(1, 3), activation=None, causal=True):
        super().__init__()
        self.convs = nn.ModuleList([Conv1d(channels, channels, kernel_size, dilation=d, causal=causal) for d in dilation])
        self.num_layers = len(self.convs)
        self.activations = nn.ModuleList([_make_act(h, channels) for _ in range(self.num_layers)])

    def forward(self, x):
        for c, a in zip(self.convs, self.activations):
            x = _axpby(c(a(x)), x, 1.0, 1.0)
        return x


class _FlowParams(nn.Module):
    """Parameter container with the keys of ResidualCouplingBlock (flows.py:622-792).  The flow only runs in the
    reference's training ``forward``; neither inference entry point touches it."""

    def __init__(self, channels, hidden, kernel_size, n_layers, n_flows=4):
        super().__init__()

        class _WN(nn.Module):
            def __init__(self):
                super().__init__()
                self.in_layers = nn.ModuleList([WNConv1d(hidden, 2 * hidden, kernel_size) for _ in range(n_layers)])
                self.res_skip_layers = nn.ModuleList([WNConv1d(hidden, 2 * hidden if i < n_layers - 1 else hidden, 1)
                                                      for i in range(n_layers)])

        class _Layer(nn.Module):
            def __init__(self):
                super().__init__()
                self.pre = nn.Conv1d(channels // 2, hidden, 1)
                self.enc = _WN()
                self.post = nn.Conv1d(hidden, channels // 2, 1)

        self.flows = nn.ModuleList()
        for _ in range(n_flows):
            self.flows.append(_Layer())
            self.flows.append(nn.Module())      # Flip: no parameters


class BigVGANFlowVAE(nn.Module):
    def __init__(self, h):
        super().__init__()
        self.h = h
        causal = h.causal
        self.audio_encoder = Encoder(out_channels=h.latent_dim, use_vae=h.use_vae, channels=list(h.downsample_channels),
                                     down_sample_factors=list(h.downsample_rates))
        self.flow = _FlowParams(h.latent_dim, h.flow_hidden_channels, 5, 4)
        self.num_kernels = len(h.resblock_kernel_sizes)
        self.num_upsamples = len(h.upsample_rates)
        self.conv_pre = Conv1d(h.latent_dim, h.upsample_initial_channel, 7, 1, causal=False)
        resblock = AMPBlock1 if h.resblock == "1" else AMPBlock2
        self.ups = nn.ModuleList()
        for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
            self.ups.append(nn.ModuleList([ConvTranspose1d(h.upsample_initial_channel // (2 ** i),
                                                           h.upsample_initial_channel // (2 ** (i + 1)), k, u, causal=causal)]))
        self.resblocks = nn.ModuleList()
        ch = h.upsample_initial_channel
        for i in range(len(self.ups)):
            ch = h.upsample_initial_channel // (2 ** (i + 1))
            for k, d in zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes):
                self.resblocks.append(resblock(h, ch, k, d, activation=h.activation, causal=causal))
        self.activation_post = _make_act(h, ch)
        self.conv_post = Conv1d(ch, 1, 7, 1, causal=causal)
        self.set_precision("fp32")
        for m in self.modules():           # inference-only model: fold weight norm once per weight version (layers.py)
            if isinstance(m, (WNConv1d, WNConvTranspose1d)):
                m._cache_fold = True

    def set_precision(self, precision: Optional[str]):
        """Arithmetic of the convolutions whose channel counts are multiples of 64 (the wide stages, > 90 % of the
        FLOPs of a production-size model): "fp32" (default) = tensor cores through the bf16x3 operand split (<= 1e-5),
        "bf16" = bf16 tensor-core operands (<= 1e-3 class), None = fp32 CUDA-core kernels everywhere."""
        if precision not in (None, "fp32", "bf16"):
            raise ValueError("precision must be None, 'fp32' or 'bf16'")
        for m in self.modules():
            if isinstance(m, (WNConv1d, WNConvTranspose1d)):
                m.tc_precision = precision
        return self

    def forward(self, x):
        raise NotImplementedError("BigVGANFlowVAE.forward (the training pass through the flow, flows.py:454-492) is "
                                  "outside the hot path; use extract_latents / inference_from_latents")

    @torch.no_grad()
    def extract_latents(self, x):
        return self.audio_encoder(x)

    @torch.no_grad()
    def inference_from_latents(self, x, do_sample=True, noise: Optional[torch.Tensor] = None):
        h = self.h
        if h.use_vae and do_sample:
            assert x.size(1) == h.latent_dim * 2, "Input must be like [B, D, H]"
            m_q, logs_q = torch.split(x, h.latent_dim, dim=1)
            if noise is None:
                noise = torch.randn_like(m_q)
            m, lg, nz = _k(m_q), _k(logs_q), _k(noise)
            z = torch.empty_like(m)
            _lib.check(_lib.lib().kvae_gauss_sample(m.data_ptr(), lg.to(m.dtype).data_ptr(), nz.to(m.dtype).data_ptr(),
                                                    z.data_ptr(), m.numel(), _lib.dtype_code(m.dtype),
                                                    _lib.stream_ptr(x.device)))
            x = z if z.dtype == x.dtype else z.to(x.dtype)
        else:
            assert x.size(1) == h.latent_dim, "Input must be like [B, D, H]"
        x = self.conv_pre(x)
        for i in range(self.num_upsamples):
            for up in self.ups[i]:
                x = up(x)
            xs = None
            for j in range(self.num_kernels):
                y = self.resblocks[i * self.num_kernels + j](x)
                xs = y if xs is None else _axpby(xs, y, 1.0, 1.0)
            x = _axpby(xs, xs, 1.0 / self.num_kernels, 0.0)
        x = self.activation_post(x)
        x = self.conv_post(x)
        return _unary(x, 1)

    def remove_weight_norm(self):
        for m in self.modules():
            if isinstance(m, (WNConv1d, WNConvTranspose1d)) and m.has_weight_norm:
                m.remove_weight_norm()
