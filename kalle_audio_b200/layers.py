"""Drop-in modules for the leaves of the Oobleck stack: SnakeBeta, WNConv1d, WNConvTranspose1d.

Same constructor arguments, attribute names, parameter names/shapes and ``state_dict`` keys as the
reference:
  * SnakeBeta                      stable_audio_tools/models/blocks.py:309-339
  * WNConv1d / WNConvTranspose1d   dac.nn.layers (un-vendored): ``weight_norm(nn.Conv1d(...))`` /
                                   ``weight_norm(nn.ConvTranspose1d(...))`` with old-style parameters
                                   ``weight_g`` / ``weight_v`` (call sites autoencoders.py:49-53,76,98)
Parameters are initialised through torch's own ``nn.Conv1d`` / ``nn.ConvTranspose1d`` constructors, so
the same ``torch.manual_seed`` yields the same random-init ``state_dict`` as the reference.

``forward`` / ``backward`` of a leaf run the layer-level C-ABI kernels (fp32 arithmetic, autograd nodes
``_SnakeFn`` / ``_ConvFn``).  The fast path is not here: ``OobleckEncoder`` / ``OobleckDecoder`` hand the whole
stack to a fused plan (see ``_plan.py``).
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib


class _SnakeFn(torch.autograd.Function):
    """SnakeBeta.forward / backward on [B, C, T] (blocks.py:331-339 and its autograd): kvae_snake_fwd / kvae_snake_bwd."""

    @staticmethod
    def forward(ctx, x, alpha, beta, logscale):
        a = alpha.detach().float().contiguous()
        b = beta.detach().float().contiguous()
        y = _snake_call(x.detach(), a, b, logscale)
        ctx.save_for_backward(x.detach(), a, b)
        ctx.logscale = bool(logscale)
        ctx.param_dtypes = (alpha.dtype, beta.dtype)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, a, b = ctx.saved_tensors
        B, Cc, T = x.shape
        # the kernel works on channels-last rows [B*T, C] (the layout of the fused plans)
        xr = x.float().transpose(1, 2).contiguous()
        gr = gy.detach().float().transpose(1, 2).contiguous()
        gx = torch.empty_like(xr)
        da = torch.empty(Cc, dtype=torch.float32, device=x.device)
        db = torch.empty(Cc, dtype=torch.float32, device=x.device)
        scratch = torch.empty(2 * Cc, dtype=torch.float32, device=x.device)
        if xr.numel():
            _lib.check(_lib.lib().kvae_snake_bwd(xr.data_ptr(), gr.data_ptr(), gx.data_ptr(), a.data_ptr(), b.data_ptr(),
                                                 int(ctx.logscale), da.data_ptr(), db.data_ptr(), B * T, Cc,
                                                 scratch.data_ptr(), _lib.stream_ptr(x.device)))
        else:
            da.zero_(); db.zero_()
        return (gx.transpose(1, 2).contiguous().to(x.dtype), da.to(ctx.param_dtypes[0]), db.to(ctx.param_dtypes[1]),
                None)


class _ConvFn(torch.autograd.Function):
    """WNConv1d / WNConvTranspose1d forward / backward on [B, C, T]: kvae_conv1d_fwd, kvae_conv1d_bwd and the
    weight-norm backward (autograd of torch.nn.utils.weight_norm + F.conv1d / F.conv_transpose1d)."""

    @staticmethod
    def forward(ctx, mod, x, bias, *wparams):
        y, xin, w = mod._forward_impl(x.detach())
        ctx.mod = mod
        ctx.save_for_backward(xin, w, *[p.detach() for p in wparams])
        ctx.has_bias = bias is not None
        ctx.x_dtype = x.dtype
        return y

    @staticmethod
    def backward(ctx, gy):
        mod = ctx.mod
        xin, w, *wparams = ctx.saved_tensors
        L = _lib.lib()
        g = gy.detach().to(xin.dtype).contiguous()
        B, _, T = xin.shape
        K, s, d, p = mod.kernel_size[0], mod.stride[0], mod.dilation[0], mod.padding[0]
        gx = torch.empty_like(xin) if ctx.needs_input_grad[1] else None
        dw = torch.empty_like(w)
        db = torch.empty(mod.out_channels, dtype=torch.float32, device=xin.device) if ctx.has_bias else None
        nscratch = L.kvae_conv1d_scratch_bytes(mod.in_channels, mod.out_channels, K)
        scratch = torch.empty(nscratch, dtype=torch.uint8, device=xin.device)
        _lib.check(L.kvae_conv1d_bwd(xin.data_ptr(), g.data_ptr(), w.data_ptr(), _lib.ptr(gx), dw.data_ptr(), _lib.ptr(db),
                                     int(mod.transposed), B, mod.in_channels, mod.out_channels, T, K, s, d, p,
                                     _lib.dtype_code(xin.dtype), scratch.data_ptr(), nscratch, _lib.stream_ptr(xin.device)))
        if len(wparams) == 2:     # (weight_g, weight_v): weight-norm backward
            gparam, v = wparams
            vf, gf = v.float().contiguous(), gparam.float().contiguous()
            dv, dg = torch.empty_like(vf), torch.empty(vf.shape[0], dtype=torch.float32, device=vf.device)
            _lib.check(L.kvae_weight_norm_bwd(vf.data_ptr(), gf.data_ptr(), dw.data_ptr(), dv.data_ptr(), dg.data_ptr(),
                                              vf.shape[0], vf[0].numel(), _lib.stream_ptr(vf.device)))
            wgrads = (dg.view_as(gparam).to(gparam.dtype), dv.to(v.dtype))
        else:                     # plain weight (weight norm removed)
            wgrads = (dw.to(wparams[0].dtype),)
        gxo = None if gx is None else gx.to(ctx.x_dtype)
        return (None, gxo, None if db is None else db, *wgrads)


def snake_beta(x, alpha, beta):
    """Functional form of blocks.py:301-302 (alpha, beta already exponentiated, broadcast [1,C,1])."""
    _lib.require_cuda(x, "snake_beta")
    a = alpha.reshape(-1).float().contiguous()
    b = beta.reshape(-1).float().contiguous()
    return _snake_call(x, a, b, logscale=False)


def _snake_call(x, alpha, beta, logscale):
    if x.dim() != 3:
        raise ValueError("SnakeBeta expects [B, C, T]")
    B, Cc, T = x.shape
    if alpha.numel() != Cc:
        raise ValueError(f"SnakeBeta has {alpha.numel()} channels, input has {Cc}")
    xin = x if x.dtype in (torch.float32, torch.bfloat16) else x.float()
    xin = xin.contiguous()
    y = torch.empty_like(xin)
    if xin.numel():
        _lib.check(_lib.lib().kvae_snake_fwd(xin.data_ptr(), y.data_ptr(), alpha.data_ptr(), beta.data_ptr(),
                                             int(logscale), B, Cc, T, _lib.dtype_code(xin.dtype),
                                             _lib.stream_ptr(x.device)))
    return y if y.dtype == x.dtype else y.to(x.dtype)


class SnakeBeta(nn.Module):
    def __init__(self, in_features, alpha=1.0, alpha_trainable=True, alpha_logscale=True):
        super().__init__()
        self.in_features = in_features
        self.alpha_logscale = alpha_logscale
        if self.alpha_logscale:
            self.alpha = nn.Parameter(torch.zeros(in_features) * alpha)
            self.beta = nn.Parameter(torch.zeros(in_features) * alpha)
        else:
            self.alpha = nn.Parameter(torch.ones(in_features) * alpha)
            self.beta = nn.Parameter(torch.ones(in_features) * alpha)
        self.alpha.requires_grad = alpha_trainable
        self.beta.requires_grad = alpha_trainable
        self.no_div_by_zero = 0.000000001

    def forward(self, x):
        _lib.require_cuda(x, "SnakeBeta.forward")
        if torch.is_grad_enabled() and (x.requires_grad or self.alpha.requires_grad or self.beta.requires_grad):
            return _SnakeFn.apply(x, self.alpha, self.beta, self.alpha_logscale)
        return _snake_call(x, self.alpha.detach().float().contiguous(), self.beta.detach().float().contiguous(),
                           self.alpha_logscale)


def _single(v):
    if isinstance(v, (tuple, list)):
        if len(v) != 1:
            raise ValueError("1-D convolution arguments must be scalars or 1-tuples")
        return int(v[0])
    return int(v)


class _WNConvBase(nn.Module):
    transposed = False

    def _init_from(self, conv: nn.Module):
        # old-style weight_norm: bias stays first, then weight_g = ||v|| over all dims but 0, weight_v = v
        w = conv.weight.detach()
        if conv.bias is not None:
            self.bias = nn.Parameter(conv.bias.detach().clone())
        else:
            self.register_parameter("bias", None)
        self.weight_g = nn.Parameter(torch.norm_except_dim(w, 2, 0).clone())
        self.weight_v = nn.Parameter(w.clone())

    # --- weight-norm handling -------------------------------------------------------------------
    @property
    def has_weight_norm(self) -> bool:
        return "weight_g" in self._parameters

    def folded_weight(self) -> torch.Tensor:
        """fp32 folded weight w = v * g/||v|| computed on the device by libkvae (fold in fp32, round once)."""
        if not self.has_weight_norm:
            return self._parameters["weight"].detach().float().contiguous()
        # Inference-only models built from stand-alone layers (BigVGANFlowVAE) set ``_cache_fold``: the folded weight is
        # reused until weight_v / weight_g change (storage or torch version counter, i.e. .to(), load_state_dict and
        # in-place optimizer steps all invalidate it).  Read-only for the callers.
        key = None
        if getattr(self, "_cache_fold", False) and not torch.is_grad_enabled():
            key = (self.weight_v.data_ptr(), self.weight_v._version, self.weight_g.data_ptr(), self.weight_g._version)
            hit = self.__dict__.get("_fold_cache")
            if hit is not None and hit[0] == key:
                return hit[1]
        v = self.weight_v.detach().float().contiguous()
        g = self.weight_g.detach().float().contiguous()
        _lib.require_cuda(v, "weight-norm fold")
        w = torch.empty_like(v)
        _lib.check(_lib.lib().kvae_weight_norm_fold(v.data_ptr(), g.data_ptr(), w.data_ptr(), v.shape[0],
                                                    v[0].numel(), _lib.stream_ptr(v.device)))
        if key is not None:
            self.__dict__["_fold_cache"] = (key, w)
        return w

    @property
    def weight(self):
        if not self.has_weight_norm:
            return self._parameters["weight"]
        return self.folded_weight().to(self.weight_v.dtype)

    def remove_weight_norm(self):
        """torch.nn.utils.remove_weight_norm equivalent (reference models/utils.py:14-20): replaces
        weight_g / weight_v by a plain ``weight`` parameter."""
        if not self.has_weight_norm:
            raise ValueError("weight_norm already removed")
        w = self.folded_weight().to(self.weight_v.dtype)
        del self._parameters["weight_g"]
        del self._parameters["weight_v"]
        self.register_parameter("weight", nn.Parameter(w))
        return self

    def conv_params(self):
        return [p for p in self.parameters(recurse=False)]

    # --- layer-level forward -------------------------------------------------------------------
    def forward(self, x):
        _lib.require_cuda(x, type(self).__name__ + ".forward")
        if x.dim() != 3 or x.shape[1] != self.in_channels:
            raise ValueError(f"expected [B, {self.in_channels}, T], got {tuple(x.shape)}")
        if x.shape[0] == 0 or x.shape[2] == 0:
            raise ValueError("empty input")
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.conv_params())):
            wparams = (self.weight_g, self.weight_v) if self.has_weight_norm else (self._parameters["weight"],)
            return _ConvFn.apply(self, x, self.bias, *wparams)
        y = self._forward_tc(x)
        return y if y is not None else self._forward_impl(x)[0]

    # Stand-alone layers run fp32 CUDA-core arithmetic by default.  ``tc_precision`` ("bf16" | "fp32") sends layers
    # whose channel counts are multiples of 64 through the tcgen05 conv instead (kvae_conv1d_tc_fwd; "fp32" = bf16x3
    # operand split, <= 1e-5); ``keep`` truncates the output to its first ``keep`` samples inside the kernel.
    tc_precision = None

    def _forward_tc(self, x, keep=0):
        L = _lib.lib()
        K, s, d, p = self.kernel_size[0], self.stride[0], self.dilation[0], self.padding[0]
        if self.tc_precision is None or getattr(self, "_same_extra_right", 0) or x.shape[2] < 64 or \
                not L.kvae_conv1d_tc_supported(self.in_channels, self.out_channels, K, s, d, int(self.transposed)):
            return None
        xin = x if x.dtype in (torch.float32, torch.bfloat16) else x.float()
        xin = xin.contiguous()
        B, _, T = xin.shape
        T_nat = (T - 1) * s - 2 * p + d * (K - 1) + 1 if self.transposed else (T + 2 * p - d * (K - 1) - 1) // s + 1
        keep = int(keep) if keep else T_nat
        if keep != T_nat and not self.transposed and s != 1:
            return None
        prec = _lib.KVAE_PREC_BF16 if self.tc_precision == "bf16" else _lib.KVAE_PREC_F32
        w = self.folded_weight()
        bias = None if self.bias is None else self.bias.detach().float().contiguous()
        y = torch.empty((B, self.out_channels, keep), dtype=xin.dtype, device=x.device)
        nscratch = L.kvae_conv1d_tc_scratch_bytes(B, self.in_channels, self.out_channels, T, K, prec)
        scratch = torch.empty(nscratch, dtype=torch.uint8, device=x.device)
        _lib.check(L.kvae_conv1d_tc_fwd(xin.data_ptr(), y.data_ptr(), w.data_ptr(), _lib.ptr(bias), int(self.transposed), B,
                                        self.in_channels, self.out_channels, T, keep, K, s, d, p,
                                        _lib.dtype_code(xin.dtype), prec, scratch.data_ptr(), nscratch,
                                        _lib.stream_ptr(x.device)))
        return y if y.dtype == x.dtype else y.to(x.dtype)

    def _forward_impl(self, x):
        """(y, the contiguous input actually used, the folded fp32 weight)"""
        xin = x if x.dtype in (torch.float32, torch.bfloat16) else x.float()
        if getattr(self, "_same_extra_right", 0):     # 'same' with an even kernel: one more zero on the right
            xin = torch.nn.functional.pad(xin, (0, self._same_extra_right))
        xin = xin.contiguous()
        w = self.folded_weight()
        bias = None if self.bias is None else self.bias.detach().float().contiguous()
        B, _, T = xin.shape
        K, s, d, p = self.kernel_size[0], self.stride[0], self.dilation[0], self.padding[0]
        if self.transposed:
            T_out = (T - 1) * s - 2 * p + d * (K - 1) + 1
        else:
            T_out = (T + 2 * p - d * (K - 1) - 1) // s + 1
        if T_out <= 0:
            raise RuntimeError("Kernel size can't be greater than actual input size")
        L = _lib.lib()
        y = torch.empty((B, self.out_channels, T_out), dtype=xin.dtype, device=x.device)
        nscratch = L.kvae_conv1d_scratch_bytes(self.in_channels, self.out_channels, K)
        scratch = torch.empty(nscratch, dtype=torch.uint8, device=x.device)
        _lib.check(L.kvae_conv1d_fwd(xin.data_ptr(), y.data_ptr(), w.data_ptr(), _lib.ptr(bias), int(self.transposed),
                                     B, self.in_channels, self.out_channels, T, K, s, d, p,
                                     _lib.dtype_code(xin.dtype), scratch.data_ptr(), nscratch,
                                     _lib.stream_ptr(x.device)))
        y = y if y.dtype == x.dtype else y.to(x.dtype)
        return y, xin, w

    def extra_repr(self):
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}, "
                f"padding={self.padding}, dilation={self.dilation}, bias={self.bias is not None}")


class WNConv1d(_WNConvBase):
    """weight_norm(nn.Conv1d(...)): weight_v [Cout, Cin, K], weight_g [Cout, 1, 1] (norm per out-channel)."""
    transposed = False

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, padding_mode="zeros", device=None, dtype=None):
        super().__init__()
        if groups != 1 or padding_mode != "zeros" or (isinstance(padding, str) and (padding != "same" or _single(stride) != 1)):
            raise NotImplementedError("kalle_audio_b200.WNConv1d: groups=1, zero padding with an integer amount or "
                                      "'same' at stride 1 only (the forms the Oobleck stack uses)")
        # padding='same' (the conv of DecoderBlock's use_nearest_upsample branch, autoencoders.py:90-95): torch pads
        # d(K-1) in total, the smaller half on the left
        self.same_padding = isinstance(padding, str)
        if self.same_padding:
            total = _single(dilation) * (_single(kernel_size) - 1)
            self._same_extra_right = total - 2 * (total // 2)
            conv_padding = padding
            padding = total // 2
        else:
            self._same_extra_right = 0
            conv_padding = padding
        conv = nn.Conv1d(in_channels, out_channels, kernel_size, stride=stride, padding=conv_padding, dilation=dilation,
                         bias=bias, device=device, dtype=dtype)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = (_single(kernel_size),)
        self.stride = (_single(stride),)
        self.padding = (_single(padding),)
        self.dilation = (_single(dilation),)
        self.groups = 1
        self._init_from(conv)


class WNConvTranspose1d(_WNConvBase):
    """weight_norm(nn.ConvTranspose1d(...)): weight_v [Cin, Cout, K], weight_g [Cin, 1, 1] -- the norm is
    taken per *in*-channel because dim 0 of a transposed-conv weight is Cin."""
    transposed = True

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0, groups=1,
                 bias=True, dilation=1, padding_mode="zeros", device=None, dtype=None):
        super().__init__()
        if groups != 1 or padding_mode != "zeros" or _single(output_padding) != 0 or _single(dilation) != 1:
            raise NotImplementedError("kalle_audio_b200.WNConvTranspose1d: groups=1, output_padding=0, dilation=1 only")
        conv = nn.ConvTranspose1d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, bias=bias,
                                  device=device, dtype=dtype)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = (_single(kernel_size),)
        self.stride = (_single(stride),)
        self.padding = (_single(padding),)
        self.dilation = (1,)
        self.output_padding = (0,)
        self.groups = 1
        self._init_from(conv)


def nearest_upsample_conv_taps(w: torch.Tensor, stride: int) -> torch.Tensor:
    """Folded taps of ``Upsample(scale_factor=stride, mode='nearest')`` followed by ``Conv1d(k = 2*stride, padding='same')``
    (DecoderBlock's use_nearest_upsample branch, autoencoders.py:87-96) as ONE transposed convolution: ``w`` [Cout, Cin,
    2s] -> ConvTranspose1d weight [Cin, Cout, 3s - 1] for stride s, padding s, output_padding 1, with
    ``w'[k'] = sum_{k = 2s-1-k'}^{3s-2-k'} w[k]`` (k clipped to 0 .. 2s-1): every group of s upsampled positions that
    reads the same input sample contributes the sum of the taps that land on it."""
    Cout, Cin, K = w.shape
    s = int(stride)
    if K != 2 * s:
        raise ValueError("the nearest-upsample conv has kernel_size == 2 * stride")
    wp = w.new_zeros(Cin, Cout, 3 * s - 1)
    for kp in range(3 * s - 1):
        lo, hi = max(0, 2 * s - 1 - kp), min(2 * s - 1, 3 * s - 2 - kp)
        wp[:, :, kp] = w[:, :, lo:hi + 1].sum(-1).t()
    return wp.contiguous()


class NearestUpsampleConv(nn.Sequential):
    """``nn.Sequential(nn.Upsample(scale_factor=stride, mode="nearest"), WNConv1d(k=2*stride, padding='same',
    bias=False))`` exactly as the reference builds it (same state_dict keys ``0`` / ``1.weight_g`` / ``1.weight_v``).
    Stand-alone it chains the two modules; inside an OobleckDecoder plan the pair runs as one tcgen05 transposed conv."""

    def __init__(self, in_channels, out_channels, stride):
        super().__init__(nn.Upsample(scale_factor=stride, mode="nearest"),
                         WNConv1d(in_channels=in_channels, out_channels=out_channels, kernel_size=2 * stride, stride=1,
                                  bias=False, padding="same"))
        self.upsample_stride = int(stride)
        self[1]._upsample_stride = int(stride)      # the plan folds this conv's weight into transposed-conv taps
