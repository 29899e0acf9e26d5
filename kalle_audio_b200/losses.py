"""Multi-resolution STFT losses of the autoencoder training wrapper, on the GPU through libkvae (csrc/mrstft.cuh).

Mirrors /root/reference/stable_audio_tools/training/losses/auraloss.py: ``MultiResolutionSTFTLoss`` (443-531),
``SumAndDifferenceSTFTLoss`` (534-606) and the A-weighting ``FIRFilter`` (70-162) they use with
``perceptual_weighting=True`` -- same constructor arguments, same call ``loss = module(input, target)`` (the training
wrapper passes input = reals, target = decoded, training/autoencoders.py:163), differentiable w.r.t. both arguments.
Value and gradients come out of ONE library call (`kvae_mrstft_loss`); nothing of spectrogram size is stored.
Not built (raise): linear-magnitude / phase terms, mel / chroma scales, scale invariance, non-mean reductions,
``output='full'`` -- none of them is used by the reference's autoencoder configs.  No CPU path."""
import ctypes as C
from typing import List, Optional

import numpy as np
import torch
from torch import nn

from . import _lib


def aw_fir_taps(fs: float, ntaps: int = 101) -> torch.Tensor:
    """A-weighting FIR of auraloss.FIRFilter(filter_type="aw") (auraloss.py:111-131): analog prototype (IEC/CD 1672)
    -> bilinear -> 512-point response -> least-squares fit.  Host-side filter design with scipy, as in the reference."""
    import scipy.signal
    if ntaps % 2 == 0:
        raise ValueError(f"ntaps must be odd (ntaps={ntaps}).")
    f1, f2, f3, f4, a1000 = 20.598997, 107.65265, 737.86223, 12194.217, 1.9997
    nums = [(2 * np.pi * f4) ** 2 * (10 ** (a1000 / 20)), 0, 0, 0, 0]
    dens = np.polymul([1, 4 * np.pi * f4, (2 * np.pi * f4) ** 2], [1, 4 * np.pi * f1, (2 * np.pi * f1) ** 2])
    dens = np.polymul(np.polymul(dens, [1, 2 * np.pi * f3]), [1, 2 * np.pi * f2])
    b, a = scipy.signal.bilinear(nums, dens, fs=fs)
    w_iir, h_iir = scipy.signal.freqz(b, a, worN=512, fs=fs)
    return torch.tensor(scipy.signal.firls(ntaps, w_iir, abs(h_iir), fs=fs).astype("float32"))


def _window(win_type: str, win_length: int) -> torch.Tensor:
    """auraloss.get_window (16-35)"""
    try:
        return getattr(torch, win_type)(win_length).float()
    except AttributeError:
        import scipy.signal
        return torch.from_numpy(scipy.signal.windows.get_window(win_type, win_length)).float()


class _MRSTFTFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, y):
        _lib.require_cuda(x, "MultiResolutionSTFTLoss")
        if x.shape != y.shape or x.dim() != 3:
            raise ValueError(f"expected two [B, C, T] tensors of the same shape, got {tuple(x.shape)} and {tuple(y.shape)}")
        B, Cc, T = x.shape
        dt = x.dtype if x.dtype in (torch.float32, torch.bfloat16) else torch.float32
        xin, yin = x.detach().to(dt).contiguous(), y.detach().to(dt).contiguous()
        need_x, need_y = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        L = _lib.lib()
        win, taps = mod._device_buffers(x.device)
        n_res = len(mod.fft_sizes)
        nscr = L.kvae_mrstft_scratch_bytes(B, Cc, T, n_res, int(mod.sum_diff), int(need_x or need_y))
        scratch = torch.empty(nscr, dtype=torch.uint8, device=x.device)
        loss = torch.empty(1, dtype=torch.float32, device=x.device)
        gx = torch.empty((B, Cc, T), dtype=torch.float32, device=x.device) if need_x else None
        gy = torch.empty((B, Cc, T), dtype=torch.float32, device=x.device) if need_y else None
        ffts = (C.c_int * n_res)(*mod.fft_sizes)
        hops = (C.c_int * n_res)(*mod.hop_sizes)
        _lib.check(L.kvae_mrstft_loss(xin.data_ptr(), yin.data_ptr(), B, Cc, T, _lib.dtype_code(dt), n_res,
                                      C.cast(ffts, C.c_void_p), C.cast(hops, C.c_void_p), win.data_ptr(), _lib.ptr(taps),
                                      0 if taps is None else taps.numel(), int(mod.sum_diff), mod.w_sum, mod.w_diff, mod.w_sc,
                                      mod.w_log_mag, loss.data_ptr(), _lib.ptr(gx), _lib.ptr(gy), scratch.data_ptr(), nscr,
                                      _lib.stream_ptr(x.device)))
        ctx.save_for_backward(*(t for t in (gx, gy) if t is not None))
        ctx.have = (need_x, need_y)
        ctx.dtypes = (x.dtype, y.dtype)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gl):
        saved = list(ctx.saved_tensors)
        gx = saved.pop(0) if ctx.have[0] else None
        gy = saved.pop(0) if ctx.have[1] else None
        return (None, None if gx is None else (gx * gl).to(ctx.dtypes[0]), None if gy is None else (gy * gl).to(ctx.dtypes[1]))


class MultiResolutionSTFTLoss(nn.Module):
    """auraloss.MultiResolutionSTFTLoss (443-531).  Every channel of a [B, C, T] input is one signal (view(-1, T))."""
    sum_diff = False
    w_sum = w_diff = 1.0

    def __init__(self, fft_sizes: List[int] = [1024, 2048, 512], hop_sizes: List[int] = [120, 240, 50],
                 win_lengths: List[int] = [600, 1200, 240], window: str = "hann_window", w_sc: float = 1.0,
                 w_log_mag: float = 1.0, w_lin_mag: float = 0.0, w_phs: float = 0.0, sample_rate: Optional[float] = None,
                 scale: Optional[str] = None, n_bins: Optional[int] = None, perceptual_weighting: bool = False,
                 scale_invariance: bool = False, **kwargs):
        super().__init__()
        assert len(fft_sizes) == len(hop_sizes) == len(win_lengths)  # must define all
        if w_lin_mag or w_phs or scale is not None or scale_invariance:
            raise NotImplementedError("kalle_audio_b200 builds the spectral-convergence + log-magnitude terms the "
                                      "reference's autoencoder configs use (w_lin_mag, w_phs, scale, scale_invariance are not)")
        if kwargs.get("output", "loss") != "loss" or kwargs.get("reduction", "mean") != "mean" or \
                kwargs.get("mag_distance", "L1") != "L1":
            raise NotImplementedError("only output='loss', reduction='mean', mag_distance='L1'")
        if perceptual_weighting and sample_rate is None:
            raise ValueError("`sample_rate` must be supplied when `perceptual_weighting = True`.")
        self.fft_sizes, self.hop_sizes, self.win_lengths = list(fft_sizes), list(hop_sizes), list(win_lengths)
        self.w_sc, self.w_log_mag = float(w_sc), float(w_log_mag)
        wins = []
        for n, wl in zip(self.fft_sizes, self.win_lengths):
            if wl > n:
                raise ValueError("win_length must be <= fft_size")
            w = torch.zeros(n)
            left = (n - wl) // 2                      # torch.stft centres a short window in the fft frame
            w[left:left + wl] = _window(window, wl)
            wins.append(w)
        self.register_buffer("windows", torch.cat(wins), persistent=False)
        self.register_buffer("fir_taps", aw_fir_taps(sample_rate) if perceptual_weighting else None, persistent=False)

    def _device_buffers(self, device):
        if self.windows.device != device:
            self.windows = self.windows.to(device)
            if self.fir_taps is not None:
                self.fir_taps = self.fir_taps.to(device)
        return self.windows, self.fir_taps

    def forward(self, x, y):
        return _MRSTFTFn.apply(self, x, y)


class SumAndDifferenceSTFTLoss(MultiResolutionSTFTLoss):
    """auraloss.SumAndDifferenceSTFTLoss (534-606): the multi-resolution loss on the sum and the difference of a stereo
    pair, (w_sum * L(sum) + w_diff * L(diff)) / 2."""
    sum_diff = True

    def __init__(self, fft_sizes: List[int], hop_sizes: List[int], win_lengths: List[int], window: str = "hann_window",
                 w_sum: float = 1.0, w_diff: float = 1.0, output: str = "loss", **kwargs):
        super().__init__(fft_sizes, hop_sizes, win_lengths, window, output=output, **kwargs)
        self.w_sum, self.w_diff = float(w_sum), float(w_diff)

    def forward(self, input, target):
        assert input.shape == target.shape  # must have same shape
        if input.size(1) != 2:
            raise ValueError(f"Input must be stereo: {input.size(1)} channel(s).")
        return _MRSTFTFn.apply(self, input, target)
