"""CPU restatement of the reference's multi-resolution STFT loss.  TEST INFRASTRUCTURE ONLY (tests/, smoke, bench CPU
legs) -- the product path never imports this.

Follows /root/reference/stable_audio_tools/training/losses/auraloss.py:
  * FIRFilter "aw" (lines 70-162): A-weighting pre-filter, a 101-tap FIR applied to input and target with zero padding
    ntaps // 2 (the taps themselves come from scipy.signal: bilinear -> freqz -> firls, lines 111-131);
  * STFTLoss.stft (360-387): torch.stft(center=True, reflect padding, periodic Hann window zero-padded to fft_size,
    onesided), magnitude = sqrt(clamp(re^2 + im^2, min = eps = 1e-8));
  * SpectralConvergenceLoss (165-175): mean over signals of ||y_mag - x_mag||_F / ||y_mag||_F;
    STFTMagnitudeLoss (177-217): mean |log x_mag - log y_mag|; STFTLoss.forward (389-441): w_sc * sc + w_log_mag * log;
  * MultiResolutionSTFTLoss.forward (511-531): mean over the resolutions;
  * SumAndDifference (37-67) + SumAndDifferenceSTFTLoss.forward (580-606): (w_sum * L(sum) + w_diff * L(diff)) / 2.
numpy float64 throughout (the reference computes in float32; the tests carry the tolerance).  Gradients w.r.t. both
arguments are written out by hand (the same formulas the CUDA path uses), checked against the reference's autograd
results in tests/golden/mrstft.npz -- parity pinned.
"""
import numpy as np


def aw_taps(fs=44100, ntaps=101):
    """auraloss.py:111-131"""
    import scipy.signal
    f1, f2, f3, f4, A1000 = 20.598997, 107.65265, 737.86223, 12194.217, 1.9997
    nums = [(2 * np.pi * f4) ** 2 * (10 ** (A1000 / 20)), 0, 0, 0, 0]
    dens = np.polymul([1, 4 * np.pi * f4, (2 * np.pi * f4) ** 2], [1, 4 * np.pi * f1, (2 * np.pi * f1) ** 2])
    dens = np.polymul(np.polymul(dens, [1, 2 * np.pi * f3]), [1, 2 * np.pi * f2])
    b, a = scipy.signal.bilinear(nums, dens, fs=fs)
    w_iir, h_iir = scipy.signal.freqz(b, a, worN=512, fs=fs)
    return scipy.signal.firls(ntaps, w_iir, abs(h_iir), fs=fs).astype("float32")


def hann_padded(win_length, n_fft):
    """torch.hann_window (periodic) centred in an n_fft frame, as torch.stft pads a short window"""
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(win_length) / win_length)
    out = np.zeros(n_fft)
    left = (n_fft - win_length) // 2
    out[left:left + win_length] = w
    return out


def _fir(sig, taps):          # F.conv1d(x, w, padding = ntaps // 2) is a cross-correlation
    pad = len(taps) // 2
    return np.stack([np.correlate(np.pad(s, (pad, pad)), taps, mode="valid") for s in sig])


def _fir_T(g, taps):          # transpose of _fir
    pad = len(taps) // 2
    return np.stack([np.convolve(np.pad(s, (pad, pad)), taps, mode="valid") for s in g])


def _frames(sig, n_fft, hop):
    T = sig.shape[-1]
    pad = n_fft // 2
    idx = np.arange(-pad, T + pad)
    idx = np.where(idx < 0, -idx, idx)
    idx = np.where(idx >= T, 2 * (T - 1) - idx, idx)
    n_frames = 1 + T // hop
    pos = np.arange(n_frames)[:, None] * hop + np.arange(n_fft)[None, :]
    return idx[pos]            # [frames, n_fft] indices into the signal


def _one_resolution(X, Y, n_fft, hop, win_length, w_sc, w_log, eps=1e-8, want_grad=True):
    """X, Y: [M, T] float64.  Returns loss, dL/dX, dL/dY."""
    M, T = X.shape
    win = hann_padded(win_length, n_fft)
    fi = _frames(X, n_fft, hop)
    xs = np.fft.rfft(X[:, fi] * win, axis=-1)                 # [M, frames, F]
    ys = np.fft.rfft(Y[:, fi] * win, axis=-1)
    px, py = xs.real ** 2 + xs.imag ** 2, ys.real ** 2 + ys.imag ** 2
    xm, ym = np.sqrt(np.maximum(px, eps)), np.sqrt(np.maximum(py, eps))
    S1 = ((ym - xm) ** 2).sum((1, 2))
    S2 = (ym ** 2).sum((1, 2))
    sc = np.mean(np.sqrt(S1) / np.sqrt(S2))
    N = xm.size
    lm = np.abs(np.log(xm) - np.log(ym)).sum() / N
    loss = w_sc * sc + w_log * lm
    if not want_grad:
        return loss, None, None
    r1, r2 = np.sqrt(S1)[:, None, None], np.sqrt(S2)[:, None, None]
    sgn = np.sign(np.log(xm) - np.log(ym))
    gxm = w_sc * (xm - ym) / (M * r1 * r2) + w_log * sgn / (N * xm)
    gym = w_sc * ((ym - xm) / (M * r1 * r2) - r1 * ym / (M * r2 ** 3)) - w_log * sgn / (N * ym)
    gxm = np.where(px > eps, gxm, 0.0)
    gym = np.where(py > eps, gym, 0.0)

    def back(gm, spec, mag):
        G = gm * spec / mag                                   # d mag / d re = re / mag, d mag / d im = im / mag
        # d re[f] / d x[n] = cos(2 pi f n / N), d im[f] / d x[n] = -sin(...):  g[n] = Re(sum_f G[f] e^{+2 pi i f n / N})
        n = np.arange(n_fft)
        f = np.arange(n_fft // 2 + 1)
        E = np.exp(2j * np.pi * np.outer(f, n) / n_fft)       # [F, n_fft]
        gfr = (G @ E).real * win                              # [M, frames, n_fft]
        out = np.zeros((M, T))
        for m in range(M):
            np.add.at(out[m], fi, gfr[m])
        return out
    return loss, back(gxm, xs, xm), back(gym, ys, ym)


def mrstft_loss(x, y, fft_sizes, hop_sizes, win_lengths, taps=None, sum_diff=False, w_sc=1.0, w_log_mag=1.0,
                w_sum=1.0, w_diff=1.0, want_grad=True):
    """x = input, y = target, [B, C, T].  Returns (loss, dL/dx, dL/dy)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    B, C, T = x.shape
    if sum_diff:
        assert C == 2
        groups = [(x[:, 0] + x[:, 1], y[:, 0] + y[:, 1], w_sum / 2.0), (x[:, 0] - x[:, 1], y[:, 0] - y[:, 1], w_diff / 2.0)]
    else:
        groups = [(x.reshape(B * C, T), y.reshape(B * C, T), 1.0)]
    total = 0.0
    gparts = []
    for X, Y, wg in groups:
        if taps is not None:
            Xf, Yf = _fir(X, np.asarray(taps, np.float64)), _fir(Y, np.asarray(taps, np.float64))
        else:
            Xf, Yf = X, Y
        gX, gY = np.zeros_like(Xf), np.zeros_like(Yf)
        for n_fft, hop, wl in zip(fft_sizes, hop_sizes, win_lengths):
            l, a, b = _one_resolution(Xf, Yf, n_fft, hop, wl, w_sc, w_log_mag, want_grad=want_grad)
            total += wg * l / len(fft_sizes)
            if want_grad:
                gX += wg * a / len(fft_sizes)
                gY += wg * b / len(fft_sizes)
        if want_grad and taps is not None:
            gX, gY = _fir_T(gX, np.asarray(taps, np.float64)), _fir_T(gY, np.asarray(taps, np.float64))
        gparts.append((gX, gY))
    if not want_grad:
        return total, None, None
    if sum_diff:
        (gsx, gsy), (gdx, gdy) = gparts
        gx = np.stack([gsx + gdx, gsx - gdx], axis=1)
        gy = np.stack([gsy + gdy, gsy - gdy], axis=1)
    else:
        gx, gy = gparts[0][0].reshape(B, C, T), gparts[0][1].reshape(B, C, T)
    return total, gx, gy
