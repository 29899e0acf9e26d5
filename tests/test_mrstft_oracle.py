"""CPU: the float64 restatement of the multi-resolution STFT loss (oracle/mrstft_oracle.py) against the reference's own
auraloss results recorded in tests/golden/mrstft.npz (value + autograd gradients w.r.t. both arguments)."""
import numpy as np
import pytest

import helpers as H
from oracle import mrstft_oracle as MO

A = dict(fft_sizes=[2048, 1024, 512, 256, 128, 64, 32], hop_sizes=[512, 256, 128, 64, 32, 16, 8],
         win_lengths=[2048, 1024, 512, 256, 128, 64, 32])


@pytest.mark.parametrize("tag,xk,yk,kw", [
    ("sd", "x2", "y2", dict(sum_diff=True, aw=True, **A)),
    ("mr_stereo", "x2", "y2", dict(aw=True, **A)),
    ("mr_mono", "x1", "y1", dict(aw=True, **A)),
    ("short_win", "x1", "y1", dict(fft_sizes=[1024, 2048, 512], hop_sizes=[120, 240, 50], win_lengths=[600, 1200, 240])),
])
def test_oracle_matches_reference(tag, xk, yk, kw):
    g = H.golden("mrstft")
    kw = dict(kw)
    taps = MO.aw_taps() if kw.pop("aw", False) else None
    if taps is not None:
        assert np.array_equal(taps, g["aw_taps_44100"])
    loss, gx, gy = MO.mrstft_loss(g[xk], g[yk], taps=taps, **kw)
    assert abs(loss - float(g[f"{tag}.loss"])) <= 2e-6 * float(g[f"{tag}.loss"])
    # the reference is fp32: near-empty bins (1 / magnitude in the log term) carry its rounding into the gradients
    assert np.abs(gx - g[f"{tag}.gx"]).max() <= 6e-3 * np.abs(g[f"{tag}.gx"]).max()
    assert np.abs(gy - g[f"{tag}.gy"]).max() <= 6e-3 * np.abs(g[f"{tag}.gy"]).max()


def test_oracle_gradient_is_the_derivative():
    """central differences on a small case (float64): the hand-written gradients are the derivative of the value"""
    rng = np.random.default_rng(0)
    x, y = 0.3 * rng.standard_normal((1, 2, 300)), 0.3 * rng.standard_normal((1, 2, 300))
    kw = dict(fft_sizes=[64, 32], hop_sizes=[16, 8], win_lengths=[64, 20], taps=MO.aw_taps()[40:61], sum_diff=True)
    _, gx, gy = MO.mrstft_loss(x, y, **kw)
    for arr, grad, which in ((x, gx, 0), (y, gy, 1)):
        for idx in [(0, 0, 5), (0, 1, 150), (0, 0, 299)]:
            e = np.zeros_like(arr)
            e[idx] = 1e-6
            args_p = (x + e, y) if which == 0 else (x, y + e)
            args_m = (x - e, y) if which == 0 else (x, y - e)
            num = (MO.mrstft_loss(*args_p, want_grad=False, **kw)[0] - MO.mrstft_loss(*args_m, want_grad=False, **kw)[0]) / 2e-6
            assert abs(num - grad[idx]) <= 1e-5 * max(1.0, abs(grad[idx])) + 1e-7
