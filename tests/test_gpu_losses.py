"""GPU parity of the multi-resolution STFT losses (kvae_mrstft_loss through kalle_audio_b200.losses) against the
reference's own auraloss module: value and autograd gradients w.r.t. both arguments, recorded in
tests/golden/mrstft.npz by tests/golden/make_golden.py (the reference computes in fp32 on the CPU)."""
import numpy as np
import pytest
import torch

import helpers as H
import kalle_audio_b200 as k

pytestmark = pytest.mark.gpu

ARGS = dict(fft_sizes=[2048, 1024, 512, 256, 128, 64, 32], hop_sizes=[512, 256, 128, 64, 32, 16, 8],
            win_lengths=[2048, 1024, 512, 256, 128, 64, 32], perceptual_weighting=True, sample_rate=44100)


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _run(mod, x, y, dev):
    with torch.enable_grad():          # other test modules of the suite switch autograd off globally
        x = H.t(x).to(dev).requires_grad_(True)
        y = H.t(y).to(dev).requires_grad_(True)
        loss = mod(x, y)
        gx, gy = torch.autograd.grad(loss, (x, y))
    return float(loss.detach()), gx.cpu().numpy(), gy.cpu().numpy()


@pytest.mark.parametrize("tag,cls,kw,xk,yk", [
    ("sd", "SumAndDifferenceSTFTLoss", ARGS, "x2", "y2"),
    ("mr_stereo", "MultiResolutionSTFTLoss", ARGS, "x2", "y2"),
    ("mr_mono", "MultiResolutionSTFTLoss", ARGS, "x1", "y1"),
    ("short_win", "MultiResolutionSTFTLoss", dict(fft_sizes=[1024, 2048, 512], hop_sizes=[120, 240, 50],
                                                   win_lengths=[600, 1200, 240]), "x1", "y1"),
])
def test_mrstft_matches_reference(dev, tag, cls, kw, xk, yk):
    g = H.golden("mrstft")
    mod = getattr(k, cls)(**kw)
    if kw.get("perceptual_weighting"):
        assert np.array_equal(mod.fir_taps.numpy(), g["aw_taps_44100"])       # same scipy design as the reference
    loss, gx, gy = _run(mod, g[xk], g[yk], dev)
    ref = float(g[f"{tag}.loss"])
    el = abs(loss - ref) / ref
    ex = float(np.abs(gx - g[f"{tag}.gx"]).max() / np.abs(g[f"{tag}.gx"]).max())
    ey = float(np.abs(gy - g[f"{tag}.gy"]).max() / np.abs(g[f"{tag}.gy"]).max())
    H.report(f"MR-STFT loss ({tag}): value rel / grad input / grad target (max-abs over max)", f"{el:.2e} / {ex:.2e} / {ey:.2e}")
    assert el <= 2e-5
    # 1 / magnitude in the log term amplifies the fp32 rounding of near-empty bins: the reference's own fp32 result is
    # 1e-3 .. 4e-3 (of the largest gradient) away from a float64 evaluation of the same formulas (oracle test)
    assert ex <= 1e-2 and ey <= 1e-2
    # value only (no gradient buffers), and a scaled upstream gradient
    with torch.no_grad():
        assert abs(float(mod(H.t(g[xk]).to(dev), H.t(g[yk]).to(dev))) - loss) <= 1e-6 * abs(loss)
    with torch.enable_grad():
        y = H.t(g[yk]).to(dev).requires_grad_(True)
        (3.0 * mod(H.t(g[xk]).to(dev), y)).backward()
    assert np.abs(y.grad.cpu().numpy() - 3.0 * gy).max() <= 1e-5 * np.abs(gy).max() * 3.0 + 1e-9


def test_mrstft_properties_and_errors(dev):
    torch.manual_seed(0)
    mod = k.SumAndDifferenceSTFTLoss(**ARGS)
    x = 0.1 * torch.randn(2, 2, 44100, device=dev)
    # identical arguments: input and target share ONE complex FFT (z = x + i y) and come apart again through the
    # Hermitian split, so their spectra differ by rounding (1e-7 relative): ~1e-6 instead of the reference's exact 0
    assert float(mod(x, x.clone())) <= 1e-5
    # bf16 inputs are read as they are
    lb = float(mod(x.bfloat16(), (0.5 * x).bfloat16()))
    lf = float(mod(x.bfloat16().float(), (0.5 * x).bfloat16().float()))
    assert abs(lb - lf) <= 1e-5 * lf
    # scaling the target by 1/2: log distance = log 2 in every bin, sc = ||y/2 - y|| / ||y/2|| = 1
    mr = k.MultiResolutionSTFTLoss(fft_sizes=[256], hop_sizes=[64], win_lengths=[256])
    assert abs(float(mr(x, 0.5 * x)) - (1.0 + np.log(2.0))) <= 2e-4
    with pytest.raises(ValueError):
        mod(x[:, :1], x[:, :1])
    with pytest.raises(k.KvaeError):
        mod(x.cpu(), x.cpu())
    with pytest.raises(NotImplementedError):
        k.MultiResolutionSTFTLoss(w_lin_mag=1.0)
    with pytest.raises(k.KvaeError):
        mr(x[:, :, :100], x[:, :, :100])             # shorter than half an fft frame
