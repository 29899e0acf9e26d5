"""GPU parity of the BigVGANFlowVAE inference path (reference backup/flows.py:396-529) through
libkvae's kernels, against the reference's recorded outputs (tests/golden/bigvgan.npz) and the oracle."""
import os
import sys

import numpy as np
import pytest
import torch

import helpers as H
import kalle_audio_b200 as k
from kalle_audio_b200 import bigvgan as BV
from oracle import alias_free_restated as AF
from oracle import bigvgan_oracle as BO

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden import BIGVGAN_H  # noqa: E402

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


class AttrDict(dict):
    __getattr__ = dict.__getitem__


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.mark.parametrize("T", [1, 5, 37, 256, 777])
def test_anti_aliased_activation_kernel(dev, T):
    """kvae_aa_act_fwd == alias_free_torch.Activation1d(SnakeBeta | Snake) (upsample x2, activation, downsample x2,
    replicate padding at both ends), incl. rows shorter than the filters and rows spanning several blocks."""
    torch.manual_seed(T)
    for cls, logscale in ((BV.SnakeBeta, True), (BV.Snake, False)):
        act = BV.Activation1d(cls(6, alpha_logscale=logscale))
        act.act.alpha.data = (0.3 * torch.randn(6)) if logscale else (1.0 + 0.3 * torch.rand(6))
        if hasattr(act.act, "beta"):
            act.act.beta.data = 0.3 * torch.randn(6)
        x = torch.randn(3, 6, T)
        u = AF.UpSample1d(2, 12)(x)
        v = BO.snake(u, act.act.alpha, getattr(act.act, "beta", None), logscale)
        ref = AF.DownSample1d(2, 12)(v)
        y = act.to(dev)(x.to(dev))
        assert y.shape == ref.shape
        assert float((y.cpu() - ref).abs().max()) <= 5e-6 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("tag,causal", [("causal", True), ("noncausal", False)])
def test_bigvgan_inference_matches_reference(dev, tag, causal):
    g = H.golden("bigvgan")
    sd = {kk[3:]: H.t(g[kk]) for kk in g.files if kk.startswith("sd.")}
    m = BV.BigVGANFlowVAE(AttrDict(BIGVGAN_H, causal=causal)).eval()
    m.load_state_dict(sd, strict=True)           # same keys as the reference's module (380 entries)
    m.to(dev)
    lat = m.extract_latents(H.t(g[f"{tag}.x"]).to(dev))
    e1 = float((lat.cpu() - H.t(g[f"{tag}.latents"])).abs().max())
    assert lat.shape == g[f"{tag}.latents"].shape and e1 <= 2e-5 * max(1.0, float(np.abs(g[f"{tag}.latents"]).max()))
    y = m.inference_from_latents(H.t(g[f"{tag}.z"]).to(dev), do_sample=False)
    e2 = float((y.cpu() - H.t(g[f"{tag}.wav"])).abs().max())
    ys = m.inference_from_latents(H.t(g[f"{tag}.latents"]).to(dev), noise=H.t(g[f"{tag}.noise"]).to(dev))
    e3 = float((ys.cpu() - H.t(g[f"{tag}.wav_sampled"])).abs().max())
    H.report(f"BigVGANFlowVAE ({tag}): extract_latents / inference_from_latents / sampled", f"{e1:.2e} / {e2:.2e} / {e3:.2e}")
    assert y.shape == g[f"{tag}.wav"].shape and e2 <= 1e-5 and e3 <= 1e-5
    # RNG-stream parity of the sampling branch, batch independence, error behaviour
    torch.manual_seed(3)
    a = m.inference_from_latents(lat)
    torch.manual_seed(3)
    nz = torch.randn(2, 16, lat.shape[2], device=dev)
    assert torch.equal(a, m.inference_from_latents(lat, noise=nz))
    assert torch.equal(y[1:2], m.inference_from_latents(H.t(g[f"{tag}.z"])[1:2].to(dev), do_sample=False))
    with pytest.raises(AssertionError):
        m.inference_from_latents(lat[:, :5])
    with pytest.raises(NotImplementedError):
        m(H.t(g[f"{tag}.x"]).to(dev))


def test_bigvgan_wider_model_vs_oracle(dev):
    """A config with tensor-core-sized channel counts (initial 256 -> 128 -> 64, AMPBlock2, snake) against the oracle."""
    h = AttrDict(BIGVGAN_H, causal=False, upsample_initial_channel=256, resblock="2", activation="snake",
                 resblock_dilation_sizes=[[1, 3], [1, 3]], snake_logscale=False)
    torch.manual_seed(1)
    m = BV.BigVGANFlowVAE(h).eval()
    sd = {n: p.clone() for n, p in m.state_dict().items()}
    z = torch.randn(2, 16, 50, generator=torch.Generator().manual_seed(2))
    ref = BO.inference_from_latents(sd, dict(h), z)
    m.to(dev)
    errs = {}
    for prec in (None, "fp32", "bf16"):          # CUDA cores / tensor cores via the bf16x3 split / bf16 tensor-core operands
        y = m.set_precision(prec).inference_from_latents(z.to(dev), do_sample=False)
        assert y.shape == ref.shape == (2, 1, 400)
        errs[prec] = float((y.cpu() - ref).abs().max())
    H.report("BigVGANFlowVAE wide config: CUDA-core / tensor-core fp32 mode / bf16 mode", " / ".join(f"{v:.2e}" for v in errs.values()))
    assert errs[None] <= 2e-5 and errs["fp32"] <= 2e-5 and errs["bf16"] <= 5e-3 * max(1.0, float(ref.abs().max()))
    assert errs["bf16"] > errs["fp32"]           # the reduced-precision path really ran
    # causal form on the tensor cores: truncation inside the kernel (no slicing copy)
    hc = AttrDict(h, causal=True, resblock="1", activation="snakebeta", resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5]],
                  snake_logscale=True)
    torch.manual_seed(2)
    mc = BV.BigVGANFlowVAE(hc).eval()
    sdc = {n: p.clone() for n, p in mc.state_dict().items()}
    zc = torch.randn(1, 16, 90, generator=torch.Generator().manual_seed(3))
    refc = BO.inference_from_latents(sdc, dict(hc), zc)
    yc = mc.to(dev).inference_from_latents(zc.to(dev), do_sample=False)
    assert float((yc.cpu() - refc).abs().max()) <= 2e-5


def test_folded_weight_cache_follows_the_weights(dev):
    """The inference-only model folds weight norm once per weight version (layers.py, ``_cache_fold``): in-place updates
    and load_state_dict must invalidate the cached fold."""
    g = H.golden("bigvgan")
    sd = {kk[3:]: H.t(g[kk]) for kk in g.files if kk.startswith("sd.")}
    m = BV.BigVGANFlowVAE(AttrDict(BIGVGAN_H, causal=True)).eval()
    m.load_state_dict(sd, strict=True)
    m.to(dev)
    z = H.t(g["causal.z"]).to(dev)
    y0 = m.inference_from_latents(z, do_sample=False)
    assert m.conv_pre.__dict__.get("_fold_cache") is not None            # the cache is in use
    assert torch.equal(y0, m.inference_from_latents(z, do_sample=False))
    with torch.no_grad():
        m.conv_pre.weight_g.mul_(1.25)                                   # bumps the version counter
    y1 = m.inference_from_latents(z, do_sample=False)
    assert not torch.equal(y0, y1)
    fresh = BV.BigVGANFlowVAE(AttrDict(BIGVGAN_H, causal=True)).eval()
    sd2 = dict(sd)
    sd2["conv_pre.weight_g"] = sd["conv_pre.weight_g"] * 1.25
    fresh.load_state_dict(sd2, strict=True)
    assert torch.equal(y1, fresh.to(dev).inference_from_latents(z, do_sample=False))
    m.load_state_dict(sd, strict=True)                                   # copy_ into the parameters: version bump again
    assert torch.equal(y0, m.inference_from_latents(z, do_sample=False))
