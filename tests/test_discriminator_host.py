"""CPU tests of the Oobleck discriminator: the oracle against the reference's recorded
outputs, the drop-in's module tree / initialisation, and the host layer's forward + hand-written backward chain run over
a CPU stand-in for libkvae (tests/_fake_disc_lib.py) against the reference's autograd gradients."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import disc_common as dc      # noqa: E402
import helpers                # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(autouse=True)
def _grad_enabled():
    """other test modules switch autograd off globally at import"""
    with torch.enable_grad():
        yield


def _oracle():
    sys.path.insert(0, ROOT)
    from oracle import discriminator_oracle as O
    return O


@pytest.mark.parametrize("tag", ["stereo", "mono"])
def test_oracle_matches_reference(tag):
    """oracle/discriminator_oracle.py vs the reference's OobleckDiscriminator (values, scores, gradients)."""
    O = _oracle()
    m, g = dc.build(tag)
    dc.check_init(m, g, tag)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    reals = torch.from_numpy(g[f"{tag}.reals"])
    fakes = torch.from_numpy(g[f"{tag}.fakes"]).requires_grad_(True)
    dis, gen, fm = O.oobleck_discriminator_loss(sd, reals, fakes)
    for k, v in (("dis", dis), ("gen", gen), ("fm", fm)):
        assert abs(float(v) - float(g[f"{tag}.{k}"])) <= 2e-6 * max(1.0, abs(float(g[f"{tag}.{k}"])))
    for k, v in (("dis", dis), ("gen", gen), ("fm", fm)):
        (gf,) = torch.autograd.grad(v, fakes, retain_graph=True)
        want = torch.from_numpy(g[f"{tag}.g_fakes.{k}"])
        assert float((gf - want).abs().max()) <= 1e-5 * float(want.abs().max()) + 1e-10
    B = reals.shape[0]
    scores, feats = O.oobleck_discriminator(sd, torch.cat([reals, fakes.detach()], 0))
    assert float((scores[:B] - torch.from_numpy(g[f"{tag}.score_reals"])).abs().max()) <= 2e-6
    assert len(feats) == int(g[f"{tag}.n_features"])
    for i, f in enumerate(feats):
        assert tuple(f[:B].shape) == tuple(int(v) for v in g[f"{tag}.feat{i}.shape"])


def test_state_dict_layout():
    """keys, shapes and registration order of the reference (discriminators.py:85-106 under old-style weight_norm)"""
    import kalle_audio_b200.discriminators as D
    m = D.OobleckDiscriminator(in_channels=2)
    keys = list(m.state_dict().keys())
    p = "multi_discriminator.discriminators.0.layers.0.net."
    assert keys[:5] == [p + "0.bias", p + "0.weight_g", p + "0.weight_v", p + "2.bias", p + "2.weight_g"]
    assert p + "8.weight" in keys and p + "8.bias" in keys
    sd = m.state_dict()
    assert tuple(sd[p + "0.weight_v"].shape) == (32, 2, 15)
    q = "multi_discriminator.discriminators.1.layers.4.net."
    assert tuple(sd[q + "6.weight_v"].shape) == (256, 128, 15, 15)
    assert tuple(sd[q + "6.weight_g"].shape) == (256, 1, 1, 1)
    assert tuple(sd[q + "8.weight"].shape) == (1, 256, 1, 1)
    assert len(keys) == 112 and sum(v.numel() for v in sd.values()) == 50403976
    assert m.multi_discriminator.discriminators[1].periods == [2, 3, 5, 7, 11]
    with pytest.raises(NotImplementedError):
        D.EncodecDiscriminator(in_channels=2)
    with pytest.raises(NotImplementedError):
        D.SharedDiscriminatorConvNet(2, torch.nn.Conv1d, activation=lambda: torch.nn.ReLU())


def test_no_cpu_path():
    import kalle_audio_b200.discriminators as D
    from kalle_audio_b200._lib import KvaeError
    m = D.OobleckDiscriminator(in_channels=1)
    with pytest.raises(KvaeError):
        m.loss(torch.zeros(1, 1, 256), torch.zeros(1, 1, 256))


@pytest.fixture
def fake_lib(monkeypatch):
    from kalle_audio_b200 import _lib
    from _fake_disc_lib import FakeLib
    fake = FakeLib()
    monkeypatch.setattr(_lib, "lib", lambda: fake)
    monkeypatch.setattr(_lib, "require_cuda", lambda t, what: None)
    monkeypatch.setattr(_lib, "stream_ptr", lambda device: 0)
    return fake


@pytest.mark.parametrize("tag", ["stereo", "mono"])
def test_host_chain_vs_reference_autograd(fake_lib, tag):
    """forward chain, folded 2-D convs and the hand-written backward chain of _SharedNetFn, over the stand-in library"""
    m, g = dc.build(tag)
    dc.check_init(m, g, tag)
    dc.check_forward_dict(m, g, tag, "cpu")
    dc.check_loss_and_grads(m, g, tag, "cpu")


def test_generator_step_skips_weight_gradients(fake_lib):
    """with the discriminator's parameters frozen only the data gradients are computed (dw = NULL at the C ABI)"""
    m, g = dc.build("mono")
    for p in m.parameters():
        p.requires_grad_(False)
    reals = torch.from_numpy(g["mono.reals"])
    fakes = torch.from_numpy(g["mono.fakes"]).requires_grad_(True)
    dis, gen, fm = m.loss(reals, fakes)
    fake_lib.calls.clear()
    (gf,) = torch.autograd.grad(gen + fm, fakes)
    bwd = [c for c in fake_lib.calls if c[0] == "conv_bwd"]
    assert len(bwd) == 40 and all(c[1] and not c[2] for c in bwd)
    want = torch.from_numpy(g["mono.g_fakes.gen"] + g["mono.g_fakes.fm"])
    assert float((gf - want).abs().max()) <= 5e-3 * float(want.abs().max())


def test_partial_gradients(fake_lib):
    """a loss that touches a single middle feature: nothing flows into the later convs, no (or an all-zero) gradient for their parameters"""
    import kalle_audio_b200.discriminators as D
    O = _oracle()
    torch.manual_seed(3)
    net = D.SharedDiscriminatorConvNet(2, torch.nn.Conv2d)
    x = (0.1 * torch.randn(2, 2, 90, 5)).requires_grad_(True)
    score, feats = net(x)
    loss = feats[1].square().sum()
    grads = torch.autograd.grad(loss, [x] + list(net.parameters()), allow_unused=True)
    sd = {"d." + k: v.detach() for k, v in net.state_dict().items()}
    x2 = x.detach().clone().requires_grad_(True)
    _, f2 = O.shared_convnet(sd, "d", x2, True)
    for a, b in zip(feats, f2):
        assert a.shape == b.shape and float((a - b).abs().max()) <= 1e-5
    (gx2,) = torch.autograd.grad(f2[1].square().sum(), x2)
    assert float((grads[0] - gx2).abs().max()) <= 1e-4 * float(gx2.abs().max())
    names = [k for k, _ in net.named_parameters()]
    for k, gp in zip(names, grads[1:]):
        layer = int(k.split(".")[1])
        assert (gp is None or not bool(gp.any())) == (layer > 2), k
