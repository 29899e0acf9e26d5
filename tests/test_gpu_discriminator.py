"""GPU parity tests of the Oobleck discriminator (discriminator half), through the Python
drop-in -> ctypes -> C ABI: the kvae_disc_* kernels one by one against torch's own ops, the strided k = 15 convolutions
on the layer-level conv kernels, and OobleckDiscriminator.loss with its hand-written backward chain against the
reference's recorded autograd gradients (tests/golden/disc.npz) and against the oracle at a longer clip."""
import ctypes as C
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import disc_common as dc      # noqa: E402
import helpers                # noqa: E402

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _grad_enabled():
    """other test modules switch autograd off globally at import; torch's own GPU convolutions are the comparison
    here, so they must not run in TF32"""
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.enable_grad():
            yield
    finally:
        torch.backends.cudnn.allow_tf32 = tf32


def _lib():
    from kalle_audio_b200 import _lib as L
    return L


def test_period_fold_and_avg_pool():
    import kalle_audio_b200.discriminators as D
    torch.manual_seed(0)
    for n, T in ((2, 1001), (3, 999), (5, 1000), (7, 64), (11, 3000), (11, 5)):
        x = torch.randn(3, 2, T, device=DEV, requires_grad=True)
        y = D._PeriodFoldFn.apply(x, n)
        pad = (n - T % n) % n
        want = F.pad(x, (0, pad)).reshape(3, 2, -1, n).permute(0, 1, 3, 2).reshape(3, 2 * n, -1)
        assert torch.equal(y, want.detach())
        r = torch.randn_like(y)
        (gx,) = torch.autograd.grad(y, x, r)
        (gw,) = torch.autograd.grad(want, x, r)
        assert torch.equal(gx, gw)
    for T in (1000, 1001, 2, 3):
        x = torch.randn(4, 2, T, device=DEV, requires_grad=True)
        y = D._AvgPool2Fn.apply(x)
        want = F.avg_pool1d(x, 2)
        assert torch.equal(y, want.detach())
        r = torch.randn_like(y)
        (gx,) = torch.autograd.grad(y, x, r)
        (gw,) = torch.autograd.grad(want, x, r)
        assert torch.equal(gx, gw)


@pytest.mark.parametrize("W", [1, 2, 3, 5, 7, 11])
def test_folded_conv2d_equals_conv2d(W):
    """Conv2d(15 x 15, stride 4, padding 7) on [N, C, H, W] == Conv1d over the folded channels with the folded weight;
    the weight unfold is the exact adjoint of the fold"""
    L = _lib()
    lib = L.lib()
    torch.manual_seed(W)
    Cin, Cout, K, s, p, N, H = 3, 8, 15, 4, 7, 2, 77
    w = torch.randn(Cout, Cin, K, K, device=DEV) * 0.05
    b = torch.randn(Cout, device=DEV)
    x = torch.randn(N, Cin, H, W, device=DEV)
    want = F.conv2d(x, w, b, stride=s, padding=p)
    Wo = lib.kvae_disc_folded_width(W, K, s, p)
    assert Wo == want.shape[3]
    wf = torch.empty(Cout * Wo, Cin * W, K, device=DEV)
    bf = torch.empty(Cout * Wo, device=DEV)
    L.check(lib.kvae_disc_fold_weight2d(w.data_ptr(), b.data_ptr(), wf.data_ptr(), bf.data_ptr(), Cout, Cin, K, s, p, W, 0,
                                        L.stream_ptr(x.device)))
    xf = x.permute(0, 1, 3, 2).reshape(N, Cin * W, H).contiguous()
    got = F.conv1d(xf, wf, bf, stride=s, padding=p).view(N, Cout, Wo, -1).transpose(2, 3)
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 2e-5 * float(want.abs().max())
    # adjoint: <fold(w), r> == <w, unfold(r)>
    r = torch.randn_like(wf)
    rb = torch.randn_like(bf)
    dw = torch.empty_like(w)
    db = torch.empty_like(b)
    L.check(lib.kvae_disc_fold_weight2d(dw.data_ptr(), db.data_ptr(), r.data_ptr(), rb.data_ptr(), Cout, Cin, K, s, p, W, 1,
                                        L.stream_ptr(x.device)))
    lhs, rhs = float((wf.double() * r.double()).sum()), float((w.double() * dw.double()).sum())
    assert abs(lhs - rhs) <= 1e-5 * max(1.0, abs(lhs))
    lhs, rhs = float((bf.double() * rb.double()).sum()), float((b.double() * db.double()).sum())
    assert abs(lhs - rhs) <= 1e-5 * max(1.0, abs(lhs))


def test_silu_score_hinge_feature_match():
    import kalle_audio_b200.discriminators as D
    L = _lib()
    lib = L.lib()
    st = L.stream_ptr(torch.device(DEV))
    torch.manual_seed(1)
    for n in (4099, 1 << 16):
        f = (3 * torch.randn(n, device=DEV)).requires_grad_(True)
        a = torch.empty(n, device=DEV)
        L.check(lib.kvae_disc_silu_fwd(f.data_ptr(), a.data_ptr(), n, st))
        want = F.silu(f)
        assert float((a - want.detach()).abs().max()) <= 1e-6
        ga, gfeat = torch.randn(n, device=DEV), torch.randn(n, device=DEV)
        gf = torch.empty(n, device=DEV)
        L.check(lib.kvae_disc_silu_bwd(f.data_ptr(), ga.data_ptr(), gfeat.data_ptr(), gf.data_ptr(), n, st))
        (gw,) = torch.autograd.grad(want, f, ga)
        assert float((gf - (gw + gfeat)).abs().max()) <= 2e-6
    # score mean + its backward
    y = torch.randn(6, 1, 331, device=DEV)
    score = torch.empty(6, device=DEV)
    L.check(lib.kvae_disc_score(y.data_ptr(), score.data_ptr(), 6, 331, 0, st))
    assert float((score - y.reshape(6, -1).mean(-1)).abs().max()) <= 1e-6
    # hinge losses, values and gradients
    for B in (1, 3, 300):
        s = (2 * torch.randn(2 * B, device=DEV)).requires_grad_(True)
        dis, gen = D._HingeFn.apply(s)
        wd = torch.relu(1 - s[:B]).mean() + torch.relu(1 + s[B:]).mean()
        wg = -s[B:].mean()
        assert abs(float(dis) - float(wd)) <= 1e-5 and abs(float(gen) - float(wg)) <= 1e-5
        (g1,) = torch.autograd.grad(0.7 * dis + 0.3 * gen, s)
        (g2,) = torch.autograd.grad(0.7 * wd + 0.3 * wg, s)
        assert float((g1 - g2).abs().max()) <= 1e-6
    # feature matching over many tensors (more than one parameter table: 60 > 56)
    feats = [torch.randn(4, 3, 5 + 17 * i, device=DEV, requires_grad=True) for i in range(60)]
    feats.append(torch.randn(2, 7, 70001, device=DEV, requires_grad=True))
    fm = D._FeatureMatchFn.apply(*feats)
    want = sum((f[: f.shape[0] // 2] - f[f.shape[0] // 2:]).abs().mean() for f in feats)
    assert abs(float(fm) - float(want)) <= 1e-5 * float(want)
    g1 = torch.autograd.grad(1.7 * fm, feats)
    g2 = torch.autograd.grad(1.7 * want, feats)
    for a, b in zip(g1, g2):
        assert float((a - b).abs().max()) <= 1e-7 + 1e-5 * float(b.abs().max())


@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("Cin,Cout,T", [(2, 32, 2999), (32, 64, 751), (96, 64, 70), (128, 256, 47), (256, 1, 12),
                                        (64, 128, 3001), (22, 96, 1030), (33, 65, 517), (128, 256, 1),
                                        (32, 64, 52001), (64, 128, 40001), (96, 64, 30002)])
def test_strided_k15_conv_fwd_bwd(Cin, Cout, T, generic, monkeypatch):
    """the discriminator's conv geometry (k = 15, stride 4, padding 7; k = 1 for the last layer) forward and all three
    gradients against torch, on the kernels written for it (kvae_disc_conv15_*: full and ragged tiles in every dimension,
    long and short layers = every tile-size instantiation)
    and on the generic layer-level conv kernels"""
    import kalle_audio_b200.discriminators as D
    monkeypatch.setattr(D, "_GENERIC", generic)
    torch.manual_seed(Cin + T)
    K, s, p = (15, 4, 7) if Cout > 1 else (1, 1, 0)
    N = 3
    x = torch.randn(N, Cin, T, device=DEV, requires_grad=True)
    w = (torch.randn(Cout, Cin, K, device=DEV) / (Cin * K) ** 0.5).requires_grad_(True)
    b = torch.randn(Cout, device=DEV, requires_grad=True)
    y = D._conv_fwd(x.detach(), w.detach(), b.detach(), Cin, Cout, K, s, p)
    want = F.conv1d(x, w, b, stride=s, padding=p)
    assert y.shape == want.shape
    assert float((y - want.detach()).abs().max()) <= 1e-5 * max(1.0, float(want.abs().max()))
    gy = torch.randn_like(y)
    gx, dw, db = D._conv_bwd(x.detach(), gy, w.detach(), Cin, Cout, K, s, p, True, True, True)
    gxw, dww, dbw = torch.autograd.grad(want, (x, w, b), gy)
    assert float((gx - gxw).abs().max()) <= 2e-5 * float(gxw.abs().max())
    assert float((dw - dww).abs().max()) <= 1e-4 * float(dww.abs().max())
    assert float((db - dbw).abs().max()) <= 1e-4 * float(dbw.abs().max())
    gx2, dw2, db2 = D._conv_bwd(x.detach(), gy, w.detach(), Cin, Cout, K, s, p, True, False, False)
    assert dw2 is None and db2 is None and torch.equal(gx2, gx)
    _, dw3, _ = D._conv_bwd(x.detach(), gy, w.detach(), Cin, Cout, K, s, p, False, True, False)
    assert float((dw3 - dww).abs().max()) <= 1e-4 * float(dww.abs().max())


def test_conv1x1_small_cout():
    """the nets' last layer on its streaming kernels with more than one output channel (a folded width > 1)"""
    import kalle_audio_b200.discriminators as D
    torch.manual_seed(9)
    N, Cin, Cout, T = 3, 70, 3, 517
    x = torch.randn(N, Cin, T, device=DEV, requires_grad=True)
    w = (torch.randn(Cout, Cin, 1, device=DEV) / Cin ** 0.5).requires_grad_(True)
    b = torch.randn(Cout, device=DEV, requires_grad=True)
    y = D._conv_fwd(x.detach(), w.detach(), b.detach(), Cin, Cout, 1, 1, 0)
    want = F.conv1d(x, w, b)
    assert float((y - want.detach()).abs().max()) <= 1e-5 * float(want.abs().max())
    gy = torch.randn_like(y)
    gx, dw, db = D._conv_bwd(x.detach(), gy, w.detach(), Cin, Cout, 1, 1, 0, True, True, True)
    gxw, dww, dbw = torch.autograd.grad(want, (x, w, b), gy)
    assert float((gx - gxw).abs().max()) <= 2e-5 * float(gxw.abs().max())
    assert float((dw - dww).abs().max()) <= 1e-4 * float(dww.abs().max())
    assert float((db - dbw).abs().max()) <= 1e-4 * float(dbw.abs().max())


@pytest.mark.parametrize("tag", ["stereo", "mono"])
def test_discriminator_vs_reference(tag):
    """OobleckDiscriminator at the reference's random init: loss values, summed scores, every feature tensor's shape and
    checksum, d / d fakes of each loss and d / d parameters -- against what the reference's module + autograd produced"""
    m, g = dc.build(tag, DEV)
    dc.check_init(m, g, tag)
    dc.check_forward_dict(m, g, tag, DEV)
    dc.check_loss_and_grads(m, g, tag, DEV, report=helpers.report)


def test_discriminator_vs_oracle_long_clip():
    """one second of 44.1 kHz stereo, batch 2: values and the generator-side gradient against the oracle run on the box"""
    sys.path.insert(0, ROOT)
    from oracle import discriminator_oracle as O
    import kalle_audio_b200.discriminators as D
    torch.manual_seed(5)
    m = D.OobleckDiscriminator(in_channels=2)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.to(DEV)
    g = torch.Generator().manual_seed(6)
    T = 44100
    t = torch.arange(T, dtype=torch.float32) / 44100.0
    base = 0.3 * torch.sin(2 * torch.pi * 220.0 * t) + 0.1 * torch.sin(2 * torch.pi * 5000.0 * t)
    reals = base + 0.05 * torch.randn(2, 2, T, generator=g)
    fakes = (0.8 * base + 0.08 * torch.randn(2, 2, T, generator=g)).requires_grad_(True)
    dis_o, gen_o, fm_o = O.oobleck_discriminator_loss(sd, reals, fakes)
    (g_o,) = torch.autograd.grad(gen_o + fm_o, fakes)
    fk = fakes.detach().to(DEV).requires_grad_(True)
    for p in m.parameters():
        p.requires_grad_(False)
    dis, gen, fm = m.loss(reals.to(DEV), fk)
    (g_k,) = torch.autograd.grad(gen + fm, fk)
    for a, b in ((dis, dis_o), (gen, gen_o), (fm, fm_o)):
        assert abs(float(a) - float(b)) <= 2e-4 * max(1e-3, abs(float(b)))
    err = float((g_k.cpu() - g_o).abs().max()) / float(g_o.abs().max())
    helpers.report("OobleckDiscriminator, 2 x 1 s stereo: d (gen + fm) / d fakes vs the oracle (max-abs over max)", err)
    assert err <= 5e-3
