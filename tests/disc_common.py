"""Shared checks of kalle_audio_b200.discriminators against the reference's recorded outputs (tests/golden/disc.npz,
made by make_golden.py from the reference's own OobleckDiscriminator).  Used on the CPU over the stand-in library
(test_discriminator_host.py: host orchestration) and on the GPU over libkvae.so (test_gpu_discriminator.py: parity)."""
import numpy as np
import torch

import helpers

TAGS = {"stereo": 2, "mono": 1}


def build(tag, device="cpu"):
    import kalle_audio_b200.discriminators as D
    g = helpers.golden("disc")
    torch.manual_seed(int(g[f"{tag}.seed"]))
    m = D.OobleckDiscriminator(in_channels=TAGS[tag])
    return m.to(device), g


def check_init(m, g, tag):
    sd = m.state_dict()
    keys = [k[len(tag) + 4:] for k in g.files if k.startswith(f"{tag}.cs.")]
    assert sorted(sd.keys()) == sorted(keys), "state_dict keys differ from the reference"
    for k in keys:
        got = float(sd[k].double().abs().sum())
        want = float(g[f"{tag}.cs.{k}"])
        assert abs(got - want) <= 1e-9 * max(1.0, abs(want)), f"parameter {k} differs from the reference init"


def check_loss_and_grads(m, g, tag, device, rel=2e-4, report=None):
    """loss(reals, fakes): the three values, the summed scores, and autograd through the hand-written backward chain --
    d / d fakes of each loss and d / d parameters of the discriminator and feature-matching terms -- against the
    reference's autograd."""
    reals = torch.from_numpy(g[f"{tag}.reals"]).to(device)
    fakes = torch.from_numpy(g[f"{tag}.fakes"]).to(device).requires_grad_(True)
    dis, gen, fm = m.loss(reals, fakes)
    vals = {"dis": dis, "gen": gen, "fm": fm}
    worst = 0.0
    for k, v in vals.items():
        want = float(g[f"{tag}.{k}"])
        err = abs(float(v) - want) / max(1e-3, abs(want))
        worst = max(worst, err)
        assert err <= rel, f"{tag}.{k}: {float(v)} vs {want}"
    names = [k for k, _ in m.named_parameters()]
    params = [p for _, p in m.named_parameters()]
    worst_g, worst_p = 0.0, 0.0
    for lname in ("dis", "gen", "fm"):
        gr = torch.autograd.grad(vals[lname], [fakes] + params, retain_graph=True, allow_unused=True)
        want = torch.from_numpy(g[f"{tag}.g_fakes.{lname}"])
        got = gr[0].detach().cpu()
        scale = float(want.abs().max())
        err = float((got - want).abs().max()) / max(scale, 1e-12)
        worst_g = max(worst_g, err)
        assert err <= 5e-3, f"{tag}: d {lname} / d fakes off by {err:.3e} of its max"
        if lname == "gen":
            continue
        for k, gp in zip(names, gr[1:]):
            wn = float(g[f"{tag}.gp.{lname}.{k}.norm"])
            ws = torch.from_numpy(g[f"{tag}.gp.{lname}.{k}.sample"])
            if gp is None:
                assert wn == 0.0, f"{tag}: no gradient for {k} ({lname}) but the reference has norm {wn}"
                continue
            gp = gp.detach().cpu()
            gn = float(gp.double().norm())
            assert abs(gn - wn) <= 2e-3 * wn + 1e-6, f"{tag}: |d {lname} / d {k}| = {gn} vs {wn}"
            flat = gp.reshape(-1)
            gs = flat[:: max(1, flat.numel() // 61)][:64]
            tol = 2e-3 * max(float(ws.abs().max()), wn / max(1.0, flat.numel() ** 0.5)) + 1e-6
            e = float((gs - ws).abs().max())
            worst_p = max(worst_p, e / tol * 2e-3)
            assert e <= tol, f"{tag}: d {lname} / d {k} sample off by {e:.3e} (tol {tol:.3e})"
    if report:
        report(f"OobleckDiscriminator.loss ({tag}): values rel / d fakes (max-abs over max) / d params (scaled)",
               f"{worst:.2e} / {worst_g:.2e} / {worst_p:.2e}")


def check_forward_dict(m, g, tag, device):
    """MultiDiscriminator.forward's dict form: summed scores and a checksum of every feature tensor in the reference's
    own shapes and order"""
    reals = torch.from_numpy(g[f"{tag}.reals"]).to(device)
    fakes = torch.from_numpy(g[f"{tag}.fakes"]).to(device)
    with torch.no_grad():
        out = m.multi_discriminator({"reals": reals, "fakes": fakes})
    for k in ("reals", "fakes"):
        want = torch.from_numpy(g[f"{tag}.score_{k}"])
        got = out[f"score_{k}"].cpu()
        assert float((got - want).abs().max()) <= 2e-5 * max(1.0, float(want.abs().max()))
    n = int(g[f"{tag}.n_features"])
    fr, ff = out["features_reals"], out["features_fakes"]
    assert len(fr) == n and len(ff) == n
    for i in range(n):
        assert tuple(fr[i].shape) == tuple(int(v) for v in g[f"{tag}.feat{i}.shape"]), f"feature {i} shape"
        sums = g[f"{tag}.feat{i}.sums"]
        got = np.array([fr[i].double().sum().item(), fr[i].double().abs().sum().item(),
                        ff[i].double().sum().item(), ff[i].double().abs().sum().item()])
        tol = 2e-4 * max(sums[1], sums[3]) + 1e-6
        assert np.all(np.abs(got - sums) <= tol), f"{tag}: feature {i} checksums {got} vs {sums}"
