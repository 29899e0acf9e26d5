"""CPU checks of the training-step oracle (oracle/oobleck_oracle.py::training_loss + torch autograd) against
gradients recorded from autograd through the REFERENCE's own modules (tests/golden/train_*.npz, made by
tests/golden/make_golden.py), and of the host-side gradient synchronisation over gloo (world size 2)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

import helpers as H
from oracle import oobleck_oracle as O


@pytest.fixture(autouse=True)
def _grad_mode():
    """Other test modules switch autograd off globally at import; the training tests need it on."""
    with torch.enable_grad():
        yield


def oracle_grads(sd, x, noise, strides, kl_weight, log_sigma):
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss, nll, kl, y = O.training_loss(sd, x, noise, strides, kl_weight, log_sigma)
    loss.backward()
    return loss.detach(), nll.detach(), kl.detach(), y.detach(), {k: v.grad for k, v in sd.items()}


def test_oracle_training_grads_tiny_all_parameters():
    g = H.golden("train_tiny")
    sd = {k[3:]: H.t(g[k]) for k in g.files if k.startswith("sd.")}
    loss, nll, kl, y, grads = oracle_grads(sd, H.t(g["x"]), H.t(g["noise"]), H.strides_of("tiny"),
                                           float(g["kl_weight"]), float(g["log_sigma"]))
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert abs(float(kl) - float(g["kl"])) <= 1e-5 * abs(float(g["kl"]))
    assert float((y - H.t(g["decoded"])).abs().max()) <= 5e-6
    names = [k[2:] for k in g.files if k.startswith("g.")]
    assert sorted(names) == sorted(grads.keys())
    for n in names:
        ref = H.t(g["g." + n])
        err = float((grads[n] - ref).abs().max())
        assert err <= 1e-4 * max(1.0, float(ref.abs().max())), (n, err)


def test_oracle_training_grads_mid_summary():
    g = H.golden("train_mid")
    torch.manual_seed(0)
    # same construction as the fixture: the weights come from the recorded seed through torch's own constructors
    from torch import nn
    from torch.nn.utils import weight_norm
    import warnings
    warnings.filterwarnings("ignore")
    cfg = H.CONFIGS["mid"]["model"]
    sd = _reference_shaped_state_dict(cfg, seed=0, snake_seed=7)
    H.check_checksums(sd, g)
    loss, nll, kl, y, grads = oracle_grads(sd, H.t(g["x"]), H.t(g["noise"]), H.strides_of("mid"), float(g["kl_weight"]),
                                           float(g["log_sigma"]))
    assert abs(float(loss) - float(g["loss"])) <= 2e-5 * abs(float(g["loss"]))
    for i, n in enumerate(str(s) for s in g["g_keys"]):
        gr = grads[n]
        ref_norm = float(g["g_norm"][i])
        assert abs(float(gr.double().norm()) - ref_norm) <= 2e-4 * max(ref_norm, 1e-6), n
        samp = gr.reshape(-1)[::max(1, gr.numel() // 256)][:256]
        ref = H.t(g[f"g_sample{i}"])
        assert float((samp - ref).abs().max()) <= 2e-4 * max(1.0, float(ref.abs().max())), n


def _reference_shaped_state_dict(model_cfg, seed, snake_seed):
    """Builds the reference-keyed random-init state_dict with plain torch modules in the reference's construction
    order (what tests/helpers.build does through kalle_audio_b200, without needing the package's GPU library)."""
    import math
    from torch import nn
    from torch.nn.utils import weight_norm

    def wn(m):
        return weight_norm(m)

    class Snake(nn.Module):
        def __init__(self, c):
            super().__init__()
            self.alpha = nn.Parameter(torch.zeros(c))
            self.beta = nn.Parameter(torch.zeros(c))

    def ru(c, d):
        m = nn.Module()
        m.layers = nn.Sequential(Snake(c), wn(nn.Conv1d(c, c, 7, dilation=d, padding=3 * d)), Snake(c), wn(nn.Conv1d(c, c, 1)))
        return m

    def enc_block(cin, cout, s):
        m = nn.Module()
        m.layers = nn.Sequential(ru(cin, 1), ru(cin, 3), ru(cin, 9), Snake(cin),
                                 wn(nn.Conv1d(cin, cout, 2 * s, stride=s, padding=math.ceil(s / 2))))
        return m

    def dec_block(cin, cout, s):
        m = nn.Module()
        m.layers = nn.Sequential(Snake(cin), wn(nn.ConvTranspose1d(cin, cout, 2 * s + s % 2, stride=s, padding=math.ceil(s / 2))),
                                 ru(cout, 1), ru(cout, 3), ru(cout, 9))
        return m

    torch.manual_seed(seed)
    e, d = model_cfg["encoder"]["config"], model_cfg["decoder"]["config"]
    cm = [1] + list(e["c_mults"])
    ch = e["channels"]
    enc = nn.Module()
    layers = [wn(nn.Conv1d(e["in_channels"], cm[0] * ch, 7, padding=3))]
    for i, s in enumerate(e["strides"]):
        layers.append(enc_block(cm[i] * ch, cm[i + 1] * ch, s))
    layers += [Snake(cm[-1] * ch), wn(nn.Conv1d(cm[-1] * ch, e["latent_dim"], 3, padding=1))]
    enc.layers = nn.Sequential(*layers)
    dec = nn.Module()
    layers = [wn(nn.Conv1d(d["latent_dim"], cm[-1] * ch, 7, padding=3))]
    for i in range(len(cm) - 1, 0, -1):
        layers.append(dec_block(cm[i] * ch, cm[i - 1] * ch, d["strides"][i - 1]))
    layers += [Snake(cm[0] * ch), wn(nn.Conv1d(cm[0] * ch, d["out_channels"], 7, padding=3, bias=False))]
    dec.layers = nn.Sequential(*layers)
    top = nn.Module()
    top.encoder, top.decoder = enc, dec
    H.randomize_snake(top, snake_seed)
    return {k: v.detach().clone() for k, v in top.state_dict().items()}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_grad_sync_gloo_world2(tmp_path):
    """GradSync (the data-parallel gradient all-reduce of the training step) on CPU tensors over gloo."""
    port = _free_port()
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_gloo_gradsync_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), "2", str(port), str(tmp_path / f"r{r}.pt")])
             for r in range(2)]
    for p in procs:
        assert p.wait(timeout=180) == 0
    res = [torch.load(tmp_path / f"r{r}.pt") for r in range(2)]
    want_a = torch.arange(1000, dtype=torch.float32) * 1.0 + torch.arange(1000, dtype=torch.float32) * 2.0
    for r in res:
        assert torch.equal(r["a"], want_a)
        assert torch.equal(r["b"], torch.full((17,), 3.0))
        assert r["scale"] == 0.5 and r["world"] == 2
    # broadcast of the initial parameters / optimizer moments: both ranks hold rank 0's values (seed 100)
    torch.manual_seed(100)
    w0, m0 = torch.randn(333), torch.randn(333)
    for r in res:
        assert torch.equal(r["w"], w0) and torch.equal(r["m1"], m0)
