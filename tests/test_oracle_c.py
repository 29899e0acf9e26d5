"""Cross-checks the two oracle restatements against each other and against the reference's golden
outputs: oracle/oobleck_ref.c (straight C loops, double accumulation) vs oracle/oobleck_oracle.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

import helpers as H
from oracle import oobleck_oracle as O

ODIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")


@pytest.fixture(scope="module")
def ref():
    so = os.path.join(ODIR, "liboracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", ODIR])
    lib = C.CDLL(so)
    lib.ref_vae_sample.restype = C.c_double
    return lib


def fp(a):
    return a.ctypes.data_as(C.c_void_p)


def test_c_conv_and_snake_match_torch_oracle(ref):
    rng = np.random.default_rng(0)
    B, Cin, Cout, T, K, d = 2, 6, 5, 70, 7, 3
    x = rng.standard_normal((B, Cin, T), dtype=np.float32)
    w = rng.standard_normal((Cout, Cin, K), dtype=np.float32)
    b = rng.standard_normal(Cout, dtype=np.float32)
    y = np.zeros((B, Cout, T), np.float32)
    ref.ref_conv1d(fp(x), fp(w), fp(b), fp(y), B, Cin, Cout, C.c_long(T), K, 1, d, 3 * d)
    yt = torch.nn.functional.conv1d(H.t(x), H.t(w), H.t(b), padding=3 * d, dilation=d).numpy()
    assert np.abs(y - yt).max() <= 1e-5
    # strided (encoder) form
    ys = np.zeros((B, Cout, 14), np.float32)
    w10 = rng.standard_normal((Cout, Cin, 10), dtype=np.float32)
    ref.ref_conv1d(fp(x), fp(w10), fp(b), fp(ys), B, Cin, Cout, C.c_long(T), 10, 5, 1, 3)
    yst = torch.nn.functional.conv1d(H.t(x), H.t(w10), H.t(b), stride=5, padding=3).numpy()
    assert ys.shape == yst.shape and np.abs(ys - yst).max() <= 1e-5
    # transposed, odd stride (k = 2s+1, pad 3)
    wt = rng.standard_normal((Cin, Cout, 11), dtype=np.float32)
    yT = np.zeros((B, Cout, T * 5), np.float32)
    ref.ref_conv_transpose1d(fp(x), fp(wt), fp(b), fp(yT), B, Cin, Cout, C.c_long(T), 11, 5, 3)
    yTt = torch.nn.functional.conv_transpose1d(H.t(x), H.t(wt), H.t(b), stride=5, padding=3).numpy()
    assert yT.shape == yTt.shape and np.abs(yT - yTt).max() <= 1e-5
    al, be = rng.standard_normal(Cin, dtype=np.float32) * 0.3, rng.standard_normal(Cin, dtype=np.float32) * 0.3
    s = np.zeros_like(x)
    ref.ref_snake_beta(fp(x), fp(s), fp(al), fp(be), 1, B, Cin, C.c_long(T))
    assert np.abs(s - O.snake_beta(H.t(x), H.t(al), H.t(be)).numpy()).max() <= 2e-6
    v = rng.standard_normal((Cout, Cin, K), dtype=np.float32)
    g = (rng.random((Cout, 1, 1), dtype=np.float32) + 0.5)
    wn = np.zeros_like(v)
    ref.ref_weight_norm(fp(v), fp(g), fp(wn), Cout, Cin * K)
    assert np.abs(wn - O.weight_norm_fold(H.t(v), H.t(g)).numpy()).max() <= 1e-6


def test_c_residual_unit_matches_reference_layer(ref):
    """First ResidualUnit of the tiny decoder's first block, against the reference's recorded activations."""
    g = H.golden("tiny_ae")
    sd = {k[len("sd.decoder."):]: g[k] for k in g.files if k.startswith("sd.decoder.")}
    st = H.strides_of("tiny")
    # input of RU1 = output of the block's transposed conv; recompute it with the torch oracle
    sdt = {k: H.t(v) for k, v in sd.items()}
    h = O._wn_conv1d(sdt, "layers.0", H.t(g["z"]), padding=3)
    h = O._snake(sdt, "layers.1.layers.0", h)
    h = O._wn_conv_transpose1d(sdt, "layers.1.layers.1", h, stride=st[-1], padding=3).numpy().copy()
    want = O.residual_unit(sdt, "layers.1.layers.2", H.t(h), 1).numpy()
    p = "layers.1.layers.2.layers."
    Cc = h.shape[1]
    w7, w1 = np.zeros_like(sd[p + "1.weight_v"]), np.zeros_like(sd[p + "3.weight_v"])
    ref.ref_weight_norm(fp(sd[p + "1.weight_v"]), fp(sd[p + "1.weight_g"]), fp(w7), Cc, Cc * 7)
    ref.ref_weight_norm(fp(sd[p + "3.weight_v"]), fp(sd[p + "3.weight_g"]), fp(w1), Cc, Cc)
    y = np.zeros_like(h)
    ref.ref_residual_unit(fp(h), fp(y), fp(sd[p + "0.alpha"]), fp(sd[p + "0.beta"]), fp(w7), fp(sd[p + "1.bias"]),
                          fp(sd[p + "2.alpha"]), fp(sd[p + "2.beta"]), fp(w1), fp(sd[p + "3.bias"]),
                          h.shape[0], Cc, C.c_long(h.shape[2]), 1)
    assert np.abs(y - want).max() <= 5e-6


def test_c_sampling_bit_exact_vs_reference(ref):
    g = H.golden("sampling")
    mean, scale, noise = (np.ascontiguousarray(g[n]) for n in ("mean", "scale", "noise"))
    out = np.zeros_like(mean)
    kl = ref.ref_vae_sample(fp(mean), fp(scale), fp(noise), fp(out), *mean.shape[:2], C.c_long(mean.shape[2]))
    assert np.array_equal(out, g["vae_latents"])
    assert abs(kl - float(g["vae_kl"])) <= 1e-4 * abs(float(g["vae_kl"]))
    ref.ref_sigma_sample(fp(mean), fp(noise), fp(out), C.c_size_t(mean.size), C.c_float(0.5))
    assert np.array_equal(out, g["fix"])
