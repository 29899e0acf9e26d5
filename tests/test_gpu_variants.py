"""Kernel-variant switches against each other (GPU): the specialised 16-warp epilogues of conv_umma2_kernel<1> must be
BIT-identical to the generic epilogue they replace; the tensor-core encoder head and the pre-activated decoder tail
change the arithmetic in the last bits only (hi/lo split products, one rounding fewer) and must stay far inside the
bf16-mode budget.  Each variant runs in its own process: the switches are read once."""
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def run_variant(tmp_path, name, **env):
    out = str(tmp_path / f"{name}.npz")
    e = dict(os.environ)
    e.update({k: str(v) for k, v in env.items()})
    r = subprocess.run([sys.executable, os.path.join(HERE, "_variant_worker.py"), out], env=e, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return np.load(out)


def test_fast_epilogue_is_bit_identical_and_edge_variants_agree(tmp_path):
    # same head / tail flow in both legs, only the epilogue instantiation differs
    old = run_variant(tmp_path, "generic", KVAE_FAST_EPI=0, KVAE_WAVE_IN_CC=1, KVAE_TAIL_RAW=1)
    new = run_variant(tmp_path, "fast", KVAE_FAST_EPI=1, KVAE_WAVE_IN_CC=1, KVAE_TAIL_RAW=1)
    for k in ("y", "e", "er"):
        assert old[k].shape == new[k].shape
        assert np.array_equal(old[k], new[k]), f"{k}: conv_umma2_kernel<1> differs from the generic epilogue"
    # shipped flow: tensor-core head (im2col, hi/lo split) and pre-activated tail
    ship = run_variant(tmp_path, "shipped")
    dy = float(np.abs(ship["y"] - new["y"]).max())
    de = float(np.abs(ship["e"] - new["e"]).max())
    der = float(np.abs(ship["er"] - new["er"]).max())
    H.report("pre-activated tail vs SnakeBeta in the tail (waveform abs max %.3f)" % float(np.abs(new["y"]).max()), dy)
    H.report("tensor-core encoder head vs CUDA-core head (latent abs max %.3f), whole / ragged" % float(np.abs(new["e"]).max()),
             max(de, der))
    assert dy <= 3e-4 * max(1.0, float(np.abs(new["y"]).max()) / 0.125)
    assert max(de, der) <= 1e-3 * max(1.0, float(np.abs(new["e"]).max()) / 0.125)
