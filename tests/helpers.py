"""Shared test helpers: the fixture architectures (identical to tests/golden/make_golden.py), model
builders on top of kalle_audio_b200, and golden-file access."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REPORT = []      # (name, value) pairs printed by conftest.pytest_terminal_summary


def report(name, value):
    REPORT.append((name, f"{value:.3e}" if isinstance(value, float) else str(value)))


def ae_config(channels, c_mults, strides, enc_latent, dec_latent, io_channels, sample_rate):
    ratio = int(np.prod(strides))
    return {
        "model_type": "autoencoder",
        "sample_rate": sample_rate,
        "model": {
            "encoder": {"type": "oobleck", "config": {"in_channels": io_channels, "channels": channels,
                                                      "c_mults": list(c_mults), "strides": list(strides),
                                                      "latent_dim": enc_latent, "use_snake": True}},
            "decoder": {"type": "oobleck", "config": {"out_channels": io_channels, "channels": channels,
                                                      "c_mults": list(c_mults), "strides": list(strides),
                                                      "latent_dim": dec_latent, "use_snake": True,
                                                      "final_tanh": False}},
            "bottleneck": {"type": "vae"},
            "latent_dim": dec_latent,
            "downsampling_ratio": ratio,
            "io_channels": io_channels,
        },
    }


CONFIGS = {
    "tiny": ae_config(8, [1, 2, 4], [2, 4, 5], 8, 4, 2, 16000),
    "tiny_sym": ae_config(8, [1, 2, 4], [2, 4, 5], 4, 4, 2, 16000),
    "mid": ae_config(64, [1, 2, 4], [2, 4, 5], 128, 64, 2, 16000),
    "sao": ae_config(128, [1, 2, 4, 8, 16], [2, 4, 4, 8, 8], 128, 64, 2, 44100),
    "o12_d512": ae_config(128, [1, 2, 4, 8, 16], [2, 4, 4, 5, 8], 1024, 512, 1, 16000),
    "o12_d256": ae_config(128, [1, 2, 4, 8, 16], [2, 4, 4, 5, 8], 512, 256, 1, 16000),
    "o12_d1024": ae_config(128, [1, 2, 4, 8, 16], [2, 4, 4, 5, 8], 2048, 1024, 1, 16000),
}


def strides_of(name):
    return CONFIGS[name]["model"]["decoder"]["config"]["strides"]


def randomize_snake(model, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith(".alpha") or name.endswith(".beta"):
                p.copy_(0.3 * torch.randn(p.shape, generator=g))


def build(name, seed=0, snake_seed=None):
    """Same seed + same construction order as the reference => same random-init state_dict."""
    import kalle_audio_b200 as k
    torch.manual_seed(seed)
    m = k.create_autoencoder_from_config(CONFIGS[name]).eval()
    if snake_seed is not None:
        randomize_snake(m, snake_seed)
    return m


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def check_checksums(state_dict, g):
    keys = [str(k) for k in g["cs_keys"]]
    vals = g["cs_vals"]
    assert list(state_dict.keys()) == keys, "state_dict keys / order differ from the reference"
    for k, v in zip(keys, vals):
        got = float(state_dict[k].double().abs().sum())
        assert abs(got - v) <= 1e-9 * max(1.0, abs(v)), f"parameter {k} differs from the reference init"


def split_sd(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def t(x):
    return torch.from_numpy(np.asarray(x))
