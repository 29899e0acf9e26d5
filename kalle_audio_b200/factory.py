"""json -> module factories for the autoencoder path (reference stable_audio_tools/models/factory.py:3-153)."""
from __future__ import annotations

import json


def create_model_from_config(model_config):
    model_type = model_config.get("model_type", None)
    assert model_type is not None, "model_type must be specified in model config"
    if model_type == "autoencoder":
        from .autoencoders import create_autoencoder_from_config
        return create_autoencoder_from_config(model_config)
    raise NotImplementedError(f"model_type {model_type!r}: only 'autoencoder' is on the sigmaVAE hot path "
                              "(diffusion / LM wrappers stay with the reference; use "
                              "create_pretransform_from_config for their .pretransform)")


def create_model_from_config_path(model_config_path):
    with open(model_config_path) as f:
        return create_model_from_config(json.load(f))


def create_pretransform_from_config(pretransform_config, sample_rate):
    pretransform_type = pretransform_config.get("type", None)
    assert pretransform_type is not None, "type must be specified in pretransform config"
    if pretransform_type != "autoencoder":
        raise NotImplementedError(f"pretransform type {pretransform_type!r} is outside the sigmaVAE hot path")
    from .autoencoders import create_autoencoder_from_config
    from .pretransforms import AutoencoderPretransform
    autoencoder_config = {"sample_rate": sample_rate, "model": pretransform_config["config"]}
    autoencoder = create_autoencoder_from_config(autoencoder_config)
    pretransform = AutoencoderPretransform(autoencoder, scale=pretransform_config.get("scale", 1.0),
                                           model_half=pretransform_config.get("model_half", False),
                                           iterate_batch=pretransform_config.get("iterate_batch", False),
                                           chunked=pretransform_config.get("chunked", False))
    enable_grad = pretransform_config.get("enable_grad", False)
    pretransform.enable_grad = enable_grad
    pretransform.eval().requires_grad_(pretransform.enable_grad)
    return pretransform


def create_bottleneck_from_config(bottleneck_config):
    from .bottleneck import create_bottleneck_from_config as f
    return f(bottleneck_config)
