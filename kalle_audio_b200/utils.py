"""Checkpoint / weight-norm / audio-shape helpers (reference stable_audio_tools/models/utils.py:6-20,
stable_audio_tools/inference/utils.py prepare_audio, stable_audio_tools/data/utils.py:8-20 PadCrop)."""
from __future__ import annotations

import torch


def load_ckpt_state_dict(ckpt_path):
    if ckpt_path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(ckpt_path)
    return torch.load(ckpt_path, map_location="cpu")["state_dict"]


def remove_weight_norm_from_model(model):
    from .layers import _WNConvBase
    for module in model.modules():
        if isinstance(module, _WNConvBase) and module.has_weight_norm:
            module.remove_weight_norm()
    return model


def set_audio_channels(audio: torch.Tensor, target_channels: int) -> torch.Tensor:
    """[B, C, L]: mono <- mean over channels; stereo <- duplicate mono / keep first two."""
    if target_channels == 1:
        return audio.mean(1, keepdim=True)
    if target_channels == 2:
        if audio.shape[1] == 1:
            return audio.repeat(1, 2, 1)
        if audio.shape[1] > 2:
            return audio[:, :2, :]
    return audio


def prepare_audio(audio, in_sr, target_sr, target_length, target_channels, device):
    """Resample, zero-pad / crop from the start to ``target_length``, add the batch dim, fix the channel count."""
    audio = audio.to(device)
    if in_sr != target_sr:
        from torchaudio import transforms as T
        audio = T.Resample(in_sr, target_sr).to(device)(audio)
    n, s = audio.shape[-2], audio.shape[-1]
    out = audio.new_zeros([n, target_length])
    out[:, :min(s, target_length)] = audio[:, :target_length]
    audio = out
    if audio.dim() == 1:
        audio = audio.unsqueeze(0).unsqueeze(0)
    elif audio.dim() == 2:
        audio = audio.unsqueeze(0)
    return set_audio_channels(audio, target_channels)
