"""Checkpoint / weight-norm / audio-shape helpers (reference stable_audio_tools/models/utils.py:6-20,
stable_audio_tools/inference/utils.py prepare_audio, stable_audio_tools/data/utils.py:8-20 PadCrop) and the
post-decode PCM tail the reference's callers repeat (infer_0828_sigma.py:298, train_offline.py:302,319)."""
from __future__ import annotations

import torch

from . import _lib


def to_pcm16(audio: torch.Tensor) -> torch.Tensor:
    """``audio.to(float32).div(max|audio|).clamp(-1, 1).mul(32767).to(int16)`` in two kernels on the device
    (global peak over the whole tensor, as in the reference); bit-identical to the torch expression."""
    _lib.require_cuda(audio, "to_pcm16")
    a = audio if audio.dtype in (torch.float32, torch.bfloat16) else audio.float()
    a = a.contiguous()
    out = torch.empty(a.shape, dtype=torch.int16, device=a.device)
    if a.numel():
        scratch = torch.empty(1, dtype=torch.int32, device=a.device)
        _lib.check(_lib.lib().kvae_pcm16(a.data_ptr(), _lib.dtype_code(a.dtype), out.data_ptr(), a.numel(),
                                         scratch.data_ptr(), _lib.stream_ptr(a.device)))
    return out


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
