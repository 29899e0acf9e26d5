"""Drop-in for the Oobleck part of the reference's ``stable_audio_tools/models/autoencoders.py``.

Module tree, constructor signatures and ``state_dict`` keys mirror the reference (ResidualUnit :39-62,
EncoderBlock :64-81, DecoderBlock :83-114, OobleckEncoder :116-147, OobleckDecoder :150-191,
AudioAutoencoder :230-560, factories :611-731).  The compute is different: ``OobleckEncoder.forward`` and
``OobleckDecoder.forward`` hand the entire stack to one libkvae plan (tensor-core implicit-GEMM convs with
SnakeBeta / residual / bias fused, see csrc/), instead of running ~75 eager kernels per direction.

Precision (``set_precision`` or automatic):
  * "fp32": fp32-accurate end to end (<= 1e-5 on the waveform vs reference fp32) -- default for fp32 modules outside
            autocast.  Tensor-core convs run as a bf16x3 operand split (hi*hi + lo*hi + hi*lo) with fp32 accumulation,
            fp32 residual stream and fp32 SnakeBeta; the io-channel convs use fp32 CUDA-core FMAs.
  * "bf16": bf16 tensor-core operands, fp32 accumulation; the residual stream is fp32 in registers / TMEM and fp16
            in HBM -- default for bf16/fp16 modules and under ``torch.autocast`` (<= 1e-3 vs reference fp32)
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Literal, Optional

import torch
from torch import nn

from . import _lib
from ._plan import PlanCache, PlanFunction
from .bottleneck import Bottleneck, create_bottleneck_from_config
from .layers import NearestUpsampleConv, SnakeBeta, WNConv1d, WNConvTranspose1d


def get_activation(activation: Literal["elu", "snake", "none"], antialias=False, channels=None) -> nn.Module:
    if antialias:
        raise NotImplementedError("antialias_activation=True (alias_free_torch Activation1d) is not built; no "
                                  "configuration of the reference uses it")
    if activation == "snake":
        return SnakeBeta(channels)
    if activation == "none":
        return nn.Identity()
    if activation == "elu":
        raise NotImplementedError("use_snake=False (ELU) is not built: every Oobleck config the reference uses sets "
                                  "use_snake=True, and this package has no eager fallback")
    raise ValueError(f"Unknown activation {activation}")


def _act(use_snake: bool, antialias: bool, channels: int) -> nn.Module:
    return get_activation("snake" if use_snake else "elu", antialias=antialias, channels=channels)


class _Sequential(nn.Module):
    """Blocks keep the reference's ``self.layers = nn.Sequential(...)`` so keys are ``layers.N...``;
    stand-alone forward chains the leaf kernels (the fused path lives in OobleckEncoder/Decoder)."""

    def forward(self, x):
        return self.layers(x)


class ResidualUnit(_Sequential):
    def __init__(self, in_channels, out_channels, dilation, use_snake=False, antialias_activation=False):
        super().__init__()
        self.dilation = dilation
        padding = (dilation * (7 - 1)) // 2
        self.layers = nn.Sequential(
            _act(use_snake, antialias_activation, out_channels),
            WNConv1d(in_channels=in_channels, out_channels=out_channels, kernel_size=7, dilation=dilation,
                     padding=padding),
            _act(use_snake, antialias_activation, out_channels),
            WNConv1d(in_channels=out_channels, out_channels=out_channels, kernel_size=1),
        )

    def forward(self, x):
        return self.layers(x) + x


class EncoderBlock(_Sequential):
    def __init__(self, in_channels, out_channels, stride, use_snake=False, antialias_activation=False):
        super().__init__()
        self.layers = nn.Sequential(
            ResidualUnit(in_channels=in_channels, out_channels=in_channels, dilation=1, use_snake=use_snake),
            ResidualUnit(in_channels=in_channels, out_channels=in_channels, dilation=3, use_snake=use_snake),
            ResidualUnit(in_channels=in_channels, out_channels=in_channels, dilation=9, use_snake=use_snake),
            _act(use_snake, antialias_activation, in_channels),
            WNConv1d(in_channels=in_channels, out_channels=out_channels, kernel_size=2 * stride, stride=stride,
                     padding=math.ceil(stride / 2)),
        )


class DecoderBlock(_Sequential):
    def __init__(self, in_channels, out_channels, stride, use_snake=False, antialias_activation=False,
                 use_nearest_upsample=False):
        super().__init__()
        if use_nearest_upsample:
            upsample_layer = NearestUpsampleConv(in_channels, out_channels, stride)
        else:
            upsample_layer = WNConvTranspose1d(in_channels=in_channels, out_channels=out_channels,
                                               kernel_size=2 * stride + stride % 2, stride=stride,
                                               padding=math.ceil(stride / 2))
        self.layers = nn.Sequential(
            _act(use_snake, antialias_activation, in_channels),
            upsample_layer,
            ResidualUnit(in_channels=out_channels, out_channels=out_channels, dilation=1, use_snake=use_snake),
            ResidualUnit(in_channels=out_channels, out_channels=out_channels, dilation=3, use_snake=use_snake),
            ResidualUnit(in_channels=out_channels, out_channels=out_channels, dilation=9, use_snake=use_snake),
        )


def _make_arch(io_channels, channels, latent_dim, c_mults, strides, final_tanh, use_nearest_upsample=False) -> _lib.KvaeArch:
    if len(c_mults) != len(strides):
        raise ValueError("c_mults and strides must have the same length")
    if len(strides) > _lib.KVAE_MAX_STAGES:
        raise ValueError(f"at most {_lib.KVAE_MAX_STAGES} stages")
    a = _lib.KvaeArch()
    a.io_channels, a.channels, a.latent_dim, a.n_stages = io_channels, channels, latent_dim, len(strides)
    for i, (c, s) in enumerate(zip(c_mults, strides)):
        a.c_mults[i], a.strides[i] = int(c), int(s)
    a.final_tanh = int(bool(final_tanh))
    a.use_nearest_upsample = int(bool(use_nearest_upsample))
    return a


class _OobleckBase(nn.Module):
    _direction = _lib.KVAE_DECODER

    def _setup(self, arch, io_channels, ratio):
        self._arch = arch
        self._ratio = ratio
        self._out_channels_for_plan = io_channels
        self._precision: Optional[str] = None
        self._plans = PlanCache()

    def set_precision(self, precision: Optional[str]):
        """None (automatic, see module docstring), "fp32" or "bf16"."""
        if precision not in (None, "fp32", "bf16"):
            raise ValueError("precision must be None, 'fp32' or 'bf16'")
        self._precision = precision
        return self

    def _resolve_precision(self) -> int:
        if self._precision == "fp32":
            return _lib.KVAE_PREC_F32
        if self._precision == "bf16":
            return _lib.KVAE_PREC_BF16
        pdtype = next(self.parameters()).dtype
        if pdtype in (torch.bfloat16, torch.float16) or torch.is_autocast_enabled():
            return _lib.KVAE_PREC_BF16
        return _lib.KVAE_PREC_F32

    def _out_dtype(self, x: torch.Tensor) -> torch.dtype:
        if torch.is_autocast_enabled():
            return torch.get_autocast_dtype("cuda")
        return next(self.parameters()).dtype

    def enable_cuda_graphs(self, enable: bool = True):
        """Replay each (batch, length) shape as one CUDA graph instead of ~38 launches (helps small batches)."""
        self._use_graphs = bool(enable)
        for r in self._plans.runners.values():
            r.use_graphs = self._use_graphs
        return self

    def runner(self, device):
        r = self._plans.get(self, self._direction, self._arch, self._resolve_precision(), device)
        r.use_graphs = getattr(self, "_use_graphs", False)
        return r

    def forward(self, x, valid_len=None):
        """``valid_len`` (optional, B ints): ragged batch.  Clip b holds valid_len[b] valid input positions (latent
        frames / audio samples) and is zero-padded to the common length (for an encoder: a multiple of the ratio);
        its output up to the matching length equals the reference's stand-alone call on that clip -- including an
        audio length that is not a multiple of the ratio, where the reference's strided convs floor -- and is zero
        beyond.  Inference only."""
        _lib.require_cuda(x, type(self).__name__ + ".forward")
        r = self.runner(x.device)
        if valid_len is not None:
            return r.run(x, self._out_channels_for_plan, self._ratio, self._out_dtype(x), valid_len=valid_len)
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            # training: saved activations + hand-written backward (kvae_forward_train / kvae_backward)
            if getattr(self._arch, "final_tanh", 0):
                raise NotImplementedError("final_tanh=True is inference-only in kalle_audio_b200 (no reference config "
                                          "trains with it)")
            if getattr(self._arch, "use_nearest_upsample", 0):
                raise NotImplementedError("use_nearest_upsample=True is inference-only in kalle_audio_b200 (no reference "
                                          "config uses it)")
            return PlanFunction.apply(r, x, self._out_channels_for_plan, self._ratio, self._out_dtype(x),
                                      *r.param_list())
        return r.run(x, self._out_channels_for_plan, self._ratio, self._out_dtype(x))

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if hasattr(self, "_plans"):
            self._plans.clear()     # .to(device/dtype) re-creates parameters; plans are rebuilt lazily
        return out


class OobleckEncoder(_OobleckBase):
    _direction = _lib.KVAE_ENCODER

    def __init__(self, in_channels=2, channels=128, latent_dim=32, c_mults=[1, 2, 4, 8], strides=[2, 4, 8, 8],
                 use_snake=False, antialias_activation=False):
        super().__init__()
        user_c_mults = list(c_mults)
        c_mults = [1] + user_c_mults
        self.depth = len(c_mults)
        layers: List[nn.Module] = [WNConv1d(in_channels=in_channels, out_channels=c_mults[0] * channels, kernel_size=7,
                                            padding=3)]
        for i in range(self.depth - 1):
            layers.append(EncoderBlock(in_channels=c_mults[i] * channels, out_channels=c_mults[i + 1] * channels,
                                       stride=strides[i], use_snake=use_snake))
        layers += [_act(use_snake, antialias_activation, c_mults[-1] * channels),
                   WNConv1d(in_channels=c_mults[-1] * channels, out_channels=latent_dim, kernel_size=3, padding=1)]
        self.layers = nn.Sequential(*layers)
        self._setup(_make_arch(in_channels, channels, latent_dim, user_c_mults, strides, False), latent_dim,
                    int(math.prod(strides)))


class OobleckDecoder(_OobleckBase):
    _direction = _lib.KVAE_DECODER

    def __init__(self, out_channels=2, channels=128, latent_dim=32, c_mults=[1, 2, 4, 8], strides=[2, 4, 8, 8],
                 use_snake=False, antialias_activation=False, use_nearest_upsample=False, final_tanh=True):
        super().__init__()
        user_c_mults = list(c_mults)
        c_mults = [1] + user_c_mults
        self.depth = len(c_mults)
        layers: List[nn.Module] = [WNConv1d(in_channels=latent_dim, out_channels=c_mults[-1] * channels, kernel_size=7,
                                            padding=3)]
        for i in range(self.depth - 1, 0, -1):
            layers.append(DecoderBlock(in_channels=c_mults[i] * channels, out_channels=c_mults[i - 1] * channels,
                                       stride=strides[i - 1], use_snake=use_snake,
                                       antialias_activation=antialias_activation,
                                       use_nearest_upsample=use_nearest_upsample))
        layers += [_act(use_snake, antialias_activation, c_mults[0] * channels),
                   WNConv1d(in_channels=c_mults[0] * channels, out_channels=out_channels, kernel_size=7, padding=3,
                            bias=False),
                   nn.Tanh() if final_tanh else nn.Identity()]
        self.layers = nn.Sequential(*layers)
        self._setup(_make_arch(out_channels, channels, latent_dim, user_c_mults, strides, final_tanh,
                               use_nearest_upsample), out_channels, int(math.prod(strides)))


# ---------------------------------------------------------------------------------------------------
def _chunk_starts(total: int, chunk: int, hop: int) -> List[int]:
    """Window starts of the reference's chunking loops (autoencoders.py:459-466 / 521-528): hop-spaced
    windows, plus one right-aligned window when the last hop does not end exactly at ``total``."""
    if total < chunk:
        # the reference hits an unbound loop variable here; keep the exception type
        raise UnboundLocalError(f"chunked processing needs at least chunk_size={chunk} frames, got {total}")
    starts = list(range(0, total - chunk + 1, hop))
    if starts[-1] + chunk != total:
        starts.append(total - chunk)
    return starts


class AudioAutoencoder(nn.Module):
    def __init__(self, encoder, decoder, latent_dim, downsampling_ratio, sample_rate, io_channels=2,
                 bottleneck: Bottleneck = None, pretransform=None, in_channels=None, out_channels=None,
                 soft_clip=False):
        super().__init__()
        self.downsampling_ratio = downsampling_ratio
        self.sample_rate = sample_rate
        self.latent_dim = latent_dim
        self.io_channels = io_channels
        self.in_channels = io_channels if in_channels is None else in_channels
        self.out_channels = io_channels if out_channels is None else out_channels
        self.min_length = self.downsampling_ratio
        self.bottleneck = bottleneck
        self.encoder = encoder
        self.decoder = decoder
        self.pretransform = pretransform
        self.soft_clip = soft_clip
        self.is_discrete = self.bottleneck is not None and self.bottleneck.is_discrete

    def set_precision(self, precision: Optional[str]):
        for m in (self.encoder, self.decoder):
            if isinstance(m, _OobleckBase):
                m.set_precision(precision)
        return self

    # -- helpers ----------------------------------------------------------------------------------
    @staticmethod
    def _per_item(fn, x):
        return torch.cat([fn(x[i:i + 1]) for i in range(x.shape[0])], dim=0)

    def _pretransform(self, fn, x, iterate_batch):
        if self.pretransform.enable_grad:
            return self._per_item(fn, x) if iterate_batch else fn(x)
        with torch.no_grad():
            return self._per_item(fn, x) if iterate_batch else fn(x)

    # -- encode / decode (autoencoders.py:275-361) ---------------------------------------------------
    def encode(self, audio, return_info=False, skip_pretransform=False, iterate_batch=False, **kwargs):
        info: Dict[str, Any] = {}
        if self.pretransform is not None and not skip_pretransform:
            audio = self._pretransform(self.pretransform.encode, audio, iterate_batch)
        if self.encoder is not None:
            # iterate_batch is a memory workaround in the reference; batch items are independent, so the
            # fused plan processes them together with identical results
            latents = self.encoder(audio)
        else:
            latents = audio
        if self.bottleneck is not None:
            latents, bottleneck_info = self.bottleneck.encode(latents, return_info=True, **kwargs)
            info.update(bottleneck_info)
        if return_info:
            return latents, info
        return latents

    @torch.no_grad()
    def encode_and_sample(self, audio, noise=None, dist_type="fix"):
        """encode -> chunk(2, dim=1) -> sample(mean, 'fix') in ONE plan call: the sampler (model_sigmaVAE.py:187-213,
        applied to the mean half of the encoder output as twj_dataset.py:251 splits it) runs in the epilogue of the
        encoder's last conv.  Returns (mean_scale [B, 2D, T], z [B, D, T]); z is bit-identical to
        ``sample(mean_scale.chunk(2, 1)[0], 'fix', noise)``.  ``noise`` defaults to ``torch.randn_like(mean)``."""
        from .sampling import _STD, sample
        if self.pretransform is not None or self.bottleneck is None or dist_type != "fix":
            ms = self.encode(audio)
            mean = ms.chunk(2, dim=1)[0].contiguous()
            return ms, sample(mean, dist_type, noise=noise)
        enc = self.encoder
        D = enc._out_channels_for_plan // 2
        if noise is None:
            noise = torch.randn((audio.shape[0], D, audio.shape[2] // self.downsampling_ratio), device=audio.device,
                                dtype=enc._out_dtype(audio))
        got = enc.runner(audio.device).run_encode_sample(audio, noise, enc._out_channels_for_plan, enc._ratio,
                                                         enc._out_dtype(audio), D, _STD)
        if got is None:       # no tensor-core output conv in this architecture: two launches
            ms = self.encode(audio)
            return ms, sample(ms.chunk(2, dim=1)[0].contiguous(), "fix", noise=noise)
        return got

    def decode(self, latents, iterate_batch=False, **kwargs):
        if self.bottleneck is not None:
            latents = self.bottleneck.decode(latents)
        decoded = self.decoder(latents, **kwargs)
        if self.pretransform is not None:
            decoded = self._pretransform(self.pretransform.decode, decoded, iterate_batch)
        if self.soft_clip:
            decoded = torch.tanh(decoded)
        return decoded

    @torch.no_grad()
    def decode_pcm16(self, latents):
        """decode + the peak-normalised int16 conversion every caller of the reference repeats
        (``output.to(float32).div(max|output|).clamp(-1, 1).mul(32767).to(int16)``, infer_0828_sigma.py:298): the peak is
        found in the tail conv's epilogue while the waveform is written.  Returns (waveform, int16 pcm)."""
        from .utils import to_pcm16
        dec = self.decoder
        plain = self.bottleneck is None or type(self.bottleneck).__name__ == "VAEBottleneck"
        if plain and self.pretransform is None and not self.soft_clip and isinstance(dec, _OobleckBase):
            got = dec.runner(latents.device).run_decode_pcm16(latents, dec._out_channels_for_plan, dec._ratio,
                                                              dec._out_dtype(latents))
            if got is not None:
                return got
        wav = self.decode(latents)
        return wav, to_pcm16(wav)

    def decode_tokens(self, tokens, **kwargs):
        raise NotImplementedError("discrete bottlenecks are outside the sigmaVAE hot path")

    # -- audio preprocessing (autoencoders.py:376-427) -------------------------------------------------
    def preprocess_audio_for_encoder(self, audio, in_sr):
        return self.preprocess_audio_list_for_encoder([audio], [in_sr])

    def preprocess_audio_list_for_encoder(self, audio_list, in_sr_list):
        from .utils import prepare_audio
        n = len(audio_list)
        if isinstance(in_sr_list, int):
            in_sr_list = [in_sr_list] * n
        assert len(in_sr_list) == n, "list of sample rates must be the same length of audio_list"
        prepared, max_length = [], 0
        for audio, in_sr in zip(audio_list, in_sr_list):
            if audio.dim() == 3 and audio.shape[0] == 1:
                audio = audio.squeeze(0)
            elif audio.dim() == 1:
                audio = audio.unsqueeze(0)
            assert audio.dim() == 2, "Audio should be shape (Channels x Length) with no batch dimension"
            if in_sr != self.sample_rate:
                from torchaudio import transforms as T
                audio = T.Resample(in_sr, self.sample_rate).to(audio.device)(audio)
            prepared.append(audio)
            max_length = max(max_length, audio.shape[-1])
        padded = max_length + (self.min_length - (max_length % self.min_length)) % self.min_length
        out = [prepare_audio(a, in_sr=sr, target_sr=sr, target_length=padded, target_channels=self.in_channels,
                             device=a.device).squeeze(0) for a, sr in zip(prepared, in_sr_list)]
        return torch.stack(out)

    # -- chunked paths (autoencoders.py:429-560) --------------------------------------------------------
    # Windows are independent clips for the convs, so a GROUP of windows goes through the plan as one batch.  The
    # reference runs one window at a time precisely to bound memory; here the group size is bounded instead
    # (``max_windows_per_call`` windows of the whole batch per plan call, 8 by default; 1 = the reference's loop) and
    # every group is pasted straight into the output, so the peak stays at one group's workspace however long the
    # clip is.  Like the reference (:475, :537) the per-window encode / decode takes no extra keyword arguments.
    max_windows_per_call = 8

    def _chunked(self, fn, x, size, hop, out_channels, out_len, to_out, ol):
        """Shared window loop of encode_audio / decode_audio (autoencoders.py:456-497, 518-560): ``size`` / ``hop`` in
        input units, ``to_out`` converts an input position to an output position, ``ol`` = trimmed overlap (output units)."""
        total, bsz = x.shape[2], x.shape[0]
        starts = _chunk_starts(total, size, hop)
        n = len(starts)
        y_final = torch.zeros((bsz, out_channels, out_len), device=x.device)
        group = max(1, int(self.max_windows_per_call))
        for g0 in range(0, n, group):
            idx = range(g0, min(n, g0 + group))
            y_all = fn(torch.cat([x[:, :, starts[i]:starts[i] + size] for i in idx], dim=0))
            if y_all.shape[1] != out_channels:
                raise RuntimeError(f"The expanded size of the tensor ({out_channels}) must match the existing size "
                                   f"({y_all.shape[1]}) at non-singleton dimension 1 (reference behaviour: chunked encode "
                                   "needs an encoder that emits latent_dim channels)")
            for j, i in enumerate(idx):
                y_chunk = y_all[j * bsz:(j + 1) * bsz]
                if i == n - 1:
                    t_end = out_len
                    t_start = t_end - y_chunk.shape[2]
                else:
                    t_start = to_out(starts[i])
                    t_end = t_start + to_out(size)
                c0, c1 = 0, y_chunk.shape[2]
                if i > 0:
                    t_start += ol
                    c0 += ol
                if i < n - 1:
                    t_end -= ol
                    c1 -= ol
                y_final[:, :, t_start:t_end] = y_chunk[:, :, c0:c1]
        return y_final

    def encode_audio(self, audio, chunked=False, overlap=32, chunk_size=128, **kwargs):
        if not chunked:
            return self.encode(audio, **kwargs)
        spl = self.downsampling_ratio
        cs, ov = chunk_size * spl, overlap * spl
        return self._chunked(self.encode, audio, cs, cs - ov, self.latent_dim, audio.shape[2] // spl,
                             lambda pos: pos // spl, ov // spl // 2)

    def decode_audio(self, latents, chunked=False, overlap=32, chunk_size=128, **kwargs):
        if not chunked:
            return self.decode(latents, **kwargs)
        spl = self.downsampling_ratio
        return self._chunked(self.decode, latents, chunk_size, chunk_size - overlap, self.out_channels,
                             latents.shape[2] * spl, lambda pos: pos * spl, (overlap // 2) * spl)


# ---------------------------------------------------------------------------------------------------
def create_encoder_from_config(encoder_config: Dict[str, Any]):
    encoder_type = encoder_config.get("type", None)
    assert encoder_type is not None, "Encoder type must be specified"
    if encoder_type != "oobleck":
        raise NotImplementedError(f"encoder type {encoder_type!r}: only 'oobleck' is on the sigmaVAE hot path")
    encoder = OobleckEncoder(**encoder_config["config"])
    if not encoder_config.get("requires_grad", True):
        for p in encoder.parameters():
            p.requires_grad = False
    return encoder


def create_decoder_from_config(decoder_config: Dict[str, Any]):
    decoder_type = decoder_config.get("type", None)
    assert decoder_type is not None, "Decoder type must be specified"
    if decoder_type != "oobleck":
        raise NotImplementedError(f"decoder type {decoder_type!r}: only 'oobleck' is on the sigmaVAE hot path")
    decoder = OobleckDecoder(**decoder_config["config"])
    if not decoder_config.get("requires_grad", True):
        for p in decoder.parameters():
            p.requires_grad = False
    return decoder


def create_autoencoder_from_config(config: Dict[str, Any]):
    ae_config = config["model"]
    encoder = create_encoder_from_config(ae_config["encoder"])
    decoder = create_decoder_from_config(ae_config["decoder"])
    bottleneck = ae_config.get("bottleneck", None)
    latent_dim = ae_config.get("latent_dim", None)
    assert latent_dim is not None, "latent_dim must be specified in model config"
    downsampling_ratio = ae_config.get("downsampling_ratio", None)
    assert downsampling_ratio is not None, "downsampling_ratio must be specified in model config"
    io_channels = ae_config.get("io_channels", None)
    assert io_channels is not None, "io_channels must be specified in model config"
    sample_rate = config.get("sample_rate", None)
    assert sample_rate is not None, "sample_rate must be specified in model config"
    in_channels = ae_config.get("in_channels", None)
    out_channels = ae_config.get("out_channels", None)
    pretransform = ae_config.get("pretransform", None)
    if pretransform is not None:
        from .factory import create_pretransform_from_config
        pretransform = create_pretransform_from_config(pretransform, sample_rate)
    if bottleneck is not None:
        bottleneck = create_bottleneck_from_config(bottleneck)
    soft_clip = ae_config["decoder"].get("soft_clip", False)
    return AudioAutoencoder(encoder, decoder, io_channels=io_channels, latent_dim=latent_dim,
                            downsampling_ratio=downsampling_ratio, sample_rate=sample_rate, bottleneck=bottleneck,
                            pretransform=pretransform, in_channels=in_channels, out_channels=out_channels,
                            soft_clip=soft_clip)
