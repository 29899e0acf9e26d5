"""Incremental latent -> waveform decoding with exact receptive-field context.

The reference streams by ``decode_audio(chunked=True, chunk_size=128, overlap=32)``
(stable_audio_tools/models/autoencoders.py:499-560; infer_stream-style callers): overlapping windows, the middle of
each pasted into the output -- 1.33x recompute, and only approximately equal to the unchunked decode (the overlap is
a guess at the receptive field).  ``StreamingDecoder`` computes the decoder's receptive field from its constructor
arguments instead (``decoder_context_frames``), so every window carries exactly the latent frames its emitted samples
depend on: the concatenated stream equals ``decoder(latents)`` of the whole sequence, the recompute is
(hop + left + right) / hop (1.21x at hop 96 for the SAO / 12.5 Hz strides), and a frame is emitted as soon as its
right context has arrived.  Windows of one shape replay as one CUDA graph (``enable_cuda_graphs``).

This is SURVEY.md section 8(f) item 1 built on the existing fused plan; it adds no kernels.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib


def decoder_context_frames(strides: Sequence[int]) -> Tuple[int, int]:
    """(left, right) latent frames an OobleckDecoder output block depends on beyond its own frames.

    Walks the stack backwards with interval arithmetic (autoencoders.py:150-191): final conv k7 pad 3; per
    DecoderBlock three ResidualUnits (k7, dilation 9 / 3 / 1) and a ConvTranspose1d(k = 2s + s % 2, stride s,
    pad ceil(s / 2)); first conv k7 pad 3.  ``strides`` in constructor (encoder) order."""
    ratio = int(math.prod(strides))
    f0, f1 = 1000, 1001                      # one frame far from the edges
    lo, hi = f0 * ratio, f1 * ratio - 1      # output sample range of that frame
    lo, hi = lo - 3, hi + 3                  # final conv
    for s in strides:                        # decoder blocks in reverse execution order = encoder stride order
        lo, hi = lo - 39, hi + 39            # ResidualUnits: 3 * (1 + 3 + 9)
        k, p = 2 * s + s % 2, math.ceil(s / 2)
        lo = -((-(lo + p - k + 1)) // s)     # ceil((lo + p - k + 1) / s)
        hi = (hi + p) // s
    lo, hi = lo - 3, hi + 3                  # first conv
    return f0 - lo, hi - (f1 - 1)


class StreamingDecoder:
    """Feed latent frames as they arrive; get waveform back as soon as it is final.

        sd = StreamingDecoder(autoencoder.decoder, hop=96)
        for z in latent_chunks:              # z: [B, D, n], any n >= 0
            wav = sd.push(z)                 # [B, C, m * ratio], m = frames that became final (possibly 0)
        tail = sd.flush()                    # the last frames (right edge zero-padded like the unchunked decode)

    ``torch.cat`` of everything returned equals ``decoder(torch.cat(latent_chunks, -1))``."""

    def __init__(self, decoder, hop: int = 96, use_cuda_graphs: bool = True):
        if not hasattr(decoder, "_arch") or decoder._direction != _lib.KVAE_DECODER:
            raise TypeError("StreamingDecoder wraps a kalle_audio_b200.OobleckDecoder")
        if hop < 1:
            raise ValueError("hop must be >= 1")
        self.decoder = decoder
        self.hop = int(hop)
        strides = [decoder._arch.strides[i] for i in range(decoder._arch.n_stages)]
        self.ratio = int(math.prod(strides))
        self.left, self.right = decoder_context_frames(strides)
        if use_cuda_graphs:
            decoder.enable_cuda_graphs(True)
        self.reset()

    def reset(self) -> None:
        self._buf: Optional[torch.Tensor] = None   # frames [emitted - left_available, received)
        self._emitted = 0                          # frames already returned
        self._start = 0                            # absolute index of _buf[..., 0]

    @property
    def recompute_factor(self) -> float:
        return (self.hop + self.left + self.right) / self.hop

    def _decode_window(self, first: int, last: int, end_of_stream: bool) -> torch.Tensor:
        """Decodes frames [first, last) with whatever context exists and returns exactly their samples."""
        buf, start = self._buf, self._start
        w0 = max(first - self.left, start)
        w1 = last if end_of_stream else last + self.right
        win = buf[:, :, w0 - start:w1 - start]
        with torch.no_grad():
            y = self.decoder(win)
        return y[:, :, (first - w0) * self.ratio:(last - w0) * self.ratio]

    def _trim(self) -> None:
        keep_from = max(self._emitted - self.left, self._start)
        if keep_from > self._start:
            self._buf = self._buf[:, :, keep_from - self._start:]
            self._start = keep_from

    def push(self, latents: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda(latents, "StreamingDecoder.push")
        if latents.dim() != 3:
            raise ValueError("expected [B, D, n]")
        self._buf = latents if self._buf is None else torch.cat([self._buf, latents], dim=2)
        received = self._start + self._buf.shape[2]
        out: List[torch.Tensor] = []
        # a frame is final once `right` frames after it have arrived; emit in hop-sized windows
        while received - self.right - self._emitted >= self.hop:
            out.append(self._decode_window(self._emitted, self._emitted + self.hop, False))
            self._emitted += self.hop
            self._trim()
        if out:
            return torch.cat(out, dim=2)
        B = latents.shape[0]
        return latents.new_zeros((B, self.decoder._out_channels_for_plan, 0))

    def flush(self) -> torch.Tensor:
        """Everything not yet emitted (end of stream).  Resets the state."""
        if self._buf is None:
            raise ValueError("nothing was pushed")
        received = self._start + self._buf.shape[2]
        out: List[torch.Tensor] = []
        while received - self._emitted > 0:
            n = min(self.hop, received - self._emitted)
            last = self._emitted + n
            end = last + self.right >= received       # no full right context left: this window ends the stream
            if end:
                last = received
            out.append(self._decode_window(self._emitted, last, end))
            self._emitted = last
            self._trim()
        res = torch.cat(out, dim=2) if out else self._buf.new_zeros((self._buf.shape[0],
                                                                     self.decoder._out_channels_for_plan, 0))
        self.reset()
        return res
