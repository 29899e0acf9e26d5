"""Host-buffer pipeline around a device-side call (encode, decode, or a whole round trip).

The reference's callers move a batch to the GPU, run the autoencoder and fetch the result, one batch after the
other on one stream (infer_0828_sigma.py:286-298, twj_dataset.py:231-256).  With pinned host buffers the three
phases of consecutive batches can overlap: `HostPipeline.submit` copies batch k+1 host->device on a copy stream
while batch k's kernels run on the caller's stream and batch k-1's result goes device->host on a third stream.
Nothing here computes: it is torch streams and events around `fn` (plumbing, like torch.distributed elsewhere).
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch


class HostPipeline:
    """`submit(x_host, y_host)` enqueues  x_host -> device -> fn -> y_host  without blocking the host.

    * x_host / y_host: pinned CPU tensors (pageable memory would serialise the copies); y_host may be reused by
      consecutive submissions (device->host copies are ordered on one stream).
    * fn: device tensor -> device tensor, launched on the stream that is current when `submit` is called.
    * depth: device-side input buffers (2 = double buffering).
    `join()` makes the caller's stream wait for all outstanding device->host copies (put a CUDA event / timer after
    it); `synchronize()` additionally blocks the host.
    """

    def __init__(self, fn: Callable[[torch.Tensor], torch.Tensor], device: torch.device, depth: int = 2):
        if torch.device(device).type != "cuda":
            raise ValueError("HostPipeline needs a CUDA device: there is no CPU path")
        self.fn = fn
        self.device = torch.device(device)
        self.depth = max(1, int(depth))
        self.s_in = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self.x_dev: List[Optional[torch.Tensor]] = [None] * self.depth
        self.ev_in = [torch.cuda.Event() for _ in range(self.depth)]
        self.ev_free: List[Optional[torch.cuda.Event]] = [None] * self.depth
        self.k = 0

    def submit(self, x_host: torch.Tensor, y_host: torch.Tensor) -> None:
        if x_host.device.type != "cpu" or y_host.device.type != "cpu":
            raise ValueError("HostPipeline.submit takes host tensors")
        compute = torch.cuda.current_stream(self.device)
        b = self.k % self.depth
        self.k += 1
        buf = self.x_dev[b]
        if buf is None or buf.shape != x_host.shape or buf.dtype != x_host.dtype:
            # allocated on the compute stream's pool; the first use below is ordered after this point by ev_free / ev_in
            buf = self.x_dev[b] = torch.empty(x_host.shape, dtype=x_host.dtype, device=self.device)
            self.s_in.wait_stream(compute)
        if self.ev_free[b] is not None:
            self.s_in.wait_event(self.ev_free[b])          # the kernels that read this buffer two batches ago are done
        with torch.cuda.stream(self.s_in):
            buf.copy_(x_host, non_blocking=True)
            self.ev_in[b].record(self.s_in)
        compute.wait_event(self.ev_in[b])
        y = self.fn(buf)
        ev = torch.cuda.Event()
        ev.record(compute)
        self.ev_free[b] = ev
        self.s_out.wait_event(ev)
        with torch.cuda.stream(self.s_out):
            y_host.copy_(y, non_blocking=True)
        y.record_stream(self.s_out)                        # the allocator must not hand y out again before the copy ran

    def join(self) -> None:
        torch.cuda.current_stream(self.device).wait_stream(self.s_out)

    def synchronize(self) -> None:
        self.s_out.synchronize()
        torch.cuda.current_stream(self.device).synchronize()
