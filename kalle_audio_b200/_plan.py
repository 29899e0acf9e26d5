"""Host side of a fused encoder / decoder plan: creates the libkvae plan for an Oobleck module,
keeps its packed weights in sync with the module's parameters, owns the workspace and issues
``kvae_encode`` / ``kvae_decode`` on the caller's current CUDA stream."""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib
from .layers import SnakeBeta, _WNConvBase, nearest_upsample_conv_taps


class PlanRunner:
    def __init__(self, module: torch.nn.Module, direction: int, arch: _lib.KvaeArch, precision: int,
                 device: torch.device):
        self.module_ref = weakref.ref(module)
        self.direction = direction
        self.precision = precision
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.KvaeError("plans exist on CUDA devices only (no CPU path)")
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        L = _lib.lib()
        handle = C.c_void_p()
        _lib.check(L.kvae_plan_create(C.byref(arch), direction, precision, self.index, C.byref(handle)))
        self.handle = handle
        self._finalizer = weakref.finalize(self, L.kvae_plan_destroy, handle)
        self.convs: List[_WNConvBase] = [m for m in module.modules() if isinstance(m, _WNConvBase)]
        self.snakes: List[SnakeBeta] = [m for m in module.modules() if isinstance(m, SnakeBeta)]
        if len(self.convs) != L.kvae_plan_num_convs(handle) or len(self.snakes) != L.kvae_plan_num_snakes(handle):
            raise _lib.KvaeError("module tree does not match the plan built from its constructor arguments")
        info = (C.c_int * 8)()
        for i, m in enumerate(self.convs):
            _lib.check(L.kvae_plan_conv_info(handle, i, C.byref(info)))
            want = (int(m.transposed), m.in_channels, m.out_channels, m.kernel_size[0], m.stride[0], m.dilation[0],
                    m.padding[0], int(m.bias is not None))
            us = getattr(m, "_upsample_stride", 0)
            if us:     # Upsample(nearest) + 'same' conv runs as ConvTranspose1d(k = 3s - 1, stride s, padding s)
                want = (1, m.in_channels, m.out_channels, 3 * us - 1, us, 1, us, 0)
            if tuple(info) != want:
                raise _lib.KvaeError(f"conv {i}: module {want} does not match plan {tuple(info)}")
        for i, m in enumerate(self.snakes):
            if L.kvae_plan_snake_channels(handle, i) != m.in_features:
                raise _lib.KvaeError(f"SnakeBeta {i}: channel count mismatch")
        self._fingerprint: Optional[Tuple] = None
        self._workspace: Optional[torch.Tensor] = None
        # optional CUDA-graph replay of the ~38 launches of one pass (small batches are launch-bound)
        self.use_graphs = False
        self._graphs: Dict[Tuple, Tuple] = {}

    # ------------------------------------------------------------------ weights
    def _params(self):
        for m in self.convs:
            yield from m.parameters(recurse=False)
        for m in self.snakes:
            yield m.alpha
            yield m.beta

    def sync_weights(self) -> None:
        """Re-folds weight norm (fp32) and re-packs when any parameter changed (in-place update, load_state_dict,
        .to()).  Load-time cost only; the steady state is a tuple comparison."""
        module = self.module_ref()
        fp = (getattr(module, "_weights_epoch", 0),) + tuple((p.data_ptr(), p._version) for p in self._params())
        if fp == self._fingerprint:
            return
        L = _lib.lib()
        st = _lib.stream_ptr(self.device)
        for i, m in enumerate(self.convs):
            if next(m.parameters(recurse=False)).device != self.device:
                raise _lib.KvaeError("module parameters moved to another device; plan is stale")
            w = m.folded_weight()
            if getattr(m, "_upsample_stride", 0):
                w = nearest_upsample_conv_taps(w, m._upsample_stride)
            b = None if m.bias is None else m.bias.detach().float().contiguous()
            _lib.check(L.kvae_plan_set_conv(self.handle, i, w.data_ptr(), _lib.ptr(b), st))
        for i, m in enumerate(self.snakes):
            a = m.alpha.detach().float().contiguous()
            b = m.beta.detach().float().contiguous()
            _lib.check(L.kvae_plan_set_snake(self.handle, i, a.data_ptr(), b.data_ptr(), int(m.alpha_logscale), st))
        self._fingerprint = fp
        self._graphs.clear()        # packed weights are rewritten in place, but keep capture state simple

    # ------------------------------------------------------------------ run
    def _get_workspace(self, B: int, T: int) -> torch.Tensor:
        need = _lib.lib().kvae_workspace_bytes(self.handle, B, T)
        if need == 0:
            raise _lib.KvaeError(_lib.lib().kvae_last_error().decode())
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = None  # release before growing
            self._graphs.clear()    # captured graphs point into the old workspace
            self._workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._workspace

    def valid_out_length(self, T: int) -> int:
        """Output length for an input of length T by the reference's conv arithmetic (floors; odd strides)."""
        return int(_lib.lib().kvae_plan_out_length(self.handle, int(T)))

    def out_length(self, T: int, ratio: int) -> int:
        return T * ratio if self.direction == _lib.KVAE_DECODER else T // ratio

    def run(self, x: torch.Tensor, out_channels: int, ratio: int, out_dtype: torch.dtype,
            valid_len=None) -> torch.Tensor:
        """``valid_len`` (B ints, host): ragged batch -- clip b holds valid_len[b] valid input positions and is
        zero-padded to the common length; its output beyond the matching length is zeroed (kvae_*_ragged)."""
        _lib.require_cuda(x, "fused plan")
        if x.device != self.device:
            raise _lib.KvaeError(f"input on {x.device}, plan on {self.device}")
        if x.dim() != 3:
            raise ValueError("expected [B, C, T]")
        B, _, T = x.shape
        if B == 0 or T == 0:
            raise ValueError("empty input")
        self.sync_weights()
        xin = x if x.dtype in (torch.float32, torch.bfloat16) else x.float()
        xin = xin.contiguous()
        kdtype = out_dtype if out_dtype in (torch.float32, torch.bfloat16) else torch.float32
        if self.direction == _lib.KVAE_ENCODER and T % ratio:
            raise ValueError(f"audio length {T} is not a multiple of the downsampling ratio {ratio} "
                             "(use preprocess_audio_for_encoder)")
        ws = self._get_workspace(B, T)
        L = _lib.lib()
        fn = L.kvae_decode if self.direction == _lib.KVAE_DECODER else L.kvae_encode
        out_shape = (B, out_channels, self.out_length(T, ratio))

        if valid_len is not None:
            lens = [int(v) for v in valid_len]
            if len(lens) != B or any(v < 0 or v > T for v in lens):
                raise ValueError("valid_len needs one length in 0..T per clip")
            arr = (C.c_int * B)(*lens)
            fn = L.kvae_decode_ragged if self.direction == _lib.KVAE_DECODER else L.kvae_encode_ragged
            out = torch.empty(out_shape, dtype=kdtype, device=self.device)
            _lib.check(fn(self.handle, xin.data_ptr(), _lib.dtype_code(xin.dtype), out.data_ptr(), _lib.dtype_code(kdtype),
                          B, T, arr, ws.data_ptr(), ws.numel(), _lib.stream_ptr(self.device)))
            out_valid = torch.tensor([self.valid_out_length(v) for v in lens], device=self.device)
            out = out * (torch.arange(out_shape[2], device=self.device)[None, :] < out_valid[:, None])[:, None, :].to(out.dtype)
            return out if out.dtype == out_dtype else out.to(out_dtype)

        def launch(src, dst):
            _lib.check(fn(self.handle, src.data_ptr(), _lib.dtype_code(src.dtype), dst.data_ptr(),
                          _lib.dtype_code(kdtype), B, T, ws.data_ptr(), ws.numel(), _lib.stream_ptr(self.device)))

        if self.use_graphs and not torch.cuda.is_current_stream_capturing():
            key = (B, T, xin.dtype, kdtype)
            entry = self._graphs.get(key)
            if entry is None:
                static_in = torch.empty_like(xin)
                static_out = torch.empty(out_shape, dtype=kdtype, device=self.device)
                static_in.copy_(xin)
                launch(static_in, static_out)          # warm-up: builds descriptors, sets kernel attributes
                torch.cuda.current_stream(self.device).synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    launch(static_in, static_out)
                entry = (graph, static_in, static_out)
                self._graphs[key] = entry
            graph, static_in, static_out = entry
            static_in.copy_(xin)
            graph.replay()
            out = static_out.clone()
        else:
            out = torch.empty(out_shape, dtype=kdtype, device=self.device)
            launch(xin, out)
        return out if out.dtype == out_dtype else out.to(out_dtype)

    def run_decode_pcm16(self, x: torch.Tensor, out_channels: int, ratio: int, out_dtype: torch.dtype):
        """(wav, int16 pcm) with the peak search fused into the tail conv (kvae_decode_pcm16); None if unsupported."""
        L = _lib.lib()
        if self.direction != _lib.KVAE_DECODER or not L.kvae_plan_fused_pcm_supported(self.handle):
            return None
        _lib.require_cuda(x, "fused plan")
        B, _, T = x.shape
        if B == 0 or T == 0:
            raise ValueError("empty input")
        self.sync_weights()
        xin = x if x.dtype in (torch.float32, torch.bfloat16) else x.float()
        xin = xin.contiguous()
        kdtype = out_dtype if out_dtype in (torch.float32, torch.bfloat16) else torch.float32
        ws = self._get_workspace(B, T)
        wav = torch.empty((B, out_channels, T * ratio), dtype=kdtype, device=self.device)
        pcm = torch.empty((B, out_channels, T * ratio), dtype=torch.int16, device=self.device)
        scratch = torch.empty(4, dtype=torch.uint8, device=self.device)
        _lib.check(L.kvae_decode_pcm16(self.handle, xin.data_ptr(), _lib.dtype_code(xin.dtype), wav.data_ptr(),
                                       _lib.dtype_code(kdtype), pcm.data_ptr(), B, T, ws.data_ptr(), ws.numel(),
                                       scratch.data_ptr(), _lib.stream_ptr(self.device)))
        return wav, pcm

    def run_encode_sample(self, x: torch.Tensor, noise: torch.Tensor, out_channels: int, ratio: int,
                          out_dtype: torch.dtype, D: int, std: float):
        """(latents [B, out_channels, T], z [B, D, T]) with the sigma-VAE sample fused into the encoder's last conv
        (kvae_encode_sample); None when this plan has no tensor-core output conv (the caller then samples separately)."""
        L = _lib.lib()
        if self.direction != _lib.KVAE_ENCODER or not L.kvae_plan_fused_sample_supported(self.handle):
            return None
        _lib.require_cuda(x, "fused plan")
        B, _, T = x.shape
        if T % ratio:
            raise ValueError(f"audio length {T} is not a multiple of the downsampling ratio {ratio}")
        self.sync_weights()
        xin = x if x.dtype in (torch.float32, torch.bfloat16) else x.float()
        xin = xin.contiguous()
        kdtype = out_dtype if out_dtype in (torch.float32, torch.bfloat16) else torch.float32
        ws = self._get_workspace(B, T)
        lat = torch.empty((B, out_channels, T // ratio), dtype=kdtype, device=self.device)
        z = torch.empty((B, D, T // ratio), dtype=kdtype, device=self.device)
        nz = noise.to(kdtype).contiguous()
        if nz.shape != z.shape:
            raise ValueError(f"noise must be {tuple(z.shape)}")
        _lib.check(L.kvae_encode_sample(self.handle, xin.data_ptr(), _lib.dtype_code(xin.dtype), lat.data_ptr(), z.data_ptr(),
                                        nz.data_ptr(), _lib.dtype_code(kdtype), D, float(std), B, T, ws.data_ptr(),
                                        ws.numel(), _lib.stream_ptr(self.device)))
        return lat, z

    # ------------------------------------------------------------------ training pass
    def param_list(self) -> List[torch.nn.Parameter]:
        """Parameters in the flat-buffer order of the plan (= module.parameters() order); checked against
        the segment sizes libkvae reports."""
        module = self.module_ref()
        params = list(module.parameters())
        L = _lib.lib()
        n = 4 * (len(self.convs) + len(self.snakes)) + 8
        sizes = (C.c_longlong * n)()
        got = L.kvae_plan_param_sizes(self.handle, sizes, n)
        if got < 0:
            raise _lib.KvaeError(L.kvae_last_error().decode())
        if [p.numel() for p in params] != [int(sizes[i]) for i in range(got)]:
            raise _lib.KvaeError("training needs weight-normalised modules whose parameters() follow the reference's "
                                 "order (alpha, beta, bias, weight_g, weight_v per layer); remove_weight_norm'ed or "
                                 "re-ordered modules are inference-only")
        return params

    def flatten(self, params) -> torch.Tensor:
        """One fp32 buffer holding every parameter in plan order.  Zero-copy when the parameters already are
        views of such a buffer (``training.flatten_parameters``), else a ``torch.cat``."""
        flat = getattr(self, "flat_master", None)
        if flat is not None and flat.device == self.device:
            off, ok = flat.data_ptr(), True
            for p in params:
                if p.data_ptr() != off or p.dtype != torch.float32:
                    ok = False
                    break
                off += p.numel() * 4
            if ok:
                return flat
        return torch.cat([p.detach().reshape(-1).float() for p in params])

    def forward_train(self, x: torch.Tensor, flat: torch.Tensor, out_channels: int, ratio: int,
                      out_dtype: torch.dtype):
        """Forward pass that keeps every layer's activations in a fresh workspace; returns (y, xin, workspace)."""
        _lib.require_cuda(x, "fused plan (training)")
        if x.device != self.device:
            raise _lib.KvaeError(f"input on {x.device}, plan on {self.device}")
        if x.dim() != 3:
            raise ValueError("expected [B, C, T]")
        B, _, T = x.shape
        if B == 0 or T == 0:
            raise ValueError("empty input")
        if self.direction == _lib.KVAE_ENCODER and T % ratio:
            raise ValueError(f"audio length {T} is not a multiple of the downsampling ratio {ratio}")
        L = _lib.lib()
        st = _lib.stream_ptr(self.device)
        logscale = {bool(m.alpha_logscale) for m in self.snakes}
        if len(logscale) > 1:
            raise _lib.KvaeError("mixed alpha_logscale settings are not supported in training")
        _lib.check(L.kvae_plan_load_params(self.handle, flat.data_ptr(), int(logscale.pop() if logscale else True), 1, st))
        self._fingerprint = None        # inference packs are refreshed lazily from the module parameters
        xin = x.detach()
        xin = xin if xin.dtype in (torch.float32, torch.bfloat16) else xin.float()
        xin = xin.contiguous()
        kdtype = out_dtype if out_dtype in (torch.float32, torch.bfloat16) else torch.float32
        need = L.kvae_train_workspace_bytes(self.handle, B, T)
        if need == 0:
            raise _lib.KvaeError(L.kvae_last_error().decode())
        ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        out = torch.empty((B, out_channels, self.out_length(T, ratio)), dtype=kdtype, device=self.device)
        _lib.check(L.kvae_forward_train(self.handle, xin.data_ptr(), _lib.dtype_code(xin.dtype), out.data_ptr(),
                                        _lib.dtype_code(kdtype), B, T, ws.data_ptr(), ws.numel(), st))
        return (out if out.dtype == out_dtype else out.to(out_dtype)), xin, ws

    def backward(self, xin: torch.Tensor, gy: torch.Tensor, ws: torch.Tensor, flat: torch.Tensor, need_gx: bool):
        """(gx or None, flat gradient buffer) for the pass whose activations live in ``ws``."""
        L = _lib.lib()
        B, _, T = xin.shape
        g = gy.detach()
        g = g if g.dtype in (torch.float32, torch.bfloat16) else g.float()
        g = g.contiguous()
        grads = torch.empty(L.kvae_plan_param_count(self.handle), dtype=torch.float32, device=self.device)
        gx = torch.empty_like(xin, dtype=torch.float32) if need_gx else None
        _lib.check(L.kvae_backward(self.handle, xin.data_ptr(), _lib.dtype_code(xin.dtype), g.data_ptr(),
                                   _lib.dtype_code(g.dtype), _lib.ptr(gx), _lib.KVAE_F32, B, T, ws.data_ptr(),
                                   ws.numel(), grads.data_ptr(), flat.data_ptr(), _lib.stream_ptr(self.device)))
        self.last_grads = grads
        hook = getattr(self, "grads_ready_hook", None)
        if hook is not None:
            hook(self, grads)
        return gx, grads

    def set_profiling(self, enable: bool) -> None:
        _lib.check(_lib.lib().kvae_plan_profile(self.handle, int(enable)))

    def step_profile(self):
        """[(ms, flops, on_tensor_cores)] per convolution step of the last profiled run (synchronises)."""
        n = 256
        ms, fl, tc = (C.c_float * n)(), (C.c_double * n)(), (C.c_int * n)()
        got = _lib.lib().kvae_plan_step_profile(self.handle, ms, fl, tc, n)
        if got < 0:
            raise _lib.KvaeError(_lib.lib().kvae_last_error().decode())
        return [(float(ms[i]), float(fl[i]), bool(tc[i])) for i in range(got)]

    def flops(self, B: int, T: int) -> float:
        return float(_lib.lib().kvae_plan_flops(self.handle, B, T))


class PlanFunction(torch.autograd.Function):
    """autograd node of OobleckEncoder.forward / OobleckDecoder.forward: one kvae_forward_train call forward,
    one kvae_backward call backward (the reference differentiates ~150 eager ops per direction instead)."""

    @staticmethod
    def forward(ctx, runner: "PlanRunner", x, out_channels, ratio, out_dtype, *params):
        flat = runner.flatten(params)
        y, xin, ws = runner.forward_train(x, flat, out_channels, ratio, out_dtype)
        ctx.runner, ctx.xin, ctx.ws, ctx.flat = runner, xin, ws, flat
        ctx.x_dtype = x.dtype
        ctx.shapes = [p.shape for p in params]
        ctx.dtypes = [p.dtype for p in params]
        return y

    @staticmethod
    def backward(ctx, gy):
        gx, grads = ctx.runner.backward(ctx.xin, gy, ctx.ws, ctx.flat, ctx.needs_input_grad[1])
        ctx.ws = None
        if gx is not None and gx.dtype != ctx.x_dtype:
            gx = gx.to(ctx.x_dtype)
        if not getattr(ctx.runner, "return_param_grads", True):
            # trainer mode: the flat buffer (runner.last_grads) is the product; see AutoencoderTrainer
            return (None, gx, None, None, None, *([None] * len(ctx.shapes)))
        outs, off = [], 0
        for shape, dt in zip(ctx.shapes, ctx.dtypes):
            n = shape.numel()
            g = grads[off:off + n].view(shape)
            outs.append(g if dt == torch.float32 else g.to(dt))
            off += n
        return (None, gx, None, None, None, *outs)


class PlanCache:
    """Per-module cache of runners keyed by (device, precision)."""

    def __init__(self):
        self.runners: Dict[Tuple[str, int], PlanRunner] = {}

    def get(self, module, direction, arch, precision, device) -> PlanRunner:
        key = (str(device), precision)
        r = self.runners.get(key)
        if r is None:
            r = PlanRunner(module, direction, arch, precision, device)
            init = getattr(module, "_runner_init", None)      # e.g. AutoencoderTrainer's flat buffer + all-reduce hook
            if init is not None:
                init(r)
            self.runners[key] = r
        return r

    def clear(self):
        self.runners.clear()

    # plans hold device handles: a copied / pickled module starts with an empty cache and rebuilds lazily
    def __deepcopy__(self, memo):
        return PlanCache()

    def __getstate__(self):
        return {}

    def __setstate__(self, state):
        self.runners = {}
