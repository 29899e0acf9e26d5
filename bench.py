#!/usr/bin/env python
"""bench.py -- throughput of the sigmaVAE / Oobleck hot path on N B200s of one node.

    python bench.py --gpus 1 --steps K --warmup W                      (this repo's CUDA path)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                               (the reference's CPU arithmetic, host cores)

Workload (BASELINE.json configs[1]): SAO-shape autoencoder (channels 128, c_mults 1/2/4/8/16, strides
2/4/4/8/8, 44.1 kHz stereo, random init, synthetic audio), one step = encode -> split mean/scale ->
sample('fix') -> decode of 16 clips x 10.03 s per GPU in bf16 mode.  Batch items are independent, so N GPUs
run N independent shards with no data-path collective (weak scaling: 16 clips per GPU).

The JSON line:
  value        audio-seconds round-tripped per second, whole job, inputs resident in HBM
  e2e          same through the public module API with pinned HOST buffers: H2D of the audio and D2H of the
               decoded waveform inside the timed region
  decode_only  decode leg alone (the metric's name), same batch
  roofline     tensor-core roofline of the dominant kernel family (conv_umma2_kernel and its fused ResidualUnit
               form conv_ru2_kernel): algorithmic FLOPs of its
               launches / their CUDA-event time, against MEASURED_PEAKS.json (sustained figure: the kernel is
               timed inside a long step)
  cpu_baseline oracle port of the reference arithmetic on the host cores, bounded sample (rank 0, N=1 only)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SAO = dict(channels=128, c_mults=[1, 2, 4, 8, 16], strides=[2, 4, 4, 8, 8], enc_latent=128, dec_latent=64,
           io_channels=2, sample_rate=44100)
CLIP_FRAMES = 216                     # 216 latent frames = 442368 samples = 10.031 s
BATCH_PER_GPU = 16
METRIC = "vae_decode_audio_sec_per_sec"
UNIT = "audio-s/s"


def sao_config():
    import helpers as H
    return H.ae_config(SAO["channels"], SAO["c_mults"], SAO["strides"], SAO["enc_latent"], SAO["dec_latent"],
                       SAO["io_channels"], SAO["sample_rate"])


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"source": "measured", "hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"])}
    return {"source": "fallback", "hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def wait_first(self, timeout=8.0):
        """nvidia-smi needs up to a few seconds before its first line on an 8-GPU box: wait for it (outside any timed
        region) so that the timed region itself is covered."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def mark(self):
        return len(self.rows)

    def stats(self, i0=0, i1=None):
        """Clocks / power / throttle reasons of the samples taken between two marks (widened by one sample on each
        side when the region was shorter than the sampling period)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        i1 = len(self.rows) if i1 is None else i1
        rows = self.rows[i0:i1]
        if len(rows) < 2:
            rows = self.rows[max(0, i0 - 1):i1 + 1]
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); smax.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()


# ------------------------------------------------------------------------------------------ reference arm
def cpu_roundtrip_sample(frames):
    """Oracle port of the reference arithmetic (fp32, torch CPU kernels -- the same library calls the
    reference's nn.Modules make) on a bounded sample: 1 clip x `frames` latent frames, encode -> sample ->
    decode.  Returns (audio_seconds, seconds, threads)."""
    import torch
    import helpers as H
    from oracle import oobleck_oracle as O
    torch.set_grad_enabled(False)
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    import kalle_audio_b200 as k
    m = k.create_autoencoder_from_config(sao_config()).eval()     # parameter container only (CPU, same init)
    sd = m.state_dict()
    enc_sd, dec_sd = H.split_sd(sd, "encoder."), H.split_sd(sd, "decoder.")
    L = frames * 2048
    x = 0.1 * torch.randn(1, 2, L, generator=torch.Generator().manual_seed(2))
    noise = torch.randn(1, 64, frames, generator=torch.Generator().manual_seed(3))

    def step():
        t0 = time.perf_counter()
        e = O.oobleck_encoder(enc_sd, x, SAO["strides"])
        mean, _ = e.chunk(2, dim=1)
        z = O.sigma_sample(mean, noise, "fix")
        y = O.oobleck_decoder(dec_sd, z, SAO["strides"])
        assert y.shape == x.shape
        return time.perf_counter() - t0

    return L / SAO["sample_rate"], step, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames = 54
    audio_s, step, threads = cpu_roundtrip_sample(frames)
    for _ in range(args.warmup):
        step()
    times = [step() for _ in range(args.steps)]
    total = sum(times)
    value = audio_s * args.steps / total
    sample = f"1 clip x {frames} latent frames ({audio_s:.2f} s audio) round trip per step, fp32, oracle port"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return {"workload": "BASELINE configs[1]: SAO-shape sigmaVAE encode -> sample('fix') -> decode round trip, "
                        f"{BATCH_PER_GPU} clips x 10.03 s per GPU, bf16 tensor-core mode, fp32 I/O",
            "arch": "oobleck channels=128 c_mults=[1,2,4,8,16] strides=[2,4,4,8,8] latent 128/64 stereo 44.1k",
            "clips_per_gpu": BATCH_PER_GPU, "clip_seconds": CLIP_FRAMES * 2048 / SAO["sample_rate"],
            "sharding": f"batch-sharded x{n_gpus}, no collective on the data path",
            "l2": "per-step working set (>10 GB of activations) is far larger than the 126 MB L2"}


# ------------------------------------------------------------------------------------------ CUDA arm
def run_cuda(args):
    import torch
    import torch.distributed as dist
    import kalle_audio_b200 as k
    from kalle_audio_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.set_grad_enabled(False)

    torch.manual_seed(0)
    ae = k.create_autoencoder_from_config(sao_config()).eval().to(dev).set_precision("bf16")
    B, L = BATCH_PER_GPU, CLIP_FRAMES * 2048
    gen = torch.Generator().manual_seed(2 + rank)
    x_host = (0.1 * torch.randn(B, 2, L, generator=gen)).pin_memory()
    y_host = torch.empty(B, 2, L).pin_memory()
    x_dev = x_host.to(dev)
    audio_s_per_step = B * L / SAO["sample_rate"]

    def roundtrip(x):
        e = ae.encode(x)
        mean, _scale = e.chunk(2, dim=1)
        z = k.sample(mean.contiguous(), "fix")
        return ae.decode(z), z

    def step_resident():
        y, _ = roundtrip(x_dev)
        return y

    # end-to-end leg: the public host-buffer pipeline (kalle_audio_b200.HostPipeline) -- every step copies its input
    # from pinned host memory and its waveform back inside the timed region; the copies of consecutive steps overlap
    # the kernels (double-buffered device input, copy streams), as a serving loop over batches would run it
    pipe = k.HostPipeline(lambda xd: roundtrip(xd)[0], dev)

    def step_e2e():
        pipe.submit(x_host, y_host)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, sampler=None):
        barrier()
        i0 = sampler.mark() if sampler else 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        clocks = sampler.stats(i0, sampler.mark()) if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), clocks

    def timed_e2e(steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step_e2e()
        pipe.join()                      # the last step's device->host copy is inside the timed region
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), None

    sampler = ClockSampler(local)        # started before the warm-up: nvidia-smi is slow to produce its first line
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler.wait_first()
    # roofline leg: per-step CUDA events inside the library, on the launching stream
    enc_r, dec_r = ae.encoder.runner(dev), ae.decoder.runner(dev)
    enc_r.set_profiling(True); dec_r.set_profiling(True)
    _lib.lib().kvae_launch_count(1)
    total_ms, clocks = timed(step_resident, args.steps, sampler)
    sampler.stop()
    launches = int(_lib.lib().kvae_launch_count(0))
    prof = enc_r.step_profile() + dec_r.step_profile()          # last step of the timed region
    enc_r.set_profiling(False); dec_r.set_profiling(False)

    _, z_dev = roundtrip(x_dev)
    for _ in range(2):
        ae.decode(z_dev)
    dec_ms, _ = timed(lambda: ae.decode(z_dev), args.steps)
    for _ in range(2):
        step_e2e()
    e2e_ms, _ = timed_e2e(args.steps)

    if rank == 0:
        peaks = measured_peaks()
        tc = [(ms, fl) for ms, fl, on_tc in prof if on_tc]
        tc_ms, tc_fl = sum(m for m, _ in tc), sum(f for _, f in tc)
        other_ms = sum(ms for ms, _, on_tc in prof if not on_tc)
        achieved = tc_fl / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
        # the single kernel that dominates the step: conv_ru2_kernel (fused ResidualUnit of the 128-channel stages).
        # It moves 1024 B per output row against 2*128*128*8 FLOPs (256 FLOP/B, right at the machine balance), and
        # no unit is above 60 % in ncu (profiles/r01_ncu_summary.txt); reported against the HBM roofline, with the
        # tensor figure in `tflops`.
        ru = [(prof[i][0], prof[i][1]) for i in range(len(prof) - 1)
              if prof[i + 1][0] < 0.01 and abs(prof[i][1] / max(prof[i + 1][1], 1.0) - 7.0) < 1e-6]
        ru_ms = sum(m for m, _ in ru)
        ru_rows = sum(f / (2.0 * 128 * 128 * 7) for _, f in ru)
        ru_bytes = ru_rows * 1024.0                       # bf16 operand in + fp16 skip in + fp16 stream out + bf16 operand out
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "conv_ru_traffic.json")
        if os.path.exists(tpath) and ru:
            with open(tpath) as f:
                tj = json.load(f)
            traffic = tj["dram_bytes"] / tj["rows"] * (ru_rows / len(ru))
        step_fl = enc_r.flops(B, L) + dec_r.flops(B, CLIP_FRAMES)
        ms_per_step = total_ms / args.steps
        value = world * audio_s_per_step * args.steps / (total_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(world),
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": world * audio_s_per_step * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": y_host.numel() * 4},
            "decode_only": {"value": world * audio_s_per_step * args.steps / (dec_ms * 1e-3), "unit": UNIT,
                            "ms_per_step": dec_ms / args.steps,
                            "tflops": dec_r.flops(B, CLIP_FRAMES) / (dec_ms / args.steps * 1e-3) / 1e12},
            "roofline": {"bound": "tensor", "kernel": "conv_umma2_kernel + conv_ru2_kernel (tcgen05 conv family)", "achieved": achieved,
                         "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["bf16_tflops_sustained"], "traffic": None,
                         "peak_source": peaks["source"] + " (sustained; burst %.1f)" % peaks["bf16_tflops"],
                         "launches_per_step": len(tc), "kernel_ms_per_step": tc_ms,
                         "other_kernels_ms_per_step": other_ms,
                         "whole_step_tflops": step_fl / (ms_per_step * 1e-3) / 1e12,
                         "whole_step_frac": step_fl / (ms_per_step * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"]},
            "roofline_dominant_kernel": {
                "kernel": "conv_ru2_kernel (one launch = one ResidualUnit of a 128-channel stage)", "bound": "hbm",
                "achieved": ru_bytes / (ru_ms * 1e-3) / 1e9 if ru_ms > 0 else 0.0, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": (ru_bytes / (ru_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]) if ru_ms > 0 else 0.0,
                "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read+write, scaled by rows)",
                "launches_per_step": len(ru), "kernel_ms_per_step": ru_ms,
                "algorithmic_bytes_per_launch": ru_bytes / len(ru) if ru else 0.0,
                "share_of_step": ru_ms / ms_per_step, "peak_source": peaks["source"] + " (copy bandwidth)",
                "tflops": sum(f for _, f in ru) * 8.0 / 7.0 / (ru_ms * 1e-3) / 1e12 if ru_ms > 0 else 0.0},
        }
        if world == 1 and not args.no_cpu_baseline:
            frames = 54
            audio_s, step, threads = cpu_roundtrip_sample(frames)
            step()
            t = min(step() for _ in range(2))
            line["cpu_baseline"] = {"value": audio_s / t, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"1 clip x {frames} latent frames ({audio_s:.2f} s audio) round trip, "
                                              "fp32 oracle port, warm-up 1 + best of 2"}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ other BASELINE configs
O12 = dict(channels=128, c_mults=[1, 2, 4, 8, 16], strides=[2, 4, 4, 5, 8], io_channels=1, sample_rate=16000)


def run_extra(args):
    """BASELINE configs[2] and [3] (not the driver's default line; same JSON keys where they apply).
       o12_decode: vae_12_5_dim1024-shape decoder (latent 512), 64 clips x 30 s, batch-sharded over the ranks
                   (strong scaling: 64 / N clips per GPU), decode only.
       stream:     vae_12_5hz_dim2048-shape decoder (latent 1024), batch 1, decode_audio(chunked=True,
                   chunk_size=128, overlap=32) semantics over T=375 -- first-chunk latency, per-chunk latency,
                   real-time factor; CUDA-graph replay on."""
    import torch
    import torch.distributed as dist
    import kalle_audio_b200 as k
    from kalle_audio_b200.sharding import shard_bounds

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.set_grad_enabled(False)
    torch.manual_seed(0)
    latent = 512 if args.workload == "o12_decode" else 1024
    dec = k.OobleckDecoder(out_channels=1, channels=O12["channels"], latent_dim=latent, c_mults=O12["c_mults"],
                           strides=O12["strides"], use_snake=True, final_tanh=False).eval().to(dev)
    dec.set_precision("bf16")

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    if args.workload == "o12_decode":
        total_clips, T = 64, 375
        lo, hi = shard_bounds(total_clips, world, rank)
        z = torch.randn(hi - lo, latent, T, generator=torch.Generator().manual_seed(1 + rank)).to(dev)
        mb = args.micro_batch or (hi - lo)

        def step():
            for i in range(0, hi - lo, mb):
                dec(z[i:i + mb])

        for _ in range(max(args.warmup, 3)):
            step()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            audio_s = total_clips * T * 1280 / O12["sample_rate"]
            r = dec.runner(dev)
            flops = r.flops(1, T) * total_clips
            t = float(ms.item()) / args.steps * 1e-3
            print(json.dumps({"metric": METRIC, "value": audio_s / t, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                              "warmup": max(args.warmup, 3), "ms_per_step": t * 1e3, "higher_is_better": True,
                              "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                              "tflops_per_gpu": flops / t / 1e12 / world,
                              "config": {"workload": "BASELINE configs[2]: O12 latent-512 (dim1024) decoder, 64 clips x 30 s, "
                                                     f"batch-sharded {total_clips // world} per GPU, micro-batch {mb}, bf16 mode"}}),
                  flush=True)
    else:
        T, chunk, overlap = 375, 128, 32
        ae = k.AudioAutoencoder(None, dec, latent_dim=latent, downsampling_ratio=1280, sample_rate=16000, io_channels=1)
        dec.enable_cuda_graphs(True)
        z = torch.randn(1, latent, T, generator=torch.Generator().manual_seed(1)).to(dev)
        for _ in range(max(args.warmup, 3)):
            ae.decode_audio(z, chunked=True, overlap=overlap, chunk_size=chunk)
            dec(z[:, :, :chunk])
        sync()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        for _ in range(args.steps):
            dec(z[:, :, :chunk])                      # one chunk, as a streaming caller would issue it
        ev[1].record()
        for _ in range(args.steps):
            ae.decode_audio(z, chunked=True, overlap=overlap, chunk_size=chunk)   # all 4 windows in one batched call
        ev[2].record()
        sync()
        t0 = time.perf_counter()
        y = dec(z[:, :, :chunk])
        y[0, 0, :8].cpu()
        first_wall = time.perf_counter() - t0
        # exact-context streaming (kalle_audio_b200.StreamingDecoder): hop 96, window 96 + 10 + 10 frames
        sdec = k.StreamingDecoder(dec, hop=chunk - overlap)
        zz = torch.randn(1, latent, 96 * 12, generator=torch.Generator().manual_seed(5)).to(dev)
        for i in range(4):
            sdec.push(zz[:, :, i * 96:(i + 1) * 96])
        sync()
        es = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        es[0].record()
        for i in range(4, 12):
            sdec.push(zz[:, :, i * 96:(i + 1) * 96])
        es[1].record()
        sync()
        stream_hop_ms = es[0].elapsed_time(es[1]) / 8
        if rank == 0:
            chunk_ms = ev[0].elapsed_time(ev[1]) / args.steps
            full_ms = ev[1].elapsed_time(ev[2]) / args.steps
            audio_s = T * 1280 / 16000
            print(json.dumps({"metric": METRIC, "value": audio_s / (full_ms * 1e-3), "unit": UNIT, "n_gpus": 1,
                              "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": full_ms,
                              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                              "data": "synthetic", "per_chunk_ms": chunk_ms,
                              "first_chunk_wall_ms_incl_d2h": first_wall * 1e3,
                              "real_time_factor_per_chunk": (chunk - overlap) * 1280 / 16000 / (chunk_ms * 1e-3),
                              "exact_context_stream": {"hop_frames": chunk - overlap, "window_frames": chunk - overlap + sdec.left + sdec.right,
                                                       "ms_per_hop": stream_hop_ms, "recompute_factor": sdec.recompute_factor,
                                                       "real_time_factor": (chunk - overlap) * 1280 / 16000 / (stream_hop_ms * 1e-3)},
                              "config": {"workload": "BASELINE configs[3]: O12 latent-1024 (dim2048) decoder, batch 1, "
                                                     "chunked decode chunk 128 / overlap 32 over T=375 (4 windows), "
                                                     "CUDA-graph replay, bf16 mode"}}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_eager_baseline(args):
    """Context number, not a bench line the driver consumes: the oracle restatement (plain torch functional ops =
    what the reference's nn.Modules execute, cuDNN convs + eager elementwise kernels) run on the same B200 in bf16
    autocast, decode only, 16 clips x 10.03 s.  SURVEY.md section 8d names this 'the only existing Blackwell path'."""
    import torch
    import helpers as H
    from oracle import oobleck_oracle as O
    import kalle_audio_b200 as k
    torch.set_grad_enabled(False)
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    m = k.create_autoencoder_from_config(sao_config()).eval()
    dec_sd = {n: p.to(dev) for n, p in H.split_sd(m.state_dict(), "decoder.").items()}
    B = args.micro_batch or 4
    z = torch.randn(B, 64, CLIP_FRAMES, device=dev)

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return O.oobleck_decoder(dec_sd, z, SAO["strides"])

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    audio_s = B * CLIP_FRAMES * 2048 / SAO["sample_rate"]
    print(json.dumps({"impl": "torch-eager-bf16 (oracle restatement on cuda:0, cuDNN + eager elementwise)",
                      "metric": METRIC, "value": audio_s / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
                      "ms_per_step": ms, "dtype": "bf16 autocast", "config": {"workload": f"SAO decode only, {B} clips x 10.03 s"}}),
          flush=True)


def run_train(args):
    """BASELINE configs[4]: SAO encoder + decoder training step (encode -> vae_sample -> decode -> Gaussian NLL +
    KL -> backward -> gradient all-reduce -> AdamW), 4 clips x 5.016 s per GPU (8 GPUs = the config's global batch
    of 32), bf16 tensor-core mode with fp32 master weights.  Reports step time, trained audio-seconds per second,
    algorithmic TFLOP/s (3 x forward FLOPs) and the exposed all-reduce time (step with the all-reduce minus the
    same step without it)."""
    import torch
    import torch.distributed as dist
    import kalle_audio_b200 as k
    from kalle_audio_b200 import _lib, training as TR

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    ae = k.create_autoencoder_from_config(sao_config()).train().to(dev)
    B, frames = args.micro_batch or 4, 108
    L = frames * 2048
    x = (0.1 * torch.randn(B, 2, L, generator=torch.Generator().manual_seed(2 + rank))).to(dev)
    noise = torch.randn(B, 64, frames, generator=torch.Generator().manual_seed(3 + rank)).to(dev)
    tr = TR.AutoencoderTrainer(ae, lr=1e-4, kl_weight=1e-4, log_sigma=-2.0, precision="bf16")

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(steps):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            info = tr.training_step(x, noise)
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps, info

    losses = []
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        losses.append(float(tr.training_step(x, noise)["loss"]))
    _lib.lib().kvae_launch_count(1)
    sampler.wait_first()
    i0 = sampler.mark()
    ms, info = timed(args.steps)
    clocks = sampler.stats(i0, sampler.mark())
    sampler.stop()
    launches = int(_lib.lib().kvae_launch_count(0))
    losses.append(float(info["loss"]))
    ms_nosync = ms
    if world > 1:
        tr.sync.enabled = False            # same step without the gradient all-reduce (ranks diverge; timing only)
        ms_nosync, _ = timed(args.steps)
        tr.sync.enabled = True
    if rank == 0:
        peaks = measured_peaks()
        enc_r, dec_r = ae.encoder.runner(dev), ae.decoder.runner(dev)
        fwd = enc_r.flops(B, L) + dec_r.flops(B, frames)
        tfl = 3.0 * fwd / (ms * 1e-3) / 1e12
        audio_s = world * B * L / SAO["sample_rate"]
        print(json.dumps({
            "metric": "vae_train_audio_sec_per_sec", "value": audio_s / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "BASELINE configs[4]: SAO sigmaVAE training step fwd+bwd (Gaussian NLL + KL), "
                                   f"{B} clips x {L / SAO['sample_rate']:.3f} s per GPU (global batch {B * world}), bf16 "
                                   "tensor-core operands, fp32 master weights / gradients / AdamW, NCCL all-reduce of "
                                   "the flat gradients overlapped with the encoder backward"},
            "clocks": clocks, "gpu_launches": launches,
            "tflops_per_gpu": tfl, "frac_of_sustained_bf16": tfl / peaks["bf16_tflops_sustained"],
            "allreduce_exposed_ms": ms - ms_nosync, "params": int(tr.flat_enc.numel() + tr.flat_dec.numel()),
            "loss_first_to_last": [losses[0], losses[-1]]}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["cuda", "reference"], default="cuda")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", choices=["roundtrip", "o12_decode", "stream", "train", "eager_baseline"], default="roundtrip",
                    help="roundtrip = BASELINE configs[1] (the driver's line); o12_decode = configs[2]; stream = configs[3]; "
                         "train = configs[4]")
    ap.add_argument("--micro-batch", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "train":
        run_train(args)
    elif args.workload == "eager_baseline":
        run_eager_baseline(args)
    elif args.workload != "roundtrip":
        run_extra(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
