"""Diagnostics on a B200: parity errors per configuration and per-step timing of the fused plans.
Writes gpurun_out/report.txt.  Not a test and not a benchmark -- a development aid."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import helpers as H
import kalle_audio_b200 as k

torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
out = open(os.path.join(ROOT, "gpurun_out", "report.txt"), "w")


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    out.write(s + "\n")
    out.flush()


def stage_names(plan_steps, direction):
    return [f"step{i}" for i in range(plan_steps)]


def profile(runner, fn, label):
    runner.set_profiling(True)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    prof = runner.step_profile()
    runner.set_profiling(False)
    tot = sum(p[0] for p in prof)
    P(f"--- {label}: {len(prof)} steps, {tot:.3f} ms total, {sum(p[1] for p in prof) / tot / 1e9:.1f} TFLOP/s")
    for i, (ms, fl, tc) in enumerate(prof):
        P(f"   step {i:2d} {'TC ' if tc else 'gen'} {ms:8.3f} ms {fl / 1e9:9.2f} GF {fl / ms / 1e9:8.1f} TFLOP/s")


try:
    g = H.golden("sao_full")
    m = H.build("sao", 0).to(dev).set_precision("bf16")
    z = torch.randn(1, 64, 216, generator=torch.Generator().manual_seed(1)).to(dev)
    x = (0.1 * torch.randn(1, 2, 442368, generator=torch.Generator().manual_seed(2))).to(dev)
    y = m.decode(z)
    idx = H.t(g["dec_idx"]).long().to(dev)
    d = (y[:, :, idx].cpu() - H.t(g["dec_out_at_idx"]))
    P(f"SAO decode bf16: max err {float(d.abs().max()):.3e} rms err {float(d.pow(2).mean().sqrt()):.3e} "
      f"ref absmax {float(g['dec_abs_max']):.3f} n={d.numel()}")
    e = m.encode(x)
    de = e.cpu() - H.t(g["enc_out"])
    P(f"SAO encode bf16: max err {float(de.abs().max()):.3e} rms err {float(de.pow(2).mean().sqrt()):.3e} "
      f"ref absmax {float(np.abs(g['enc_out']).max()):.3f}")
    profile(m.decoder.runner(dev), lambda: m.decode(z), "SAO decode B=1 T=216 bf16")
    m32 = H.build("sao", 0).to(dev).set_precision("fp32")
    z4 = torch.randn(4, 64, 216, device=dev)
    profile(m32.decoder.runner(dev), lambda: m32.decode(z4), "SAO decode B=4 T=216 fp32 mode B=8-style (bf16x3 split on tensor cores)")
    y32 = m32.decode(z)
    d32 = (y32[:, :, idx].cpu() - H.t(g["dec_out_at_idx"]))
    P(f"SAO decode fp32 mode, full clip: max err {float(d32.abs().max()):.3e}")
    del m32
    profile(m.encoder.runner(dev), lambda: m.encode(x), "SAO encode B=1 L=442368 bf16")
    zb = torch.randn(8, 64, 216, device=dev)
    profile(m.decoder.runner(dev), lambda: m.decode(zb), "SAO decode B=8 T=216 bf16")
    xb = 0.1 * torch.randn(8, 2, 442368, device=dev)
    profile(m.encoder.runner(dev), lambda: m.encode(xb), "SAO encode B=8 L=442368 bf16")
    del xb
    t0 = time.time()
    m.set_precision("fp32")
    z6 = z[:, :, :24].contiguous()
    y32 = m.decode(z6)
    torch.cuda.synchronize()
    P(f"SAO fp32-mode decode T=24 ran in {time.time() - t0:.2f} s (incl. plan build)")
    profile(m.decoder.runner(dev), lambda: m.decode(z6), "SAO decode B=1 T=24 fp32 mode")
    # fp32 mode vs bf16 mode on the same short clip = bf16 error without needing the CPU
    m.set_precision("bf16")
    y16 = m.decode(z6)
    P(f"SAO T=24: bf16 vs fp32 mode max diff {float((y16 - y32).abs().max()):.3e} "
      f"rms {float((y16 - y32).pow(2).mean().sqrt()):.3e}")
except Exception as ex:  # noqa
    import traceback
    P("REPORT FAILED:", traceback.format_exc())
out.close()
