"""Per-step CUDA-event profile of the SAO encoder + decoder at the bench shape (B = 16 x 216 frames, bf16 mode), grouped
by layer class.  Usage: python tools/prof_sao_b16.py [B]"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import helpers as H
import kalle_audio_b200 as k
from kalle_audio_b200 import _lib
import ctypes as C
torch.set_grad_enabled(False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
m = H.build("sao", 0).to(dev).set_precision("bf16")
z = torch.randn(B, 64, 216, device=dev)
x = 0.1 * torch.randn(B, 2, 216 * 2048, device=dev)
groups = collections.OrderedDict()
tot_all = 0.0
for name, mod, inp in (("decoder", m.decoder, z), ("encoder", m.encoder, x)):
    for _ in range(3):
        mod(inp)
    r = mod.runner(dev)
    r.set_profiling(True)
    acc = None
    N = 10
    for _ in range(N):
        mod(inp)
        p = r.step_profile()
        acc = [a + q[0] for a, q in zip(acc, p)] if acc else [q[0] for q in p]
    L = _lib.lib()
    info = (C.c_int * 8)()
    tot = sum(acc) / N
    tot_all += tot
    print(f"--- SAO {name} B={B}: {len(p)} steps, {tot:.3f} ms, {sum(q[1] for q in p) / tot / 1e9:.0f} TFLOP/s")
    for i, (q, a) in enumerate(zip(p, acc)):
        ms = a / N
        L.kvae_plan_conv_info(r.handle, i, C.byref(info))
        kind, cin, cout, K, s = info[0], info[1], info[2], info[3], info[4]
        if os.environ.get("PROF_STEPS"):
            print(f"    step {i:2d} kind {kind} C{cin}->{cout} k{K} s{s}  {ms:7.3f} ms  {q[1] / (ms * 1e-3 + 1e-12) / 1e12:7.0f} TFLOP/s")
        if ms < 0.005:
            cls = None          # second half of a fused unit
        elif min(cin, cout) <= 2:
            cls = "edge (io conv)"
        elif kind == 1:
            cls = f"convT C{cin}->{cout}"
        elif s > 1:
            cls = f"strided conv C{cin}->{cout}"
        elif K == 7 and cin == cout == 128:
            cls = "fused RU C128"
        elif K == 7:
            cls = f"k7 C{cin}"
        elif K == 1:
            cls = f"k1 C{cin}"
        else:
            cls = f"k{K} C{cin}->{cout}"
        if cls:
            g = groups.setdefault(cls, [0.0, 0.0, 0])
            fl = q[1] + (p[i + 1][1] if cls == "fused RU C128" else 0.0)
            g[0] += ms; g[1] += fl; g[2] += 1
print(f"=== by layer class (enc + dec = {tot_all:.3f} ms)")
for cls, (ms, fl, n) in sorted(groups.items(), key=lambda kv: -kv[1][0]):
    print(f"  {cls:28s} {n:3d} launches {ms:8.3f} ms {100 * ms / tot_all:5.1f} %  {fl / (ms * 1e-3) / 1e12:7.0f} TFLOP/s")
