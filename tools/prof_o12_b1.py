"""Per-step CUDA-event profile of the O12 latent-1024 decoder at batch 1 (BASELINE config 4's chunk), T = argv[1] (128)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import kalle_audio_b200 as k
torch.set_grad_enabled(False)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda:0")
torch.manual_seed(0)
dec = k.OobleckDecoder(out_channels=1, channels=128, latent_dim=1024, c_mults=[1, 2, 4, 8, 16], strides=[2, 4, 4, 5, 8],
                       use_snake=True, final_tanh=False).eval().to(dev).set_precision("bf16")
z = torch.randn(B, 1024, T, device=dev)
for _ in range(5):
    dec(z)
r = dec.runner(dev)
r.set_profiling(True)
acc = None
for _ in range(20):
    dec(z)
    p = r.step_profile()
    acc = [a + q[0] for a, q in zip(acc, p)] if acc else [q[0] for q in p]
tot = sum(acc) / 20
print(f"O12 latent-1024 decode B={B} T={T}: {len(p)} steps, {tot:.4f} ms, {sum(q[1] for q in p) / tot / 1e9:.1f} TFLOP/s")
for i, (q, a) in enumerate(zip(p, acc)):
    ms = a / 20
    if ms > 0.0015:
        print(f"  step {i:2d} {ms * 1e3:7.1f} us {q[1] / 1e9:8.2f} GF {q[1] / (ms * 1e-3) / 1e12:8.1f} TFLOP/s")
