"""Tiny driver for ncu: SAO training steps (B clips x 5.016 s, bf16 mode).  argv[1] = B (default 4), argv[2] = steps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import helpers as H
from kalle_audio_b200 import training as TR
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
m = H.build("sao", 0).to("cuda").train()
x = 0.1 * torch.randn(B, 2, 108 * 2048, device="cuda")
noise = torch.randn(B, 64, 108, device="cuda")
tr = TR.AutoencoderTrainer(m, precision="bf16")
for _ in range(steps):
    info = tr.training_step(x, noise)
torch.cuda.synchronize()
print("ok", float(info["loss"]))
