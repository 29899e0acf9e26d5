"""Tiny driver for ncu: SAO decode at batch B (argv[1], default 2), a few iterations."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import helpers as H
torch.set_grad_enabled(False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
m = H.build("sao", 0).to("cuda").set_precision("bf16")
z = torch.randn(B, 64, 216, device="cuda")
for _ in range(3):
    y = m.decode(z)
torch.cuda.synchronize()
print("ok", tuple(y.shape))
