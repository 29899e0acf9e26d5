"""Two SAO encodes at B = argv[1] (16) for ncu launch lists / captures of single encoder layers."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import helpers as H
torch.set_grad_enabled(False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
m = H.build("sao", 0).to("cuda").set_precision("bf16")
x = 0.1 * torch.randn(B, 2, 216 * 2048, device="cuda")
for _ in range(2):
    y = m.encoder(x)
torch.cuda.synchronize()
print("ok", float(y.abs().max()))
