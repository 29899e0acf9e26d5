"""Two O12 latent-1024 decodes at batch 1 (T = argv[1] frames, default 128) for ncu captures of single layers at the
streaming shape (conv_umma2 launch order per decode: 0 head conv, 1 convT 2048->1024, 2 k7 C1024, 3 k1 C1024, ...)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import helpers as H
torch.set_grad_enabled(False)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 128
m = H.build("o12_d1024", 0).to("cuda").set_precision("bf16")
z = torch.randn(1, 1024, T, device="cuda")
for _ in range(2):
    y = m.decode(z)
torch.cuda.synchronize()
print("ok", float(y.abs().max()))
