"""Tiny driver for ncu: SAO decode, B clips (default 8), twice.  conv_umma2_kernel launch order inside one decode:
0 latent conv, 1 convT, 2-7 ResidualUnits @1024 (k7, k1 alternating), 8 convT, 9-14 @512, 15 convT, 16-21 @256, 22 convT,
(fused units @128: conv_ru2_kernel), 23 convT 128->128."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import helpers as H
torch.set_grad_enabled(False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
m = H.build("sao", 0).to("cuda").set_precision("bf16")
z = torch.randn(B, 64, 216, device="cuda")
for _ in range(2):
    y = m.decode(z)
torch.cuda.synchronize()
print("ok", tuple(y.shape))
