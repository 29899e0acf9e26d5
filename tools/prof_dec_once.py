"""Two SAO decodes at B = argv[1] (16) for ncu captures of single layers (conv_umma2 launch order per decode: 0 first conv,
1 convT, 2..7 k7/k1 x3 (C=1024), 8 convT, 9..14 (C=512), 15 convT, 16 k7 C256, 17 k1 C256, ..., 22 convT 128->128)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import helpers as H
torch.set_grad_enabled(False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
m = H.build("sao", 0).to("cuda").set_precision("bf16")
z = torch.randn(B, 64, 216, device="cuda")
for _ in range(2):
    y = m.decode(z)
torch.cuda.synchronize()
print("ok", float(y.abs().max()))
