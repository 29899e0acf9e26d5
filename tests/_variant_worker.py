"""Runs one SAO decode + encode (+ a ragged encode and a training step's gradients) under whatever KVAE_* switches the
parent test put into the environment and saves the outputs; tests/test_gpu_variants.py compares the files.  The kernel
selection switches are read once per process, hence the subprocess."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import helpers as H  # noqa: E402

out = sys.argv[1]
dev = torch.device("cuda:0")
torch.set_grad_enabled(False)
m = H.build("sao", 0, snake_seed=3).to(dev).set_precision("bf16")
z = torch.randn(2, 64, 24, generator=torch.Generator().manual_seed(1)).to(dev)
x = (0.1 * torch.randn(2, 2, 2048 * 12 + 777, generator=torch.Generator().manual_seed(2))).to(dev)
y = m.decode(z)
e = m.encoder(x[:, :, :2048 * 12])
valid = [2048 * 12 + 777, 2048 * 7 + 5]
xr = torch.zeros(2, 2, 2048 * 13, device=dev)
xr[:, :, :x.shape[-1]] = x
er = m.encoder(xr, valid_len=valid)
np.savez(out, y=y.float().cpu().numpy(), e=e.float().cpu().numpy(), er=er.float().cpu().numpy())
print("ok")
