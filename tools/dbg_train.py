import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, helpers as H
from oracle import oobleck_oracle as O
import test_gpu_training as TG
dev=torch.device('cuda:0')
with torch.enable_grad():
    m = H.build("sao", 0).to(dev).train()
    H.randomize_snake(m, 7)
    x = 0.1 * torch.randn(1, 2, 2048 * 6, generator=torch.Generator().manual_seed(2))
    noise = torch.randn(1, 64, 6, generator=torch.Generator().manual_seed(3))
    sd = {kk: v.detach().cpu().clone().requires_grad_(True) for kk, v in m.state_dict().items()}
    ref_loss, _, _, _ = O.training_loss(sd, x, noise, H.strides_of("sao"), 1e-2, -1.0)
    ref_loss.backward()
    loss, kl, dec, grads = TG._loss_and_grads(m, x.to(dev), noise.to(dev), 1e-2, -1.0, "bf16")
bad=[]
for n, gr in grads.items():
    e = TG.rel_l2(gr, sd[n].grad)
    if e > 2.5e-2: bad.append((n, round(e,4)))
print(len(bad), 'of', len(grads)); print(bad[:40])
