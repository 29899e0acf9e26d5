"""Worker for tests/test_gpu_ddp.py: one rank of a 2-GPU NCCL group runs one AutoencoderTrainer step on its half
of a global batch; rank 0 also runs the same step single-process on the whole batch for comparison.
usage: python _nccl_train_worker.py <rank> <world> <port> <out.pt>"""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import helpers as H  # noqa: E402
from kalle_audio_b200 import training as TR  # noqa: E402

rank, world, port = (int(a) for a in sys.argv[1:4])
os.environ["MASTER_ADDR"] = "127.0.0.1"
os.environ["MASTER_PORT"] = str(port)
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

x = 0.1 * torch.randn(4, 2, 40 * 16, generator=torch.Generator().manual_seed(9))
noise = torch.randn(4, 64, 16, generator=torch.Generator().manual_seed(10))
per = 4 // world


def one_step(xb, nb, group_enabled):
    # data-parallel run: every rank initialises from its OWN seed; the trainer broadcasts rank 0's parameters
    # (seed 0) at construction, so the ranks still end up with the single-process result
    m = H.build("mid", rank if group_enabled else 0, snake_seed=7 + (rank if group_enabled else 0)).to(dev).train()
    tr = TR.AutoencoderTrainer(m, lr=1e-3, kl_weight=1e-2, log_sigma=-1.0, precision="fp32", data_parallel=group_enabled)
    assert tr.replica_checksum_spread() == 0.0
    tr.training_step(xb.to(dev), nb.to(dev))
    torch.cuda.synchronize(dev)
    assert tr.replica_checksum_spread() == 0.0
    return tr.flat_enc.detach().cpu().clone(), tr.flat_dec.detach().cpu().clone()


fe, fd = one_step(x[rank * per:(rank + 1) * per], noise[rank * per:(rank + 1) * per], True)
out = {"flat_enc": fe, "flat_dec": fd}
if rank == 0:
    se, sdd = one_step(x, noise, False)
    out["single"] = {"flat_enc": se, "flat_dec": sdd}
dist.barrier()
torch.save(out, sys.argv[4])
dist.destroy_process_group()
