"""Worker for tests/test_training_oracle.py::test_grad_sync_gloo_world2: one rank of a gloo group on CPU.
usage: python _gloo_gradsync_worker.py <rank> <world> <port> <out.pt>"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kalle_audio_b200.training import GradSync  # noqa: E402

rank, world, port = (int(a) for a in sys.argv[1:4])
os.environ["MASTER_ADDR"] = "127.0.0.1"
os.environ["MASTER_PORT"] = str(port)
dist.init_process_group("gloo", rank=rank, world_size=world)
sync = GradSync()
a = torch.arange(1000, dtype=torch.float32) * (rank + 1)     # "decoder" gradients
b = torch.full((17,), float(rank + 1))                       # "encoder" gradients
sync.launch(a)      # launched first, overlaps the work that produces b
sync.launch(b)
sync.wait()
# initial-state broadcast (AutoencoderTrainer.broadcast_parameters): ranks seeded differently continue from rank 0
torch.manual_seed(100 + rank)
w = torch.randn(333)
m1 = torch.randn(333)
sync.broadcast([w, m1], 0)
torch.save({"a": a, "b": b, "scale": sync.grad_scale, "world": sync.world, "w": w, "m1": m1}, sys.argv[4])
dist.destroy_process_group()
