"""Worker for tests/test_sharding_gloo.py: one rank of a gloo process group on CPU.
usage: python _gloo_worker.py <rank> <world> <port> <n_items> <out.pt>"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kalle_audio_b200.sharding import max_over_ranks, run_sharded  # noqa: E402

rank, world, port, n_items = (int(a) for a in sys.argv[1:5])
os.environ["MASTER_ADDR"] = "127.0.0.1"
os.environ["MASTER_PORT"] = str(port)
dist.init_process_group("gloo", rank=rank, world_size=world)
items = torch.arange(n_items * 3, dtype=torch.float32).view(n_items, 3)
seen = []


def fake_decode(x):            # stands in for ae.decode: per-item, no cross-batch term
    seen.append(x.shape[0])
    return x.repeat_interleave(2, dim=1) * 2.0


out = run_sharded(fake_decode, items, gather=True, micro_batch=2)
slowest = max_over_ranks(10.0 + rank)
torch.save({"rank": rank, "out": out, "slowest": slowest, "seen": sum(seen)}, sys.argv[5])
dist.destroy_process_group()
