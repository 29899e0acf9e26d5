"""World-size-2 (and 3) gloo tests of the batch-sharding host logic on CPU (no GPU, no compute kernels)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

import helpers  # noqa: F401
from kalle_audio_b200.sharding import max_over_ranks, run_sharded, shard_bounds


def test_shard_bounds_partition():
    for n in (0, 1, 7, 16, 64, 65):
        for w in (1, 2, 3, 4, 8):
            cover = []
            for r in range(w):
                lo, hi = shard_bounds(n, w, r)
                assert 0 <= lo <= hi <= n
                cover += list(range(lo, hi))
            assert cover == list(range(n))
            sizes = [shard_bounds(n, w, r)[1] - shard_bounds(n, w, r)[0] for r in range(w)]
            assert max(sizes) - min(sizes) <= 1
    assert [shard_bounds(64, 8, r) for r in (0, 7)] == [(0, 8), (56, 64)]       # BASELINE config 3: 64 clips / 8 GPUs
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,n_items", [(2, 7), (2, 8), (3, 5)])
def test_run_sharded_gloo(world, n_items, tmp_path):
    port = _free_port()
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_gloo_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), str(world), str(port), str(n_items),
                               str(tmp_path / f"r{r}.pt")]) for r in range(world)]
    for p in procs:
        assert p.wait(timeout=180) == 0
    res = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    items = torch.arange(n_items * 3, dtype=torch.float32).view(n_items, 3)
    want = items.repeat_interleave(2, dim=1) * 2.0
    assert torch.equal(res[0]["out"], want)                      # rank 0 holds every clip, in order
    assert all(r["out"] is None for r in res[1:])
    assert all(r["slowest"] == 10.0 + world - 1 for r in res)    # max over ranks
    assert sum(r["seen"] for r in res) == n_items                # every clip processed exactly once
