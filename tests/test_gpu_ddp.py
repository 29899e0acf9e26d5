"""Two-GPU check of the data-parallel training step (BASELINE config 5: DDP with an NCCL gradient all-reduce).
Skipped on boxes with fewer than two GPUs (the driver's `pytest -m gpu` box has one)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_training_step_matches_single_rank_on_the_global_batch(tmp_path):
    port = _free_port()
    worker = os.path.join(HERE, "_nccl_train_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), "2", str(port), str(tmp_path / f"r{r}.pt")])
             for r in range(2)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    r0, r1 = (torch.load(tmp_path / f"r{r}.pt") for r in range(2))
    # both ranks hold identical parameters after the step ...
    assert torch.equal(r0["flat_enc"], r1["flat_enc"]) and torch.equal(r0["flat_dec"], r1["flat_dec"])
    # ... and they equal a single-process step on the concatenated batch (mean over the global batch), fp32 mode.
    # The first AdamW step moves every parameter by ~lr * g / (|g| + eps) = up to 1e-3; summation order differs
    # (two half-batch sums + all-reduce vs one sum, float atomics), which only matters where |g| ~ eps.
    ref = r0["single"]
    for name in ("flat_enc", "flat_dec"):
        diff = (r0[name] - ref[name]).abs()
        assert float(diff.max()) <= 5e-5, (name, float(diff.max()))
        assert float(diff.mean()) <= 1e-7, (name, float(diff.mean()))
