"""2-rank NCCL test of the sharded evaluation on real GPUs: one process per GPU via torchrun, each
rank evaluating its shard through the C ABI; the all-reduced loss and gradient must equal the reference's golden on
every rank, bit-identical across ranks, and L-BFGS must take the same branches everywhere (tests/nccl_worker.py).
Needs >= 2 GPUs (the driver's 1-GPU test box skips it; profiles/r2_nccl_2gpu.log holds a run)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (one process per GPU)")
def test_two_ranks_reproduce_the_single_process_answer():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    print(r.stdout[-3000:])
    print(r.stderr[-3000:])
    assert r.returncode == 0
    assert "NCCL_WORKER_OK" in r.stdout
