"""global matching of ONE sequence with the memory bank sharded over two GPUs -- local segmented min in the
tcgen05 matching kernel, partial minima exchanged from inside the kernel over NVLink (peer stores + arrival counters).
min is associative, so the sharded logits must be bit-identical to the single-GPU ones.  Needs two GPUs
(`gpurun --gpus 2`); skipped on a one-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_bank_sharded_matching_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    worker = os.path.join(os.path.dirname(__file__), "shard_match_worker.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", worker],
                       capture_output=True, text=True, timeout=900)
    print(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
