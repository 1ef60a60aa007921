"""Multi-GPU partition on real GPUs: the partitioned model (one rank per GPU, NCCL) must reproduce the single-GPU model —
log-density, Gibbs draws given the same z, prediction, beta step and a lock-step chain.  Needs >= 2 GPUs; on a
single-GPU box the partition logic is covered by tests/test_partition_cpu.py (gloo) instead."""
import os
import subprocess
import sys

import pytest

from common import ROOT

pytestmark = pytest.mark.gpu


# 4 ranks cut the tree below level 1 (4 subtrees), 8 ranks below level 2 (16 subtrees, two per rank): the frontier shapes of
# the 4- and 8-GPU bench runs
@pytest.mark.parametrize("nranks,q,n,limited", [(2, 3, 20000, False), (2, 1, 6000, False), (4, 3, 40000, False), (8, 3, 60000, False),
                                                (2, 3, 20000, True), (4, 2, 30000, True)])   # limited_tree = TRUE on a partition
def test_partitioned_model_matches_single_gpu(nranks, q, n, limited):
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    port = 29700 + os.getpid() % 200
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "run_partition.py"), str(q), str(n)] + (["limited"] if limited else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "PARTITION PARITY OK" in r.stdout
