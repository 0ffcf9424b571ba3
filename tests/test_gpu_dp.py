"""Data-parallel correctness on real GPUs (-m gpu, skipped below 2 devices): launches tests/dp_check.py under torchrun -
averaged gradient == CPU-oracle gradient on the concatenated global batch, identical gradients on every rank, through the
single captured graph that contains both NCCL all-reduces (model._dp_step)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("graph", ["1", "0"], ids=["one_graph_with_nccl", "eager_phases"])
def test_data_parallel_gradient_matches_oracle(graph):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    env = dict(os.environ, SSHSLIE_DP_GRAPH=graph)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "dp_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT, env=env)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-3000:])
    assert "DP_CHECK ok" in out.stdout
