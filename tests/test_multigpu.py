"""-m gpu, needs >= 2 devices (skipped on the 1-GPU box): the NCCL data-parallel MLP step equals the oracle's
full-batch step and leaves bit-identical parameters on every rank."""
import os
import subprocess
import sys

import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("path,batch", [("fp32", "2000"), ("3xtf32", "4096"), ("fp32", "1001")])
def test_data_parallel_step_matches_oracle(path, batch):
    n = min(_ngpu(), 2 if batch != "4096" else 4)
    env = dict(os.environ, DP_PATH=path, DP_BATCH=batch)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "dp_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert "DP_CHECK_OK" in p.stdout, (p.stdout[-3000:], p.stderr[-3000:])
    assert "SHARDED_GEMM_OK" in p.stdout, (p.stdout[-3000:], p.stderr[-3000:])
    assert "UNET_DP_OK" in p.stdout, (p.stdout[-3000:], p.stderr[-3000:])


@pytest.mark.skipif(_ngpu() < 2 or not os.environ.get("BLA_TEST_PEER"),
                    reason="peer-window all-reduce (comm.cu) is opt-in until it has run on hardware: BLA_TEST_PEER=1")
def test_data_parallel_step_over_peer_windows():
    """The same data-parallel checks with the gradient all-reduce done by the library's own kernel over NVLink peer windows
    (BLA_PEER_ALLREDUCE=1) instead of NCCL."""
    env = dict(os.environ, DP_PATH="fp32", DP_BATCH="2000", BLA_PEER_ALLREDUCE="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(_ngpu(), 2)}", "--master-addr", "127.0.0.1",
           "--master-port", "29519", os.path.join(ROOT, "tests", "dp_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=180, env=env, cwd=ROOT)
    assert "DP_CHECK_OK" in p.stdout, (p.stdout[-3000:], p.stderr[-3000:])
    assert "peer windows unavailable" not in p.stderr
