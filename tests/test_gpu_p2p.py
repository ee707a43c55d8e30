"""glf_p2p_allreduce (one kernel over NVLink peer memory, csrc/glf_p2p.cu) against NCCL: needs >= 2 GPUs on the box
(skipped on the single-GPU runs; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_p2p.py`)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_p2p_allreduce_matches_nccl():
    n = min(torch.cuda.device_count(), 8)
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "p2p_worker.py")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", worker],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "P2P_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
