"""GPU (-m gpu, needs 2 GPUs: skipped on the 1-GPU round-end box; run with `gpurun --gpus 2`): one data-parallel training step
under DistributedDataParallel + NCCL exactly as train.py:139-148 drives it (scripts/ddp_train_step.py)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("freeze", [0, 4])
def test_ddp_training_step_two_gpus(freeze):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29540 + freeze), os.path.join(ROOT, "scripts", "ddp_train_step.py"), "--batch", "4", "--steps", "2",
           "--freeze", str(freeze)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    print(line)
    assert d["n_gpus"] == 2 and d["params_identical_after_step"] and d["grad_allreduce_err_vs_mean_of_local_grads"] < 2e-3
