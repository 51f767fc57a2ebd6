"""N > 1: one process per GPU (torchrun), frames sharded, ONE NCCL all-reduce of the int64 confusion matrix through
pcls_confusion_allreduce.  Skipped on boxes with a single GPU (the gloo test covers the host logic on CPU)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_eval_allreduce_equals_single_gpu():
  n = min(torch.cuda.device_count(), 8)
  cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
         "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "multi_gpu_eval_worker.py")]
  r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
  assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
  line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
  rep = json.loads(line)
  assert rep["ok"] and rep["world"] == n
