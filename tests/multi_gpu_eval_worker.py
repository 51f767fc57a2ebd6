"""torchrun worker for tests/test_gpu_multi.py: BASELINE config 5 in miniature.  Every rank evaluates its contiguous
shard of a synthetic val split (SqueezeSegV2, nuScenes config), the int64 confusion matrices are summed with ONE
ncclAllReduce through the C ABI, and rank 0 checks the result against the matrix it computes alone over all frames."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pclsegmentation_b200.pipeline import Evaluator  # noqa: E402
from pclsegmentation_b200.sharding import Communicator, env_rank, shard_range  # noqa: E402
from pclsegmentation_b200.utils.args_loader import config_map, model_map  # noqa: E402
from tests.util import synth_range_images  # noqa: E402


def main():
  rank, local_rank, world = env_rank()
  torch.cuda.set_device(local_rank)
  dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
  comm = Communicator()
  mc = config_map["squeezesegv2nuscenes"]()
  mc.AZIMUTH_LEVEL = 256                      # keep the test small; H = 32, NC = 11, None = 10
  model = model_map["squeezesegv2"](mc)
  model.randomize_batch_norm(1)
  n_frames = 13                               # ragged shards
  frames = synth_range_images(np.random.default_rng(77), n_frames, 32, 256, valid_rate=0.6, num_classes=10)
  lo, hi = shard_range(n_frames, rank, world)
  ev = Evaluator(model, comm)
  for i in range(lo, hi, 4):
    ev.update(frames[i:min(i + 4, hi)])
  rep = ev.finish()
  ok = True
  if rank == 0:
    solo = Evaluator(model, None)
    solo.update(frames)
    ref = solo.finish()
    ok = bool(np.array_equal(rep["confusion_matrix"], ref["confusion_matrix"])) and \
        int(rep["confusion_matrix"].sum()) == n_frames * 32 * 256 and abs(rep["miou"] - ref["miou"]) < 1e-9
    print(json.dumps({"ok": ok, "world": world, "pixels": int(rep["confusion_matrix"].sum()), "miou": rep["miou"]}))
  # every rank must hold the same reduced matrix
  cm = ev.miou_tracker.total_cm.clone()
  cm0 = cm.clone()
  dist.broadcast(cm0, src=0)
  same = bool(torch.equal(cm, cm0))
  comm.close()
  dist.destroy_process_group()
  sys.exit(0 if (ok and same) else 1)


if __name__ == "__main__":
  main()
