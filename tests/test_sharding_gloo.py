"""world_size-2 gloo test of the N>1 host logic: contiguous frame shards + the confusion-matrix all-reduce give the
single-process matrix (the CUDA path does the same exchange with ncclAllReduce through the C ABI)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import confusion as C
from pclsegmentation_b200.sharding import Communicator, shard_range


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


def _worker(rank, world, port, label, pred, nc, out):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    lo, hi = shard_range(label.shape[0], rank, world)
    cm = torch.from_numpy(C.confusion_matrix(label[lo:hi], pred[lo:hi], nc))  # per-rank matrix of its frame shard
    comm = Communicator()
    assert (comm.rank, comm.world_size) == (rank, world)
    comm.allreduce_confusion(cm)
    out[rank] = cm.numpy().copy()
  finally:
    dist.destroy_process_group()


def test_two_rank_confusion_allreduce():
  rng = np.random.default_rng(0)
  nc, frames = 11, 7                                  # odd frame count: ragged shards (4 + 3)
  label = rng.integers(0, nc, (frames, 8, 32)).astype(np.int32)
  pred = rng.integers(0, nc, (frames, 8, 32)).astype(np.int32)
  mgr = mp.Manager()
  out = mgr.dict()
  mp.spawn(_worker, args=(2, _free_port(), label, pred, nc, out), nprocs=2, join=True)
  full = C.confusion_matrix(label, pred, nc)
  assert np.array_equal(out[0], full) and np.array_equal(out[1], full)
  assert full.sum() == label.size
