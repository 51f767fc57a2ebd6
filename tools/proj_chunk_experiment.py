"""Experiment: projection of 64 scans in chunks that keep the winner keys (and the chunk's points) L2-resident."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pclsegmentation_b200 import _lib
from tests.util import synth_scan
lib = _lib.load()
B, H, W = 64, 64, 2048
rng = np.random.default_rng(4321)
sizes = rng.integers(115000, 125001, B)
dev = torch.device("cuda", 0)
bufs = [torch.from_numpy(np.concatenate([synth_scan(rng, int(n)) for n in sizes])).to(dev) for _ in range(3)]
off_np = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
image = torch.empty((B, H, W, 6), dtype=torch.float32, device=dev)
idx = torch.empty((B, H, W), dtype=torch.int32, device=dev)
s = torch.cuda.current_stream().cuda_stream
def run(chunk, it):
  pts = bufs[it % 3]
  keys = run.keys[chunk]
  for c0 in range(0, B, chunk):
    c1 = min(B, c0 + chunk)
    offs = run.offs[(chunk, c0)]
    start = int(off_np[c0]); total = int(off_np[c1] - off_np[c0])
    _lib.check(lib.pcls_project_scatter(pts.data_ptr() + start * 16, None, offs.data_ptr(), c1 - c0, total, H, W, 3.0, -25.0,
                                        keys.data_ptr(), None, None, None, s))
    _lib.check(lib.pcls_project_resolve(pts.data_ptr() + start * 16, None, offs.data_ptr(), c1 - c0, H, W, keys.data_ptr(), None, 0,
                                        0.0, image.data_ptr() + c0 * H * W * 24, idx.data_ptr() + c0 * H * W * 4, None, s))
run.keys, run.offs = {}, {}
ref = None
for chunk in (64, 32, 16, 8, 4):
  run.keys[chunk] = torch.empty((chunk, H, W), dtype=torch.int64, device=dev)
  for c0 in range(0, B, chunk):
    c1 = min(B, c0 + chunk)
    run.offs[(chunk, c0)] = torch.as_tensor(off_np[c0:c1 + 1] - off_np[c0]).to(dev)
  for i in range(3): run(chunk, i)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for i in range(30): run(chunk, i)
  e1.record(); torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / 30
  run(chunk, 0); torch.cuda.synchronize()
  chk = (int(idx.sum().item()), float(image.double().sum().item()))
  if ref is None: ref = chk
  alg = 16 * int(sizes.sum()) + B * H * W * 28
  print("chunk %2d: %.4f ms  %.0f scans/s  algorithmic %.0f GB/s  same result: %s" % (chunk, ms, B / ms * 1e3, alg / ms / 1e6, chk == ref))
