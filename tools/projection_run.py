"""Projection (config 4 shape) + confusion update, plain launches: the target of `ncu -k regex:project|confusion` captures
and of `compute-sanitizer --tool memcheck|racecheck` runs.   python tools/projection_run.py [scans] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pclsegmentation_b200.laserscan import SphericalProjector
from pclsegmentation_b200.metrics import MeanIoU
from tests.util import synth_scan

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
H, W = 64, 2048
rng = np.random.default_rng(4321)
sizes = rng.integers(115000, 125001, B)
scans = [synth_scan(rng, int(n)) for n in sizes]
labels = [rng.integers(0, 260, int(n)).astype(np.uint32) for n in sizes]
dev = torch.device("cuda", 0)
pts = torch.from_numpy(np.concatenate(scans)).to(dev)
lab = torch.from_numpy(np.concatenate(labels).view(np.int32)).to(dev)
offsets = torch.as_tensor(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)).to(dev)
proj = SphericalProjector(H, W, 3.0, -25.0, label_lut=np.arange(300, dtype=np.int32) % 20)
for _ in range(reps):
  out = proj.project(pts, offsets, labels=lab, empty_fill=0.0, want_sem=True, want_point_outputs=True)
m = MeanIoU(20)
label = (out["proj_sem_label"] % 20).to(torch.int32)
pred = torch.randint(0, 20, label.shape, dtype=torch.int32, device=dev)
for _ in range(reps):
  m.update_state(label, pred)
torch.cuda.synchronize()
print("ok", int(m.total_cm.sum().item()), int((out["proj_idx"] >= 0).sum().item()))
