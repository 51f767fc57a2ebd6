"""Two plain (no CUDA graph) forwards of the bench workload, for `ncu -k regex:... --launch-skip N -c M` captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pclsegmentation_b200.utils.args_loader import model_map

workload = sys.argv[1] if len(sys.argv) > 1 else "squeezesegv2_kitti_64x2048_b32"
name, mc, B = bench.make_config(workload)
model = model_map[name](mc)
model.randomize_batch_norm(1)
model.set_option("use_graph", 0)
raw = torch.from_numpy(bench.synth_raw(1, B, mc.ZENITH_LEVEL, mc.AZIMUTH_LEVEL)).cuda()
for _ in range(2):
  model.forward_device(raw, None, mean=mc.INPUT_MEAN, std=mc.INPUT_STD)
torch.cuda.synchronize()
