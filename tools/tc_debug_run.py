"""Prints the per-role wait-cycle counters of every conv_tc_kernel launch of one forward (producer / issuer / epilogue).
Needs a library built with the counters compiled in:

    PCLS_NVCC_FLAGS=-DPCLS_TC_DEBUG=1 python -m pclsegmentation_b200.build --force && python tools/tc_debug_run.py

(rebuild without the flag afterwards: the counters cost 2.6 % on SqueezeSegV2).
"""
import sys, ctypes, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pclsegmentation_b200 import _lib
from pclsegmentation_b200.utils.args_loader import model_map
name, mc, B = bench.make_config(sys.argv[1] if len(sys.argv) > 1 else "squeezesegv2_kitti_64x2048_b32")
if len(sys.argv) > 2: B = int(sys.argv[2])
model = model_map[name](mc); model.randomize_batch_norm(1)
model.set_option("tc_debug", 1); model.set_option("use_graph", 0)
raw = torch.from_numpy(bench.synth_raw(1, B, 64, 2048)).cuda()
for _ in range(2): model.forward_device(raw, None, mean=mc.INPUT_MEAN, std=mc.INPUT_STD)
torch.cuda.synchronize()
lib = _lib.load(); net = model._net
n = lib.pcls_net_num_ops(net); ms = (ctypes.c_float * n)()
mean = (ctypes.c_double*5)(*mc.INPUT_MEAN.reshape(-1)); std=(ctypes.c_double*5)(*mc.INPUT_STD.reshape(-1))
preds = torch.empty((B,64,2048), dtype=torch.int32, device="cuda")
_lib.check(lib.pcls_net_profile_ops(net, raw.data_ptr(), 5, None, mean, std, B, None, None, preds.data_ptr(), ms, torch.cuda.current_stream().cuda_stream))
