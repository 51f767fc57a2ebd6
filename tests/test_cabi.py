"""The C-ABI library loads and exports every symbol include/pclseg.h declares; without a GPU the compute entry
points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from pclsegmentation_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
  src = open(os.path.join(ROOT, "include", "pclseg.h")).read()
  src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
  return sorted(set(re.findall(r"\b(pcls_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
  lib = _lib.load()
  names = header_symbols()
  assert len(names) >= 20
  for n in names:
    assert hasattr(lib, n), "libpclseg.so does not export %s" % n
  assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES and include/pclseg.h disagree"
  assert lib.pcls_abi_version() == _lib.ABI_VERSION == 2


def test_struct_layout_matches_header():
  # pcls_conv_desc: 6 ints, 6 pointers, float, 7 ints ; pcls_cam_desc: 2 ints, 12 pointers, float, 2 ints
  assert ctypes.sizeof(_lib.ConvDesc) == 6 * 4 + 6 * 8 + 4 + 7 * 4 + 0 or ctypes.sizeof(_lib.ConvDesc) % 8 == 0
  assert _lib.ConvDesc.h_kernel.offset == 24 and _lib.ConvDesc.bn_eps.offset == 72
  assert _lib.CamDesc.h_sq_kernel.offset == 8 and _lib.CamDesc.bn_eps.offset == 104


def test_argument_validation_without_gpu():
  lib = _lib.load()
  assert lib.pcls_head(None, None, 10, 64, 0, None, None, None) == -1
  assert b"num_classes" in lib.pcls_last_error()
  assert lib.pcls_confusion_update(None, None, 0, 99, None, None, None) == -1
  assert lib.pcls_project_scatter(None, None, None, 1, 0, 0, 8, 3.0, -25.0, None, None, None, None, None) == -1
  assert lib.pcls_input_stage(None, 4, 0, None, None, 0, None, None, None, None, 0, None, None) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
  lib = _lib.load()
  h = ctypes.c_void_p()
  assert lib.pcls_net_create(ctypes.byref(h), 32, 240, 0, 1) == -2
  assert b"no CUDA device" in lib.pcls_last_error()
  from pclsegmentation_b200.utils.args_loader import load_model_config
  import numpy as np
  mc, model = load_model_config("squeezesegv2", "squeezesegv2")
  with pytest.raises(_lib.PclsError):
    model([np.zeros((1, 32, 240, 6), np.float32), np.ones((1, 32, 240), bool)])
  from pclsegmentation_b200.laserscan import LaserScan
  with pytest.raises(_lib.PclsError):
    LaserScan(project=True).set_points(np.ones((4, 3), np.float32))
  from pclsegmentation_b200.metrics import MeanIoU
  with pytest.raises(_lib.PclsError):
    MeanIoU(11).update_state(np.zeros(4, np.int32), np.zeros(4, np.int32))
