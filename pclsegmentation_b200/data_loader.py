"""Device-side ``DataLoader.parse_sample`` (pcl_segmentation/data_loader/data_loader.py:153-187), batched.

The reference parses ONE ``.npy`` sample on the host with numpy: mask = depth > 0, float64 normalisation with
INPUT_MEAN / INPUT_STD, zero-fill, append the mask channel, ``label[~mask] = CLASSES.index("None")`` and the class-weight
map ``weight[label == l] = CLS_LOSS_WEIGHT[l]``.  ``parse_samples`` produces the same four arrays for a batch with one
kernel (pcls_input_stage); the TFRecord / tf.data plumbing around it (:54-136, :252-330) is out of scope.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .device import ptr, samples_to_device, stream_handle


def parse_samples(samples, mc):
  """samples [B,H,W,6] (x,y,z,intensity,depth,label; numpy float64 / float32 or a CUDA tensor) ->
  (lidar [B,H,W,6] f32, mask [B,H,W] bool, label [B,H,W] i32, weight [B,H,W] f32), CUDA tensors."""
  x = samples_to_device(samples)
  B, H, W, C = x.shape
  if C != 6:
    raise ValueError("samples need 6 channels (x,y,z,intensity,depth,label)")
  dev = x.device
  lidar = torch.empty((B, H, W, 6), dtype=torch.float32, device=dev)
  mask = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
  label = torch.empty((B, H, W), dtype=torch.int32, device=dev)
  weight = torch.empty((B, H, W), dtype=torch.float32, device=dev)
  mean = (ctypes.c_double * 5)(*np.asarray(mc.INPUT_MEAN, np.float64).reshape(-1))
  std = (ctypes.c_double * 5)(*np.asarray(mc.INPUT_STD, np.float64).reshape(-1))
  cw = (ctypes.c_double * int(mc.NUM_CLASS))(*np.asarray(mc.CLS_LOSS_WEIGHT, np.float64).reshape(-1))
  _lib.check(_lib.load().pcls_input_stage(ptr(x), 6, B * H * W, mean, std, int(mc.CLASSES.index("None")), ptr(lidar),
                                          ptr(mask), ptr(label), cw, int(mc.NUM_CLASS), ptr(weight), stream_handle()),
             "pcls_input_stage")
  return lidar, mask.view(torch.bool), label, weight
