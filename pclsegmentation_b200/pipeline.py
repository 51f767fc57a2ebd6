"""Device-resident pipelines over the C-ABI kernels:

* ``ScanSegmenter``  raw scans -> spherical projection -> network forward -> labels (BASELINE config 4: the converter's
  per-scan loop dataset_convert/semantic_kitti.py:152-179 fused with inference.py:44-78; range images never leave HBM)
* ``Evaluator``      the eval.py hot loop (eval.py:45-50): forward -> head -> confusion update per batch on this rank's
  shard of the frames, one all-reduce of the int64 matrix at the end, then IoU / recall / precision / mIoU.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .device import ptr, samples_to_device, stream_handle, to_device
from .laserscan import SphericalProjector
from .metrics import MeanIoU
from .sharding import Communicator, shard_range
from .utils.util import confusion_matrix_to_iou_recall_precision


class ScanSegmenter:
  def __init__(self, model, fov_up=3.0, fov_down=-25.0, label_lut=None):
    mc = model.mc
    self.model, self.mc = model, mc
    self.projector = SphericalProjector(mc.ZENITH_LEVEL, mc.AZIMUTH_LEVEL, fov_up, fov_down, label_lut=label_lut)

  def segment_device(self, points, offsets, labels=None, want_probabilities=False, out=None, want_image=False,
                     want_logits=False):
    """points [total,4] float32 CUDA (x,y,z,remission), offsets [B+1] int64 CUDA -> dict(predictions, proj_idx,
    probabilities?, logits?, image?) - keys and activations stay on the device.

    Default (no image, no labels wanted): the resolve pass writes the NORMALISED 16-bit network input and the mask
    straight into the net's own buffers (pcls_project_resolve_net_input): no float32 [B,H,W,6] range image is written
    and read back, and the forward runs without its input kernel.  With want_image / labels the range image is
    materialised (the converters' output) and the forward reads it."""
    B = int(offsets.numel()) - 1
    if want_image or labels is not None or B > self._staged_capacity(B):
      proj = self.projector.project(points, offsets, labels=labels, empty_fill=0.0)
      res = self.model.forward_device(proj["image"], None, mean=self.mc.INPUT_MEAN, std=self.mc.INPUT_STD,
                                      want_probabilities=want_probabilities, want_logits=want_logits, out=out)
      res.update(image=proj["image"], proj_idx=proj["proj_idx"])
      return res
    lib = _lib.load()
    H, W = self.projector.H, self.projector.W
    inp, msk, _ = self.model.input_buffers(B)
    keys = torch.empty((B, H, W), dtype=torch.int64, device=points.device)
    idx = torch.empty((B, H, W), dtype=torch.int32, device=points.device)
    s = stream_handle()
    _lib.check(lib.pcls_project_scatter(ptr(points), None, ptr(offsets), B, int(points.shape[0]), H, W,
                                        float(self.projector.fov_up), float(self.projector.fov_down), ptr(keys), None, None,
                                        None, s), "pcls_project_scatter")
    mean = (ctypes.c_double * 5)(*np.asarray(self.mc.INPUT_MEAN, np.float64).reshape(-1))
    std = (ctypes.c_double * 5)(*np.asarray(self.mc.INPUT_STD, np.float64).reshape(-1))
    _lib.check(lib.pcls_project_resolve_net_input(ptr(points), ptr(offsets), B, H, W, ptr(keys), mean, std,
                                                  self.model.precision, inp, msk, ptr(idx), s),
               "pcls_project_resolve_net_input")
    res = self.model.forward_staged(B, want_probabilities=want_probabilities, want_logits=want_logits, out=out)
    res.update(proj_idx=idx)
    return res

  def _staged_capacity(self, B):
    return self.model.input_buffers(B)[2]

  def segment(self, scans, labels=None, **kw):
    """scans: list of [N_i,4] float32 numpy arrays."""
    lens = [int(s.shape[0]) for s in scans]
    dev = self.projector._dev
    offsets = torch.as_tensor(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).to(dev)
    pts = to_device(np.concatenate(scans, axis=0), torch.float32)
    lab = None
    if labels is not None:
      lab = to_device(np.concatenate(labels).astype(np.uint32).view(np.int32), torch.int32)
    kw.setdefault("want_image", True)     # host convenience: the projected range images come back with the labels
    return self.segment_device(pts, offsets, labels=lab, **kw)

  def point_labels(self, res, points_per_scan):
    """Back-projects the per-pixel predictions to the points (each point takes the label of its pixel)."""
    raise NotImplementedError("per-point label back-projection is not part of the reference path")


class Evaluator:
  """eval.py:33-58 on one rank; ``comm`` sums the confusion matrix over ranks (sharding.Communicator) or is None."""

  def __init__(self, model, comm=None):
    self.model, self.mc = model, model.mc
    self.comm = comm
    self.miou_tracker = MeanIoU(num_classes=self.mc.NUM_CLASS, name="MeanIoU")
    self._none = self.mc.CLASSES.index("None")
    self._mean = (ctypes.c_double * 5)(*np.asarray(self.mc.INPUT_MEAN).reshape(-1))
    self._std = (ctypes.c_double * 5)(*np.asarray(self.mc.INPUT_STD).reshape(-1))

  def update(self, samples, return_label=False):
    """samples: [B,H,W,6] float64 / float32 (x,y,z,i,d,label) numpy or CUDA tensor - the .npy frames of the dataset.
    Runs the fused input stage + forward + head, fixes the labels up (label[~mask] = None, data_loader.py:176) and
    accumulates the confusion matrix; nothing returns to the host."""
    x = samples_to_device(samples)
    B, H, W, C = x.shape
    if C != 6:
      raise ValueError("evaluation samples need 6 channels (x,y,z,intensity,depth,label)")
    res = self.model.forward_device(x, None, mean=self.mc.INPUT_MEAN, std=self.mc.INPUT_STD, want_probabilities=False)
    label = torch.empty((B, H, W), dtype=torch.int32, device=x.device)
    _lib.check(_lib.load().pcls_input_stage(ptr(x), 6, B * H * W, self._mean, self._std, self._none, None, None,
                                            ptr(label), None, 0, None, stream_handle()), "pcls_input_stage")
    self.miou_tracker.update_state(label, res["predictions"])
    return (res["predictions"], label) if return_label else res["predictions"]

  def finish(self):
    """All-reduce (if sharded) and compute the report of eval.py:50-58."""
    if self.comm is not None:
      self.miou_tracker.allreduce(self.comm)
    cm = self.miou_tracker.total_cm
    iou, recall, precision = confusion_matrix_to_iou_recall_precision(cm)
    return dict(confusion_matrix=cm.cpu().numpy(), iou=iou, recall=recall, precision=precision,
                miou=float(self.miou_tracker.result()))


def shard_files(files, comm=None):
  files = sorted(files)
  if comm is None or comm.world_size == 1:
    return files
  lo, hi = shard_range(len(files), comm.rank, comm.world_size)
  return files[lo:hi]
