"""``MeanIoU`` - the slice of ``tf.keras.metrics.MeanIoU`` the reference uses (eval.py:41,48,50,58;
nets/SegmentationNetwork.py:52): ``update_state(label, pred)``, ``total_cm``, ``result()``, ``reset_states()``.

The confusion matrix lives on the GPU as int64 (csrc/confusion.cu: shared-memory histogram kernel); TF keeps it in
float32, which is exact only below 2^24 per cell.  ``allreduce()`` is the one multi-GPU exchange step of the path:
one ncclAllReduce(int64, sum) over the per-GPU matrices.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .device import ptr, require_cuda, stream_handle, to_device


class MeanIoU:
  def __init__(self, num_classes, name="MeanIoU"):
    self.num_classes = int(num_classes)
    self.name = name
    self._cm = None
    self._cmw = None      # float64 matrix of weighted updates (test_step); None until the first weighted update
    self._dropped = None

  def _ensure(self):
    if self._cm is None:
      dev = require_cuda()
      self._cm = torch.zeros((self.num_classes, self.num_classes), dtype=torch.int64, device=dev)
      self._dropped = torch.zeros(1, dtype=torch.int64, device=dev)
    return self._cm

  def update_state(self, y_true, y_pred, sample_weight=None):
    """cm[label, pred] += 1 over every element (eval.py passes no weights: masked pixels count as None/None)."""
    cm = self._ensure()
    label = to_device(y_true, torch.int32).reshape(-1)
    pred = to_device(y_pred, torch.int32).reshape(-1)
    if label.numel() != pred.numel():
      raise ValueError("label and prediction sizes differ: %d vs %d" % (label.numel(), pred.numel()))
    if sample_weight is not None:
      # test_step (nets/SegmentationNetwork.py:129): tf.math.confusion_matrix(..., weights=w) sums the weights per cell
      # (float64 shared-memory histogram, csrc/validation.cu)
      w = to_device(sample_weight, torch.float32).reshape(-1)
      if w.numel() != label.numel():
        raise ValueError("weights and label sizes differ: %d vs %d" % (w.numel(), label.numel()))
      _lib.check(_lib.load().pcls_validation_update(None, ptr(label), ptr(pred), None, ptr(w), label.numel(),
                                                    self.num_classes, 0, 0.0, 0.0, None, ptr(self.weighted_cm()),
                                                    ptr(self._dropped), stream_handle()), "pcls_validation_update")
      return
    _lib.check(_lib.load().pcls_confusion_update(ptr(label), ptr(pred), label.numel(), self.num_classes, ptr(cm),
                                                 ptr(self._dropped), stream_handle()), "pcls_confusion_update")

  def weighted_cm(self):
    """The float64 [NC,NC] matrix of weighted updates (created on first use)."""
    cm = self._ensure()
    if self._cmw is None:
      self._cmw = torch.zeros((self.num_classes, self.num_classes), dtype=torch.float64, device=cm.device)
    return self._cmw

  @property
  def total_cm(self):
    """[NC,NC] CUDA tensor, rows = label, cols = prediction (tf.math.confusion_matrix layout): int64 counts, or
    float64 once weighted updates have been made."""
    cm = self._ensure()
    return cm if self._cmw is None else cm.to(torch.float64) + self._cmw

  @property
  def dropped(self):
    return int(self._dropped.item()) if self._dropped is not None else 0

  def allreduce(self, comm):
    """Sum the matrix over all ranks of ``comm`` (sharding.Communicator), in place, on the current stream: the int64
    counts with ONE ncclAllReduce through the C ABI; the float64 matrix of weighted updates (test_step) and the dropped
    counter, when present on any rank, with ``torch.distributed`` (off the hot path)."""
    comm.allreduce_confusion(self._ensure())
    comm.allreduce_aux(self)

  def result(self):
    """Mean IoU over the classes whose denominator is non-zero (tf.keras.metrics.MeanIoU.result)."""
    cm = self.total_cm.cpu().numpy()
    tp = np.diag(cm).astype(np.float64)
    denom = (cm.sum(0) + cm.sum(1)).astype(np.float64) - tp
    valid = denom != 0
    if not valid.any():
      return np.float32(0.0)
    iou = np.zeros_like(tp)
    np.divide(tp, denom, out=iou, where=valid)
    return np.float32(iou.sum() / valid.sum())

  def reset_states(self):
    if self._cm is not None:
      self._cm.zero_()
      self._dropped.zero_()
    self._cmw = None

  reset_state = reset_states
