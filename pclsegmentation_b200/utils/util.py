"""Metric helpers - mirrors the non-plotting part of pcl_segmentation/utils/util.py (plots need matplotlib and are
visualisation, out of the hot path)."""
import numpy as np
import torch


def _to_numpy(cm):
  if torch.is_tensor(cm):
    return cm.detach().cpu().numpy()
  return np.asarray(cm)


def confusion_matrix_to_iou_recall_precision(cm):
  """
  Computes the classwise iou, recall and precision from a confusion matrix (utils/util.py:64-79).
  cm: nxn confusion matrix as kept by ``MeanIoU.total_cm`` (rows = true label, cols = prediction).
  Division by zero yields 0 (tf.math.divide_no_nan).  Returns float64 numpy arrays.
  """
  cm = _to_numpy(cm).astype(np.float64)
  sum_over_col = cm.sum(axis=1)
  sum_over_row = cm.sum(axis=0)
  tp = np.diag(cm)
  fp = sum_over_row - tp
  fn = sum_over_col - tp

  def div(a, b):
    out = np.zeros_like(a)
    np.divide(a, b, out=out, where=b != 0)
    return out

  return div(tp, tp + fp + fn), div(tp, tp + fn), div(tp, tp + fp)


def normalize(x):
  return (x - x.min()) / (x.max() - x.min())
