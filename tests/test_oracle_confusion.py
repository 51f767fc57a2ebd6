import numpy as np

from oracle import confusion as C


def test_confusion_layout_and_iou():
  label = np.array([0, 0, 1, 2, 2, 2])
  pred = np.array([0, 1, 1, 2, 2, 0])
  cm = C.confusion_matrix(label, pred, 4)
  assert cm.dtype == np.int64 and cm.sum() == 6
  assert cm[0, 1] == 1 and cm[2, 0] == 1 and cm[2, 2] == 2      # rows = label, cols = prediction
  iou, rec, prec = C.iou_recall_precision(cm)
  assert np.allclose(iou[:3], [1 / 3, 1 / 2, 2 / 3]) and iou[3] == 0  # class 3 absent -> divide_no_nan -> 0
  assert np.allclose(rec[:3], [1 / 2, 1, 2 / 3]) and np.allclose(prec[:3], [1 / 2, 1 / 2, 1])
  assert abs(C.mean_iou(cm) - np.mean([1 / 3, 1 / 2, 2 / 3])) < 1e-12  # absent class excluded from the mean
  assert C.mean_iou(np.zeros((3, 3), np.int64)) == 0.0
