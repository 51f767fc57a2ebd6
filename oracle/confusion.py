"""CPU restatement of the confusion-matrix / IoU accumulation (TEST INFRASTRUCTURE ONLY).

Follows ``tf.keras.metrics.MeanIoU.update_state / total_cm / result`` as used at
pcl_segmentation/eval.py:41,48,50,58 and nets/SegmentationNetwork.py:52,113,129, and
``confusion_matrix_to_iou_recall_precision`` pcl_segmentation/utils/util.py:64-79.

PARITY UNPINNED (TF 2.9.1 is a third-party dependency that is absent here; the reference has
no tests for it).  TF accumulates in float32; this oracle and the kernel use int64, equal to TF
wherever TF itself is exact (every cell < 2^24).
"""
import numpy as np


def confusion_matrix(label, pred, num_classes):
    """tf.math.confusion_matrix(labels, predictions): rows = label, cols = prediction."""
    label = np.asarray(label).astype(np.int64).ravel()
    pred = np.asarray(pred).astype(np.int64).ravel()
    return np.bincount(label * num_classes + pred, minlength=num_classes * num_classes
                       ).astype(np.int64).reshape(num_classes, num_classes)


def _div_no_nan(a, b):
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    out = np.zeros_like(a)
    np.divide(a, b, out=out, where=b != 0)
    return out


def iou_recall_precision(cm):
    """utils/util.py:64-79."""
    cm = np.asarray(cm)
    sum_over_col = cm.sum(axis=1)
    sum_over_row = cm.sum(axis=0)
    tp = np.diag(cm)
    fp = sum_over_row - tp
    fn = sum_over_col - tp
    return _div_no_nan(tp, tp + fp + fn), _div_no_nan(tp, tp + fn), _div_no_nan(tp, tp + fp)


def mean_iou(cm):
    """tf.keras.metrics.MeanIoU.result(): mean IoU over classes whose denominator is non-zero."""
    cm = np.asarray(cm)
    tp = np.diag(cm).astype(np.float64)
    denom = (cm.sum(axis=0) + cm.sum(axis=1)).astype(np.float64) - tp
    valid = denom != 0
    if not valid.any():
        return 0.0
    iou = np.zeros_like(tp)
    np.divide(tp, denom, out=iou, where=valid)
    return float(iou.sum() / valid.sum())
