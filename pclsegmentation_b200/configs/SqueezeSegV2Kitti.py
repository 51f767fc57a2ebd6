"""``SqueezeSegV2KittiConfig`` - SemanticKITTI 64x1024, 20 classes, "None" = 0
(reference: pcl_segmentation/configs/SqueezeSegV2Kitti.py:32-120)."""
from ._tables import KITTI_CLASSES, KITTI_COLORS_BGR, KITTI_MEAN, KITTI_STD, make_config


def rgb(bgr):
  return [bgr[2], bgr[1], bgr[0]]


def SqueezeSegV2KittiConfig():
  return make_config(classes=KITTI_CLASSES, colors=[rgb(c) for c in KITTI_COLORS_BGR], batch=64, height=64,
                     width=1024, mean=KITTI_MEAN, std=KITTI_STD, lr=0.001, lr_steps=500, lr_factor=0.99,
                     grad_norm=100.0, l2=0.05, drop=0.1, bn_momentum=0.9, reduction=16)
