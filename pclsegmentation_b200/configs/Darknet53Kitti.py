"""``Darknet53Kitti`` - SemanticKITTI 64x1024 (reference: pcl_segmentation/configs/Darknet53Kitti.py:32-121)."""
from ._tables import KITTI_CLASSES, KITTI_COLORS_BGR, KITTI_MEAN, KITTI_STD, make_config
from .SqueezeSegV2Kitti import rgb


def Darknet53Kitti():
  return make_config(classes=KITTI_CLASSES, colors=[rgb(c) for c in KITTI_COLORS_BGR], batch=16, height=64,
                     width=1024, mean=KITTI_MEAN, std=KITTI_STD, lr=0.001, lr_steps=500, lr_factor=0.99,
                     grad_norm=100.0, drop=0.01, bn_momentum=0.9, num_layers=53, output_stride=16)
