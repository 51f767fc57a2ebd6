"""``SqueezeSegV2ConfigNuScenes`` - nuScenes 32x1024, 11 classes, loss weight of "None" = 0
(reference: pcl_segmentation/configs/SqueezeSegV2NuScenes.py:30-101)."""
from ._tables import IKA_CLASSES, IKA_COLORS, NUSC_MEAN, NUSC_STD, make_config


def SqueezeSegV2ConfigNuScenes():
  return make_config(classes=IKA_CLASSES, colors=IKA_COLORS, loss_weight=[1.0] * 10 + [0.0], batch=32, height=32,
                     width=1024, mean=NUSC_MEAN, std=NUSC_STD, lr=0.003, lr_steps=1000, lr_factor=0.99,
                     grad_norm=100.0, l2=0.05, drop=0.1, bn_momentum=0.99, reduction=16)
