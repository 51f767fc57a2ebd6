"""``SqueezeSegV2Config`` - ika 32x240 config (reference: pcl_segmentation/configs/SqueezeSegV2.py:30-99)."""
from ._tables import IKA_CLASSES, IKA_COLORS, IKA_MEAN, IKA_STD, make_config


def SqueezeSegV2Config():
  return make_config(classes=IKA_CLASSES, colors=IKA_COLORS, loss_weight=[1.0] * 11, batch=32, height=32, width=240,
                     mean=IKA_MEAN, std=IKA_STD, lr=0.003, lr_steps=1000, lr_factor=0.97, grad_norm=100.0,
                     l2=0.05, drop=0.1, bn_momentum=0.99, reduction=16)
