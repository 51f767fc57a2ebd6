"""``Darknet21`` - ika 32x240 config (reference: pcl_segmentation/configs/Darknet21.py:30-94)."""
from ._tables import IKA_COLORS, IKA_MEAN, IKA_STD, make_config

DARKNET_IKA_CLASSES = ['Road', 'Sidewalk', 'Building', 'Pole', 'Vegetation', 'Person', 'TwoWheeler', 'Car', 'Truck',
                       'Bus', "None"]


def _darknet_ika(num_layers, lr):
  return make_config(classes=DARKNET_IKA_CLASSES, colors=IKA_COLORS, color_dtype="f64", loss_weight=[1.0] * 11,
                     batch=16, height=32, width=240, mean=IKA_MEAN, std=IKA_STD, lr=lr, lr_steps=500,
                     lr_factor=0.99, grad_norm=1.0, drop=0.01, bn_momentum=0.9, num_layers=num_layers,
                     output_stride=16)


def Darknet21():
  return _darknet_ika(21, 0.01)
