"""``Darknet53`` - ika 32x240 config (reference: pcl_segmentation/configs/Darknet53.py:30-94)."""
from .Darknet21 import _darknet_ika


def Darknet53():
  return _darknet_ika(53, 0.005)
