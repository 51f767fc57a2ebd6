"""The constants behind the six ``mc`` factories of the reference (pcl_segmentation/configs/*.py).

These values are part of the call-surface contract: class lists (the index of
"None" drives the head's mask fill), colour maps, input shapes, INPUT_MEAN / INPUT_STD (float64
``[1,1,5]`` arrays), Darknet NUM_LAYERS / OUTPUT_STRIDE.  Training-only fields are kept so that code
written against the reference's ``mc`` objects finds every attribute.
"""
import numpy as np

from .easydict import EasyDict

IKA_CLASSES = ["Road", "Sidewalk", "Building", "Pole", "Vegetation", "Person", "Two-wheeler", "Car", "Truck",
               "Bus", "None"]                                     # configs/SqueezeSegV2.py:33-44
IKA_COLORS = [[128, 64, 128], [244, 35, 232], [70, 70, 70], [153, 153, 153], [107, 142, 35], [220, 20, 60],
              [255, 0, 0], [0, 0, 142], [0, 0, 70], [0, 60, 100], [0, 0, 0]]   # configs/SqueezeSegV2.py:48-59

KITTI_CLASSES = ["None", "car", "bicycle", "motorcycle", "truck", "other-vehicle", "person", "bicyclist",
                 "motorcyclist", "road", "parking", "sidewalk", "other-ground", "building", "fence", "vegetation",
                 "trunk", "terrain", "pole", "traffic-sign"]      # configs/SqueezeSegV2Kitti.py:35-54
# BGR triples as listed at configs/SqueezeSegV2Kitti.py:60-80 (the factory flips them to RGB)
KITTI_COLORS_BGR = [[0, 0, 0], [245, 150, 100], [245, 230, 100], [150, 60, 30], [180, 30, 80], [255, 0, 0],
                    [30, 30, 255], [200, 40, 255], [90, 30, 150], [255, 0, 255], [255, 150, 255], [75, 0, 75],
                    [75, 0, 175], [0, 200, 255], [50, 120, 255], [0, 175, 0], [0, 60, 135], [80, 240, 150],
                    [150, 240, 255], [0, 0, 255]]

IKA_MEAN, IKA_STD = [24.810, 0.819, 0.000, 16.303, 25.436], [30.335, 7.807, 2.058, 25.208, 30.897]
KITTI_MEAN, KITTI_STD = [-0.047, 0.365, -0.855, 0.2198, 8.3568], [10.154, 7.627, 0.8651, 0.1764, 9.6474]
NUSC_MEAN, NUSC_STD = [-0.1090, -0.1645, -0.6275, 17.2574, 11.5727], [11.4001, 12.9684, 1.9548, 20.2257, 12.9454]


def _color_map_f32(colors):
  cmap = np.zeros((len(colors), 3), dtype=np.float32)
  for i, c in enumerate(colors):
    cmap[i] = np.array(c, np.float32) / 255.0
  return cmap


def make_config(*, classes, colors, color_dtype="f32", loss_weight=None, batch, height, width, mean, std,
                lr, lr_steps, lr_factor, grad_norm, drop, bn_momentum, l2=None, reduction=None,
                num_layers=None, output_stride=None):
  mc = EasyDict()
  mc.CLASSES = list(classes)
  mc.NUM_CLASS = len(mc.CLASSES)
  mc.CLS_2_ID = dict(zip(mc.CLASSES, range(len(mc.CLASSES))))
  mc.CLS_LOSS_WEIGHT = np.ones(mc.NUM_CLASS) if loss_weight is None else np.array(loss_weight, dtype=np.float64)
  mc.CLS_COLOR_MAP = _color_map_f32(colors) if color_dtype == "f32" else np.array(colors) / 255.0
  mc.BATCH_SIZE, mc.AZIMUTH_LEVEL, mc.ZENITH_LEVEL, mc.NUM_FEATURES = batch, width, height, 6
  mc.USE_FOCAL_LOSS, mc.FOCAL_GAMMA, mc.CLS_LOSS_COEF, mc.DENOM_EPSILON = False, 2.0, 15.0, 1e-12
  mc.LEARNING_RATE, mc.LR_DECAY_STEPS, mc.LR_DECAY_FACTOR, mc.MAX_GRAD_NORM = lr, lr_steps, lr_factor, grad_norm
  if l2 is not None:
    mc.L2_WEIGHT_DECAY = l2
  mc.DROP_RATE, mc.BN_MOMENTUM = drop, bn_momentum
  if reduction is not None:
    mc.REDUCTION = reduction
  if num_layers is not None:
    mc.NUM_LAYERS, mc.OUTPUT_STRIDE = num_layers, output_stride
  mc.DATA_AUGMENTATION, mc.RANDOM_FLIPPING, mc.SHIFT_UP_DOWN, mc.SHIFT_LEFT_RIGHT = True, True, 0, 70
  mc.INPUT_MEAN = np.array([[list(mean)]])   # float64 [1,1,5], like the reference
  mc.INPUT_STD = np.array([[list(std)]])
  return mc
