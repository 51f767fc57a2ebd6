"""Model / config registries - mirrors pcl_segmentation/utils/args_loader.py:36-55."""
from ..nets.Darknet import Darknet
from ..nets.SqueezeSegV2 import SqueezeSegV2

from ..configs.SqueezeSegV2 import SqueezeSegV2Config
from ..configs.SqueezeSegV2Kitti import SqueezeSegV2KittiConfig
from ..configs.SqueezeSegV2NuScenes import SqueezeSegV2ConfigNuScenes
from ..configs.Darknet53 import Darknet53
from ..configs.Darknet21 import Darknet21
from ..configs.Darknet53Kitti import Darknet53Kitti

model_map = {"squeezesegv2": SqueezeSegV2, "darknet53": Darknet, "darknet21": Darknet}

config_map = {
  "squeezesegv2": SqueezeSegV2Config,
  "darknet53": Darknet53,
  "darknet21": Darknet21,
  "darknet53kitti": Darknet53Kitti,
  "squeezesegv2kitti": SqueezeSegV2KittiConfig,
  "squeezesegv2nuscenes": SqueezeSegV2ConfigNuScenes,
}


def load_model_config(model_name, config_name):
  """-> (config, model), like the reference; the model is built for mc.ZENITH_LEVEL x mc.AZIMUTH_LEVEL."""
  config = config_map[config_name.lower()]()
  model = model_map[model_name.lower()](config)
  return config, model
