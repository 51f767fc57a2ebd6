"""``mc`` config factories with the reference's names (pcl_segmentation/configs/__init__.py exports only
``SqueezeSegV2Config``; the others are imported by module path in utils/args_loader.py:27-32)."""
from .SqueezeSegV2 import SqueezeSegV2Config
from .SqueezeSegV2Kitti import SqueezeSegV2KittiConfig
from .SqueezeSegV2NuScenes import SqueezeSegV2ConfigNuScenes
from .Darknet21 import Darknet21
from .Darknet53 import Darknet53
from .Darknet53Kitti import Darknet53Kitti
