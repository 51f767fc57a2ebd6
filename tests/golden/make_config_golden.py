"""Dumps every field of the reference's six `mc` factories (pcl_segmentation/configs/*.py) to
tests/golden/reference_configs.json.  Run in the build container only (needs /root/reference):

    python tests/golden/make_config_golden.py

The reference files are imported UNMODIFIED; `easydict` (absent from the image) is satisfied by the repo's
attribute-dict stand-in.  numpy arrays are stored as {"dtype", "shape", "data"} so that dtype and shape are pinned too.
"""
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from pclsegmentation_b200.configs import easydict as shim  # noqa: E402

sys.modules.setdefault("easydict", shim)
sys.path.insert(0, "/root/reference/pcl_segmentation")

FACTORIES = {  # registry key (utils/args_loader.py:43-49) -> (module, factory)
  "squeezesegv2": ("configs.SqueezeSegV2", "SqueezeSegV2Config"),
  "squeezesegv2kitti": ("configs.SqueezeSegV2Kitti", "SqueezeSegV2KittiConfig"),
  "squeezesegv2nuscenes": ("configs.SqueezeSegV2NuScenes", "SqueezeSegV2ConfigNuScenes"),
  "darknet21": ("configs.Darknet21", "Darknet21"),
  "darknet53": ("configs.Darknet53", "Darknet53"),
  "darknet53kitti": ("configs.Darknet53Kitti", "Darknet53Kitti"),
}


def encode(v):
  if isinstance(v, np.ndarray):
    return {"dtype": str(v.dtype), "shape": list(v.shape), "data": v.tolist()}
  if isinstance(v, (np.floating, np.integer)):
    return v.item()
  if isinstance(v, dict):
    return {str(k): encode(x) for k, x in v.items()}
  if isinstance(v, (list, tuple)):
    return [encode(x) for x in v]
  return v


if __name__ == "__main__":
  out = {}
  for key, (mod, fn) in FACTORIES.items():
    mc = getattr(importlib.import_module(mod), fn)()
    out[key] = {k: encode(v) for k, v in mc.items()}
    print(key, len(out[key]), "fields")
  with open(os.path.join(HERE, "reference_configs.json"), "w") as f:
    json.dump(out, f, indent=0, sort_keys=True)
