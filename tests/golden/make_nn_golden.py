"""Generates tests/golden/nn_sample_dataset.npz (BASELINE config 1: the three ``dataset_samples/sample_dataset/val``
frames through SqueezeSegV2 / Darknet21 / Darknet53 with seeded weights).

Run in the build container only (needs /root/reference for the fixture frames):
    python tests/golden/make_nn_golden.py

TensorFlow is not installable here, so the expected outputs come from the CPU oracle (oracle/nn.py), NOT from the
reference's own arithmetic: this fixture pins the data format and guards against regressions of oracle and kernels;
NN parity with TF itself stays UNPINNED (see oracle/__init__.py).
"""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import nn as O  # noqa: E402
from pclsegmentation_b200.utils.args_loader import load_model_config  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def weight_checksum(model):
  return float(sum(np.abs(v.astype(np.float64)).sum() for _, v in sorted(model.variables.items())))


if __name__ == "__main__":
  files = sorted(glob.glob("/root/reference/dataset_samples/sample_dataset/val/*.npy"))
  frames = np.stack([np.load(f) for f in files])
  assert frames.dtype == np.float64 and frames.shape[1:] == (32, 240, 6)
  out = dict(frames=frames.astype(np.float32), names=np.array([os.path.basename(f) for f in files]))
  for name, cfg in (("squeezesegv2", "squeezesegv2"), ("darknet21", "darknet21"), ("darknet53", "darknet53")):
    mc, model = load_model_config(name, cfg)
    model.randomize_batch_norm(1)
    none = mc.CLASSES.index("None")
    lid, msk, lab = zip(*[O.input_stage(f, mc.INPUT_MEAN, mc.INPUT_STD, none) for f in frames])
    arch = "squeezesegv2" if name == "squeezesegv2" else "darknet"
    lg, pr, pd = O.forward(arch, model.variables, np.stack(lid), np.stack(msk), none,
                           num_layers=getattr(mc, "NUM_LAYERS", 53), output_stride=getattr(mc, "OUTPUT_STRIDE", 16))
    out[name + "_pred"] = pd.astype(np.int8)
    out[name + "_logits0"] = lg[0].astype(np.float32)
    out[name + "_wsum"] = np.float64(weight_checksum(model))
    out["label"] = np.stack(lab).astype(np.int8)
    out["mask"] = np.stack(msk)
    print(name, "valid", np.stack(msk).mean(), "logit absmax", np.abs(lg).max(), "wsum", out[name + "_wsum"])
  np.savez_compressed(os.path.join(HERE, "nn_sample_dataset.npz"), **out)
