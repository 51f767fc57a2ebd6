"""Host side of the drop-in: config contract, registries, model builders' traces, error behaviour (no GPU)."""
import os

import numpy as np
import pytest

from oracle import nn as O
from pclsegmentation_b200 import _lib
from pclsegmentation_b200.configs import (Darknet21, Darknet53, Darknet53Kitti, SqueezeSegV2Config,
                                          SqueezeSegV2ConfigNuScenes, SqueezeSegV2KittiConfig)
from pclsegmentation_b200.laserscan import LaserScan, SemLaserScan
from pclsegmentation_b200.nets.Darknet import rewrite_strides
from pclsegmentation_b200.sharding import shard_range
from pclsegmentation_b200.utils.args_loader import config_map, load_model_config, model_map
from pclsegmentation_b200.utils.util import confusion_matrix_to_iou_recall_precision


def test_config_contract_values():
  rows = [(SqueezeSegV2Config, 32, 240, 11, 10, 32, 24.810, 30.897),
          (SqueezeSegV2KittiConfig, 64, 1024, 20, 0, 64, -0.047, 9.6474),
          (SqueezeSegV2ConfigNuScenes, 32, 1024, 11, 10, 32, -0.1090, 12.9454),
          (Darknet21, 32, 240, 11, 10, 16, 24.810, 30.897),
          (Darknet53, 32, 240, 11, 10, 16, 24.810, 30.897),
          (Darknet53Kitti, 64, 1024, 20, 0, 16, -0.047, 9.6474)]
  for fn, H, W, nc, none, batch, mean0, std4 in rows:
    mc = fn()
    assert (mc.ZENITH_LEVEL, mc.AZIMUTH_LEVEL, mc.NUM_FEATURES, mc.NUM_CLASS) == (H, W, 6, nc)
    assert mc.CLASSES.index("None") == none and mc.BATCH_SIZE == batch
    assert mc.INPUT_MEAN.shape == (1, 1, 5) and mc.INPUT_MEAN.dtype == np.float64
    assert mc.INPUT_MEAN[0, 0, 0] == mean0 and mc.INPUT_STD[0, 0, 4] == std4
    assert mc.CLS_COLOR_MAP.shape == (nc, 3)
  assert Darknet21().NUM_LAYERS == 21 and Darknet53().NUM_LAYERS == 53 and Darknet53Kitti().OUTPUT_STRIDE == 16
  assert SqueezeSegV2ConfigNuScenes().CLS_LOSS_WEIGHT[10] == 0.0
  mc = SqueezeSegV2Config()
  mc.FOO = 3
  assert mc["FOO"] == 3  # attribute-dict behaviour of EasyDict


def test_configs_equal_reference_field_by_field(golden_dir):
  """a11: every field of the six `mc` factories equals the reference's (tests/golden/reference_configs.json, dumped by
  tests/golden/make_config_golden.py from the unmodified pcl_segmentation/configs/*.py): same field set, same values,
  and for arrays the same dtype and shape (INPUT_MEAN / INPUT_STD are float64 [1,1,5])."""
  import json
  import os
  ref = json.load(open(os.path.join(golden_dir, "reference_configs.json")))
  assert set(ref) == set(config_map)
  for key, fields in ref.items():
    mc = config_map[key]()
    assert set(mc.keys()) == set(fields), (key, set(mc.keys()) ^ set(fields))
    for name, want in fields.items():
      got = mc[name]
      if isinstance(want, dict) and set(want) == {"dtype", "shape", "data"}:
        assert isinstance(got, np.ndarray), (key, name)
        assert str(got.dtype) == want["dtype"] and list(got.shape) == want["shape"], (key, name, got.dtype, got.shape)
        assert np.array_equal(got, np.array(want["data"], dtype=got.dtype)), (key, name)
      elif isinstance(want, dict):
        assert {str(k): v for k, v in got.items()} == want, (key, name)
      else:
        assert type(got) is type(want) and got == want, (key, name, got, want)


def test_registries():
  assert set(model_map) == {"squeezesegv2", "darknet53", "darknet21"}
  assert set(config_map) == {"squeezesegv2", "darknet53", "darknet21", "darknet53kitti", "squeezesegv2kitti",
                             "squeezesegv2nuscenes"}
  with pytest.raises(KeyError):
    load_model_config("nope", "squeezesegv2")


def test_squeezesegv2_trace_matches_weight_inventory():
  mc, model = load_model_config("SqueezeSegV2", "squeezesegv2kitti")
  v = model.variables
  assert v["conv1/kernel"].shape == (3, 3, 6, 64) and v["conv1/bias"].shape == (64,)
  assert v["cam1/squeeze/kernel"].shape == (1, 1, 64, 4) and v["cam2/excitation/kernel"].shape == (1, 1, 8, 128)
  assert v["conv1_skip/kernel"].shape == (1, 1, 6, 64) and v["bn1_skip/moving_variance"].shape == (64,)
  fires = {2: (64, 16, 64), 3: (128, 16, 64), 4: (128, 32, 128), 5: (256, 32, 128), 6: (256, 48, 192),
           7: (384, 48, 192), 8: (384, 64, 256), 9: (512, 64, 256), 10: (512, 64, 128), 11: (256, 32, 64),
           12: (128, 16, 32), 13: (64, 16, 32)}
  for n, (cin, s, e) in fires.items():
    assert v[f"fire{n}/squeeze/kernel"].shape == (1, 1, cin, s)
    assert v[f"fire{n}/expand1x1/kernel"].shape == (1, 1, s, e) and v[f"fire{n}/expand3x3/kernel"].shape == (3, 3, s, e)
    assert (f"fire{n}/upconv/kernel" in v) == (n >= 10)
    if n >= 10:
      assert v[f"fire{n}/upconv/kernel"].shape == (1, 4, s, s) and f"fire{n}/upconv_bn/gamma" not in v
  assert v["conv14/kernel"].shape == (3, 3, 64, 20)
  prog = model._graph.program
  # 43 Keras convolutions; the expand1x1 || expand3x3 pairs of fire2-5 and fire10-13 (N <= 256) are merged into one op
  merged = [o for o in prog if o["op"] == "conv" and o.get("merge")]
  assert sum(o["op"] == "conv" for o in prog) == 43 - 8 and len(merged) == 8
  assert sum(o["op"] == "cam" for o in prog) == 3 and sum(o["op"] == "pool" for o in prog) == 3
  for o in merged:
    a, b = o["merge"]
    assert a["kernel"].endswith("expand1x1/kernel") and b["kernel"].endswith("expand3x3/kernel")
    assert o["cout"] == a["cout"] + b["cout"] and (o["kh"], o["kw"]) == (3, 3)
    k = model._graph._merged_arrays(o)[0]
    assert k.shape == (3, 3, a["cin"], o["cout"]) and np.array_equal(k[1, 1, :, :a["cout"]], model.variables[a["kernel"]][0, 0])
    assert not k[0, 0, :, :a["cout"]].any() and np.array_equal(k[..., a["cout"]:], model.variables[b["kernel"]])
  # every tf.add skip was folded into the (merged) expand convolution of its FireDeconv
  for n in (10, 11, 12, 13):
    ops = [o for o in merged if o["merge"][0]["kernel"].startswith(f"fire{n}/")]
    assert len(ops) == 1 and len(ops[0]["res"]) == 1 and ops[0]["off"] == 0
  # fire6-9 (N = 384 / 512) keep two convolutions writing channel slices of the concat tensor
  for n in (6, 7, 8, 9):
    ops = [o for o in prog if o["op"] == "conv" and (o["kernel"] or "").startswith(f"fire{n}/expand")]
    assert len(ops) == 2 and sorted(o["off"] for o in ops) == [0, ops[0]["cout"]]
  skip = [o for o in prog if o["op"] == "conv" and o["kernel"] == "conv1_skip/kernel"][0]
  assert skip["act"] == _lib.ACT_NONE and skip["bn"] == "bn1_skip"


@pytest.mark.parametrize("cfg,layers,nconv", [("darknet21", 21, 36), ("darknet53kitti", 53, 68)])
def test_darknet_trace(cfg, layers, nconv):
  mc, model = load_model_config("darknet%d" % layers, cfg)
  v, prog = model.variables, model._graph.program
  assert len(prog) == nconv and all(o["op"] == "conv" for o in prog)
  assert v["conv1/kernel"].shape == (3, 3, 6, 32) and "conv1/bias" not in v
  assert v["enc5/conv1/kernel"].shape == (3, 3, 512, 1024)
  assert v["enc3/residual_1/conv1/kernel"].shape == (1, 1, 256, 128)
  assert v["dec5/conv1/kernel"].shape == (3, 3, 1024, 512) and v["dec5/conv1/bias"].shape == (512,)
  # decoder quirk: block expands 1x1 out->in, 3x3 in->out
  assert v["dec5/block/conv1/kernel"].shape == (1, 1, 512, 1024) and v["dec5/block/conv2/kernel"].shape == (3, 3, 1024, 512)
  assert v["dec4/upconv1/kernel"].shape == (1, 4, 256, 512) and v["head/kernel"].shape == (3, 3, 32, mc.NUM_CLASS)
  # widths: encoder strides [2,2,2,2,1], decoder [1,2,2,2,2]
  widths = {o["kernel"]: o["dst"].width for o in prog}
  W = mc.AZIMUTH_LEVEL
  assert widths["enc4/conv1/kernel"] == W // 16 and widths["enc5/conv1/kernel"] == W // 16
  assert widths["dec5/conv1/kernel"] == W // 16 and widths["dec1/upconv1/kernel"] == W
  # dec4..dec1: block residual + encoder skip -> two residual adds on the last conv; dec5: one
  for i in (4, 3, 2, 1):
    assert len([o for o in prog if o["kernel"] == f"dec{i}/block/conv2/kernel"][0]["res"]) == 2
  assert len([o for o in prog if o["kernel"] == "dec5/block/conv2/kernel"][0]["res"]) == 1


def test_rewrite_strides_matches_reference_loop():
  for os_ in (8, 16, 32):
    assert rewrite_strides(os_) == O.darknet_strides(os_)
  with pytest.raises(ValueError):
    rewrite_strides(12)


def test_weight_roundtrip(tmp_path):
  mc, model = load_model_config("squeezesegv2", "squeezesegv2")
  model.randomize_batch_norm(5)
  w = model.get_weights_dict()
  p = str(tmp_path / "w.npz")
  model.save_weights_npz(p)
  mc2, model2 = load_model_config("squeezesegv2", "squeezesegv2")
  model2.load_weights_npz(p)
  assert all(np.array_equal(w[k], model2.variables[k]) for k in w)
  with pytest.raises(ValueError):
    model2.set_weights_dict({**w, "conv1/bias": np.zeros(3)})
  with pytest.raises(KeyError):
    model2.set_weights_dict({"conv1/bias": w["conv1/bias"]})
  # TF checkpoint style keys are accepted
  model2.set_weights_dict({k + "/.ATTRIBUTES/VARIABLE_VALUE": v for k, v in w.items()})


def test_laserscan_error_behaviour():
  # same exception types for the same conditions as laserscan_semantic_kitti.py:64-70, 88-93, 249-252
  scan = LaserScan(project=False)
  with pytest.raises(TypeError):
    scan.open_scan(123)
  with pytest.raises(RuntimeError):
    scan.open_scan("scan.txt")
  with pytest.raises(TypeError):
    scan.set_points([[1, 2, 3]])
  with pytest.raises(TypeError):
    scan.set_points(np.zeros((2, 3), np.float32), remissions=[0, 1])
  sem = SemLaserScan(20, {0: [0, 0, 0], 10: [245, 150, 100]}, project=False)
  sem.set_points(np.zeros((3, 3), np.float32))
  assert len(sem) == 3 and sem.proj_range.shape == (64, 1024) and (sem.proj_idx == -1).all()
  with pytest.raises(TypeError):
    sem.set_label([1, 2, 3])
  with pytest.raises(ValueError):
    sem.set_label(np.zeros(5, np.uint32))
  with pytest.raises(RuntimeError):
    sem.open_label("x.bin")
  sem.set_label(np.array([10 | (7 << 16), 0, 1], np.uint32))
  assert sem.sem_label.tolist() == [10, 0, 1] and sem.inst_label.tolist() == [7, 0, 0]


def test_shard_range_partitions():
  for n in (0, 1, 7, 8, 8192, 8193):
    for world in (1, 2, 3, 8):
      parts = [shard_range(n, r, world) for r in range(world)]
      assert parts[0][0] == 0 and parts[-1][1] == n
      assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
      sizes = [b - a for a, b in parts]
      assert max(sizes) - min(sizes) <= 1
  with pytest.raises(ValueError):
    shard_range(4, 2, 2)


def test_iou_helper_matches_oracle():
  from oracle import confusion as C
  rng = np.random.default_rng(0)
  cm = rng.integers(0, 50, (11, 11))
  cm[:, 3] = 0
  cm[3, :] = 0
  a = confusion_matrix_to_iou_recall_precision(cm)
  b = C.iou_recall_precision(cm)
  assert all(np.allclose(x, y) for x, y in zip(a, b))


def test_cli_weight_loading_never_falls_back_silently(tmp_path):
  """--path_to_model must always be loaded when given (a checkpoint PREFIX never exists as a file), a wrong
  path is an error (the reference's tf.keras.models.load_model raises, inference.py:39), and running on random weights
  needs the explicit --random_init."""
  import argparse
  from pclsegmentation_b200.inference import load_model_weights
  mc, model = load_model_config("squeezesegv2", "squeezesegv2")
  model.randomize_batch_norm(3)
  want = model.get_weights_dict()
  prefix = str(tmp_path / "ckpt")
  model.save_weights_bundle(prefix)
  assert not os.path.exists(prefix) and os.path.exists(prefix + ".index")
  _, fresh = load_model_config("squeezesegv2", "squeezesegv2")
  load_model_weights(fresh, argparse.Namespace(path_to_model=prefix, random_init=False))
  assert all(np.array_equal(fresh.variables[k], want[k]) for k in want)
  for bad in (str(tmp_path / "nope"), str(tmp_path / "nope.npz")):
    with pytest.raises(FileNotFoundError):
      load_model_weights(fresh, argparse.Namespace(path_to_model=bad, random_init=False))
  with pytest.raises(SystemExit):
    load_model_weights(fresh, argparse.Namespace(path_to_model=None, random_init=False))
  load_model_weights(fresh, argparse.Namespace(path_to_model=None, random_init=True), verbose=False)
