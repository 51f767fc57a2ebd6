"""BASELINE config 1 through the reference-facing CLIs: inference.py over the sample_dataset val frames (golden fixture
tests/golden/nn_sample_dataset.npz) and eval.py's report, plus the fused projection -> forward pipeline (config 4)."""
import os

import numpy as np
import pytest
import torch

from oracle import confusion as C
from oracle import nn as O
from oracle import projection as P
from tests.util import synth_scan

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden(golden_dir):
  with np.load(os.path.join(golden_dir, "nn_sample_dataset.npz")) as f:
    return {k: f[k] for k in f.files}


def _wsum(model):
  return float(sum(np.abs(v.astype(np.float64)).sum() for _, v in sorted(model.variables.items())))


@pytest.mark.parametrize("name,cfg", [("squeezesegv2", "squeezesegv2"), ("darknet21", "darknet21"), ("darknet53", "darknet53")])
def test_inference_cli_on_sample_dataset(tmp_path, golden, name, cfg):
  from pclsegmentation_b200 import inference
  from pclsegmentation_b200.utils.args_loader import load_model_config
  mc, model = load_model_config(name, cfg)
  model.randomize_batch_norm(1)
  assert abs(_wsum(model) - float(golden[name + "_wsum"])) < 1e-6 * float(golden[name + "_wsum"]), \
      "seeded weights differ from the ones the fixture was generated with"
  wpath = str(tmp_path / "weights.npz")
  model.save_weights_npz(wpath)
  indir, outdir = tmp_path / "in", tmp_path / "out"
  indir.mkdir()
  for n, f in zip(golden["names"], golden["frames"]):
    np.save(str(indir / str(n)), f.astype(np.float64))          # the converter writes float64 [H,W,6]
  inference.main(["-d", str(indir / "*.npy"), "-m", name, "-n", cfg, "-t", str(outdir), "-p", wpath, "-b", "2"])
  mask = golden["mask"]
  agree = []
  for i, n in enumerate(golden["names"]):
    stem = os.path.splitext(str(n))[0]
    pred = np.load(str(outdir / ("pred_" + stem + ".npy")))
    assert pred.dtype == np.int32 and pred.shape == (32, 240)
    assert (pred[~mask[i]] == mc.CLASSES.index("None")).all()
    agree.append((pred == golden[name + "_pred"][i])[mask[i]].mean())
    assert os.path.exists(str(outdir / ("plot_" + stem + ".png"))) and os.path.exists(str(outdir / ("plot_gt_" + stem + ".png")))
  assert min(agree) >= 0.998 and np.mean(agree) >= 0.999, agree
  # logits of frame 0
  res = model.forward_device(torch.from_numpy(golden["frames"][:1]).cuda(), None, mean=mc.INPUT_MEAN, std=mc.INPUT_STD,
                             want_logits=True)
  ref = golden[name + "_logits0"]
  err = np.abs(res["logits"][0].cpu().numpy() - ref).max()
  assert err <= 1e-2 * max(1.0, float(np.abs(ref).max()) / 4.0), err


def test_eval_cli_report(tmp_path, golden, capsys):
  from pclsegmentation_b200 import eval as ev
  from pclsegmentation_b200.utils.args_loader import load_model_config
  mc, model = load_model_config("squeezesegv2", "squeezesegv2")
  model.randomize_batch_norm(1)
  wpath = str(tmp_path / "weights.npz")
  model.save_weights_npz(wpath)
  d = tmp_path / "data" / "val"
  d.mkdir(parents=True)
  for n, f in zip(golden["names"], golden["frames"]):
    np.save(str(d / str(n)), f.astype(np.float64))
  rep = ev.main(["-d", str(tmp_path / "data"), "-i", "val", "-m", "squeezesegv2", "-n", "squeezesegv2", "-p", wpath, "-b", "2"])
  out = capsys.readouterr().out
  assert "MIoU:" in out and "ROAD" in out and "NONE" in out
  # the matrix must be exactly the histogram of (fixed-up labels, OUR predictions)
  res = model.forward_device(torch.from_numpy(golden["frames"]).cuda(), None, mean=mc.INPUT_MEAN, std=mc.INPUT_STD,
                             want_probabilities=False)
  cm = C.confusion_matrix(golden["label"].astype(np.int32), res["predictions"].cpu().numpy(), mc.NUM_CLASS)
  assert np.array_equal(rep["confusion_matrix"], cm) and cm.sum() == 3 * 32 * 240
  assert abs(rep["miou"] - C.mean_iou(cm)) < 1e-6
  assert np.allclose(rep["iou"], C.iou_recall_precision(cm)[0])


def test_projection_to_labels_pipeline():
  """Config 4 shape of work on a small case: raw scans -> range images -> Darknet forward, all on the device, equals
  oracle projection -> oracle forward."""
  from pclsegmentation_b200.pipeline import ScanSegmenter
  from pclsegmentation_b200.utils.args_loader import config_map, model_map
  mc = config_map["darknet53kitti"]()
  mc.ZENITH_LEVEL, mc.AZIMUTH_LEVEL, mc.NUM_LAYERS = 64, 256, 21
  model = model_map["darknet21"](mc)
  model.randomize_batch_norm(1)
  rng = np.random.default_rng(9)
  scans = [synth_scan(rng, n) for n in (9000, 7000)]
  seg = ScanSegmenter(model, 3.0, -25.0)
  res = seg.segment(scans, want_probabilities=False)
  none = mc.CLASSES.index("None")
  lid, msk = [], []
  for b, s in enumerate(scans):
    o = P.range_projection(s[:, :3], s[:, 3], 64, 256, 3.0, -25.0, trig="cr")
    img = P.assemble_range_image(o).astype(np.float32)
    assert np.array_equal(res["image"][b].cpu().numpy(), img)
    l, m, _ = O.input_stage(img, mc.INPUT_MEAN, mc.INPUT_STD, none)
    lid.append(l)
    msk.append(m)
  lg, pr, pd = O.forward("darknet", model.variables, np.stack(lid), np.stack(msk), none, num_layers=21, output_stride=16)
  agree = (res["predictions"].cpu().numpy() == pd)[np.stack(msk)].mean()
  assert agree >= 0.999, agree


def test_fused_projection_into_network_input_equals_two_step_path():
  """ScanSegmenter.segment_device default: the resolve pass writes the normalised 16-bit network input in place
  (pcls_project_resolve_net_input) and the forward skips its input kernel - same arithmetic, so logits and labels are
  IDENTICAL to projecting to a float32 [B,H,W,6] image first."""
  from pclsegmentation_b200.pipeline import ScanSegmenter
  from pclsegmentation_b200.utils.args_loader import load_model_config
  mc, model = load_model_config("squeezesegv2", "squeezesegv2kitti")
  mc.AZIMUTH_LEVEL = 512
  mc, model = mc, type(model)(mc)
  model.randomize_batch_norm(2)
  rng = np.random.default_rng(8)
  sizes = [20000, 0, 15000]
  scans = [synth_scan(rng, n) for n in sizes]
  pts = torch.from_numpy(np.concatenate(scans)).cuda()
  offsets = torch.as_tensor(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)).cuda()
  seg = ScanSegmenter(model, 3.0, -25.0)
  two = seg.segment_device(pts, offsets, want_image=True, want_logits=True)
  one = seg.segment_device(pts, offsets, want_logits=True)
  assert "image" not in one
  assert torch.equal(one["proj_idx"], two["proj_idx"])
  assert torch.equal(one["logits"], two["logits"]) and torch.equal(one["predictions"], two["predictions"])
  assert (one["predictions"][1] == mc.CLASSES.index("None")).all()        # the empty scan
