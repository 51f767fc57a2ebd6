"""batched raw-file ingest (KITTI [N,4] .bin + uint32 .label, nuScenes 5-float .bin + uint8 lidarseg) ->
device projection -> [H,W,6] range images, against the oracle evaluated scan by scan (the converters' per-scan loop:
dataset_convert/semantic_kitti.py:150-179, nu_dataset.py:131-168)."""
import os

import numpy as np
import pytest
import torch

from oracle import projection as P
from tests.util import synth_scan

pytestmark = pytest.mark.gpu


def _write_kitti(tmp, rng, n_scans, H):
  scans, labels, sf, lf = [], [], [], []
  for i in range(n_scans):
    n = int(rng.integers(3000, 9000)) if i != 1 else 1
    s = synth_scan(rng, n, 3.0, -25.0, H)
    lab = (rng.integers(0, 260, n).astype(np.uint32) | (rng.integers(0, 50, n).astype(np.uint32) << 16))
    a, b = os.path.join(tmp, "%06d.bin" % i), os.path.join(tmp, "%06d.label" % i)
    s.tofile(a)
    lab.tofile(b)
    scans.append(s); labels.append(lab); sf.append(a); lf.append(b)
  return scans, labels, sf, lf


def test_kitti_batched_conversion_matches_per_scan_oracle(tmp_path):
  from pclsegmentation_b200.dataset_convert import convert_scans, learning_map_lut
  rng = np.random.default_rng(3)
  H, W = 64, 512
  scans, labels, sf, lf = _write_kitti(str(tmp_path), rng, 7, H)
  mapping = {k: (k * 7) % 20 for k in range(260)}
  lut = learning_map_lut(mapping)
  got = []
  for files, images in convert_scans(sf, lf, "kitti", H, W, 3.0, -25.0, lut, batch=3):     # batches of 3, 3, 1
    got.append(images.cpu().numpy())
  got = np.concatenate(got)
  assert got.shape == (7, H, W, 6)
  for i, (s, lab) in enumerate(zip(scans, labels)):
    o = P.range_projection(s[:, :3], s[:, 3], H, W, 3.0, -25.0, trig="cr")
    ref = P.assemble_range_image(o, P.label_projection(o["proj_idx"], lab), lut).astype(np.float32)
    assert np.array_equal(got[i], ref), i


def test_nuscenes_records_ring_and_fov_projection(tmp_path):
  from pclsegmentation_b200.dataset_convert import ScanBatchLoader, convert_scans, learning_map_lut
  rng = np.random.default_rng(4)
  H, W = 32, 1024
  recs, labs, sf, lf = [], [], [], []
  for i in range(4):
    n = int(rng.integers(2000, 6000))
    s = synth_scan(rng, n, 12.0, -30.0, H)
    ring = rng.integers(0, H, n)
    rec = np.concatenate([s, ring[:, None].astype(np.float32)], 1).astype(np.float32)     # x, y, z, intensity, ring
    lab = rng.integers(0, 32, n).astype(np.uint8)
    a, b = os.path.join(str(tmp_path), "s%d.pcd.bin" % i), os.path.join(str(tmp_path), "s%d_lidarseg.bin" % i)
    rec.tofile(a)
    lab.tofile(b)
    recs.append(rec); labs.append(lab); sf.append(a); lf.append(b)
  lut = learning_map_lut({k: k % 11 for k in range(32)})
  # the loader itself: points / ring split on the device
  data = ScanBatchLoader("nuscenes").load(sf, lf)
  allrec = np.concatenate(recs)
  assert np.array_equal(data["points"].cpu().numpy(), allrec[:, :4])
  assert np.array_equal(data["ring"].cpu().numpy(), allrec[:, 4].astype(np.int32))
  assert np.array_equal(data["labels"].cpu().numpy(), np.concatenate(labs).astype(np.int32))
  # nu_dataset.py flow (use_ring_projection=False: fov 12 / -30) and the ring variant (pcd_dataset.py flow)
  for use_ring in (False, True):
    got = np.concatenate([im.cpu().numpy() for _, im in convert_scans(sf, lf, "nuscenes", H, W, 12.0, -30.0, lut,
                                                                     use_ring_projection=use_ring, batch=4)])
    for i, (rec, lab) in enumerate(zip(recs, labs)):
      if use_ring:
        o = P.range_projection_ring(rec[:, :3], rec[:, 3], rec[:, 4].astype(np.int32), H, W)
      else:
        o = P.range_projection(rec[:, :3], rec[:, 3], H, W, 12.0, -30.0, trig="cr")
      ref = P.assemble_range_image(o, P.label_projection(o["proj_idx"], lab.astype(np.uint32)), lut).astype(np.float32)
      assert np.array_equal(got[i], ref), (use_ring, i)


def test_converter_cli_writes_float64_samples(tmp_path):
  from pclsegmentation_b200 import dataset_convert
  rng = np.random.default_rng(5)
  _, _, sf, lf = _write_kitti(str(tmp_path), rng, 3, 64)
  out = tmp_path / "out"
  dataset_convert.main(["--format", "kitti", "--scans", str(tmp_path / "*.bin"), "--labels", str(tmp_path / "*.label"),
                        "--output_dir", str(out), "--width", "256", "--batch", "2"])
  files = sorted(os.listdir(out))
  assert files == ["0.npy", "1.npy", "2.npy"]
  a = np.load(out / "0.npy")
  assert a.shape == (64, 256, 6) and a.dtype == np.float64          # like the reference's np.save of final_data
  with pytest.raises(RuntimeError):
    list(dataset_convert.convert_scans([str(tmp_path / "x.txt")], None))
