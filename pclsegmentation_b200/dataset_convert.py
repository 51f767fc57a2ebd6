"""Batched scan ingest + conversion: raw ``.bin`` / ``.label`` files -> ``[H,W,6]`` range images on the GPU.

Mirrors the per-scan loops of the reference's converters - ``dataset_convert/semantic_kitti.py:150-179`` (KITTI:
``SemLaserScan.open_scan`` + ``open_label``, projection, ``mask = proj_range > 0``, zero-fill, ``learning_map`` through
``np.vectorize``, ``np.concatenate`` -> ``np.save``) and ``dataset_convert/nu_dataset.py:131-168`` (nuScenes: 5-float
records, uint8 lidarseg labels, ``label_map``) - but for a BATCH of scans at a time:

* the files of a batch are read into ONE pinned host buffer (``np.fromfile`` straight into slices of it) and uploaded with
  one asynchronous copy; nuScenes records (x,y,z,intensity,ring) are split into points / ring index on the device
  (pcls_unpack_xyzir);
* projection, label gather, learning-map LUT and the ``[B,H,W,6]`` assembly run in the two projection kernels
  (pcls_project_scatter / pcls_project_resolve); nothing but the finished images returns to the host.

Dataset walking (KITTI ``sequences`` layout, the nuScenes devkit tables) stays outside: the functions take file lists.

    python -m pclsegmentation_b200.dataset_convert --format kitti --scans 'seq/08/velodyne/*.bin' \\
        --labels 'seq/08/labels/*.label' --learning_map semantic-kitti.yaml --output_dir out/val
"""
import argparse
import glob
import os

import numpy as np
import torch

from . import _lib
from .device import ptr, require_cuda, stream_handle
from .laserscan import SphericalProjector

KITTI, NUSCENES = "kitti", "nuscenes"


def learning_map_lut(mapping):
  """dict raw label -> class id (semantic-kitti.yaml ``learning_map`` :109-143, nu_dataset.py ``label_map`` :48-102) as a
  dense int32 LUT; ids the dict does not know map to 0."""
  size = max(int(k) for k in mapping) + 1
  lut = np.zeros(size, np.int32)
  for k, v in mapping.items():
    lut[int(k)] = int(v)
  return lut


class ScanBatchLoader:
  """Reads scan (+ label) files of one batch into pinned staging buffers and uploads them with one copy each."""

  def __init__(self, fmt=KITTI, max_points_per_batch=1 << 23):
    if fmt not in (KITTI, NUSCENES):
      raise ValueError("format must be 'kitti' or 'nuscenes'")
    self.fmt = fmt
    self.rec = 4 if fmt == KITTI else 5           # floats per point record (laserscan_semantic_kitti.py:73-74 / laserscan_nuscenes.py:27-28)
    self.cap = int(max_points_per_batch)
    self._dev = require_cuda()
    self._pin_pts = torch.empty((self.cap, self.rec), dtype=torch.float32, pin_memory=True)
    self._pin_lab = torch.empty(self.cap, dtype=torch.int32, pin_memory=True)
    self._busy = None

  def load(self, scan_files, label_files=None):
    """-> dict(points [total,4] f32 CUDA, offsets [B+1] i64 CUDA, ring [total] i32 CUDA or None, labels [total] i32 CUDA
    (uint32 label words / uint8 lidarseg ids as int32 bit patterns) or None, sizes list)."""
    for f in scan_files:
      if not isinstance(f, str):
        raise TypeError("Filename should be string type, but was {type}".format(type=str(type(f))))
      if not f.endswith(".bin"):
        raise RuntimeError("Filename extension is not valid scan file.")
    if self._busy is not None:
      self._busy.synchronize()                      # the previous batch's upload has left the staging buffers
    sizes = [os.path.getsize(f) // (4 * self.rec) for f in scan_files]
    total = int(sum(sizes))
    if total > self.cap:
      raise ValueError("batch holds %d points, staging capacity is %d" % (total, self.cap))
    host = self._pin_pts.numpy()
    pos = 0
    for f, n in zip(scan_files, sizes):
      host[pos:pos + n] = np.fromfile(f, dtype=np.float32, count=n * self.rec).reshape(n, self.rec)
      pos += n
    lab_host = None
    if label_files is not None:
      lab_host = self._pin_lab.numpy()
      pos = 0
      for f, n in zip(label_files, sizes):
        if self.fmt == KITTI:                       # uint32 words, sem = lower 16 bits (laserscan_semantic_kitti.py:232-246)
          words = np.fromfile(f, dtype=np.uint32)
        else:                                       # nuScenes lidarseg: one uint8 class id per point (load_bin_file)
          words = np.fromfile(f, dtype=np.uint8).astype(np.uint32)
        if words.shape[0] != n:
          raise ValueError("Scan and Label don't contain same number of points")
        lab_host[pos:pos + n] = words.view(np.int32) if words.dtype == np.uint32 else words
        pos += n
    dev = self._dev
    raw = self._pin_pts[:total].to(dev, non_blocking=True)
    labels = self._pin_lab[:total].to(dev, non_blocking=True) if lab_host is not None else None
    self._busy = torch.cuda.Event()
    self._busy.record(torch.cuda.current_stream(dev))
    ring = None
    if self.fmt == NUSCENES:
      points = torch.empty((total, 4), dtype=torch.float32, device=dev)
      ring = torch.empty(total, dtype=torch.int32, device=dev)
      _lib.check(_lib.load().pcls_unpack_xyzir(ptr(raw), total, ptr(points), ptr(ring), stream_handle()), "pcls_unpack_xyzir")
    else:
      points = raw
    offsets = torch.as_tensor(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)).to(dev)
    return dict(points=points, offsets=offsets, ring=ring, labels=labels, sizes=sizes)


def convert_scans(scan_files, label_files=None, fmt=KITTI, H=64, W=1024, fov_up=3.0, fov_down=-25.0, label_lut=None,
                  use_ring_projection=False, batch=64, loader=None, projector=None):
  """Generator over batches: yields (files of the batch, images [b,H,W,6] float32 CUDA tensor) - the converters'
  ``final_data`` (x, y, z, remission, range, mapped label; zeros where no point was projected)."""
  loader = loader or ScanBatchLoader(fmt)
  projector = projector or SphericalProjector(H, W, fov_up, fov_down, label_lut=label_lut)
  for i in range(0, len(scan_files), batch):
    sf = scan_files[i:i + batch]
    lf = label_files[i:i + batch] if label_files is not None else None
    data = loader.load(sf, lf)
    out = projector.project(data["points"], data["offsets"], labels=data["labels"],
                            ring=data["ring"] if use_ring_projection else None, empty_fill=0.0, want_idx=False)
    yield sf, out["image"]


def main(argv=None):
  ap = argparse.ArgumentParser(description="raw scans (+ labels) -> [H,W,6] range-image .npy files, batched on the GPU")
  ap.add_argument("--format", choices=[KITTI, NUSCENES], default=KITTI)
  ap.add_argument("--scans", required=True, help="glob of .bin scan files")
  ap.add_argument("--labels", default=None, help="glob of .label (KITTI) / lidarseg .bin (nuScenes) files, same order")
  ap.add_argument("--learning_map", default=None, help="yaml with a `learning_map` dict (semantic-kitti.yaml) to reduce the raw labels")
  ap.add_argument("--output_dir", "-p", required=True)
  ap.add_argument("--height", type=int, default=None)
  ap.add_argument("--width", type=int, default=1024)
  ap.add_argument("--fov_up", type=float, default=None)
  ap.add_argument("--fov_down", type=float, default=None)
  ap.add_argument("--ring", action="store_true", help="nuScenes: rows from the ring index (do_range_projection_ring)")
  ap.add_argument("--batch", type=int, default=64)
  a = ap.parse_args(argv)
  kitti = a.format == KITTI
  H = a.height or (64 if kitti else 32)
  fu = a.fov_up if a.fov_up is not None else (3.0 if kitti else 12.0)      # laserscan defaults / nu_dataset.py:134
  fd = a.fov_down if a.fov_down is not None else (-25.0 if kitti else -30.0)
  scans = sorted(glob.glob(a.scans))
  labels = sorted(glob.glob(a.labels)) if a.labels else None
  if labels is not None and len(labels) != len(scans):
    raise SystemExit("%d scans but %d label files" % (len(scans), len(labels)))
  lut = None
  if a.learning_map:
    import yaml
    lut = learning_map_lut(yaml.safe_load(open(a.learning_map))["learning_map"])
  os.makedirs(a.output_dir, exist_ok=True)
  index = 0
  for files, images in convert_scans(scans, labels, a.format, H, a.width, fu, fd, lut, a.ring, a.batch):
    host = images.cpu().numpy().astype(np.float64)      # the reference saves float64 (int64 labels promote the concat)
    for k in range(len(files)):
      np.save(os.path.join(a.output_dir, str(index)), host[k])
      index += 1
  print("converted %d scans -> %s" % (index, a.output_dir))


if __name__ == "__main__":
  main()
