"""Spherical projection of raw LiDAR scans: the reference's ``LaserScan`` / ``SemLaserScan`` call surface over the
CUDA scatter kernels (csrc/projection.cu).

Mirrors

* ``LaserScan``    dataset_convert/laserscan_semantic_kitti.py:5-166 and the ring-index variant of
                   dataset_convert/laserscan_nuscenes.py:71-286 (``use_ring_projection``)
* ``SemLaserScan`` dataset_convert/laserscan_semantic_kitti.py:169-279 (labels; colours are visualisation only)

``SphericalProjector`` is the batched device-side entry the converters / the fused projection+inference
pipeline use: B scans in, ``[B,H,W,6]`` range images out, everything resident in HBM.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .device import ptr, require_cuda, stream_handle, to_device


class SphericalProjector:
  """Batched projection on the current CUDA device.

  project(points, offsets, ...) takes the concatenated ``[total,4]`` float32 points (x,y,z,remission) of B scans
  and the ``[B+1]`` int64 offsets and returns device tensors.
  """

  def __init__(self, H=64, W=1024, fov_up=3.0, fov_down=-25.0, label_lut=None):
    self.H, self.W = int(H), int(W)
    self.fov_up, self.fov_down = fov_up, fov_down
    self._lib = _lib.load()
    self._dev = require_cuda()
    self.label_lut = None
    if label_lut is not None:
      self.label_lut = torch.as_tensor(np.asarray(label_lut, dtype=np.int32)).to(self._dev)

  def project(self, points, offsets, labels=None, ring=None, empty_fill=0.0, want_image=True, want_idx=True,
              want_sem=False, want_point_outputs=False):
    lib, H, W = self._lib, self.H, self.W
    B = int(offsets.numel()) - 1
    total = int(points.shape[0])
    if points.dtype != torch.float32 or points.dim() != 2 or points.shape[1] != 4:
      raise TypeError("points must be a float32 [total,4] CUDA tensor")
    if ring is None and (self.fov_up is None or self.fov_down is None):
      raise NotImplementedError("projection needs either fov_up/fov_down or a ring index")
    dev = points.device
    keys = torch.empty((B, H, W), dtype=torch.int64, device=dev)
    out = {"keys": keys}
    px = py = ur = None
    if want_point_outputs:
      px = torch.empty(total, dtype=torch.int32, device=dev)
      py = torch.empty(total, dtype=torch.int32, device=dev)
      ur = torch.empty(total, dtype=torch.float32, device=dev)
      out.update(proj_x=px, proj_y=py, unproj_range=ur)
    s = stream_handle()
    _lib.check(lib.pcls_project_scatter(ptr(points), ptr(ring), ptr(offsets), B, total, H, W,
                                        float(self.fov_up if self.fov_up is not None else 0.0),
                                        float(self.fov_down if self.fov_down is not None else 0.0),
                                        ptr(keys), ptr(px), ptr(py), ptr(ur), s), "pcls_project_scatter")
    image = idx = sem = None
    if want_image:
      image = torch.empty((B, H, W, 6), dtype=torch.float32, device=dev)
      out["image"] = image
    if want_idx:
      idx = torch.empty((B, H, W), dtype=torch.int32, device=dev)
      out["proj_idx"] = idx
    if want_sem:
      sem = torch.empty((B, H, W), dtype=torch.int32, device=dev)
      out["proj_sem_label"] = sem
    lut = self.label_lut
    _lib.check(lib.pcls_project_resolve(ptr(points), ptr(labels), ptr(offsets), B, H, W, ptr(keys), ptr(lut),
                                        int(lut.numel()) if lut is not None else 0, float(empty_fill), ptr(image),
                                        ptr(idx), ptr(sem), s), "pcls_project_resolve")
    return out

  def project_scans(self, scans, labels=None, rings=None, **kw):
    """Host convenience: list of ``[N_i,4]`` float32 numpy scans (+ optional uint32 label / int32 ring arrays)."""
    lens = [int(s.shape[0]) for s in scans]
    offsets = torch.as_tensor(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).to(self._dev)
    pts = to_device(np.concatenate(scans, axis=0) if scans else np.zeros((0, 4), np.float32), torch.float32)
    lab = rng = None
    if labels is not None:
      # uint32 words travel as int32 bit patterns
      lab = to_device(np.concatenate(labels).astype(np.uint32).view(np.int32), torch.int32)
    if rings is not None:
      rng = to_device(np.concatenate(rings).astype(np.int32), torch.int32)
    return self.project(pts, offsets, labels=lab, ring=rng, **kw)


class LaserScan:
  """Class that contains LaserScan with x,y,z,r (reference API; arithmetic on the GPU)."""
  EXTENSIONS_SCAN = ['.bin']

  def __init__(self, project=False, H=64, W=1024, fov_up=3.0, fov_down=-25.0, use_ring_projection=False):
    self.project = project
    self.proj_H = H
    self.proj_W = W
    self.proj_fov_up = fov_up
    self.proj_fov_down = fov_down
    self.use_ring_projection = use_ring_projection
    self._projector = None
    self.reset()

  def reset(self):
    """ Reset scan members (same initial values as laserscan_semantic_kitti.py:17-48). """
    H, W = self.proj_H, self.proj_W
    self.points = np.zeros((0, 3), dtype=np.float32)
    self.remissions = np.zeros((0, 1), dtype=np.float32)
    self.ring_index = np.zeros((0, 1), dtype=np.int32)
    self.proj_range = np.full((H, W), -1, dtype=np.float32)
    self.unproj_range = np.zeros((0, 1), dtype=np.float32)
    self.proj_xyz = np.full((H, W, 3), -1, dtype=np.float32)
    self.proj_remission = np.full((H, W), -1, dtype=np.float32)
    self.proj_idx = np.full((H, W), -1, dtype=np.int32)
    self.proj_x = np.zeros((0, 1), dtype=np.float32)
    self.proj_y = np.zeros((0, 1), dtype=np.float32)
    self.proj_mask = np.zeros((H, W), dtype=np.int32)
    self._dev_points = None
    self._dev_offsets = None
    self._dev_keys = None

  def size(self):
    return self.points.shape[0]

  def __len__(self):
    return self.size()

  def open_scan(self, filename):
    """ Open raw scan ([N,4] float32 KITTI .bin, laserscan_semantic_kitti.py:57-79) and fill in attributes """
    self.reset()
    if not isinstance(filename, str):
      raise TypeError("Filename should be string type, "
                      "but was {type}".format(type=str(type(filename))))
    if not any(filename.endswith(ext) for ext in self.EXTENSIONS_SCAN):
      raise RuntimeError("Filename extension is not valid scan file.")
    scan = np.fromfile(filename, dtype=np.float32).reshape((-1, 4))
    self.set_points(scan[:, 0:3], scan[:, 3])

  def set_points(self, points, remissions=None, ring_index=None):
    """ Set scan attributes (instead of opening from file) """
    self.reset()
    if not isinstance(points, np.ndarray):
      raise TypeError("Scan should be numpy array")
    if remissions is not None and not isinstance(remissions, np.ndarray):
      raise TypeError("Remissions should be numpy array")
    self.points = points
    if remissions is not None:
      self.remissions = remissions
    else:
      self.remissions = np.zeros((points.shape[0]), dtype=np.float32)
    if ring_index is not None:
      self.ring_index = ring_index
    else:
      self.ring_index = np.zeros((points.shape[0]), dtype=np.int32)
    if self.project:
      if ring_index is not None and self.use_ring_projection:
        self.do_range_projection_ring()
      elif self.proj_fov_up is not None and self.proj_fov_down is not None and not self.use_ring_projection:
        self.do_range_projection()
      else:
        raise NotImplementedError

  # -- projection --------------------------------------------------------------------------------
  def _run(self, ring):
    if self._projector is None or (self._projector.H, self._projector.W) != (self.proj_H, self.proj_W):
      self._projector = SphericalProjector(self.proj_H, self.proj_W, self.proj_fov_up, self.proj_fov_down)
    self._projector.fov_up, self._projector.fov_down = self.proj_fov_up, self.proj_fov_down
    n = self.points.shape[0]
    packed = np.empty((n, 4), dtype=np.float32)
    packed[:, 0:3] = self.points
    packed[:, 3] = np.asarray(self.remissions, dtype=np.float32).reshape(-1)
    dev = self._projector._dev
    pts = to_device(packed, torch.float32)
    offsets = torch.tensor([0, n], dtype=torch.int64, device=dev)
    rng = to_device(np.asarray(ring, dtype=np.int32), torch.int32) if ring is not None else None
    out = self._projector.project(pts, offsets, ring=rng, empty_fill=-1.0, want_point_outputs=True)
    img = out["image"][0].cpu().numpy()
    self.proj_xyz = np.ascontiguousarray(img[:, :, 0:3])
    self.proj_remission = np.ascontiguousarray(img[:, :, 3])
    self.proj_range = np.ascontiguousarray(img[:, :, 4])
    self.proj_idx = out["proj_idx"][0].cpu().numpy()
    self.proj_x = out["proj_x"].cpu().numpy()
    self.unproj_range = out["unproj_range"].cpu().numpy()
    if ring is None:
      self.proj_y = out["proj_y"].cpu().numpy()
    self.proj_mask = (self.proj_idx > 0).astype(np.float32)  # off-by-one kept (laserscan_semantic_kitti.py:166)
    self._dev_points, self._dev_offsets, self._dev_keys = pts, offsets, out["keys"]

  def do_range_projection(self):
    """ Project a pointcloud into a spherical projection image (laserscan_semantic_kitti.py:106-166). """
    self._run(None)

  def do_range_projection_ring(self):
    """ Range projection based on ring index (laserscan_nuscenes.py:191-223). """
    self._run(self.ring_index)


class SemLaserScan(LaserScan):
  """Class that contains LaserScan with x,y,z,r,sem_label,inst_label (colour members are kept for API parity;
  they are visualisation only and are filled on the host)."""
  EXTENSIONS_LABEL = ['.label']

  def __init__(self, nclasses, sem_color_dict=None, project=False, H=64, W=1024, fov_up=3.0, fov_down=-25.0,
               use_ring_projection=False):
    super(SemLaserScan, self).__init__(project, H, W, fov_up, fov_down, use_ring_projection)
    self.reset()
    self.nclasses = nclasses
    sem_color_dict = sem_color_dict or {}
    max_sem_key = max([k + 1 for k in sem_color_dict] + [0])
    self.sem_color_lut = np.zeros((max_sem_key + 100, 3), dtype=np.float32)
    for key, value in sem_color_dict.items():
      self.sem_color_lut[key] = np.array(value, np.float32) / 255.0

  def reset(self):
    super(SemLaserScan, self).reset()
    H, W = self.proj_H, self.proj_W
    self.sem_label = np.zeros((0, 1), dtype=np.uint32)
    self.sem_label_color = np.zeros((0, 3), dtype=np.float32)
    self.inst_label = np.zeros((0, 1), dtype=np.uint32)
    self.proj_sem_label = np.zeros((H, W), dtype=np.int32)
    self.proj_sem_color = np.zeros((H, W, 3), dtype=float)
    self.proj_inst_label = np.zeros((H, W), dtype=np.int32)

  def open_label(self, filename):
    if not isinstance(filename, str):
      raise TypeError("Filename should be string type, "
                      "but was {type}".format(type=str(type(filename))))
    if not any(filename.endswith(ext) for ext in self.EXTENSIONS_LABEL):
      raise RuntimeError("Filename extension is not valid label file.")
    label = np.fromfile(filename, dtype=np.uint32).reshape((-1))
    self.set_label(label)

  def set_label(self, label):
    if not isinstance(label, np.ndarray):
      raise TypeError("Label should be numpy array")
    if label.shape[0] == self.points.shape[0]:
      self.sem_label = label & 0xFFFF
      self.inst_label = label >> 16
    else:
      print("Points shape: ", self.points.shape)
      print("Label shape: ", label.shape)
      raise ValueError("Scan and Label don't contain same number of points")
    if self.project:
      self.do_label_projection()

  def colorize(self):
    self.sem_label_color = self.sem_color_lut[self.sem_label].reshape((-1, 3))

  def do_label_projection(self):
    """laserscan_semantic_kitti.py:269-279: gather labels through the winner index (on the GPU, from the keys the
    range projection left on the device)."""
    if self._dev_keys is None:
      raise RuntimeError("do_label_projection needs a projected scan (set_points with project=True first)")
    H, W = self.proj_H, self.proj_W
    lib = self._projector._lib
    dev = self._projector._dev
    sem = torch.empty((1, H, W), dtype=torch.int32, device=dev)

    def gather(words_np):  # the kernel gathers (word & 0xFFFF) through the winner index
      words = to_device(np.ascontiguousarray(words_np).astype(np.uint32).view(np.int32), torch.int32)
      _lib.check(lib.pcls_project_resolve(ptr(self._dev_points), ptr(words), ptr(self._dev_offsets), 1, H, W,
                                          ptr(self._dev_keys), None, 0, ctypes.c_float(-1.0), None, None, ptr(sem),
                                          stream_handle()), "pcls_project_resolve")
      return sem[0].cpu().numpy()

    self.proj_sem_label = gather(self.sem_label)
    self.proj_inst_label = gather(self.inst_label)  # instance id = upper 16 bits, already shifted down
    mask = self.proj_idx >= 0
    self.proj_sem_color[mask] = self.sem_color_lut[self.proj_sem_label[mask]]
