"""Generates tests/golden/projection_*.npz by running the REFERENCE's own projection classes.

Run in the build container only (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_projection_golden.py

Imports LaserScan / SemLaserScan from /root/reference/dataset_convert/laserscan_semantic_kitti.py unmodified
(np.float shim for the removed numpy alias, :211,:217) and stores, for a few seeded synthetic scans, the inputs and
every attribute the reference produces, plus the converter assembly (dataset_convert/semantic_kitti.py:162-173)
evaluated with the reference's own statements.
"""
import os
import sys

import numpy as np
import yaml

REF = "/root/reference/dataset_convert"
sys.path.insert(0, REF)
np.float = float  # noqa: removed alias used by SemLaserScan.reset
import laserscan_semantic_kitti as ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CFG = yaml.safe_load(open(os.path.join(REF, "semantic-kitti.yaml")))


def synth_scan(rng, n, fov_up, fov_down, rings, quantum=0.002):
  """ring-structured synthetic scan, 2 mm range quantisation (realistic depth ties)."""
  k = rng.integers(0, rings, n)
  pitch = np.deg2rad(fov_down + (fov_up - fov_down) * (k + rng.random(n)) / rings)
  yaw = rng.uniform(-np.pi, np.pi, n)
  r = np.round(rng.uniform(2, 80, n) / quantum) * quantum
  pts = np.stack([r * np.cos(pitch) * np.cos(yaw), r * np.cos(pitch) * np.sin(yaw), r * np.sin(pitch)], 1)
  rem = rng.random(n)
  keys = np.array(sorted(CFG["learning_map"].keys()), dtype=np.uint32)
  sem = keys[rng.integers(0, len(keys), n)]
  inst = rng.integers(0, 500, n).astype(np.uint32)
  return pts.astype(np.float32), rem.astype(np.float32), (sem | (inst << 16)).astype(np.uint32)


def run_case(name, seed, n, H, W, fov_up, fov_down, dup=0, out_of_fov=0):
  rng = np.random.default_rng(seed)
  pts, rem, lab = synth_scan(rng, n, fov_up, fov_down, H)
  if dup:  # exact duplicates -> exact depth ties inside a pixel
    src = rng.integers(0, n, dup)
    dst = rng.integers(0, n, dup)
    pts[dst] = pts[src]
  if out_of_fov:  # points above / below the FOV are clamped into the first / last row, not dropped
    sel = rng.integers(0, n, out_of_fov)
    pts[sel, 2] += rng.choice([-1.0, 1.0], out_of_fov).astype(np.float32) * 60.0
  scan = ref.SemLaserScan(len(CFG["color_map"]), CFG["color_map"], project=True, H=H, W=W, fov_up=fov_up,
                          fov_down=fov_down)
  scan.set_points(pts, rem)
  scan.set_label(lab)
  out = dict(points=pts, remissions=rem, label=lab, H=H, W=W, fov_up=fov_up, fov_down=fov_down,
             proj_range=scan.proj_range.copy(), proj_xyz=scan.proj_xyz.copy(),
             proj_remission=scan.proj_remission.copy(), proj_idx=scan.proj_idx.copy(), proj_x=scan.proj_x.copy(),
             proj_y=scan.proj_y.copy(), unproj_range=scan.unproj_range.copy(),
             proj_sem_label=scan.proj_sem_label.copy(), proj_inst_label=scan.proj_inst_label.copy(),
             proj_mask=scan.proj_mask.copy())
  # converter assembly, the reference's statements (semantic_kitti.py:162-173)
  vfunc = np.vectorize(CFG["learning_map"].get)
  mask = scan.proj_range > 0
  scan.proj_range[~mask] = 0.0
  scan.proj_xyz[~mask] = 0.0
  scan.proj_remission[~mask] = 0.0
  sem = vfunc(scan.proj_sem_label)
  final = np.concatenate([scan.proj_xyz, scan.proj_remission.reshape((H, W, 1)), scan.proj_range.reshape((H, W, 1)),
                          sem.reshape((H, W, 1))], axis=2)
  out["final_data"] = final.astype(np.float32)  # exactly representable
  np.savez_compressed(os.path.join(HERE, "projection_%s.npz" % name), **out)
  print(name, "n", n, "valid px", int(mask.sum()), "final dtype", final.dtype)


def run_ring_case(name, seed, n, H, W):
  """The nuScenes variant: `laserscan_nuscenes.LaserScan` imported UNMODIFIED (its top-level `from nuscenes...` imports,
  :3-4, are satisfied by empty stand-in modules: the devkit classes are only used by open_scan, not by set_points /
  do_range_projection_ring :191-223 / do_range_projection :226-286)."""
  import types
  for mod, attrs in (("nuscenes", {}), ("nuscenes.utils", {}),
                     ("nuscenes.utils.data_classes", {"PointCloud": type("PointCloud", (), {})}),
                     ("nuscenes.utils.data_io", {"load_bin_file": lambda path: None})):
    m = types.ModuleType(mod)
    m.__dict__.update(attrs)
    sys.modules.setdefault(mod, m)
  import laserscan_nuscenes as refn
  rng = np.random.default_rng(seed)
  pts, rem, _ = synth_scan(rng, n, 12.0, -30.0, H)
  ring = rng.integers(0, H, n).astype(np.int32)       # several points per (ring, column): the last one written wins
  src, dst = rng.integers(0, n, 200), rng.integers(0, n, 200)
  pts[dst] = pts[src]
  scan = refn.LaserScan(project=True, H=H, W=W, fov_up=12.0, fov_down=-30.0, use_ring_projection=True)
  scan.set_points(pts, rem, ring)
  out = dict(points=pts, remissions=rem, ring=ring, H=H, W=W, proj_range=scan.proj_range.copy(),
             proj_xyz=scan.proj_xyz.copy(), proj_remission=scan.proj_remission.copy(), proj_idx=scan.proj_idx.copy(),
             proj_x=scan.proj_x.copy(), proj_mask=scan.proj_mask.copy())
  # the same class with use_ring_projection=False runs its copy of do_range_projection (:226-286)
  scan2 = refn.LaserScan(project=True, H=H, W=W, fov_up=12.0, fov_down=-30.0, use_ring_projection=False)
  scan2.set_points(pts, rem)
  out.update(fov_proj_range=scan2.proj_range.copy(), fov_proj_idx=scan2.proj_idx.copy(), fov_proj_x=scan2.proj_x.copy(),
             fov_proj_y=scan2.proj_y.copy(), fov_unproj_range=scan2.unproj_range.copy())
  np.savez_compressed(os.path.join(HERE, "projection_%s.npz" % name), **out)
  print(name, "n", n, "occupied px", int((scan.proj_idx >= 0).sum()))


if __name__ == "__main__":
  run_ring_case("nusc_ring_32x1024", 14, 30000, 32, 1024)
  run_case("kitti_64x512", 11, 20000, 64, 512, 3.0, -25.0, dup=300, out_of_fov=50)
  run_case("kitti_64x2048", 12, 40000, 64, 2048, 3.0, -25.0, dup=100)
  run_case("nusc_32x1024", 13, 12000, 32, 1024, 12.0, -30.0, dup=100, out_of_fov=20)
  lm = CFG["learning_map"]
  np.savez(os.path.join(HERE, "semantic_kitti_learning_map.npz"), keys=np.array(list(lm.keys()), np.int32),
           values=np.array(list(lm.values()), np.int32))
