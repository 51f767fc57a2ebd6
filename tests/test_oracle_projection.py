"""The projection oracle is pinned against the reference's own outputs (tests/golden/projection_*.npz, produced by
tests/golden/make_projection_golden.py from dataset_convert/laserscan_semantic_kitti.py)."""
import os

import numpy as np
import pytest

from oracle import projection as P

CASES = ["kitti_64x512", "kitti_64x2048", "nusc_32x1024"]


def load(golden_dir, name):
  with np.load(os.path.join(golden_dir, "projection_%s.npz" % name)) as f:
    return {k: f[k] for k in f.files}


def lut(golden_dir):
  with np.load(os.path.join(golden_dir, "semantic_kitti_learning_map.npz")) as f:
    return P.learning_map_lut(dict(zip(f["keys"].tolist(), f["values"].tolist())))


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("trig", ["libm", "cr"])
def test_oracle_matches_reference_fixture(golden_dir, name, trig):
  g = load(golden_dir, name)
  H, W = int(g["H"]), int(g["W"])
  o = P.range_projection(g["points"], g["remissions"], H, W, float(g["fov_up"]), float(g["fov_down"]), trig)
  # depth is libm independent: bit exact everywhere
  assert np.array_equal(o["unproj_range"], g["unproj_range"])
  amb = P.ambiguous_points(g["points"], H, W, float(g["fov_up"]), float(g["fov_down"]))
  assert amb.mean() < 2e-3
  # per point pixel coordinates: equal on every non-ambiguous point (on the generating host 'libm' matches 100 %)
  assert np.array_equal(o["proj_x"][~amb], g["proj_x"][~amb])
  assert np.array_equal(o["proj_y"][~amb], g["proj_y"][~amb])
  # images: equal on every pixel no ambiguous point can touch
  touched = np.zeros((H, W), bool)
  for src in (o, g):
    touched[src["proj_y"][amb], src["proj_x"][amb]] = True
  ok = ~touched
  assert np.array_equal(o["proj_range"][ok], g["proj_range"][ok])   # min depth per pixel: exact even on ties
  # index / xyz / remission may differ from the reference only where the winning depth is tied (unstable argsort)
  diff = ok & (o["proj_idx"] != g["proj_idx"])
  if diff.any():
    d = P.point_depth(g["points"])
    assert np.array_equal(d[o["proj_idx"][diff]], d[g["proj_idx"][diff]])
    assert (o["proj_idx"][diff] < g["proj_idx"][diff]).all()          # our rule: lowest index
  same = ok & ~diff
  assert np.array_equal(o["proj_xyz"][same], g["proj_xyz"][same])
  assert np.array_equal(o["proj_remission"][same], g["proj_remission"][same])
  assert diff.sum() <= 0.02 * ok.sum()


@pytest.mark.parametrize("name", CASES)
def test_label_projection_and_assembly(golden_dir, name):
  g = load(golden_dir, name)
  H, W = int(g["H"]), int(g["W"])
  o = P.range_projection(g["points"], g["remissions"], H, W, float(g["fov_up"]), float(g["fov_down"]), "libm")
  sem = P.label_projection(o["proj_idx"], g["label"])
  same = o["proj_idx"] == g["proj_idx"]
  assert np.array_equal(sem[same], g["proj_sem_label"][same])
  final = P.assemble_range_image(o, sem, lut(golden_dir)).astype(np.float32)
  assert final.shape == (H, W, 6)
  assert np.array_equal(final[same], g["final_data"][same])
  assert same.mean() > 0.98
  # empty pixels: zeros and the LUT image of raw label 0
  empty = o["proj_idx"] < 0
  assert (final[empty][:, :5] == 0).all() and (final[empty][:, 5] == lut(golden_dir)[0]).all()


def test_tie_rule_and_clamping():
  # two points in the same pixel with identical depth: lowest index wins; a nearer third point beats both
  pts = np.array([[10, 0, 0], [10, 0, 0], [5, 0, 0], [0, 0, 50.0], [0, 0, -50.0]], np.float32)
  o = P.range_projection(pts[:2], None, 8, 16, 3.0, -25.0)
  r, c = o["proj_y"][0], o["proj_x"][0]
  assert o["proj_idx"][r, c] == 0 and o["proj_range"][r, c] == 10
  o = P.range_projection(pts[:3], None, 8, 16, 3.0, -25.0)
  assert o["proj_idx"][r, c] == 2 and o["proj_range"][r, c] == 5
  # straight up / straight down are clamped into the first / last row, not dropped
  o = P.range_projection(pts[3:], None, 8, 16, 3.0, -25.0)
  assert o["proj_y"].tolist() == [0, 7]
  assert (o["proj_idx"] >= 0).sum() == 2


def test_empty_scan_and_ring_variant():
  o = P.range_projection(np.zeros((0, 3), np.float32), None, 4, 8, 3.0, -25.0)
  assert (o["proj_idx"] == -1).all() and (o["proj_range"] == -1).all() and o["proj_x"].shape == (0,)
  pts = np.array([[10, 0, 0], [20, 0, 0], [0, 5, 0]], np.float32)
  o = P.range_projection_ring(pts, np.array([.1, .2, .3], np.float32), np.array([0, 0, 3]), 4, 8)
  # ring 0 -> last row; in-order scatter: the HIGHER index (1) wins even though it is farther
  assert o["proj_idx"][3, o["proj_x"][0]] == 1 and o["proj_range"][3, o["proj_x"][0]] == 20
  assert o["proj_idx"][0, o["proj_x"][2]] == 2


def test_ring_oracle_matches_reference_fixture(golden_dir):
  """`do_range_projection_ring` (laserscan_nuscenes.py:191-223) as run by the reference class itself
  (tests/golden/make_projection_golden.py: run_ring_case): the column depends on atan2 (libm-ambiguous points excluded
  like above), the row on the ring index; the highest point index written to a pixel wins."""
  g = load(golden_dir, "nusc_ring_32x1024")
  H, W = int(g["H"]), int(g["W"])
  assert np.unique(np.stack([(H - 1) - g["ring"], g["proj_x"]]), axis=1).shape[1] < g["ring"].shape[0]  # collisions exist
  for trig in ("libm", "cr"):
    o = P.range_projection_ring(g["points"], g["remissions"], g["ring"], H, W, trig)
    amb = P.ambiguous_points(g["points"], H, W, 12.0, -30.0, columns_only=True)
    assert amb.mean() < 2e-3
    assert np.array_equal(o["proj_x"][~amb], g["proj_x"][~amb])
    touched = np.zeros((H, W), bool)
    rows = (H - 1) - g["ring"]
    for src in (o, g):
      touched[rows[amb], src["proj_x"][amb]] = True
    ok = ~touched
    for k in ("proj_idx", "proj_range", "proj_xyz", "proj_remission"):
      assert np.array_equal(o[k][ok], g[k][ok]), (trig, k)
    assert ok.mean() > 0.99
  # the nuScenes file's own copy of do_range_projection (:226-286) equals the KITTI one the oracle restates
  o = P.range_projection(g["points"], g["remissions"], H, W, 12.0, -30.0, "libm")
  assert np.array_equal(o["unproj_range"], g["fov_unproj_range"])
  amb = P.ambiguous_points(g["points"], H, W, 12.0, -30.0)
  assert np.array_equal(o["proj_x"][~amb], g["fov_proj_x"][~amb]) and np.array_equal(o["proj_y"][~amb], g["fov_proj_y"][~amb])
  touched = np.zeros((H, W), bool)
  for src_x, src_y in ((o["proj_x"], o["proj_y"]), (g["fov_proj_x"], g["fov_proj_y"])):
    touched[src_y[amb], src_x[amb]] = True
  assert np.array_equal(o["proj_range"][~touched], g["fov_proj_range"][~touched])
