"""Shared helpers for the tests: seeded synthetic inputs of SURVEY.md §8(d)."""
import numpy as np


def synth_scan(rng, n, fov_up=3.0, fov_down=-25.0, rings=64, quantum=0.002):
  """Ring-structured synthetic scan with 2 mm range quantisation; returns [n,4] float32 (x,y,z,remission)."""
  k = rng.integers(0, rings, n)
  pitch = np.deg2rad(fov_down + (fov_up - fov_down) * (k + rng.random(n)) / rings)
  yaw = rng.uniform(-np.pi, np.pi, n)
  r = np.round(rng.uniform(2, 80, n) / quantum) * quantum
  pts = np.stack([r * np.cos(pitch) * np.cos(yaw), r * np.cos(pitch) * np.sin(yaw), r * np.sin(pitch),
                  rng.random(n)], 1)
  return pts.astype(np.float32)


def synth_range_images(rng, B, H, W, valid_rate=0.78, fov_up=3.0, fov_down=-25.0, channels=6, num_classes=20):
  """Synthetic RAW range images [B,H,W,channels] float32 (x,y,z,intensity,depth,label): valid ~ Bernoulli(rate),
  depth ~ U(2,80), direction from the pixel centre, invalid pixels all-zero."""
  valid = rng.random((B, H, W)) < valid_rate
  depth = rng.uniform(2, 80, (B, H, W))
  pitch = np.deg2rad(fov_up - (fov_up - fov_down) * (np.arange(H) + 0.5) / H)[None, :, None]
  yaw = (np.pi - 2 * np.pi * (np.arange(W) + 0.5) / W)[None, None, :]
  x = depth * np.cos(pitch) * np.cos(yaw)
  y = depth * np.cos(pitch) * np.sin(yaw)
  z = depth * np.sin(pitch) * np.ones_like(yaw)
  inten = rng.uniform(0, 0.99, (B, H, W))
  chans = [x, y, z, inten, depth]
  if channels == 6:
    chans.append(rng.integers(0, num_classes, (B, H, W)).astype(np.float64))
  img = np.stack(chans, -1) * valid[..., None]
  return img.astype(np.float32)
