"""Self-checks of the NN oracle against hand-computed TF-2.9 semantics.  The reference pins
nothing here (TensorFlow absent, no tests, no weights): parity for this part is UNPINNED and these cases are what
anchors the restatement."""
import numpy as np
import torch

from oracle import nn as O


def test_same_padding_rule():
  assert O.same_pad(8, 3, 1) == (1, 1)
  assert O.same_pad(8, 3, 2) == (0, 1)      # even width, stride 2: 0 left / 1 right
  assert O.same_pad(7, 3, 2) == (1, 1)
  assert O.same_pad(8, 7, 1) == (3, 3)
  assert O.same_pad(8, 1, 1) == (0, 0)


def test_conv_stride2_even_width_covers_2j_2j1_2j2():
  x = torch.arange(1, 9, dtype=torch.float32).view(1, 1, 1, 8)
  k = torch.zeros(3, 3, 1, 1)
  k[1, :, 0, 0] = torch.tensor([1.0, 10.0, 100.0])   # only the middle kernel row sees data (H = 1)
  y = O.conv2d_same(x, k, None, (1, 2)).view(-1)
  # out[j] = x[2j] + 10 x[2j+1] + 100 x[2j+2], right edge zero padded
  assert y.tolist() == [1 + 20 + 300, 3 + 40 + 500, 5 + 60 + 700, 7 + 80 + 0]


def test_conv_transpose_index_map():
  # out[m] = sum_{m = 2j + k - 1} in[j] w[k]
  x = torch.tensor([1.0, 2.0, 3.0]).view(1, 1, 1, 3)
  w = torch.tensor([1.0, 10.0, 100.0, 1000.0]).view(1, 4, 1, 1)  # Keras [1,4,Cout,Cin]
  y = O.conv2d_transpose_1x4_s2(x, w, None).view(-1)
  exp = np.zeros(6)
  for j, xv in enumerate([1.0, 2.0, 3.0]):
    for k, wv in enumerate([1.0, 10.0, 100.0, 1000.0]):
      m = 2 * j + k - 1
      if 0 <= m < 6:
        exp[m] += xv * wv
  assert y.tolist() == exp.tolist()
  # channel roles: kernel[0,k,co,ci]
  x2 = torch.tensor([[1.0], [2.0]]).view(1, 2, 1, 1)           # Cin = 2, W = 1
  w2 = torch.zeros(1, 4, 3, 2)
  w2[0, 1, 2, 1] = 5.0                                          # k=1, co=2, ci=1
  y2 = O.conv2d_transpose_1x4_s2(x2, w2, None)
  assert y2.shape == (1, 3, 1, 2) and y2[0, 2, 0, 0] == 10.0 and y2.abs().sum() == 10.0


def test_maxpool_same_padding_never_wins():
  x = -torch.arange(1, 9, dtype=torch.float32).view(1, 1, 1, 8)
  y = O.max_pool_same(x, 3, (1, 2)).view(-1)
  assert y.tolist() == [-1, -3, -5, -7]
  y7 = O.max_pool_same(x, 7, (1, 1)).view(-1)
  assert y7.tolist() == [-1, -1, -1, -1, -2, -3, -4, -5]


def test_batch_norm_eps_and_head_rules():
  p = {"bn/gamma": torch.tensor([2.0]), "bn/beta": torch.tensor([1.0]), "bn/moving_mean": torch.tensor([3.0]),
       "bn/moving_variance": torch.tensor([0.999])}
  y = O.batch_norm(torch.tensor([5.0]).view(1, 1, 1, 1), p, "bn")
  assert abs(y.item() - (2.0 * (5 - 3) / 1.0 + 1.0)) < 1e-6      # sqrt(0.999 + 1e-3) = 1
  logits = torch.tensor([[[[1.0, 3.0, 3.0], [0.0, 0.0, 0.0]]]])   # ties -> lowest index
  prob, pred = O.segmentation_head(logits, torch.tensor([[[True, False]]]), none_index=2)
  assert pred.tolist() == [[[1, 2]]]
  assert abs(prob[0, 0, 1].sum().item() - 1.0) < 1e-6 and abs(prob[0, 0, 1, 0].item() - 1 / 3) < 1e-6


def test_darknet_stride_rewriting():
  assert O.darknet_strides(16) == ([2, 2, 2, 2, 1], [1, 2, 2, 2, 2])
  assert O.darknet_strides(8) == ([2, 2, 2, 1, 1], [1, 1, 2, 2, 2])
  assert O.darknet_strides(32) == ([2, 2, 2, 2, 2], [2, 2, 2, 2, 2])


def test_input_stage_float64_normalisation():
  s = np.zeros((1, 2, 6), np.float32)
  s[0, 0] = [1, 2, 3, 0.5, 10, 7]
  lidar, mask, label = O.input_stage(s, [1, 1, 1, 0, 5], [2, 2, 2, 1, 5], none_index=4)
  assert mask.tolist() == [[True, False]] and label.tolist() == [[7, 4]]
  assert lidar[0, 0].tolist() == [0.0, 0.5, 1.0, 0.5, 1.0, 1.0] and lidar[0, 1].tolist() == [0] * 6


def test_skip_shapes_and_fp32_vs_fp64(tmp_path):
  from pclsegmentation_b200.configs import Darknet21, SqueezeSegV2Config
  from pclsegmentation_b200.nets.Darknet import Darknet
  from pclsegmentation_b200.nets.SqueezeSegV2 import SqueezeSegV2
  rng = np.random.default_rng(0)
  for arch, cls, mc, widths in (("squeezesegv2", SqueezeSegV2, SqueezeSegV2Config(), (240, 1024, 2048)),
                                ("darknet", Darknet, Darknet21(), (240, 512))):
    for W in widths:
      mc.ZENITH_LEVEL, mc.AZIMUTH_LEVEL = 2, W
      model = cls(mc)
      model.randomize_batch_norm(3)
      x = rng.normal(size=(1, 2, W, 6)).astype(np.float32)
      m = np.ones((1, 2, W), bool)
      kw = dict(num_layers=getattr(mc, "NUM_LAYERS", 53), output_stride=getattr(mc, "OUTPUT_STRIDE", 16))
      lg32, pr, pd = O.forward(arch, model.variables, x, m, mc.CLASSES.index("None"), **kw)
      assert lg32.shape == (1, 2, W, mc.NUM_CLASS) and pd.dtype == np.int32
      if W == 240:
        lg64, _, _ = O.forward(arch, model.variables, x, m, mc.CLASSES.index("None"), dtype=torch.float64, **kw)
        assert np.abs(lg32 - lg64).max() < 1e-4   # oracle noise is far below the 1e-2 budget
