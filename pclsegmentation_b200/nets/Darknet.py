"""Darknet21 / Darknet53 model builder - mirrors pcl_segmentation/nets/Darknet.py (BasicBlock :29-66, EncoderLayer
:69-103, DecoderLayer :106-138, model_blocks :142-145, Darknet :147-314 incl. the output-stride rewriting :158-181 /
:216-231 and the skip bookkeeping of run_enc_block / run_dec_block :263-277).  ``call`` is traced symbolically
(nets/layers.py): every conv+BN+LeakyReLU(+residual,+skip) chain becomes one fused implicit-GEMM op."""
from . import layers as L
from .SegmentationNetwork import PCLSegmentationNetwork


class BasicBlock(L.Layer):
  """Basic Block of the Darknet Architecture: 1x1 -> 3x3, both conv+BN+LeakyReLU(0.1), then += residual"""

  def __init__(self, path, inplanes, planes):
    super().__init__(path)
    self.conv1 = L.Conv2D(path + "/conv1", planes[0], 1, use_bias=False)
    self.bn1 = L.BatchNormalization(path + "/bn1")
    self.leaky_relu1 = L.LeakyReLU(0.1)
    self.conv2 = L.Conv2D(path + "/conv2", planes[1], 3, use_bias=False)
    self.bn2 = L.BatchNormalization(path + "/bn2")
    self.leaky_relu2 = L.LeakyReLU(0.1)

  def __call__(self, inputs, training=False):
    residual = inputs
    x = self.conv1(inputs)
    x = self.bn1(x)
    x = self.leaky_relu1(x)
    x = self.conv2(x)
    x = self.bn2(x)
    x = self.leaky_relu2(x)
    x += residual
    return x


class EncoderLayer(L.Layer):
  """Basic Encoder Layer of the Darknet Architecture"""

  def __init__(self, path, block, planes, num_blocks, stride):
    super().__init__(path)
    self.num_blocks = num_blocks
    self.conv1 = L.Conv2D(path + "/conv1", planes[1], 3, strides=[1, stride], use_bias=False)  # downsample
    self.bn1 = L.BatchNormalization(path + "/bn1")
    self.leaky_relu1 = L.LeakyReLU(0.1)
    inplanes = planes[1]
    for i in range(0, self.num_blocks):
      setattr(self, "residual_{}".format(i), block("{}/residual_{}".format(path, i), inplanes=inplanes, planes=planes))

  def __call__(self, inputs, training=False):
    x = self.conv1(inputs)
    x = self.bn1(x)
    x = self.leaky_relu1(x)
    for i in range(0, self.num_blocks):
      x = getattr(self, "residual_{}".format(i))(x)
    return x


class DecoderLayer(L.Layer):
  """Basic Decoder Layer of the Darknet Architecture"""

  def __init__(self, path, block, planes, stride):
    super().__init__(path)
    self.stride = stride
    if self.stride == 2:
      self.upconv1 = L.Conv2DTranspose(path + "/upconv1", planes[1], kernel_size=[1, 4], strides=[1, 2])  # upsample
    else:
      self.conv1 = L.Conv2D(path + "/conv1", planes[1], 3)  # keep constant
    self.bn1 = L.BatchNormalization(path + "/bn1")
    self.leaky_relu1 = L.LeakyReLU(0.1)
    # quirk kept: block(planes[1], planes) expands 1x1 out->in, then 3x3 in->out (nets/Darknet.py:128)
    self.block = block(path + "/block", planes[1], planes)

  def __call__(self, inputs, training=False):
    if self.stride == 2:
      x = self.upconv1(inputs)
    else:
      x = self.conv1(inputs)
    x = self.bn1(x)
    x = self.leaky_relu1(x)
    x = self.block(x)
    return x


# number of layers per model
model_blocks = {
  21: [1, 1, 2, 2, 1],
  53: [1, 2, 8, 8, 4],
}


def rewrite_strides(output_stride):
  """Net effect of the 'stride play' at nets/Darknet.py:158-181 (encoder, rewritten from the back) and :216-231
  (decoder, rewritten from the front): with k = log2(OUTPUT_STRIDE) the first k encoder layers halve the width and
  the last k decoder layers double it.  OUTPUT_STRIDE 16 (every shipped config) -> [2,2,2,2,1] / [1,2,2,2,2].
  The reference silently builds an inconsistent net for other values; they are rejected here."""
  if output_stride not in (1, 2, 4, 8, 16, 32):
    raise ValueError("OUTPUT_STRIDE must be a power of two <= 32, got %r" % (output_stride,))
  k = output_stride.bit_length() - 1
  return [2] * k + [1] * (5 - k), [1] * (5 - k) + [2] * k


class Darknet(PCLSegmentationNetwork):
  """Implements the Darknet Segmentation Model"""

  def __init__(self, mc):
    super(Darknet, self).__init__(mc)
    self.drop_rate = mc.DROP_RATE
    self.output_stride = mc.OUTPUT_STRIDE  # Output stride only horizontally
    self.num_layers = mc.NUM_LAYERS
    self.last_channels_encoder = 1024
    self.encoder_strides, self.decoder_strides = rewrite_strides(self.output_stride)
    self.num_blocks = model_blocks[self.num_layers]

    self.conv1 = L.Conv2D("conv1", 32, 3, use_bias=False)
    self.bn1 = L.BatchNormalization("bn1")
    self.leaky_relu1 = L.LeakyReLU(0.1)

    enc_planes = [[32, 64], [64, 128], [128, 256], [256, 512], [512, self.last_channels_encoder]]
    for i, planes in enumerate(enc_planes):
      setattr(self, "enc%d" % (i + 1), EncoderLayer("enc%d" % (i + 1), block=BasicBlock, planes=planes,
                                                    num_blocks=self.num_blocks[i], stride=self.encoder_strides[i]))
    self.dropout = L.Dropout(self.drop_rate)

    dec_planes = [[self.last_channels_encoder, 512], [512, 256], [256, 128], [128, 64], [64, 32]]
    for j, planes in enumerate(dec_planes):
      setattr(self, "dec%d" % (5 - j), DecoderLayer("dec%d" % (5 - j), BasicBlock, planes=planes,
                                                    stride=self.decoder_strides[j]))

    self.head = L.Conv2D("head", mc.NUM_CLASS, 3)

    self._trace()

  def run_enc_block(self, x, layer, skips, os):
    y = layer(x)
    if y.shape[1] < x.shape[1] or y.shape[2] < x.shape[2]:
      skips[os] = x
      os *= 2
    x = y
    return x, skips, os

  def run_dec_block(self, x, layer, skips, os):
    y = layer(x)  # up
    if y.shape[2] > x.shape[2]:
      os //= 2  # match skip
      y = y + skips[os]  # add skip (folded into the block's last conv epilogue)
    x = y
    return x, skips, os

  def call(self, inputs, training=False, mask=None):
    lidar_input, lidar_mask = inputs[0], inputs[1]
    skips, os = {}, 1  # skip connections keyed by the output stride at which they were taken

    x, skips, os = self.run_enc_block(lidar_input, self.conv1, skips, os)
    x = self._tap("conv1", self.leaky_relu1(self.bn1(x)))
    for i in (1, 2, 3, 4, 5):  # encoder, dropout (identity at inference) after every layer
      x, skips, os = self.run_enc_block(x, getattr(self, "enc%d" % i), skips, os)
      x = self._tap("enc%d" % i, self.dropout(x, training))
    for i in (5, 4, 3, 2, 1):  # decoder
      x, skips, os = self.run_dec_block(x, getattr(self, "dec%d" % i), skips, os)
      self._tap("dec%d" % i, x)

    logits = self.head(self.dropout(x, training))
    return self.segmentation_head(logits, lidar_mask)
