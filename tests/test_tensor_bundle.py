"""TensorBundle reader (reference weight ingestion without TensorFlow).

No TensorFlow and no reference checkpoint exist in the build container, so the container format is exercised through
round trips with the writer of the same module (parity unpinned, stated in the module header); the primitives that have
published known answers (CRC32C, its LevelDB mask, the table magic, protobuf varints) are pinned here."""
import os
import struct

import numpy as np
import pytest

from pclsegmentation_b200.utils import tensor_bundle as tb


def test_crc32c_known_answers():
  assert tb.crc32c(b"123456789") == 0xE3069283           # the standard CRC-32C check value
  assert tb.crc32c(b"") == 0
  assert tb.crc32c(b"\x00" * 32) == 0x8A9136AA             # RFC 3720 B.4 test vectors
  assert tb.crc32c(b"\xff" * 32) == 0x62A8AB43
  assert tb.crc32c(bytes(range(32))) == 0x46DD794E
  # leveldb/util/crc32c.h: Mask(crc) = ((crc >> 15) | (crc << 17)) + 0xa282ead8
  assert tb.mask_crc(0) == 0xA282EAD8
  assert tb.mask_crc(0xE3069283) == ((((0xE3069283 >> 15) | (0xE3069283 << 17)) + 0xA282EAD8) & 0xFFFFFFFF)


def test_varint_roundtrip():
  for v in (0, 1, 127, 128, 300, 2 ** 32 - 1, 2 ** 63 + 5):
    enc = tb._varint(v)
    assert tb._read_varint(enc, 0) == (v, len(enc))
  assert tb._varint(300) == b"\xac\x02"                    # protobuf documentation example


def _tensors(rng, n=150):
  out = {}
  for i in range(n):
    shape = tuple(int(s) for s in rng.integers(1, 5, size=int(rng.integers(0, 5))))
    out["layer%03d/sub/kernel/.ATTRIBUTES/VARIABLE_VALUE" % i] = rng.normal(size=shape).astype(np.float32)
  out["save_counter/.ATTRIBUTES/VARIABLE_VALUE"] = np.array(7, dtype=np.int64)
  out["x/half"] = rng.normal(size=(3, 2)).astype(np.float16)
  out["x/empty"] = np.zeros((0, 4), np.float32)
  return out


def test_roundtrip_many_blocks(tmp_path):
  rng = np.random.default_rng(0)
  tensors = _tensors(rng)
  prefix = str(tmp_path / "variables" / "variables")
  tb.write_bundle(prefix, tensors, block_entries=16)       # several data blocks + prefix-compressed keys
  header, entries = tb.read_index(prefix)
  assert header["num_shards"] == 1 and set(entries) == set(tensors)
  got = tb.read_bundle(prefix, verify_data=True)
  assert set(got) == set(tensors)
  for k, v in tensors.items():
    assert got[k].dtype == v.dtype and got[k].shape == v.shape
    np.testing.assert_array_equal(got[k], v)
  # footer layout: 48 bytes, magic last
  raw = open(prefix + ".index", "rb").read()
  assert struct.unpack("<Q", raw[-8:])[0] == 0xDB4775248B80FB57
  # resolve from the SavedModel directory, the variables directory and the .index file
  for p in (str(tmp_path), str(tmp_path / "variables"), prefix, prefix + ".index"):
    assert tb.resolve_prefix(p) == prefix
  only = tb.read_bundle(prefix, names=["x/half"])
  assert list(only) == ["x/half"]


def test_corruption_is_detected(tmp_path):
  rng = np.random.default_rng(1)
  prefix = str(tmp_path / "ckpt")
  tb.write_bundle(prefix, _tensors(rng, 20))
  raw = bytearray(open(prefix + ".index", "rb").read())
  raw[10] ^= 0x40
  open(prefix + ".index", "wb").write(bytes(raw))
  with pytest.raises(ValueError, match="checksum"):
    tb.read_index(prefix)
  prefix2 = str(tmp_path / "ckpt2")
  tb.write_bundle(prefix2, {"a": np.arange(6, dtype=np.float32)})
  data = bytearray(open(prefix2 + ".data-00000-of-00001", "rb").read())
  data[3] ^= 1
  open(prefix2 + ".data-00000-of-00001", "wb").write(bytes(data))
  with pytest.raises(ValueError, match="data checksum"):
    tb.read_bundle(prefix2, verify_data=True)
  with pytest.raises(FileNotFoundError):
    tb.resolve_prefix(str(tmp_path / "nothing"))
  open(str(tmp_path / "bad.index"), "wb").write(b"\x00" * 64)
  with pytest.raises(ValueError, match="magic"):
    tb.read_index(str(tmp_path / "bad"))


def test_model_loads_savedmodel_layout(tmp_path):
  """A SavedModel-shaped directory written from one model restores into a second one (plus the optimizer / metric
  entries a real checkpoint carries, which must be ignored); a checkpoint that lacks variables is an error."""
  from pclsegmentation_b200.utils.args_loader import load_model_config
  _, src = load_model_config("squeezesegv2", "squeezesegv2")
  src.randomize_batch_norm(3)
  ref = src.get_weights_dict()
  extra = {k + "/.ATTRIBUTES/VARIABLE_VALUE": v for k, v in ref.items()}
  some = sorted(ref)[0]
  extra[some + "/.OPTIMIZER_SLOT/optimizer/m/.ATTRIBUTES/VARIABLE_VALUE"] = np.zeros_like(ref[some])
  extra["optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE"] = np.array(5, np.int64)
  extra["miou_tracker/total_confusion_matrix/.ATTRIBUTES/VARIABLE_VALUE"] = np.zeros((11, 11), np.float32)
  tb.write_bundle(str(tmp_path / "model" / "variables" / "variables"), extra)
  _, dst = load_model_config("squeezesegv2", "squeezesegv2")
  dst.load_weights(str(tmp_path / "model"))
  got = dst.get_weights_dict()
  assert set(got) == set(ref)
  for k in ref:
    np.testing.assert_array_equal(got[k], ref[k])
  # checkpoint-prefix form through save_weights_bundle
  src.save_weights_bundle(str(tmp_path / "ckpt" / "checkpoint"))
  _, dst2 = load_model_config("squeezesegv2", "squeezesegv2")
  dst2.load_weights(str(tmp_path / "ckpt" / "checkpoint"))
  np.testing.assert_array_equal(dst2.get_weights_dict()[some], ref[some])
  # incomplete checkpoint
  del extra[some + "/.ATTRIBUTES/VARIABLE_VALUE"]
  tb.write_bundle(str(tmp_path / "partial" / "variables" / "variables"), extra)
  with pytest.raises(KeyError, match="missing"):
    dst.load_weights(str(tmp_path / "partial"))
