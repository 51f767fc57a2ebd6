"""Pure-Python reader (and a minimal writer) for TensorFlow's TensorBundle, the `variables/variables.{index,data-*}` pair
inside a Keras SavedModel / checkpoint - the weight hand-off format of the reference (`train.py:42,60` writes it,
`inference.py:39` / `eval.py:40` read it through `tf.keras.models.load_model`).  TensorFlow itself is not needed.

PARITY UNPINNED: neither TensorFlow nor a reference checkpoint exists in the build container, so the format below is a
restatement from TensorFlow's published sources (tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/table*,
tensorflow/core/protobuf/tensor_bundle.proto) and is exercised by round trips through `write_bundle` only.

Format:
  <prefix>.index   an SSTable (LevelDB table format): data blocks | metaindex block | index block | 48-byte footer.
                   block   = contents | 1 byte compression (0 none, 1 snappy) | 4 bytes masked CRC32C(contents + type)
                   contents = entries | uint32 restarts[n] | uint32 n
                   entry   = varint shared | varint unshared | varint value_len | key suffix | value
                   footer  = BlockHandle(metaindex) | BlockHandle(index) | zero padding to 40 bytes | magic 0xdb4775248b80fb57
                   key ""  -> BundleHeaderProto {1: num_shards, 2: endianness, 3: version}
                   key k   -> BundleEntryProto  {1: dtype, 2: TensorShapeProto{2: Dim{1: size}}, 3: shard_id, 4: offset,
                                                 5: size, 6: fixed32 masked crc32c}
  <prefix>.data-SSSSS-of-NNNNN   raw little-endian row-major tensor bytes at [offset, offset + size)
"""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
_MASK_DELTA = 0xA282EAD8

# tensorflow/core/framework/types.proto
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
          17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
DT_STRING, DT_BFLOAT16 = 7, 14
_DTYPE_IDS = {np.dtype(v): k for k, v in DTYPES.items()}


# ---- CRC32C (Castagnoli), masked the LevelDB way -------------------------------------------------------------------
def _make_table():
  tbl = []
  for i in range(256):
    c = i
    for _ in range(8):
      c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
    tbl.append(c)
  return tbl


_CRC_TABLE = _make_table()


def crc32c(data, crc=0):
  crc ^= 0xFFFFFFFF
  tbl = _CRC_TABLE
  for b in bytes(data):
    crc = tbl[(crc ^ b) & 0xFF] ^ (crc >> 8)
  return crc ^ 0xFFFFFFFF


def mask_crc(crc):
  return (((crc >> 15) | (crc << 17)) + _MASK_DELTA) & 0xFFFFFFFF


# ---- varints / the three protobuf wire types that occur -----------------------------------------------------------
def _read_varint(buf, pos):
  result, shift = 0, 0
  while True:
    b = buf[pos]
    pos += 1
    result |= (b & 0x7F) << shift
    if not b & 0x80:
      return result, pos
    shift += 7


def _varint(v):
  out = bytearray()
  while True:
    b = v & 0x7F
    v >>= 7
    if v:
      out.append(b | 0x80)
    else:
      out.append(b)
      return bytes(out)


def _parse_proto(buf):
  """-> {field number: [values]} with varint ints, bytes for length-delimited, raw 4/8 bytes for fixed fields."""
  fields, pos = {}, 0
  while pos < len(buf):
    tag, pos = _read_varint(buf, pos)
    field, wire = tag >> 3, tag & 7
    if wire == 0:
      v, pos = _read_varint(buf, pos)
    elif wire == 1:
      v, pos = buf[pos:pos + 8], pos + 8
    elif wire == 2:
      n, pos = _read_varint(buf, pos)
      v, pos = buf[pos:pos + n], pos + n
    elif wire == 5:
      v, pos = buf[pos:pos + 4], pos + 4
    else:
      raise ValueError("unsupported protobuf wire type %d" % wire)
    fields.setdefault(field, []).append(v)
  return fields


def _signed64(v):
  return v - (1 << 64) if v >= (1 << 63) else v


# ---- SSTable ---------------------------------------------------------------------------------------------------------
def _read_block(buf, offset, size, verify):
  contents = buf[offset:offset + size]
  ctype = buf[offset + size]
  if verify:
    stored, = struct.unpack_from("<I", buf, offset + size + 1)
    if mask_crc(crc32c(buf[offset:offset + size + 1])) != stored:
      raise ValueError("TensorBundle index: block checksum mismatch at offset %d" % offset)
  if ctype != 0:
    raise ValueError("TensorBundle index: compressed blocks (type %d) are not supported; TensorFlow writes the bundle "
                     "index uncompressed" % ctype)
  return contents


def _block_entries(block):
  n_restarts, = struct.unpack_from("<I", block, len(block) - 4)
  end = len(block) - 4 - 4 * n_restarts
  pos, key = 0, b""
  while pos < end:
    shared, pos = _read_varint(block, pos)
    unshared, pos = _read_varint(block, pos)
    vlen, pos = _read_varint(block, pos)
    key = key[:shared] + bytes(block[pos:pos + unshared])
    pos += unshared
    yield key, bytes(block[pos:pos + vlen])
    pos += vlen


def read_index(prefix, verify=True):
  """-> (header dict, {tensor name: entry dict(dtype, shape, shard_id, offset, size, crc32c)})"""
  with open(prefix + ".index", "rb") as f:
    buf = f.read()
  if len(buf) < 48:
    raise ValueError("%s.index is too short to be a TensorBundle index" % prefix)
  magic, = struct.unpack_from("<Q", buf, len(buf) - 8)
  if magic != TABLE_MAGIC:
    raise ValueError("%s.index: bad table magic %#x" % (prefix, magic))
  footer = buf[len(buf) - 48:]
  _, pos = _read_varint(footer, 0)          # metaindex handle (unused)
  _, pos = _read_varint(footer, pos)
  ioff, pos = _read_varint(footer, pos)
  isize, pos = _read_varint(footer, pos)
  header, entries = {}, {}
  for _, handle in _block_entries(_read_block(buf, ioff, isize, verify)):
    boff, p = _read_varint(handle, 0)
    bsize, p = _read_varint(handle, p)
    for key, value in _block_entries(_read_block(buf, boff, bsize, verify)):
      f = _parse_proto(value)
      if key == b"":
        header = {"num_shards": f.get(1, [1])[0], "endianness": f.get(2, [0])[0]}
        continue
      shape = []
      if 2 in f:
        for dim in _parse_proto(f[2][0]).get(2, []):
          shape.append(_signed64(_parse_proto(dim).get(1, [0])[0]))
      entries[key.decode()] = {
          "dtype": f.get(1, [0])[0], "shape": tuple(shape), "shard_id": f.get(3, [0])[0],
          "offset": f.get(4, [0])[0], "size": f.get(5, [0])[0],
          "crc32c": struct.unpack("<I", f[6][0])[0] if 6 in f else None, "sliced": 7 in f}
  if header.get("endianness", 0) != 0:
    raise ValueError("big-endian TensorBundles are not supported")
  return header, entries


def read_bundle(prefix, names=None, verify_data=False):
  """Reads the numeric tensors of a TensorBundle -> {name: ndarray}.  `names`: optional iterable restricting the read;
  string tensors (the object graph), sliced (partitioned) variables and bfloat16 are skipped unless asked for by name,
  in which case they raise."""
  header, entries = read_index(prefix)
  n_shards = header.get("num_shards", 1)
  wanted = set(names) if names is not None else None
  out, files = {}, {}
  try:
    for name, e in entries.items():
      if wanted is not None and name not in wanted:
        continue
      if e["dtype"] not in DTYPES or e["sliced"]:
        if wanted is not None:
          raise ValueError("tensor %r has unsupported dtype %d / slicing" % (name, e["dtype"]))
        continue
      shard = e["shard_id"]
      if shard not in files:
        files[shard] = open("%s.data-%05d-of-%05d" % (prefix, shard, n_shards), "rb")
      f = files[shard]
      f.seek(e["offset"])
      raw = f.read(e["size"])
      dt = np.dtype(DTYPES[e["dtype"]])
      count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
      if len(raw) != e["size"] or count * dt.itemsize != e["size"]:
        raise ValueError("tensor %r: %d bytes on disk, shape %s of %s needs %d" % (name, len(raw), e["shape"], dt,
                                                                                   count * dt.itemsize))
      if verify_data and e["crc32c"] is not None and mask_crc(crc32c(raw)) != e["crc32c"]:
        raise ValueError("tensor %r: data checksum mismatch" % name)
      out[name] = np.frombuffer(raw, dtype=dt).reshape(e["shape"]).copy()
  finally:
    for f in files.values():
      f.close()
  return out


def resolve_prefix(path):
  """SavedModel directory, `variables` directory, checkpoint prefix or `*.index` file -> bundle prefix."""
  cands = [path, os.path.join(path, "variables", "variables"), os.path.join(path, "variables")]
  if path.endswith(".index"):
    cands.insert(0, path[:-len(".index")])
  for c in cands:
    if os.path.isfile(c + ".index"):
      return c
  raise FileNotFoundError("no TensorBundle (<prefix>.index) found at %r" % path)


def load_keras_variables(path, verify_data=False):
  """{Keras attribute path: ndarray} of a SavedModel / checkpoint: the object-graph keys
  '<attr>/<attr>/.../.ATTRIBUTES/VARIABLE_VALUE' with the suffix removed; optimizer slots are dropped."""
  suffix = "/.ATTRIBUTES/VARIABLE_VALUE"
  out = {}
  for name, arr in read_bundle(resolve_prefix(path), verify_data=verify_data).items():
    if not name.endswith(suffix) or "/.OPTIMIZER_SLOT/" in name:
      continue
    out[name[:-len(suffix)]] = arr
  return out


# ---- writer (tests, and exporting weights a TensorFlow user can restore) -------------------------------------------------
def _field(num, wire, payload):
  return _varint((num << 3) | wire) + payload


def _entry_proto(dtype_id, shape, offset, size, crc):
  dims = b"".join(_field(2, 2, _varint(len(d)) + d) for d in (_field(1, 0, _varint(int(s))) for s in shape))
  msg = _field(1, 0, _varint(dtype_id))
  msg += _field(2, 2, _varint(len(dims)) + dims)
  if offset:
    msg += _field(4, 0, _varint(offset))
  msg += _field(5, 0, _varint(size))
  msg += _field(6, 5, struct.pack("<I", crc))
  return msg


def _build_block(items, restart_interval=16):
  body, restarts, prev = bytearray(), [], b""
  for i, (key, value) in enumerate(items):
    shared = 0
    if i % restart_interval == 0:
      restarts.append(len(body))
    else:
      while shared < min(len(prev), len(key)) and prev[shared] == key[shared]:
        shared += 1
    body += _varint(shared) + _varint(len(key) - shared) + _varint(len(value)) + key[shared:] + value
    prev = key
  if not restarts:
    restarts = [0]
  for r in restarts:
    body += struct.pack("<I", r)
  body += struct.pack("<I", len(restarts))
  return bytes(body)


def write_bundle(prefix, tensors, block_entries=64):
  """Writes {name: ndarray} as a single-shard TensorBundle (names are stored sorted, as the table format requires)."""
  os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
  items = [(b"", _field(1, 0, _varint(1)) + _field(3, 2, _varint(2) + _field(1, 0, _varint(1))))]  # 1 shard, producer 1
  with open(prefix + ".data-00000-of-00001", "wb") as f:
    offset = 0
    for name in sorted(tensors):
      arr = np.asarray(tensors[name], order="C")   # (ascontiguousarray would turn scalars into shape (1,))
      if arr.dtype not in _DTYPE_IDS:
        raise ValueError("unsupported dtype %s for %r" % (arr.dtype, name))
      raw = arr.tobytes()
      f.write(raw)
      items.append((name.encode(), _entry_proto(_DTYPE_IDS[arr.dtype], arr.shape, offset, len(raw),
                                                mask_crc(crc32c(raw)))))
      offset += len(raw)
  out, index_items = bytearray(), []

  def emit(contents):
    handle = _varint(len(out)) + _varint(len(contents))
    out.extend(contents + b"\x00")
    out.extend(struct.pack("<I", mask_crc(crc32c(contents + b"\x00"))))
    return handle

  for i in range(0, len(items), block_entries):
    chunk = items[i:i + block_entries]
    index_items.append((chunk[-1][0], emit(_build_block(chunk))))
  meta_handle = emit(_build_block([]))
  index_handle = emit(_build_block(index_items, restart_interval=1))
  footer = meta_handle + index_handle
  out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC))
  with open(prefix + ".index", "wb") as f:
    f.write(bytes(out))
