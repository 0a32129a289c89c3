"""Pure-Python reader / writer of TensorFlow's checkpoint "tensor bundle" (the `model-<step>.index` +
`model-<step>.data-00000-of-00001` pair that the reference's `tf.train.Saver` writes, models/network.py:131-145, 223-226,
254-262), so that variables trained with the reference can be loaded -- and this package's variables handed back -- without
TensorFlow.  No torch, no CUDA: numpy and the standard library only.

Format (tensorflow/core/util/tensor_bundle + tensorflow/core/lib/io/table, the LevelDB table format):
  * `<prefix>.index`: an immutable sorted string table.  Blocks of prefix-compressed entries
    (varint32 shared, varint32 unshared, varint32 value_len, key suffix, value) followed by a uint32 restart array and its
    length; every block is followed by a 5-byte trailer (compression type, masked CRC32C).  The file ends with a 48-byte
    footer: BlockHandle(metaindex), BlockHandle(index) as varint64 (offset, size), zero padding, the magic
    0xdb4775248b80fb57.  The index block maps a key >= the last key of each data block to that block's handle.
    TF writes the bundle index uncompressed (BundleWriter sets table::kNoCompression).
  * keys: "" -> BundleHeaderProto {1: num_shards, 2: endianness, 3: version}; "<variable name>" -> BundleEntryProto
    {1: dtype, 2: TensorShapeProto {2: Dim {1: size}}, 3: shard_id, 4: offset, 5: size, 6: fixed32 masked crc32c}.
  * `<prefix>.data-0000S-of-0000N`: the raw little-endian tensor bytes at [offset, offset + size).
PARITY NOTE: no TensorFlow build exists in this image, so the reader is tested against this module's own writer, which
follows the published format byte for byte (tests/test_tf_bundle.py); it has not met a file written by TensorFlow."""
import os
import struct

import numpy as np

MAGIC = 0xDB4775248B80FB57
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_INT64 = 1, 2, 3, 9
_DTYPES = {DT_FLOAT: np.dtype("<f4"), DT_DOUBLE: np.dtype("<f8"), DT_INT32: np.dtype("<i4"), DT_INT64: np.dtype("<i8")}
_DT_OF = {np.dtype("float32"): DT_FLOAT, np.dtype("float64"): DT_DOUBLE, np.dtype("int32"): DT_INT32,
          np.dtype("int64"): DT_INT64}


# ---- CRC32C (Castagnoli), masked as LevelDB / TF do ------------------------------------------------------------------------
def _crc_table():
    tab = np.zeros(256, np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab[i] = c
    return tab


def _slice_tables():
    t0 = [int(v) for v in _crc_table()]
    tabs = [t0]
    for _ in range(3):
        prev = tabs[-1]
        tabs.append([(prev[i] >> 8) ^ t0[prev[i] & 0xFF] for i in range(256)])
    return tabs


_T0, _T1, _T2, _T3 = _slice_tables()


def crc32c(data, crc=0):
    """CRC-32C of a bytes-like object (slicing-by-4 in pure Python: ~10 MB/s)."""
    data = bytes(data)
    crc ^= 0xFFFFFFFF
    n4 = len(data) // 4
    t0, t1, t2, t3 = _T0, _T1, _T2, _T3
    for (w,) in struct.iter_unpack("<I", data[:4 * n4]):
        crc ^= w
        crc = t3[crc & 0xFF] ^ t2[(crc >> 8) & 0xFF] ^ t1[(crc >> 16) & 0xFF] ^ t0[crc >> 24]
    for b in data[4 * n4:]:
        crc = t0[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def mask_crc(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---- varints / minimal protobuf ---------------------------------------------------------------------------------------------
def _put_varint(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _get_varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _pb_fields(buf):
    """Yield (field number, wire type, value) of a serialized message (varint, fixed32, fixed64, length-delimited)."""
    pos = 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        f, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield f, wt, v


def _pb_varint(f, v):
    return _put_varint(f << 3) + _put_varint(v)


def _pb_bytes(f, b):
    return _put_varint((f << 3) | 2) + _put_varint(len(b)) + b


def _parse_shape(buf):
    dims = []
    for f, _, v in _pb_fields(buf):
        if f == 2:
            size = 0
            for f2, _, v2 in _pb_fields(v):
                if f2 == 1:
                    size = v2 if v2 < (1 << 63) else v2 - (1 << 64)
            dims.append(size)
    return tuple(dims)


def _parse_entry(buf):
    e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "sliced": False}
    for f, _, v in _pb_fields(buf):
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            e["shape"] = _parse_shape(v)
        elif f == 3:
            e["shard_id"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 6:
            e["crc32c"] = v
        elif f == 7:
            e["sliced"] = True
    return e


# ---- table reader -------------------------------------------------------------------------------------------------------------
def _read_block(data, offset, size, verify):
    block = data[offset:offset + size]
    ctype = data[offset + size]
    if verify:
        want = struct.unpack_from("<I", data, offset + size + 1)[0]
        if mask_crc(crc32c(data[offset:offset + size + 1])) != want:
            raise ValueError("tensor bundle index: block checksum mismatch")
    if ctype != 0:
        raise ValueError("tensor bundle index: compressed block (TF writes bundle indices uncompressed)")
    return block


def _block_entries(block):
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        unshared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + unshared])
        pos += unshared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_index(prefix, verify=True):
    """-> (header dict, {variable name: entry dict})."""
    data = open(prefix + ".index", "rb").read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != MAGIC:
        raise ValueError(f"{prefix}.index is not a TensorFlow tensor-bundle index (bad magic)")
    footer = data[-48:]
    pos = 0
    _, pos = _get_varint(footer, pos)          # metaindex handle (unused)
    _, pos = _get_varint(footer, pos)
    ioff, pos = _get_varint(footer, pos)
    isize, pos = _get_varint(footer, pos)
    entries, header = {}, {}
    for _, handle in _block_entries(_read_block(data, ioff, isize, verify)):
        boff, p2 = _get_varint(handle, 0)
        bsize, _ = _get_varint(handle, p2)
        for key, value in _block_entries(_read_block(data, boff, bsize, verify)):
            if key == b"":
                for f, _, v in _pb_fields(value):
                    header[{1: "num_shards", 2: "endianness"}.get(f, f"field{f}")] = v
            else:
                entries[key.decode()] = _parse_entry(value)
    return header, entries


def load_checkpoint(prefix, names=None, verify=False):
    """{variable name: numpy array} of a TF checkpoint `<prefix>.index` / `<prefix>.data-*` (names: optional subset).
    verify=True also checks every tensor's CRC32C (pure Python: ~1 s per MB)."""
    header, entries = read_index(prefix, verify=True)
    if header.get("endianness", 0) != 0:
        raise ValueError("big-endian tensor bundles are not supported")
    nshards = max(1, header.get("num_shards", 1))
    out, files = {}, {}
    for name, e in entries.items():
        if names is not None and name not in names:
            continue
        if e["sliced"]:
            raise ValueError(f"{name}: partitioned (sliced) variables are not supported")
        if e["dtype"] not in _DTYPES:
            continue                                                   # e.g. string tensors of the Saver's bookkeeping
        path = f"{prefix}.data-{e['shard_id']:05d}-of-{nshards:05d}"
        if path not in files:
            files[path] = open(path, "rb")
        f = files[path]
        f.seek(e["offset"])
        raw = f.read(e["size"])
        if len(raw) != e["size"]:
            raise ValueError(f"{name}: truncated data shard")
        if verify and e["crc32c"] is not None and mask_crc(crc32c(raw)) != e["crc32c"]:
            raise ValueError(f"{name}: tensor checksum mismatch")
        out[name] = np.frombuffer(raw, _DTYPES[e["dtype"]]).reshape(e["shape"]).copy()
    for f in files.values():
        f.close()
    return out


# ---- writer (one shard, uncompressed; what a TF 1.x Saver(write_version=V2) produces for dense variables) -----------------------
class _BlockBuilder:
    def __init__(self, restart_interval=16):
        self.buf, self.restarts, self.count, self.last, self.interval = bytearray(), [0], 0, b"", restart_interval

    def add(self, key, value):
        shared = 0
        if self.count % self.interval == 0 and self.count:
            self.restarts.append(len(self.buf))
        elif self.count:
            while shared < min(len(key), len(self.last)) and key[shared] == self.last[shared]:
                shared += 1
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        self.last, self.count = key, self.count + 1

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def _emit_block(out, content):
    off = len(out)
    out += content + b"\x00" + struct.pack("<I", mask_crc(crc32c(content + b"\x00")))
    return _put_varint(off) + _put_varint(len(content))


def save_checkpoint(prefix, tensors, block_size=4096):
    """Write {name: array} as `<prefix>.index` + `<prefix>.data-00000-of-00001`."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items = sorted((k.encode(), np.asarray(v).copy(order="C")) for k, v in tensors.items())
    header = _pb_varint(1, 1) + _pb_varint(2, 0) + _pb_bytes(3, _pb_varint(1, 1))
    kv = [(b"", header)]
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        off = 0
        for key, arr in items:
            dt = _DT_OF.get(arr.dtype)
            if dt is None:
                raise ValueError(f"{key.decode()}: dtype {arr.dtype} not supported")
            raw = arr.astype(arr.dtype.newbyteorder("<")).tobytes()
            f.write(raw)
            shape = b"".join(_pb_bytes(2, _pb_varint(1, int(d))) for d in arr.shape)
            entry = _pb_varint(1, dt) + _pb_bytes(2, shape) + _pb_varint(4, off) + _pb_varint(5, len(raw)) + \
                _put_varint((6 << 3) | 5) + struct.pack("<I", mask_crc(crc32c(raw)))
            kv.append((key, entry))
            off += len(raw)
    out = bytearray()
    index = _BlockBuilder(restart_interval=1)
    blk = _BlockBuilder()
    for key, value in kv:
        blk.add(key, value)
        if len(blk.buf) >= block_size:
            index.add(key, _emit_block(out, blk.finish()))
            blk = _BlockBuilder()
    if blk.count:
        index.add(blk.last, _emit_block(out, blk.finish()))
    meta = _emit_block(out, _BlockBuilder().finish())
    idx = _emit_block(out, index.finish())
    footer = meta + idx
    out += footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))
    return prefix
