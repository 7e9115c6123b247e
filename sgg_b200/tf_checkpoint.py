"""Reader (and minimal writer) for TensorFlow V2 checkpoints ("tensor bundles"), in pure Python + numpy.

The reference builds a ``tf.train.Saver`` (train.py:280) whose ``save`` / ``restore`` (train.py:288-292) use this format:
``<prefix>.index`` is a LevelDB-style sorted string table mapping variable names to ``BundleEntryProto`` records (dtype,
shape, shard, offset, size, crc32c) and ``<prefix>.data-00000-of-0000N`` holds the raw little-endian tensor bytes.  TF itself
is not available in this image, so the format is restated here from its published layout (tensorflow/core/util/tensor_bundle,
tensorflow/core/lib/io/{format,block,table_builder}.cc, snappy's format description):

* table file  = data blocks | metaindex block | index block | 48-byte footer (two varint64 block handles, padding, magic
  0xdb4775248b80fb57);
* block       = prefix-compressed entries (shared, non-shared, value-length varint32s, key suffix, value), a restart array
  of uint32 offsets and its length; every block is followed by a 1-byte compression tag (0 none, 1 snappy) and a masked crc32c;
* the entry with the empty key is the ``BundleHeaderProto``; every other key is a variable name.

PARITY UNPINNED: without TensorFlow in the image no real checkpoint could be produced here; ``write_checkpoint`` emits the
same layout (uncompressed blocks) and the tests round-trip through it, check prefix compression, snappy decoding and the
crc32c known answers.  The name mapping onto the parameter buckets is the one of ``sgg_param_table`` (TF variable names).
"""
from __future__ import annotations

import math
import os
import struct
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_INT64 = 1, 2, 3, 9
_DTYPES = {DT_FLOAT: np.dtype("<f4"), DT_DOUBLE: np.dtype("<f8"), DT_INT32: np.dtype("<i4"), DT_INT64: np.dtype("<i8")}
_DTYPE_IDS = {np.dtype("float32"): DT_FLOAT, np.dtype("float64"): DT_DOUBLE, np.dtype("int32"): DT_INT32, np.dtype("int64"): DT_INT64}


# ------------------------------------------------------------------------------------------------ crc32c (Castagnoli)
def _crc_table() -> List[int]:
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC = _crc_table()


def crc32c(data: bytes, crc: int = 0) -> int:
    c = crc ^ 0xFFFFFFFF
    for b in data:
        c = _CRC[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    """leveldb/TF store crcs rotated and offset so that a crc of bytes that contain crcs stays well distributed."""
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------ varints / protobuf
def _get_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    shift = out = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _proto_fields(buf: bytes) -> Iterator[Tuple[int, int, object]]:
    """(field number, wire type, value) of a serialised protobuf message; nested messages stay bytes."""
    pos = 0
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield field, wt, v


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _parse_entry(buf: bytes) -> dict:
    """BundleEntryProto: dtype=1, shape=2 (TensorShapeProto: dim=2 {size=1}), shard_id=3, offset=4, size=5, crc32c=6, slices=7."""
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "sliced": False}
    for f, _, v in _proto_fields(buf):
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            for f2, _, v2 in _proto_fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, v3 in _proto_fields(v2):
                        if f3 == 1:
                            size = _signed64(v3)
                    e["shape"].append(size)
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


def _build_entry(dtype: int, shape, shard_id: int, offset: int, size: int, crc: int) -> bytes:
    dims = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(int(s)) for s in shape))
    out = b"\x08" + _put_varint(dtype) + b"\x12" + _put_varint(len(dims)) + dims
    if shard_id:
        out += b"\x18" + _put_varint(shard_id)
    if offset:
        out += b"\x20" + _put_varint(offset)
    out += b"\x28" + _put_varint(size) + b"\x35" + struct.pack("<I", crc)
    return out


# ------------------------------------------------------------------------------------------------ snappy (decode only)
def snappy_decompress(buf: bytes) -> bytes:
    """Raw snappy block format: varint32 uncompressed length, then literal / copy elements."""
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:                                   # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:                                   # copy, 1-byte offset
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:                                 # copy, 2-byte offset
            ln = (tag >> 2) + 1
            off = buf[pos] | (buf[pos + 1] << 8)
            pos += 2
        else:                                           # copy, 4-byte offset
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError("snappy: bad copy offset")
        for _ in range(ln):                             # may overlap its own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError(f"snappy: expected {n} bytes, produced {len(out)}")
    return bytes(out)


# ------------------------------------------------------------------------------------------------ sorted string table
def _read_block(data: bytes, offset: int, size: int, verify: bool = True) -> bytes:
    raw = data[offset:offset + size]
    ctype = data[offset + size]
    if verify:
        stored = struct.unpack_from("<I", data, offset + size + 1)[0]
        if mask_crc(crc32c(data[offset:offset + size + 1])) != stored:
            raise ValueError(f"table block at {offset}: crc32c mismatch")
    if ctype == 0:
        return raw
    if ctype == 1:
        return snappy_decompress(raw)
    raise ValueError(f"table block at {offset}: unknown compression type {ctype}")


def _block_entries(block: bytes) -> Iterator[Tuple[bytes, bytes]]:
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_table(path: str, verify: bool = True) -> List[Tuple[bytes, bytes]]:
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != TABLE_MAGIC:
        raise ValueError(f"{path}: not a TensorFlow/LevelDB table (bad magic)")
    footer = data[-48:]
    _, pos = _get_varint(footer, 0)             # metaindex handle (unused)
    _, pos = _get_varint(footer, pos)
    ioff, pos = _get_varint(footer, pos)
    isize, pos = _get_varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(data, ioff, isize, verify)):
        boff, p2 = _get_varint(handle, 0)
        bsize, _ = _get_varint(handle, p2)
        out.extend(_block_entries(_read_block(data, boff, bsize, verify)))
    return out


def _make_block(entries: List[Tuple[bytes, bytes]], restart_interval: int = 16) -> bytes:
    out, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    out += b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))
    return bytes(out)


def write_table(path: str, entries: List[Tuple[bytes, bytes]], block_entries: int = 64) -> None:
    """Uncompressed table with `block_entries` keys per data block (keys must be sorted)."""
    body, index = bytearray(), []

    def emit(block: bytes) -> Tuple[int, int]:
        off = len(body)
        body.extend(block + b"\x00")
        body.extend(struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
        return off, len(block)
    for i in range(0, max(1, len(entries)), block_entries):
        chunk = entries[i:i + block_entries]
        off, size = emit(_make_block(chunk))
        index.append(((chunk[-1][0] if chunk else b"") + b"\x00", _put_varint(off) + _put_varint(size)))
    moff, msize = emit(_make_block([]))
    ioff, isize = emit(_make_block(index, restart_interval=1))
    footer = _put_varint(moff) + _put_varint(msize) + _put_varint(ioff) + _put_varint(isize)
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    with open(path, "wb") as f:
        f.write(bytes(body) + footer)


# ------------------------------------------------------------------------------------------------ tensor bundle
def list_variables(prefix: str) -> Dict[str, dict]:
    out = {}
    for k, v in read_table(prefix + ".index"):
        if k == b"":
            continue                                   # BundleHeaderProto (num_shards, endianness, version)
        out[k.decode()] = _parse_entry(v)
    return out


def _num_shards(prefix: str) -> int:
    for k, v in read_table(prefix + ".index"):
        if k == b"":
            for f, _, val in _proto_fields(v):
                if f == 1:
                    return int(val)
    return 1


def read_checkpoint(prefix: str, names: Optional[List[str]] = None, verify: bool = True,
                    verify_data_limit: int = 1 << 22) -> Dict[str, np.ndarray]:
    """{variable name: array} of a V2 checkpoint ``<prefix>.index`` + ``<prefix>.data-*``.  The index blocks are always
    crc-checked; the per-tensor crc32c of the data file is checked for tensors up to ``verify_data_limit`` bytes (the
    pure-Python crc runs at a few MB/s; pass a larger limit to check the 79 MB attention kernels too)."""
    entries = list_variables(prefix)
    shards = _num_shards(prefix)
    files: Dict[int, bytes] = {}
    out = {}
    for name, e in entries.items():
        if names is not None and name not in names:
            continue
        if e["sliced"]:
            raise ValueError(f"{name}: partitioned (sliced) variables are not supported")
        if e["dtype"] not in _DTYPES:
            continue                                   # strings / resources: not parameters
        sid = e["shard_id"]
        if sid not in files:
            with open(f"{prefix}.data-{sid:05d}-of-{shards:05d}", "rb") as f:
                files[sid] = f.read()
        raw = files[sid][e["offset"]:e["offset"] + e["size"]]
        if verify and e["crc32c"] is not None and len(raw) <= verify_data_limit and mask_crc(crc32c(raw)) != e["crc32c"]:
            raise ValueError(f"{name}: crc32c mismatch in the data file")
        out[name] = np.frombuffer(raw, dtype=_DTYPES[e["dtype"]]).reshape(e["shape"]).copy()
    return out


def write_checkpoint(prefix: str, tensors: Dict[str, np.ndarray]) -> None:
    """Single-shard V2 checkpoint with the layout ``read_checkpoint`` (and tf.train.Saver) reads."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    data, entries = bytearray(), []
    for name in sorted(tensors, key=lambda s: s.encode()):
        a = np.asarray(tensors[name])                  # (ascontiguousarray would turn a scalar into shape (1,))
        if a.dtype not in _DTYPE_IDS:
            raise ValueError(f"{name}: dtype {a.dtype} not supported")
        raw = a.astype(a.dtype.newbyteorder("<")).tobytes(order="C")
        entries.append((name.encode(), _build_entry(_DTYPE_IDS[a.dtype], a.shape, 0, len(data), len(raw), mask_crc(crc32c(raw)))))
        data += raw
    header = b"\x08\x01" + b"\x1a\x02\x08\x01"       # num_shards = 1, little endian (default), version { producer: 1 }
    write_table(prefix + ".index", [(b"", header)] + entries)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))


# ------------------------------------------------------------------------------------------------ onto the parameter buckets
def split_for_buckets(tensors: Dict[str, np.ndarray], beta1: float = 0.5):
    """Splits a reference checkpoint into {"generator", "discriminator"} parameter dicts (TF variable names, as
    sgg_param_table lists them), the Adam slots tf.train.AdamOptimizer saves next to them (``<var>/Adam`` = m,
    ``<var>/Adam_1`` = v) and the optimiser step recovered from ``beta1_power`` (= beta1 ** t).  Variables of the
    convolutional front-end (gen:29-68) are returned untouched under "other"."""
    out = {"generator": {}, "discriminator": {}, "adam_m": {}, "adam_v": {}, "other": {}, "step": {}}
    hot = ("attention_perceptron/", "layer_norm_basic_lstm_cell/", "decoder/")
    for name, a in tensors.items():
        base, slot = name, None
        if name.endswith("/Adam"):
            base, slot = name[:-5], "adam_m"
        elif name.endswith("/Adam_1"):
            base, slot = name[:-7], "adam_v"
        is_hot = base == "Discriminator/W" or any(h in base for h in hot)
        if name.endswith("beta1_power") or name.endswith("beta1_power_1"):
            t = math.log(float(a)) / math.log(beta1) if 0.0 < float(a) < 1.0 else 0.0
            out["step"][name] = int(round(t))
        elif not is_hot:
            out["other"][name] = a
        elif slot:
            out[slot][base] = a
        elif base.startswith("Generator/"):
            out["generator"][base] = a
        else:
            out["discriminator"][base] = a
    return out
