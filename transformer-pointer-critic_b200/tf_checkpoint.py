"""TensorFlow checkpoint (TensorBundle) interchange for the ACTOR weights the reference exports with
`agent.save_weights(location)` -> `bin_actor.save_weights(location)` (agents/agent.py:289-294, agents/trainer.py:178-184):
Keras writes `<location>.index` + `<location>.data-00000-of-00001` in the object-based TF-checkpoint format because the
path has no `.h5` suffix.

Host-side file I/O only (NumPy + the standard library): no TensorFlow, nothing here is on the hot path.

Format restated (tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/table*, tensorflow/core/protobuf/
{tensor_bundle,trackable_object_graph}.proto as published for TF 2.3; TensorFlow is not installable here, so the files
this module writes are checked by reading them back with the reader below, NOT by TensorFlow itself -- stated in
DESIGN.md as "unpinned"):
  .index   a LevelDB-style table (no compression): data blocks of prefix-compressed (key, value) entries + restart array,
           a 5-byte trailer per block (type, masked crc32c), an index block, an empty meta-index block, a 48-byte footer
           with the magic 0xdb4775248b80fb57. Key "" -> BundleHeaderProto, every other key -> BundleEntryProto
           (dtype, shape, shard, offset, size, masked crc32c of the bytes).
  .data    the tensors' raw little-endian bytes at the recorded offsets.
  keys     `<attribute path>/.ATTRIBUTES/VARIABLE_VALUE`, e.g. `encoder/enc_layers/0/mha/wq/kernel/.ATTRIBUTES/...`;
           list attributes index by position, a Sequential's layers are `layer_with_weights-<i>`, a TimeDistributed wraps
           its Dense as `layer`; `_CHECKPOINTABLE_OBJECT_GRAPH` holds the serialized TrackableObjectGraph Keras walks when
           it restores (nodes = trackable objects, children by attribute name, one VARIABLE_VALUE attribute per variable).
The variable <-> flat-buffer mapping is the reference's own `trainable_weights` order, pinned by
tests/golden/model_*.npz (attribute paths recorded from the unmodified reference classes)."""
from __future__ import annotations

import struct

import numpy as np

MAGIC = 0xDB4775248B80FB57
VAR_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"
GRAPH_KEY = "_CHECKPOINTABLE_OBJECT_GRAPH"
DT_FLOAT, DT_STRING = 1, 7


# ---------------------------------------------------------------------------------------------- crc32c (Castagnoli)
def _make_table():
    tbl = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tbl.append(c)
    return np.array(tbl, dtype=np.uint32)


_TBL = _make_table()


def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    tbl = _TBL
    for b in data:
        c = int(tbl[(c ^ b) & 0xFF]) ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------- protobuf wire helpers
def _varint(n: int) -> bytes:
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf: bytes, pos: int):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _field(num: int, wire: int, payload: bytes) -> bytes:
    return _varint((num << 3) | wire) + payload


def _len_field(num: int, payload: bytes) -> bytes:
    return _field(num, 2, _varint(len(payload)) + payload)


def _parse(buf: bytes):
    """-> list of (field number, wire type, value) of one message."""
    out, pos = [], 0
    while pos < len(buf):
        tag, pos = _read_varint(buf, pos)
        num, wire = tag >> 3, tag & 7
        if wire == 0:
            v, pos = _read_varint(buf, pos)
        elif wire == 2:
            n, pos = _read_varint(buf, pos)
            v, pos = buf[pos:pos + n], pos + n
        elif wire == 5:
            v, pos = struct.unpack_from("<I", buf, pos)[0], pos + 4
        elif wire == 1:
            v, pos = struct.unpack_from("<Q", buf, pos)[0], pos + 8
        else:
            raise ValueError(f"unsupported wire type {wire}")
        out.append((num, wire, v))
    return out


def _entry_proto(dtype: int, shape, offset: int, size: int, crc: int) -> bytes:
    dims = b"".join(_len_field(2, _field(1, 0, _varint(int(d)))) for d in shape)
    msg = _field(1, 0, _varint(dtype)) + _len_field(2, dims)
    if offset:
        msg += _field(4, 0, _varint(offset))
    msg += _field(5, 0, _varint(size)) + _field(6, 5, struct.pack("<I", crc))
    return msg


def _header_proto() -> bytes:
    return _field(1, 0, _varint(1)) + _len_field(3, _field(1, 0, _varint(1)))     # num_shards = 1, version.producer = 1


# ---------------------------------------------------------------------------------------------- table (.index)
def _block(entries, restart_interval=16) -> bytes:
    buf, restarts, last = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(buf))
        else:
            while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                shared += 1
        buf += _varint(shared) + _varint(len(k) - shared) + _varint(len(v)) + k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        buf += struct.pack("<I", r)
    buf += struct.pack("<I", len(restarts))
    return bytes(buf)


def _with_trailer(block: bytes) -> bytes:
    return block + b"\x00" + struct.pack("<I", masked_crc(block + b"\x00"))


def _write_table(entries) -> bytes:
    """entries: sorted list of (key bytes, value bytes). One data block (the actor has ~90 variables)."""
    out = bytearray()
    data = _block(entries)
    data_handle = _varint(0) + _varint(len(data))
    out += _with_trailer(data)
    meta = _block([])
    meta_handle = _varint(len(out)) + _varint(len(meta))
    out += _with_trailer(meta)
    index = _block([(entries[-1][0] if entries else b"", data_handle)], restart_interval=1)
    index_handle = _varint(len(out)) + _varint(len(index))
    out += _with_trailer(index)
    footer = meta_handle + index_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC)
    return bytes(out + footer)


def _read_block(buf: bytes, offset: int, size: int):
    raw = buf[offset:offset + size]
    trailer = buf[offset + size:offset + size + 5]
    if trailer[0] != 0:
        raise ValueError("compressed table blocks are not supported (TensorBundle writes none)")
    if struct.unpack("<I", trailer[1:])[0] != masked_crc(raw + trailer[:1]):
        raise ValueError("table block checksum mismatch")
    n_restarts = struct.unpack_from("<I", raw, len(raw) - 4)[0]
    end = len(raw) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _read_varint(raw, pos)
        non_shared, pos = _read_varint(raw, pos)
        vlen, pos = _read_varint(raw, pos)
        key = key[:shared] + raw[pos:pos + non_shared]
        pos += non_shared
        out.append((key, raw[pos:pos + vlen]))
        pos += vlen
    return out


def _read_table(buf: bytes):
    if len(buf) < 48 or struct.unpack("<Q", buf[-8:])[0] != MAGIC:
        raise ValueError("not a TensorBundle index (bad magic)")
    footer = buf[-48:-8]
    _, p = _read_varint(footer, 0)
    _, p = _read_varint(footer, p)
    ioff, p = _read_varint(footer, p)
    isize, p = _read_varint(footer, p)
    entries = []
    for _, handle in _read_block(buf, ioff, isize):
        off, q = _read_varint(handle, 0)
        size, q = _read_varint(handle, q)
        entries += _read_block(buf, off, size)
    return entries


# ---------------------------------------------------------------------------------------------- object graph
def _object_graph(paths) -> bytes:
    """TrackableObjectGraph for variables at attribute paths like 'encoder/enc_layers/0/mha/wq/kernel'."""
    nodes = [{"children": {}, "attr": None}]           # node 0 = the model

    def child(node_id, name):
        ch = nodes[node_id]["children"]
        if name not in ch:
            nodes.append({"children": {}, "attr": None})
            ch[name] = len(nodes) - 1
        return ch[name]
    for path in paths:
        nid = 0
        for part in path.split("/"):
            nid = child(nid, part)
        nodes[nid]["attr"] = path
    out = b""
    for n in nodes:
        msg = b""
        for name, cid in n["children"].items():
            msg += _len_field(1, _field(1, 0, _varint(cid)) + _len_field(2, name.encode()))
        if n["attr"] is not None:
            full = n["attr"].replace("/", "_")
            msg += _len_field(2, _len_field(1, b"VARIABLE_VALUE") + _len_field(2, full.encode()) +
                              _len_field(3, (n["attr"] + VAR_SUFFIX).encode()))
        out += _len_field(1, msg)
    return out


def _string_tensor_bytes(s: bytes) -> bytes:
    """Scalar DT_STRING as the bundle stores it: varint64 lengths, masked crc32c of the lengths, then the bytes."""
    lengths = _varint(len(s))
    return lengths + struct.pack("<I", masked_crc(lengths)) + s


# ---------------------------------------------------------------------------------------------- public API
def keras_path(spec_name: str, time_distributed: bool = True) -> str:
    """Oracle / engine parameter name ('encoder.enc_layers.0.ffn.1.kernel') -> Keras attribute path of the reference's
    classes ('encoder/enc_layers/0/ffn/layer_with_weights-1/kernel'). Pinned by the recorded reference paths in
    tests/golden/model_*.npz (tests/test_tf_checkpoint.py)."""
    parts = spec_name.split(".")
    out = []
    i = 0
    while i < len(parts):
        p = parts[i]
        if p == "ffn" and i + 1 < len(parts) and parts[i + 1].isdigit():
            out += ["ffn", f"layer_with_weights-{parts[i + 1]}"]
            i += 2
            continue
        if p == "pointer":
            p = "last_decoder_layer"
        out.append(p)
        wrapped = p in ("embedding1", "embedding2", "embedding") and len(out) >= 2 and out[-2] in ("encoder", "decoder")
        if time_distributed and wrapped:
            out.append("layer")
        i += 1
    return "/".join(out)


def write_checkpoint(prefix: str, tensors: dict) -> None:
    """tensors: {attribute path: float32 ndarray}. Writes `<prefix>.index` and `<prefix>.data-00000-of-00001`."""
    data = bytearray()
    entries = [(b"", _header_proto())]
    items = {p + VAR_SUFFIX: np.ascontiguousarray(a, dtype="<f4") for p, a in tensors.items()}
    graph = _string_tensor_bytes(_object_graph(list(tensors.keys())))
    for key in sorted(list(items.keys()) + [GRAPH_KEY]):
        if key == GRAPH_KEY:
            raw, dtype, shape = graph, DT_STRING, ()
        else:
            raw, dtype, shape = items[key].tobytes(), DT_FLOAT, items[key].shape
        entries.append((key.encode(), _entry_proto(dtype, shape, len(data), len(raw), masked_crc(raw))))
        data += raw
    with open(prefix + ".index", "wb") as f:
        f.write(_write_table(entries))
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))


def read_checkpoint(prefix: str) -> dict:
    """-> {attribute path: float32 ndarray} of every VARIABLE_VALUE float tensor in `<prefix>.index` / `.data-*`."""
    with open(prefix + ".index", "rb") as f:
        entries = _read_table(f.read())
    with open(prefix + ".data-00000-of-00001", "rb") as f:
        data = f.read()
    out = {}
    for key, val in entries:
        k = key.decode()
        if not k.endswith(VAR_SUFFIX):
            continue
        dtype, shape, offset, size, crc = 0, [], 0, 0, None
        for num, wire, v in _parse(val):
            if num == 1:
                dtype = v
            elif num == 2:
                shape = [_parse(d)[0][2] if d else 0 for n2, _, d in _parse(v) if n2 == 2]
            elif num == 4:
                offset = v
            elif num == 5:
                size = v
            elif num == 6:
                crc = v
        if dtype != DT_FLOAT:
            continue
        raw = data[offset:offset + size]
        if crc is not None and masked_crc(raw) != crc:
            raise ValueError(f"checksum mismatch for {k}")
        out[k[:-len(VAR_SUFFIX)]] = np.frombuffer(raw, dtype="<f4").reshape(shape).copy()
    return out
