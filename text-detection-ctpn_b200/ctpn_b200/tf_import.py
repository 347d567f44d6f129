"""Readers for the weight files the reference restores from, without TensorFlow (SURVEY.md §8 f rank 1).

* TF checkpoint V2 ("tensor bundle"): `saver.restore(sess, ckpt.model_checkpoint_path)` in ctpn/demo.py:88-90 and
  ctpn/generate_pb.py:22-24.  `<prefix>.index` is an SSTable in the LevelDB table format (tensorflow/core/lib/io/table*)
  mapping "" -> BundleHeaderProto and variable name -> BundleEntryProto (dtype, shape, shard, offset, size, crc32c);
  `<prefix>.data-SSSSS-of-NNNNN` holds the raw little-endian tensor bytes.
* `checkpoint` state file of a directory (`tf.train.get_checkpoint_state`, demo.py:88): text proto with
  `model_checkpoint_path: "<name>"`.
* Frozen GraphDef `.pb` written by ctpn/generate_pb.py:36-40 (`convert_variables_to_constants`) and consumed by
  ctpn/demo_pb.py: every variable is a `Const` node named like the variable, its value in attr["value"].tensor.

Pure Python + numpy; protobuf messages are decoded from the wire format directly (no .proto files needed).  PARITY NOTE:
there is no TensorFlow in this environment, so these parsers are exercised against files produced by the independent
writer in tests/tf_format_writer.py (same published formats), not against TF-written files.
"""
import os
import re
import struct

import numpy as np

_TABLE_MAGIC = 0xDB4775248B80FB57
_MASK_DELTA = 0xA282EAD8

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 4: np.dtype("u1"), 5: np.dtype("<i2"),
           6: np.dtype("i1"), 9: np.dtype("<i8"), 10: np.dtype("?"), 17: np.dtype("<u2"), 19: np.dtype("<f2"),
           22: np.dtype("<u4"), 23: np.dtype("<u8")}
_DT_BFLOAT16 = 14


class TFFormatError(ValueError):
    pass


# ---- protobuf wire format -------------------------------------------------------------------------------------------
def _varint(buf, pos):
    result = shift = 0
    while True:
        if pos >= len(buf):
            raise TFFormatError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise TFFormatError("varint too long")


def _fields(buf):
    """Yield (field_number, wire_type, value) of one serialized message; value is an int (varint / fixed) or bytes."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            if pos + ln > n:
                raise TFFormatError("truncated length-delimited field")
            val = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            val = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise TFFormatError("unsupported protobuf wire type %d" % wt)
        yield field, wt, val


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _shape(buf):
    """TensorShapeProto -> tuple (dim = field 2, Dim.size = field 1)."""
    dims = []
    for f, _, v in _fields(buf):
        if f == 2:
            size = 0
            for ff, _, vv in _fields(v):
                if ff == 1:
                    size = _signed64(vv)
            dims.append(size)
        elif f == 3 and v:
            raise TFFormatError("tensor of unknown rank")
    return tuple(dims)


# ---- crc32c (Castagnoli), masked as in LevelDB / TF -----------------------------------------------------------------
def _make_crc_table():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC_TABLE = _make_crc_table()


def crc32c(data, crc=0):
    crc ^= 0xFFFFFFFF
    tab = _CRC_TABLE
    for b in bytes(data):
        crc = tab[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def crc32c_fast(data):
    """crc32c through the library's C implementation (ctpn_crc32c_host, ~1 GB/s) when it is loadable, else the pure-Python
    table loop above (~2 MB/s; only reached when tf_import is used stand-alone without the built library)."""
    try:
        from . import _native as N
    except Exception:
        return crc32c(data)
    buf = bytes(data)
    return int(N.lib.ctpn_crc32c_host(buf, len(buf), 0))


def mask_crc(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + _MASK_DELTA) & 0xFFFFFFFF


# ---- snappy (raw format) decompression: LevelDB blocks may be compressed ----------------------------------------------
def snappy_decompress(buf):
    n, pos = _varint(buf, 0)
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
            raise TFFormatError("corrupt snappy stream")
        for _ in range(ln):                             # may overlap its own output
            out.append(out[-off])
    if len(out) != n:
        raise TFFormatError("snappy length mismatch (%d != %d)" % (len(out), n))
    return bytes(out)


# ---- LevelDB table (SSTable) --------------------------------------------------------------------------------------------
def _block_handle(buf, pos):
    off, pos = _varint(buf, pos)
    size, pos = _varint(buf, pos)
    return off, size, pos


def _read_block(data, off, size, verify):
    if off + size + 5 > len(data):
        raise TFFormatError("block handle outside the file")
    body = data[off:off + size]
    ctype = data[off + size]
    if verify:
        stored = struct.unpack_from("<I", data, off + size + 1)[0]
        if mask_crc(crc32c(data[off:off + size + 1])) != stored:
            raise TFFormatError("block checksum mismatch at offset %d" % off)
    if ctype == 0:
        return body
    if ctype == 1:
        return snappy_decompress(body)
    raise TFFormatError("unknown block compression type %d" % ctype)


def _block_entries(block):
    if len(block) < 4:
        raise TFFormatError("block too small")
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 * (num_restarts + 1)
    if limit < 0:
        raise TFFormatError("corrupt restart array")
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        if shared > len(key) or pos + non_shared + vlen > limit:
            raise TFFormatError("corrupt block entry")
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_table(path, verify=True):
    """All (key, value) pairs of a LevelDB-format table file, in key order."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 48:
        raise TFFormatError("%s: too short for an SSTable" % path)
    footer = data[-48:]
    if struct.unpack_from("<Q", footer, 40)[0] != _TABLE_MAGIC:
        raise TFFormatError("%s: not an SSTable (bad magic)" % path)
    _, _, pos = _block_handle(footer, 0)              # metaindex (unused)
    ioff, isize, _ = _block_handle(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(data, ioff, isize, verify)):
        off, size, _ = _block_handle(handle, 0)
        out.extend(_block_entries(_read_block(data, off, size, verify)))
    return out


# ---- checkpoint V2 ------------------------------------------------------------------------------------------------------
def _bundle_entry(buf):
    e = dict(dtype=0, shape=(), shard_id=0, offset=0, size=0, crc32c=None, sliced=False)
    for f, _, v in _fields(buf):
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            e["shape"] = _shape(v)
        elif f == 3:
            e["shard_id"] = v
        elif f == 4:
            e["offset"] = _signed64(v)
        elif f == 5:
            e["size"] = _signed64(v)
        elif f == 6:
            e["crc32c"] = v
        elif f == 7:
            e["sliced"] = True
    return e


def checkpoint_prefix(path):
    """Accept a prefix, a '<prefix>.index' / '.data-...' file name, or a directory holding a 'checkpoint' state file."""
    if os.path.isdir(path):
        p = latest_checkpoint(path)
        if p is None:
            raise FileNotFoundError("no 'checkpoint' state file (or no .index file) in %s" % path)
        return p
    for suffix in (".index", ".meta"):
        if path.endswith(suffix):
            return path[:-len(suffix)]
    m = re.match(r"(.*)\.data-\d{5}-of-\d{5}$", path)
    return m.group(1) if m else path


def latest_checkpoint(directory):
    """tf.train.get_checkpoint_state(dir).model_checkpoint_path (demo.py:88); falls back to the newest *.index."""
    state = os.path.join(directory, "checkpoint")
    if os.path.isfile(state):
        with open(state, "r") as f:
            for line in f:
                m = re.match(r'\s*model_checkpoint_path\s*:\s*"(.*)"\s*$', line)
                if m:
                    p = m.group(1)
                    return p if os.path.isabs(p) else os.path.join(directory, p)
    idx = sorted((os.path.getmtime(os.path.join(directory, n)), n) for n in os.listdir(directory) if n.endswith(".index"))
    return os.path.join(directory, idx[-1][1][:-len(".index")]) if idx else None


def list_checkpoint(prefix, verify=True):
    """{variable name: (numpy dtype or 'bfloat16', shape)} of a V2 checkpoint."""
    out = {}
    for key, val in read_table(checkpoint_prefix(prefix) + ".index", verify):
        if key == b"":
            continue
        e = _bundle_entry(val)
        out[key.decode("utf-8")] = ("bfloat16" if e["dtype"] == _DT_BFLOAT16 else _DTYPES.get(e["dtype"]), e["shape"])
    return out


def read_checkpoint(prefix, names=None, verify=True):
    """Variables of a TF checkpoint V2 as {name: ndarray}.  `names`: optional iterable restricting what is read (a
    missing name raises KeyError).  bfloat16 tensors are widened to float32."""
    prefix = checkpoint_prefix(prefix)
    entries, num_shards = {}, 1
    for key, val in read_table(prefix + ".index", verify):
        if key == b"":
            for f, _, v in _fields(val):
                if f == 1:
                    num_shards = v
                elif f == 2 and v != 0:
                    raise TFFormatError("big-endian checkpoints are not supported")
            continue
        entries[key.decode("utf-8")] = _bundle_entry(val)
    wanted = list(entries) if names is None else list(names)
    shards, out = {}, {}
    try:
        for name in wanted:
            if name not in entries:
                raise KeyError("variable '%s' is not in checkpoint %s" % (name, prefix))
            e = entries[name]
            if e["sliced"]:
                raise TFFormatError("partitioned variable '%s' is not supported" % name)
            if e["dtype"] == _DT_BFLOAT16:
                dt = np.dtype("<u2")
            elif e["dtype"] in _DTYPES:
                dt = _DTYPES[e["dtype"]]
            else:
                raise TFFormatError("variable '%s': unsupported dtype enum %d" % (name, e["dtype"]))
            count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
            if count * dt.itemsize != e["size"]:
                raise TFFormatError("variable '%s': %d bytes stored, shape %s needs %d" % (name, e["size"], e["shape"], count * dt.itemsize))
            sid = e["shard_id"]
            if sid not in shards:
                shards[sid] = open("%s.data-%05d-of-%05d" % (prefix, sid, num_shards), "rb")
            f = shards[sid]
            f.seek(e["offset"])
            raw = f.read(e["size"])
            if len(raw) != e["size"]:
                raise TFFormatError("variable '%s': data shard truncated" % name)
            if verify and e["crc32c"] is not None:
                # every tensor is checked: a truncated / corrupt shard or a mis-parsed offset must not load as weights
                if mask_crc(crc32c_fast(raw)) != e["crc32c"]:
                    raise TFFormatError("variable '%s': checksum mismatch" % name)
            arr = np.frombuffer(raw, dtype=dt).reshape(e["shape"])
            if e["dtype"] == _DT_BFLOAT16:
                arr = (arr.astype(np.uint32) << 16).view(np.float32)
            out[name] = np.array(arr)
    finally:
        for f in shards.values():
            f.close()
    return out


# ---- frozen GraphDef ----------------------------------------------------------------------------------------------------
def _tensor_proto(buf):
    dtype, shape, content = 0, (), b""
    floats, ints, doubles, int64s = [], [], [], []
    for f, wt, v in _fields(buf):
        if f == 1:
            dtype = v
        elif f == 2:
            shape = _shape(v)
        elif f == 4:
            content = v
        elif f == 5:        # float_val (packed or not)
            floats.extend(np.frombuffer(v, "<f4") if wt == 2 else [struct.unpack("<f", struct.pack("<I", v))[0]])
        elif f == 6:
            doubles.extend(np.frombuffer(v, "<f8") if wt == 2 else [struct.unpack("<d", struct.pack("<Q", v))[0]])
        elif f == 7:        # int_val
            if wt == 2:
                pos = 0
                while pos < len(v):
                    x, pos = _varint(v, pos)
                    ints.append(_signed64(x))
            else:
                ints.append(_signed64(v))
        elif f == 10:       # int64_val
            if wt == 2:
                pos = 0
                while pos < len(v):
                    x, pos = _varint(v, pos)
                    int64s.append(_signed64(x))
            else:
                int64s.append(_signed64(v))
    if dtype == _DT_BFLOAT16 or dtype not in _DTYPES:
        return None
    dt = _DTYPES[dtype]
    count = int(np.prod(shape, dtype=np.int64)) if shape else 1
    if content:
        if len(content) != count * dt.itemsize:
            raise TFFormatError("tensor_content has %d bytes, shape %s needs %d" % (len(content), shape, count * dt.itemsize))
        return np.array(np.frombuffer(content, dt).reshape(shape))
    vals = {1: floats, 2: doubles, 9: int64s}.get(dtype, ints)
    if len(vals) == 0:
        return np.zeros(shape, dt)
    arr = np.asarray(vals, dtype=dt)
    if arr.size == count:
        return arr.reshape(shape)
    if arr.size < count:     # TensorProto semantics: the last value repeats
        return np.concatenate([arr, np.full(count - arr.size, arr[-1], dt)]).reshape(shape)
    raise TFFormatError("more values than the shape %s holds" % (shape,))


def read_frozen_graph(path, names=None):
    """{node name: ndarray} for the Const nodes of a serialized GraphDef (demo_pb.py loads it with
    graph_def.ParseFromString).  `names` restricts / checks the result like read_checkpoint."""
    with open(path, "rb") as f:
        data = f.read()
    out = {}
    for f_, wt, node in _fields(data):
        if f_ != 1 or wt != 2:
            continue
        name, op, value = None, None, None
        for nf, _, nv in _fields(node):
            if nf == 1:
                name = nv.decode("utf-8")
            elif nf == 2:
                op = nv.decode("utf-8")
            elif nf == 5:       # map<string, AttrValue> entry
                key, attr = None, None
                for mf, _, mv in _fields(nv):
                    if mf == 1:
                        key = mv
                    elif mf == 2:
                        attr = mv
                if key == b"value" and attr is not None:
                    for af, awt, av in _fields(attr):
                        if af == 8 and awt == 2:
                            value = av
        if op == "Const" and name is not None and value is not None and (names is None or name in names):
            t = _tensor_proto(value)
            if t is not None:
                out[name] = t
    if names is not None:
        for n in names:
            if n not in out:
                raise KeyError("constant '%s' is not in graph %s" % (n, path))
    return out


def _main(argv=None):
    """python -m ctpn_b200.tf_import <checkpoint prefix | directory | frozen.pb> [out.npz]
    Lists the tensors of a TF checkpoint / frozen graph, or converts them to an .npz keyed by variable name."""
    import argparse
    ap = argparse.ArgumentParser(prog="python -m ctpn_b200.tf_import", description=_main.__doc__)
    ap.add_argument("source")
    ap.add_argument("out", nargs="?", help=".npz to write (omit to list)")
    a = ap.parse_args(argv)
    tensors = read_frozen_graph(a.source) if a.source.endswith(".pb") else read_checkpoint(a.source)
    if a.out:
        np.savez(a.out, **tensors)
        print("wrote %d tensors to %s" % (len(tensors), a.out))
    else:
        for k in sorted(tensors):
            print("%-60s %-8s %s" % (k, tensors[k].dtype, tuple(tensors[k].shape)))
    return 0


if __name__ == "__main__":
    raise SystemExit(_main())
