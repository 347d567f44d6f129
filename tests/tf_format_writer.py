"""Test-side WRITER for the TensorFlow file formats that ctpn_b200.tf_import reads (there is no TensorFlow here):
checkpoint V2 (LevelDB-format .index table + .data shard), the directory 'checkpoint' state file, and a frozen GraphDef
with Const / Identity nodes.  Written independently of the reader (own varint / proto / crc code) from the published
formats: tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/table_builder.cc + format.cc, graph.proto, tensor.proto."""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
DT = {np.dtype("float32"): 1, np.dtype("float64"): 2, np.dtype("int32"): 3, np.dtype("int64"): 9, np.dtype("float16"): 19}


def varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def tag(field, wt):
    return varint((field << 3) | wt)


def f_varint(field, v):
    return tag(field, 0) + varint(v)


def f_bytes(field, b):
    return tag(field, 2) + varint(len(b)) + bytes(b)


def f_fixed32(field, v):
    return tag(field, 5) + struct.pack("<I", v)


def crc32c_bitwise(data):
    crc = 0xFFFFFFFF
    for byte in bytes(data):
        crc ^= byte
        for _ in range(8):
            crc = (crc >> 1) ^ (0x82F63B78 & -(crc & 1))
    return crc ^ 0xFFFFFFFF


_TABLE = None


def crc32c_table(data):
    """Byte-wise table version of the same polynomial (for tensors too large for the bitwise loop); a real TF Saver stores
    the checksum of EVERY tensor."""
    global _TABLE
    if _TABLE is None:
        _TABLE = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ (0x82F63B78 & -(c & 1))
            _TABLE.append(c)
    crc = 0xFFFFFFFF
    t = _TABLE
    for byte in data:
        crc = t[(crc ^ byte) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def masked(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def shape_proto(shape):
    return b"".join(f_bytes(2, f_varint(1, int(d))) for d in shape)


def snappy_literal_only(data):
    """A valid raw-snappy stream made of literals (<= 60 bytes each)."""
    out = bytearray(varint(len(data)))
    for i in range(0, len(data), 60):
        chunk = data[i:i + 60]
        out.append((len(chunk) - 1) << 2)
        out += chunk
    return bytes(out)


class TableBuilder:
    def __init__(self, block_size=4096, restart_interval=16, compress=False):
        self.block_size, self.restart_interval, self.compress = block_size, restart_interval, compress
        self.file = bytearray()
        self.index = []            # (last key of block, offset, size)
        self._reset()

    def _reset(self):
        self.buf, self.restarts, self.count, self.last_key = bytearray(), [0], 0, b""

    def add(self, key, value):
        shared = 0
        if self.count % self.restart_interval == 0 and self.count:
            self.restarts.append(len(self.buf))
        elif self.count:
            while shared < min(len(key), len(self.last_key)) and key[shared] == self.last_key[shared]:
                shared += 1
        self.buf += varint(shared) + varint(len(key) - shared) + varint(len(value)) + key[shared:] + value
        self.last_key = key
        self.count += 1
        if len(self.buf) >= self.block_size:
            self._flush()

    def _emit(self, body):
        ctype = 0
        if self.compress:
            body, ctype = snappy_literal_only(body), 1
        off = len(self.file)
        self.file += body + bytes([ctype])
        self.file += struct.pack("<I", masked(crc32c_bitwise(body + bytes([ctype]))))
        return off, len(body)

    def _finish_block(self, buf, restarts):
        return bytes(buf) + b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))

    def _flush(self):
        if not self.count:
            return
        off, size = self._emit(self._finish_block(self.buf, self.restarts))
        self.index.append((self.last_key, off, size))
        self._reset()

    def finish(self):
        self._flush()
        moff, msize = self._emit(self._finish_block(b"", [0]))
        ibuf, irestarts = bytearray(), []
        for key, off, size in self.index:       # restart interval 1: no prefix sharing in the index block
            irestarts.append(len(ibuf))
            handle = varint(off) + varint(size)
            ibuf += varint(0) + varint(len(key)) + varint(len(handle)) + key + handle
        ioff, isize = self._emit(self._finish_block(ibuf, irestarts or [0]))
        footer = varint(moff) + varint(msize) + varint(ioff) + varint(isize)
        footer += b"\0" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
        self.file += footer
        return bytes(self.file)


def write_checkpoint(prefix, tensors, block_size=4096, compress=False, write_state=True):
    """tensors: {name: ndarray}.  One data shard."""
    data = bytearray()
    tb = TableBuilder(block_size=block_size, compress=compress)
    header = f_varint(1, 1) + f_varint(2, 0) + f_bytes(3, f_varint(1, 1))      # num_shards, LITTLE endian, version.producer
    tb.add(b"", header)
    for name in sorted(tensors, key=lambda s: s.encode("utf-8")):
        a = np.asarray(tensors[name])           # (ascontiguousarray would turn a 0-d scalar into shape (1,))
        raw = a.astype(a.dtype.newbyteorder("<")).tobytes()
        entry = f_varint(1, DT[a.dtype]) + f_bytes(2, shape_proto(a.shape)) + f_varint(3, 0) + f_varint(4, len(data)) + \
            f_varint(5, len(raw))
        entry += f_fixed32(6, masked(crc32c_bitwise(raw) if len(raw) <= (1 << 12) else crc32c_table(raw)))
        tb.add(name.encode("utf-8"), entry)
        data += raw
    with open(prefix + ".index", "wb") as f:
        f.write(tb.finish())
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))
    if write_state:
        with open(os.path.join(os.path.dirname(prefix), "checkpoint"), "w") as f:
            base = os.path.basename(prefix)
            f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))


def tensor_proto(a, use_content=True):
    a = np.asarray(a)
    t = f_varint(1, DT[a.dtype]) + f_bytes(2, shape_proto(a.shape))
    if use_content:
        return t + f_bytes(4, a.astype(a.dtype.newbyteorder("<")).tobytes())
    assert a.dtype == np.float32
    return t + f_bytes(5, a.astype("<f4").tobytes())          # packed float_val


def write_frozen_graph(path, tensors, scalar_fill=None):
    """GraphDef with, per variable, a Const node `name` and an Identity node `name/read` (what
    convert_variables_to_constants leaves behind), plus a Placeholder.  scalar_fill: {name: (shape, value)} Const nodes
    stored as a single float_val that fills the shape."""
    g = bytearray()
    g += f_bytes(1, f_bytes(1, b"Placeholder") + f_bytes(2, b"Placeholder") +
                 f_bytes(5, f_bytes(1, b"dtype") + f_bytes(2, f_varint(6, 1))))
    for name, a in tensors.items():
        a = np.asarray(a)
        node = f_bytes(1, name.encode()) + f_bytes(2, b"Const")
        node += f_bytes(5, f_bytes(1, b"dtype") + f_bytes(2, f_varint(6, DT[a.dtype])))
        node += f_bytes(5, f_bytes(1, b"value") + f_bytes(2, f_bytes(8, tensor_proto(a, use_content=(a.size != 3)))))
        g += f_bytes(1, node)
        ident = f_bytes(1, (name + "/read").encode()) + f_bytes(2, b"Identity") + f_bytes(3, name.encode())
        ident += f_bytes(5, f_bytes(1, b"T") + f_bytes(2, f_varint(6, DT[a.dtype])))
        g += f_bytes(1, ident)
    for name, (shape, value) in (scalar_fill or {}).items():
        t = f_varint(1, 1) + f_bytes(2, shape_proto(shape)) + tag(5, 5) + struct.pack("<f", value)   # one unpacked float_val
        node = f_bytes(1, name.encode()) + f_bytes(2, b"Const") + f_bytes(5, f_bytes(1, b"value") + f_bytes(2, f_bytes(8, t)))
        g += f_bytes(1, node)
    g += f_bytes(4, f_varint(1, 24))     # versions.producer
    with open(path, "wb") as f:
        f.write(bytes(g))
