"""TEST INFRASTRUCTURE: writes TensorFlow V2 checkpoint bundles without TensorFlow, from the published
format (tensorflow/core/util/tensor_bundle/tensor_bundle.cc BundleWriter, tensorflow/core/lib/io/
table_builder.cc, block_builder.cc, format.cc), so that the reader in nsynth_wavenet_b200/tf_bundle.py
can be exercised on multi-block, prefix-compressed, multi-shard indices.  It shares only the CRC routine
with the reader; every structure below is built independently of the parsing code."""
import struct

import numpy as np

from nsynth_wavenet_b200.tf_bundle import crc32c, mask_crc, TABLE_MAGIC

DTYPE_ENUM = {np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3, np.dtype(np.int64): 9,
              np.dtype(np.float16): 19, np.dtype(np.uint8): 4, np.dtype(np.bool_): 10}


def varint(v):
    out = bytearray()
    while True:
        b = v & 0x7f
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def field_varint(num, v):
    return varint((num << 3) | 0) + varint(v)


def field_bytes(num, b):
    return varint((num << 3) | 2) + varint(len(b)) + b


def field_fixed32(num, v):
    return varint((num << 3) | 5) + struct.pack('<I', v)


def header_proto(num_shards):
    version = field_varint(1, 1)                                   # VersionDef.producer = 1
    return field_varint(1, num_shards) + field_bytes(3, version)   # endianness LITTLE (0) is the default: omitted


def entry_proto(dtype_enum, shape, shard_id, offset, size, crc_masked):
    dims = b''.join(field_bytes(2, field_varint(1, d)) for d in shape)
    out = field_varint(1, dtype_enum) + field_bytes(2, dims)
    if shard_id:
        out += field_varint(3, shard_id)
    if offset:
        out += field_varint(4, offset)
    out += field_varint(5, size) + field_fixed32(6, crc_masked)
    return out


class BlockBuilder:
    def __init__(self, restart_interval):
        self.interval = restart_interval
        self.buf = bytearray()
        self.restarts = [0]
        self.counter = 0
        self.last_key = b''

    def add(self, key, value):
        shared = 0
        if self.counter < self.interval:
            while shared < min(len(key), len(self.last_key)) and key[shared] == self.last_key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.counter = 0
        self.buf += varint(shared) + varint(len(key) - shared) + varint(len(value)) + key[shared:] + value
        self.last_key = key
        self.counter += 1

    def size(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def finish(self):
        return bytes(self.buf) + b''.join(struct.pack('<I', r) for r in self.restarts) + \
            struct.pack('<I', len(self.restarts))


def write_table(path, items, block_size=4096, restart_interval=16):
    """items: sorted list of (key bytes, value bytes)."""
    out = bytearray()
    index = BlockBuilder(1)

    def emit(block_bytes):
        off = len(out)
        out.extend(block_bytes)
        out.append(0)                                                   # kNoCompression
        out.extend(struct.pack('<I', mask_crc(crc32c(block_bytes + b'\x00'))))
        return varint(off) + varint(len(block_bytes))

    blk = BlockBuilder(restart_interval)
    for key, value in items:
        blk.add(key, value)
        if blk.size() >= block_size:
            index.add(blk.last_key, emit(blk.finish()))
            blk = BlockBuilder(restart_interval)
    if blk.buf:
        index.add(blk.last_key, emit(blk.finish()))
    meta_handle = emit(BlockBuilder(restart_interval).finish())        # empty metaindex block
    index_handle = emit(index.finish())
    footer = meta_handle + index_handle
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', TABLE_MAGIC)
    out.extend(footer)
    with open(path, 'wb') as f:
        f.write(bytes(out))


def write_bundle(prefix, tensors, num_shards=1, block_size=4096, restart_interval=16):
    """tensors: dict name -> ndarray.  Variables are dealt round-robin over `num_shards` data files."""
    names = sorted(tensors)
    shard_bufs = [bytearray() for _ in range(num_shards)]
    items = [(b'', header_proto(num_shards))]
    for i, name in enumerate(names):
        a = np.asarray(tensors[name])          # (ascontiguousarray would turn a scalar into shape (1,))
        raw = a.astype(a.dtype.newbyteorder('<')).tobytes()
        sid = i % num_shards
        off = len(shard_bufs[sid])
        shard_bufs[sid] += raw
        items.append((name.encode('utf-8'),
                      entry_proto(DTYPE_ENUM[a.dtype], a.shape, sid, off, len(raw), mask_crc(crc32c(raw)))))
    write_table(prefix + '.index', items, block_size, restart_interval)
    for sid, buf in enumerate(shard_bufs):
        with open('{}.data-{:05d}-of-{:05d}'.format(prefix, sid, num_shards), 'wb') as f:
            f.write(bytes(buf))
