"""Assembles tests/golden/handmade.ckpt.{index,data-00000-of-00003,data-00002-of-00003} byte by byte from the published
TensorFlow V2 bundle / LevelDB table format -- NOT with the repo's writers (neither tests/tf_bundle_writer.py nor
tf_bundle.write_bundle): every structure is laid out literally below, and the CRC-32C is a bitwise implementation local
to this script.  The committed files pin nsynth_wavenet_b200/tf_bundle.py against something it did not produce.

Sources: tensorflow/core/lib/io/format.{h,cc} (BlockHandle, Footer, kTableMagicNumber, block trailer),
block_builder.cc (entry = varint shared | varint non_shared | varint value_len | key delta | value; restart array;
num_restarts), table_builder.cc (index block: one entry per data block, restart interval 1),
tensorflow/core/protobuf/tensor_bundle.proto (BundleHeaderProto, BundleEntryProto), tensor_shape.proto, types.proto,
tensorflow/core/lib/hash/crc32c.h (Mask).

Covers: footer magic, block trailer CRCs, two data blocks, prefix-compressed keys with several restart points, varint
edge cases (127 / 128 / 16383 / 16384 / a 5-byte offset beyond 2 GiB), three shards (one of them absent on purpose: its
entries must be parseable and skippable), a scalar, float16 and int64 tensors.

    python tests/golden/make_handmade_bundle.py        # rewrites the three files next to this script
"""
import os
import struct

HERE = os.path.dirname(os.path.abspath(__file__))


def crc32c_bitwise(data):
    c = 0xffffffff
    for b in data:
        c ^= b
        for _ in range(8):
            c = (c >> 1) ^ 0x82f63b78 if c & 1 else c >> 1
    return c ^ 0xffffffff


def masked(crc):                     # crc32c::Mask: rotate right 15, add kMaskDelta
    return ((((crc >> 15) | (crc << 17)) & 0xffffffff) + 0xa282ead8) & 0xffffffff


def trailer(block):                  # 1 byte compression type (0 = none) + fixed32 masked crc of (block + type)
    return b'\x00' + struct.pack('<I', masked(crc32c_bitwise(block + b'\x00')))


# ---- tensor bytes --------------------------------------------------------------------------------------------------
scalar_i64 = struct.pack('<q', 200000)                                  # b/scalar_i64, shard 0, offset 0
mat_f16 = struct.pack('<4e', 0.5, -1.0, 2.0, 65504.0)                   # c/mat_f16 [2,2], shard 0, offset 8
shard0 = scalar_i64 + mat_f16
vec_f32 = struct.pack('<3f', 1.0, -2.5, 3.25)                           # b/vec_f32 [3], shard 2, offset 0
vec_ema = struct.pack('<3f', 0.125, 0.25, -0.5)                         # b/vec_f32/ExponentialMovingAverage, offset 12
shard2 = vec_f32 + vec_ema

# ---- BundleHeaderProto: num_shards (field 1, varint) = 3; version (field 3, message) { producer (field 1) = 1 } ------
header = bytes([0x08, 0x03, 0x1a, 0x02, 0x08, 0x01])


def fixed32(v):
    return struct.pack('<I', v)


# ---- BundleEntryProto: dtype (1) | shape (2: TensorShapeProto { dim (2) { size (1) } }) | shard_id (3) | offset (4) |
#      size (5) | crc32c (6, fixed32 => tag 0x35) --------------------------------------------------------------------
# a/big_offset: DT_FLOAT (1), shape [2], shard 1, offset 2^31 + 16 = 0x80000010 -> varint 90 80 80 80 08, size 8
e_big = bytes([0x08, 0x01,
               0x12, 0x04, 0x12, 0x02, 0x08, 0x02,
               0x18, 0x01,
               0x20, 0x90, 0x80, 0x80, 0x80, 0x08,
               0x28, 0x08,
               0x35]) + fixed32(masked(0x12345678))
# a/dims: DT_UINT8 (4), shape [127, 128, 16383, 16384]: dim sizes 7f | 80 01 | ff 7f | 80 80 01; shard 1, offset 0,
# size 127 * 128 * 16383 * 16384 = 4363420434432 -> varint 80 80 80 81 ff 7e
e_dims = bytes([0x08, 0x04,
                0x12, 0x14,
                0x12, 0x02, 0x08, 0x7f,
                0x12, 0x03, 0x08, 0x80, 0x01,
                0x12, 0x03, 0x08, 0xff, 0x7f,
                0x12, 0x04, 0x08, 0x80, 0x80, 0x01,
                0x18, 0x01,
                0x28, 0x80, 0x80, 0x80, 0x81, 0xff, 0x7e,
                0x35]) + fixed32(masked(0))
assert 127 * 128 * 16383 * 16384 == 4363420434432
# b/scalar_i64: DT_INT64 (9), empty shape message, shard 0 (default, omitted), offset 0 (omitted), size 8
e_scalar = bytes([0x08, 0x09, 0x12, 0x00, 0x28, 0x08, 0x35]) + fixed32(masked(crc32c_bitwise(scalar_i64)))
# b/vec_f32: DT_FLOAT, shape [3], shard 2, offset 0, size 12
e_vec = bytes([0x08, 0x01, 0x12, 0x04, 0x12, 0x02, 0x08, 0x03, 0x18, 0x02, 0x28, 0x0c, 0x35]) + \
    fixed32(masked(crc32c_bitwise(vec_f32)))
# b/vec_f32/ExponentialMovingAverage: same, offset 12
e_ema = bytes([0x08, 0x01, 0x12, 0x04, 0x12, 0x02, 0x08, 0x03, 0x18, 0x02, 0x20, 0x0c, 0x28, 0x0c, 0x35]) + \
    fixed32(masked(crc32c_bitwise(vec_ema)))
# c/mat_f16: DT_HALF (19 = 0x13), shape [2, 2], shard 0, offset 8, size 8
e_f16 = bytes([0x08, 0x13, 0x12, 0x08, 0x12, 0x02, 0x08, 0x02, 0x12, 0x02, 0x08, 0x02, 0x20, 0x08, 0x28, 0x08, 0x35]) + \
    fixed32(masked(crc32c_bitwise(mat_f16)))


def entry(shared, key_delta, value):
    assert shared < 128 and len(key_delta) < 128 and len(value) < 128      # one-byte varints in this fixture
    return bytes([shared, len(key_delta), len(value)]) + key_delta + value


# ---- data block 1: "", "a/big_offset", "a/dims" -- restart interval 16: one restart point, keys prefix-compressed -----
b1 = entry(0, b'', header)
b1 += entry(0, b'a/big_offset', e_big)
b1 += entry(2, b'dims', e_dims)                                         # shares "a/"
b1 += fixed32(0) + fixed32(1)                                           # restart[0] = 0, num_restarts = 1

# ---- data block 2: restart interval 2 => restart points at entries 0 and 2 ------------------------------------------
r0 = 0
b2 = entry(0, b'b/scalar_i64', e_scalar)
b2 += entry(2, b'vec_f32', e_vec)                                       # shares "b/"
r1 = len(b2)
b2 += entry(0, b'b/vec_f32/ExponentialMovingAverage', e_ema)            # restart point: full key
b2 += entry(0, b'c/mat_f16', e_f16)                                     # nothing shared with the previous key
b2 += fixed32(r0) + fixed32(r1) + fixed32(2)

index_file = bytearray()
h1 = (len(index_file), len(b1))
index_file += b1 + trailer(b1)
h2 = (len(index_file), len(b2))
index_file += b2 + trailer(b2)


def handle(off, size):
    assert off < 16384 and size < 16384
    out = bytearray()
    for v in (off, size):
        if v < 128:
            out.append(v)
        else:
            out += bytes([(v & 0x7f) | 0x80, v >> 7])
    return bytes(out)


# ---- metaindex block: empty ------------------------------------------------------------------------------------------
meta = fixed32(0) + fixed32(1)
hm = (len(index_file), len(meta))
index_file += meta + trailer(meta)
# ---- index block: separator keys >= last key of each data block, restart interval 1 ----------------------------------
ib = entry(0, b'a/dims', handle(*h1))
ir1 = len(ib)
ib += entry(0, b'c/mat_f16', handle(*h2))
ib += fixed32(0) + fixed32(ir1) + fixed32(2)
hi = (len(index_file), len(ib))
index_file += ib + trailer(ib)
# ---- footer: metaindex handle | index handle | zero padding to 40 bytes | magic 0xdb4775248b80fb57 (little endian) ----
footer = handle(*hm) + handle(*hi)
footer += b'\x00' * (40 - len(footer)) + bytes([0x57, 0xfb, 0x80, 0x8b, 0x24, 0x75, 0x47, 0xdb])
assert len(footer) == 48
index_file += footer

if __name__ == '__main__':
    for name, data in (('handmade.ckpt.index', bytes(index_file)), ('handmade.ckpt.data-00000-of-00003', shard0),
                       ('handmade.ckpt.data-00002-of-00003', shard2)):
        with open(os.path.join(HERE, name), 'wb') as f:
            f.write(data)
        print(name, len(data), 'bytes')
