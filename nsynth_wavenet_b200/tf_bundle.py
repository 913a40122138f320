"""Reader for TensorFlow V2 checkpoint bundles (``<prefix>.index`` + ``<prefix>.data-SSSSS-of-NNNNN``)
without TensorFlow.

Replaces what ``tf.train.Saver.restore`` does for the generation path (wavenet/fastgen.py:81-84,
wavenet/parallelgen.py:30-41; eval_wavenet.py:22 resolves the prefix with tf.train.latest_checkpoint).

TensorFlow is an un-vendored, un-pinned dependency of the reference, so the format is restated from
its published definition (tensorflow/core/util/tensor_bundle/tensor_bundle.{h,cc},
tensorflow/core/protobuf/tensor_bundle.proto, tensorflow/core/lib/io/{table,block,format}.cc — the
LevelDB table format):

* the ``.index`` file is an immutable sorted string table: data blocks of prefix-compressed
  (key, value) entries with a restart array, each block followed by a 5-byte trailer (compression
  type, masked CRC-32C), an index block mapping last-keys to block handles, and a 48-byte footer
  ending in the magic 0xdb4775248b80fb57;
* key "" holds a ``BundleHeaderProto`` (num_shards, endianness, version), every other key is a
  variable name whose value is a ``BundleEntryProto`` (dtype, shape, shard_id, offset, size,
  masked crc32c of the raw bytes);
* tensor bytes sit at ``offset`` in the shard file, little-endian, row-major.

PARITY UNPINNED: no TensorFlow and no reference checkpoint exist in this environment, so the reader
is tested against bundles fabricated by ``tests/tf_bundle_writer.py`` (written from the same
definition) — block structure, prefix compression, multi-block indices, shards and CRCs — not
against a file TensorFlow wrote."""
from __future__ import annotations

import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
FOOTER_LEN = 48
BLOCK_TRAILER_LEN = 5
MASK_DELTA = 0xa282ead8

# tensorflow/core/framework/types.proto
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8,
          9: np.int64, 10: np.bool_, 17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}


class BundleError(IOError):
    pass


# ---------------------------------------------------------------- CRC-32C (Castagnoli) ----------
def _make_crc_table():
    poly = 0x82f63b78
    tab = np.zeros(256, np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        tab[i] = c
    return tab


_CRC_TABLE = _make_crc_table()
_CRC_LIST = [int(v) for v in _CRC_TABLE]


_fast_crc = None


def _native_crc():
    """nsw_crc32c from libnsw_b200.so when the library is there (slice-by-8, ~1 GB/s); the pure-Python loop below
    gives the same value and is what runs on a machine that only has the Python package."""
    global _fast_crc
    if _fast_crc is None:
        try:
            from . import _lib
            _fast_crc = _lib.load(build_if_stale=False).nsw_crc32c
        except Exception:
            _fast_crc = False
    return _fast_crc


def crc32c(data, crc=0):
    """CRC-32C of ``data`` (bytes-like), continuing from ``crc``."""
    if len(data) >= 256 and _native_crc():
        arr = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data).view(np.uint8)
        return int(_fast_crc(arr.ctypes.data, arr.size, crc))
    return crc32c_py(data, crc)


def crc32c_py(data, crc=0):
    c = crc ^ 0xffffffff
    tab = _CRC_LIST
    for b in bytes(data):
        c = tab[(c ^ b) & 0xff] ^ (c >> 8)
    return c ^ 0xffffffff


def mask_crc(crc):
    """crc32c::Mask (tensorflow/core/lib/hash/crc32c.h): rotate right by 15, add a constant."""
    return ((((crc >> 15) | (crc << 17)) & 0xffffffff) + MASK_DELTA) & 0xffffffff


def unmask_crc(masked):
    rot = (masked - MASK_DELTA) & 0xffffffff
    return ((rot >> 17) | (rot << 15)) & 0xffffffff


# ---------------------------------------------------------------- varints / protobuf wire -------
def _varint(buf, pos):
    result, shift = 0, 0
    while True:
        if pos >= len(buf):
            raise BundleError('truncated varint')
        b = buf[pos]
        pos += 1
        result |= (b & 0x7f) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise BundleError('varint too long')


def _proto_fields(buf):
    """Yield (field_number, wire_type, value) of one protobuf message (wire types 0, 1, 2, 5)."""
    pos = 0
    while pos < len(buf):
        key, pos = _varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = struct.unpack_from('<Q', buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = bytes(buf[pos:pos + ln])
            if len(val) != ln:
                raise BundleError('truncated length-delimited field')
            pos += ln
        elif wt == 5:
            val = struct.unpack_from('<I', buf, pos)[0]
            pos += 4
        else:
            raise BundleError('unsupported protobuf wire type {}'.format(wt))
        yield field, wt, val


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def parse_header(buf):
    """BundleHeaderProto -> dict(num_shards, endianness, producer)."""
    out = {'num_shards': 0, 'endianness': 0, 'producer': 0}
    for field, _, val in _proto_fields(buf):
        if field == 1:
            out['num_shards'] = val
        elif field == 2:
            out['endianness'] = val
        elif field == 3:
            for f2, _, v2 in _proto_fields(val):
                if f2 == 1:
                    out['producer'] = v2
    return out


def parse_entry(buf):
    """BundleEntryProto -> dict(dtype, shape, shard_id, offset, size, crc32c, sliced)."""
    out = {'dtype': 0, 'shape': (), 'shard_id': 0, 'offset': 0, 'size': 0, 'crc32c': None, 'sliced': False}
    for field, _, val in _proto_fields(buf):
        if field == 1:
            out['dtype'] = val
        elif field == 2:
            dims = []
            for f2, _, v2 in _proto_fields(val):
                if f2 == 2:                      # TensorShapeProto.Dim
                    size = 0
                    for f3, _, v3 in _proto_fields(v2):
                        if f3 == 1:
                            size = _signed64(v3)
                    dims.append(size)
                elif f2 == 3 and v2:
                    raise BundleError('tensor of unknown rank in a checkpoint')
            out['shape'] = tuple(dims)
        elif field == 3:
            out['shard_id'] = val
        elif field == 4:
            out['offset'] = val
        elif field == 5:
            out['size'] = val
        elif field == 6:
            out['crc32c'] = val
        elif field == 7:
            out['sliced'] = True
    return out


# ---------------------------------------------------------------- table (.index) ----------------
def _read_block(data, offset, size, verify):
    end = offset + size + BLOCK_TRAILER_LEN
    if end > len(data):
        raise BundleError('block handle ({}, {}) runs past the end of the index file'.format(offset, size))
    contents = data[offset:offset + size]
    ctype = data[offset + size]
    if verify:
        stored = struct.unpack_from('<I', data, offset + size + 1)[0]
        actual = crc32c(data[offset:offset + size + 1])
        if unmask_crc(stored) != actual:
            raise BundleError('index block at {} fails its CRC-32C'.format(offset))
    if ctype == 1:
        raise BundleError('snappy-compressed index block (TensorFlow writes bundle indices uncompressed)')
    if ctype != 0:
        raise BundleError('unknown block compression type {}'.format(ctype))
    return contents


def _block_entries(block):
    """Yield (key, value) of one table block (tensorflow/core/lib/io/block.cc)."""
    if len(block) < 4:
        raise BundleError('table block too small')
    num_restarts = struct.unpack_from('<I', block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    if limit < 0:
        raise BundleError('bad restart array')
    pos, key = 0, b''
    while pos < limit:
        shared, pos = _varint(block, pos)
        unshared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        if shared > len(key) or pos + unshared + vlen > limit:
            raise BundleError('corrupt table entry')
        key = key[:shared] + bytes(block[pos:pos + unshared])
        pos += unshared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_index(index_path, verify=True):
    """-> (header dict, {name: entry dict}) of ``<prefix>.index``."""
    with open(index_path, 'rb') as f:
        data = f.read()
    if len(data) < FOOTER_LEN:
        raise BundleError('{}: too short for a table footer'.format(index_path))
    footer = data[-FOOTER_LEN:]
    if struct.unpack_from('<Q', footer, FOOTER_LEN - 8)[0] != TABLE_MAGIC:
        raise BundleError('{}: not a TensorFlow V2 checkpoint index (bad table magic)'.format(index_path))
    pos = 0
    _, pos = _varint(footer, pos)       # metaindex handle (unused)
    _, pos = _varint(footer, pos)
    ioff, pos = _varint(footer, pos)    # index block handle
    isize, pos = _varint(footer, pos)
    header, entries = None, {}
    for _, handle in _block_entries(_read_block(data, ioff, isize, verify)):
        boff, p2 = _varint(handle, 0)
        bsize, _ = _varint(handle, p2)
        for key, value in _block_entries(_read_block(data, boff, bsize, verify)):
            if key == b'':
                header = parse_header(value)
            else:
                entries[key.decode('utf-8')] = parse_entry(value)
    if header is None:
        raise BundleError('{}: bundle header entry is missing'.format(index_path))
    if header['endianness'] != 0:
        raise BundleError('big-endian bundles are not supported')
    return header, entries


# ---------------------------------------------------------------- public API --------------------
def is_bundle_prefix(prefix):
    return os.path.isfile(os.fspath(prefix) + '.index')


def latest_checkpoint(ckpt_dir):
    """tf.train.latest_checkpoint on the ``checkpoint`` state file (a text CheckpointState proto);
    falls back to the newest ``*.index`` in the directory.  -> prefix or None."""
    state = os.path.join(ckpt_dir, 'checkpoint')
    if os.path.isfile(state):
        with open(state, 'rt') as f:
            for line in f:
                line = line.strip()
                if line.startswith('model_checkpoint_path:'):
                    name = line.split(':', 1)[1].strip().strip('"')
                    prefix = name if os.path.isabs(name) else os.path.join(ckpt_dir, name)
                    if is_bundle_prefix(prefix):
                        return prefix
    cands = [os.path.join(ckpt_dir, f[:-len('.index')]) for f in os.listdir(ckpt_dir) if f.endswith('.index')]
    if not cands:
        return None
    return max(cands, key=lambda p: os.path.getmtime(p + '.index'))


AUTO_VERIFY_BYTES = 1 << 20


def read_bundle(prefix, names=None, verify_data='auto'):
    """-> {variable name: ndarray} of the bundle at ``prefix``.  ``names``: optional predicate or container
    selecting variables (optimizer slots of a training checkpoint need not be read).  ``verify_data``: True
    checks every tensor's CRC-32C, False none, 'auto' every tensor when libnsw_b200.so provides the fast CRC, else
    tensors up to 1 MiB (the pure-Python CRC runs at ~5 MB/s); the index blocks are always verified."""
    prefix = os.fspath(prefix)
    header, entries = read_index(prefix + '.index')
    shards = {}
    out = {}
    for name in sorted(entries):
        if names is not None and not (names(name) if callable(names) else name in names):
            continue
        e = entries[name]
        if e['sliced']:
            raise BundleError('{}: partitioned (sliced) variables are not supported'.format(name))
        if e['dtype'] not in DTYPES:
            if e['dtype'] == 7:
                continue                      # DT_STRING (e.g. a saved config): not a weight
            raise BundleError('{}: unsupported dtype enum {}'.format(name, e['dtype']))
        dt = np.dtype(DTYPES[e['dtype']])
        count = int(np.prod(e['shape'], dtype=np.int64)) if e['shape'] else 1
        if count * dt.itemsize != e['size']:
            raise BundleError('{}: {} bytes recorded for shape {} of {}'.format(name, e['size'], e['shape'], dt))
        sid = e['shard_id']
        if sid not in shards:
            path = '{}.data-{:05d}-of-{:05d}'.format(prefix, sid, header['num_shards'])
            if not os.path.isfile(path):
                raise BundleError('missing shard file ' + path)
            shards[sid] = np.memmap(path, dtype=np.uint8, mode='r')
        raw = shards[sid][e['offset']:e['offset'] + e['size']]
        if len(raw) != e['size']:
            raise BundleError('{}: shard {} is truncated'.format(name, sid))
        check = verify_data is True or (verify_data == 'auto' and (e['size'] <= AUTO_VERIFY_BYTES or _native_crc()))
        if check and e['crc32c'] is not None and unmask_crc(e['crc32c']) != crc32c(raw):
            raise BundleError('{}: tensor bytes fail their CRC-32C'.format(name))
        out[name] = np.frombuffer(bytes(raw), dtype=dt.newbyteorder('<')).reshape(e['shape']).astype(dt)
    return out


# ---------------------------------------------------------------- writer ------------------------
# What tf.train.Saver.save produces for a list of variables, needed by tools/make_eval_model.py (strip a training
# checkpoint to its EMA shadows).  One shard, uncompressed blocks, restart interval 16, block size 4 KiB -- the
# BundleWriter / TableBuilder defaults.
_DTYPE_ENUM = {np.dtype(v): k for k, v in DTYPES.items()}


def _put_varint(out, v):
    while v >= 0x80:
        out.append((v & 0x7f) | 0x80)
        v >>= 7
    out.append(v)


def _pb_varint(num, v):
    out = bytearray()
    _put_varint(out, (num << 3) | 0)
    _put_varint(out, v)
    return bytes(out)


def _pb_bytes(num, b):
    out = bytearray()
    _put_varint(out, (num << 3) | 2)
    _put_varint(out, len(b))
    return bytes(out) + bytes(b)


def _entry_bytes(arr, offset, crc):
    dims = b''.join(_pb_bytes(2, _pb_varint(1, int(d))) for d in arr.shape)
    out = _pb_varint(1, _DTYPE_ENUM[arr.dtype]) + _pb_bytes(2, dims)
    if offset:
        out += _pb_varint(4, offset)
    out += _pb_varint(5, arr.nbytes)
    tag = bytearray()
    _put_varint(tag, (6 << 3) | 5)
    return out + bytes(tag) + struct.pack('<I', mask_crc(crc))


def _finish_block(entries, restart_interval=16):
    """entries: sorted [(key bytes, value bytes)] -> block contents (prefix compression + restart array)."""
    buf, restarts, prev = bytearray(), [], b''
    for i, (key, val) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(buf))
        else:
            while shared < min(len(key), len(prev)) and key[shared] == prev[shared]:
                shared += 1
        _put_varint(buf, shared)
        _put_varint(buf, len(key) - shared)
        _put_varint(buf, len(val))
        buf += key[shared:] + val
        prev = key
    for r in restarts or [0]:
        buf += struct.pack('<I', r)
    buf += struct.pack('<I', max(1, len(restarts)))
    return bytes(buf)


def write_bundle(prefix, tensors, block_size=4096):
    """Write {name: ndarray} as a single-shard TF-V2 bundle at ``prefix`` (-> the prefix)."""
    prefix = os.fspath(prefix)
    names = sorted(tensors)
    offset, recs = 0, []
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        for n in names:
            arr = np.asarray(tensors[n])
            arr = arr if arr.ndim == 0 else np.ascontiguousarray(arr)       # (ascontiguousarray would turn a scalar into [1])
            if arr.dtype not in _DTYPE_ENUM:
                raise BundleError('{}: dtype {} cannot be stored'.format(n, arr.dtype))
            raw = arr.astype(arr.dtype.newbyteorder('<'), copy=False).tobytes()
            f.write(raw)
            recs.append((n.encode('utf-8'), _entry_bytes(arr, offset, crc32c(raw))))
            offset += len(raw)
    header = _pb_varint(1, 1) + _pb_bytes(3, _pb_varint(1, 1))     # num_shards = 1, little endian, producer 1
    items = [(b'', header)] + recs
    out, index_entries, cur, cur_size = bytearray(), [], [], 0

    def flush():
        nonlocal cur, cur_size
        if not cur:
            return
        block = _finish_block(cur)
        handle = bytearray()
        _put_varint(handle, len(out))
        _put_varint(handle, len(block))
        out.extend(block)
        out.append(0)                                               # no compression
        out.extend(struct.pack('<I', mask_crc(crc32c(block + b'\x00'))))
        index_entries.append((cur[-1][0], bytes(handle)))            # (a key >= every key of the block)
        cur, cur_size = [], 0

    for kv in items:
        cur.append(kv)
        cur_size += len(kv[0]) + len(kv[1]) + 3
        if cur_size >= block_size:
            flush()
    flush()
    meta = _finish_block([])
    meta_handle = bytearray()
    _put_varint(meta_handle, len(out))
    _put_varint(meta_handle, len(meta))
    out.extend(meta + b'\x00' + struct.pack('<I', mask_crc(crc32c(meta + b'\x00'))))
    idx = _finish_block(index_entries, restart_interval=1)
    idx_handle = bytearray()
    _put_varint(idx_handle, len(out))
    _put_varint(idx_handle, len(idx))
    out.extend(idx + b'\x00' + struct.pack('<I', mask_crc(crc32c(idx + b'\x00'))))
    footer = bytes(meta_handle) + bytes(idx_handle)
    footer += b'\x00' * (FOOTER_LEN - 8 - len(footer)) + struct.pack('<Q', TABLE_MAGIC)
    out.extend(footer)
    with open(prefix + '.index', 'wb') as f:
        f.write(bytes(out))
    return prefix
