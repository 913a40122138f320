"""TF-V2 checkpoint bundles read without TensorFlow (nsynth_wavenet_b200/tf_bundle.py), exercised on bundles
fabricated by tests/tf_bundle_writer.py from the published format.  CRC-32C is pinned to the RFC 3720 vectors."""
import os
import struct

import numpy as np
import pytest

from nsynth_wavenet_b200 import checkpoint, tf_bundle
from tf_bundle_writer import write_bundle


def test_crc32c_known_answers():
    # RFC 3720 appendix B.4 + the classic check value
    assert tf_bundle.crc32c(b'\x00' * 32) == 0x8a9136aa
    assert tf_bundle.crc32c(b'\xff' * 32) == 0x62a8ab43
    assert tf_bundle.crc32c(bytes(range(32))) == 0x46dd794e
    assert tf_bundle.crc32c(b'123456789') == 0xe3069283
    assert tf_bundle.crc32c(b'6789', tf_bundle.crc32c(b'12345')) == 0xe3069283       # continuation
    for v in (0, 1, 0x12345678, 0xffffffff):
        assert tf_bundle.unmask_crc(tf_bundle.mask_crc(v)) == v
    assert tf_bundle.mask_crc(0) == 0xa282ead8


def teacher_like_variables(rng, n_layers=30, ema=True, slots=True):
    out = {}
    for i in range(1, n_layers + 1):
        for nm, shape in (('dilated_conv_%d/W' % i, (1, 3, 8, 8)), ('dilated_conv_%d/biases' % i, (8,)),
                          ('res_%d/W' % i, (1, 1, 4, 8)), ('skip_%d/W' % i, (1, 1, 4, 4))):
            out[nm] = rng.normal(0, 1, shape).astype(np.float32)
            if ema:
                out[nm + '/ExponentialMovingAverage'] = rng.normal(0, 1, shape).astype(np.float32)
            if slots:
                out[nm + '/Adam'] = np.zeros(shape, np.float32)
                out[nm + '/Adam_1'] = np.zeros(shape, np.float32)
    out['global_step'] = np.asarray(200000, np.int64)
    out['beta1_power'] = np.asarray(0.0, np.float32)
    return out


@pytest.mark.parametrize('num_shards,block_size,restart', [(1, 4096, 16), (3, 256, 4), (2, 64, 1), (1, 1 << 20, 16)])
def test_round_trip_multi_block_multi_shard(tmp_path, num_shards, block_size, restart):
    rng = np.random.default_rng(5)
    vars_ = teacher_like_variables(rng)
    vars_['misc/f64'] = rng.normal(0, 1, (3, 2)).astype(np.float64)
    vars_['misc/i32'] = rng.integers(-5, 5, (7,)).astype(np.int32)
    vars_['misc/empty'] = np.zeros((0, 4), np.float32)
    prefix = str(tmp_path / 'model.ckpt-200000')
    write_bundle(prefix, vars_, num_shards=num_shards, block_size=block_size, restart_interval=restart)
    header, entries = tf_bundle.read_index(prefix + '.index')
    assert header['num_shards'] == num_shards and header['endianness'] == 0
    assert set(entries) == set(vars_)
    got = tf_bundle.read_bundle(prefix, verify_data=True)
    assert set(got) == set(vars_)
    for k, v in vars_.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape, k
        assert np.array_equal(got[k], v), k
    only = tf_bundle.read_bundle(prefix, names=lambda n: n.startswith('res_1/'))
    assert sorted(only) == sorted(k for k in vars_ if k.startswith('res_1/'))


def test_corruption_is_detected(tmp_path):
    rng = np.random.default_rng(6)
    vars_ = teacher_like_variables(rng, n_layers=3)
    prefix = str(tmp_path / 'model.ckpt-1')
    write_bundle(prefix, vars_, block_size=256)
    data_path = prefix + '.data-00000-of-00001'
    raw = bytearray(open(data_path, 'rb').read())
    raw[10] ^= 0x40
    open(data_path, 'wb').write(bytes(raw))
    with pytest.raises(tf_bundle.BundleError, match='CRC'):
        tf_bundle.read_bundle(prefix, verify_data=True)
    tf_bundle.read_bundle(prefix, verify_data=False)                       # explicit opt-out still reads
    raw[10] ^= 0x40
    open(data_path, 'wb').write(bytes(raw[:-8]))                           # truncated shard
    with pytest.raises(tf_bundle.BundleError, match='truncated'):
        tf_bundle.read_bundle(prefix)
    idx = bytearray(open(prefix + '.index', 'rb').read())
    idx[20] ^= 0x01                                                        # inside the first data block
    open(prefix + '.index', 'wb').write(bytes(idx))
    with pytest.raises(tf_bundle.BundleError, match='CRC'):
        tf_bundle.read_index(prefix + '.index')
    idx[20] ^= 0x01
    idx[-1] ^= 0xff                                                        # magic
    open(prefix + '.index', 'wb').write(bytes(idx))
    with pytest.raises(tf_bundle.BundleError, match='magic'):
        tf_bundle.read_index(prefix + '.index')
    os.remove(data_path)
    idx[-1] ^= 0xff
    open(prefix + '.index', 'wb').write(bytes(idx))
    with pytest.raises(tf_bundle.BundleError, match='missing shard'):
        tf_bundle.read_bundle(prefix)


def test_load_weights_from_bundle_applies_the_ema_map_like_the_saver(tmp_path):
    """fastgen.py:12-14,81-84: the Saver restores `<name>/ExponentialMovingAverage` INTO `<name>`; parallelgen.py:32-39
    reads the teacher deconv stack un-shadowed.  Same behaviour from a bundle, a directory, or an .npz."""
    rng = np.random.default_rng(7)
    vars_ = teacher_like_variables(rng, n_layers=2)
    vars_['deconv_0/W'] = rng.normal(0, 1, (1, 4, 3, 3)).astype(np.float32)
    vars_['deconv_0/W/ExponentialMovingAverage'] = rng.normal(0, 1, (1, 4, 3, 3)).astype(np.float32)
    d = tmp_path / 'ns_wn-eval'
    d.mkdir()
    write_bundle(str(d / 'model.ckpt-100'), {k: v * 0 for k, v in vars_.items()})
    write_bundle(str(d / 'model.ckpt-200'), vars_, num_shards=2, block_size=512)
    (d / 'checkpoint').write_text('model_checkpoint_path: "model.ckpt-200"\n'
                                  'all_model_checkpoint_paths: "model.ckpt-100"\n'
                                  'all_model_checkpoint_paths: "model.ckpt-200"\n')
    for path in (str(d), str(d / 'model.ckpt-200'), str(d / 'model.ckpt-200.index')):
        w = checkpoint.load_weights(path)
        assert 'global_step' not in w and not any(k.endswith('/Adam') for k in w)
        assert not any(k.endswith(checkpoint.EMA_SUFFIX) for k in w)
        assert np.array_equal(w['res_1/W'], vars_['res_1/W/ExponentialMovingAverage'])
        assert np.array_equal(w['deconv_0/W'], vars_['deconv_0/W/ExponentialMovingAverage'])
        w2 = checkpoint.load_weights(path, unshadowed_substrings=('deconv_',))
        assert np.array_equal(w2['deconv_0/W'], vars_['deconv_0/W'])
        assert np.array_equal(w2['res_1/W'], vars_['res_1/W/ExponentialMovingAverage'])
    # the .npz route gives the same dict
    npz = checkpoint.save_weights(str(tmp_path / 'export'), {k: v for k, v in vars_.items()}, ema=False)
    w3 = checkpoint.load_weights(npz)
    ref = checkpoint.load_weights(str(d))
    assert set(k for k in w3 if checkpoint._is_model_variable(k)) == set(ref)
    for k in ref:
        assert np.array_equal(w3[k], ref[k])
    with pytest.raises(FileNotFoundError):
        checkpoint.load_weights(str(tmp_path / 'nothing-here'))


def test_footer_and_block_layout_are_the_published_ones(tmp_path):
    """Byte-level checks of a fabricated index against the format definition (format.cc / block_builder.cc), so the
    writer the reader is tested with is itself anchored: 48-byte footer, magic, 5-byte block trailers, restart array."""
    prefix = str(tmp_path / 'm')
    write_bundle(prefix, {'a/W': np.arange(6, dtype=np.float32).reshape(2, 3), 'a/b': np.ones(2, np.float32)})
    data = open(prefix + '.index', 'rb').read()
    assert struct.unpack('<Q', data[-8:])[0] == 0xdb4775248b80fb57
    footer = data[-48:]
    moff, p = tf_bundle._varint(footer, 0)
    msize, p = tf_bundle._varint(footer, p)
    ioff, p = tf_bundle._varint(footer, p)
    isize, p = tf_bundle._varint(footer, p)
    assert footer[p:40] == b'\x00' * (40 - p)
    assert ioff + isize + 5 == len(data) - 48                     # index block is the last block before the footer
    assert moff + msize + 5 == ioff
    first = data[:moff - 5]                                       # the single data block
    assert data[moff - 5] == 0                                    # kNoCompression
    n_restarts = struct.unpack('<I', first[-4:])[0]
    assert n_restarts == 1 and struct.unpack('<I', first[-8:-4])[0] == 0
    # first entry: key "" (shared 0, unshared 0) -> header proto with num_shards = 1
    assert first[0] == 0 and first[1] == 0
    vlen = first[2]
    assert tf_bundle.parse_header(first[3:3 + vlen])['num_shards'] == 1
    # second entry shares no prefix with "", third shares "a/" with the second
    pos = 3 + vlen
    assert first[pos] == 0 and first[pos + 1] == 3                # "a/W": shared 0, unshared 3
    ents = list(tf_bundle._block_entries(first))
    assert [k for k, _ in ents] == [b'', b'a/W', b'a/b']
    e = tf_bundle.parse_entry(ents[1][1])
    assert e['dtype'] == 1 and e['shape'] == (2, 3) and e['size'] == 24 and e['offset'] == 0
    assert tf_bundle.unmask_crc(e['crc32c']) == tf_bundle.crc32c(np.arange(6, dtype='<f4').tobytes())


# ---- a bundle the repo's writers did NOT produce -------------------------------------------------------------------
HANDMADE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'handmade.ckpt')


def test_handmade_byte_level_fixture_parses_to_the_values_it_was_assembled_from():
    """tests/golden/handmade.ckpt.* were laid out byte by byte from the TensorFlow / LevelDB format sources by
    tests/golden/make_handmade_bundle.py (own bitwise CRC, literal protobuf bytes): two data blocks, prefix-compressed
    keys over several restart points, varint edges (127/128/16383/16384, a 5-byte offset past 2 GiB, a 6-byte size),
    three shards with shard 1 absent."""
    header, entries = tf_bundle.read_index(HANDMADE + '.index')
    assert header == {'num_shards': 3, 'endianness': 0, 'producer': 1}
    assert sorted(entries) == ['a/big_offset', 'a/dims', 'b/scalar_i64', 'b/vec_f32',
                               'b/vec_f32/ExponentialMovingAverage', 'c/mat_f16']
    big = entries['a/big_offset']
    assert (big['dtype'], big['shape'], big['shard_id'], big['offset'], big['size']) == (1, (2,), 1, 2 ** 31 + 16, 8)
    assert tf_bundle.unmask_crc(big['crc32c']) == 0x12345678
    dims = entries['a/dims']
    assert dims['shape'] == (127, 128, 16383, 16384) and dims['size'] == 127 * 128 * 16383 * 16384 and dims['dtype'] == 4
    assert entries['b/scalar_i64']['shape'] == () and entries['c/mat_f16']['dtype'] == 19
    got = tf_bundle.read_bundle(HANDMADE, names=lambda n: not n.startswith('a/'), verify_data=True)
    assert got['b/scalar_i64'].dtype == np.int64 and int(got['b/scalar_i64']) == 200000
    assert np.array_equal(got['b/vec_f32'], np.asarray([1.0, -2.5, 3.25], np.float32))
    assert np.array_equal(got['b/vec_f32/ExponentialMovingAverage'], np.asarray([0.125, 0.25, -0.5], np.float32))
    assert np.array_equal(got['c/mat_f16'], np.asarray([[0.5, -1.0], [2.0, 65504.0]], np.float16))
    with pytest.raises(tf_bundle.BundleError, match='missing shard'):
        tf_bundle.read_bundle(HANDMADE)                       # shard 1 is absent on purpose


def test_handmade_fixture_is_what_its_generator_script_writes(tmp_path):
    import importlib.util
    spec = importlib.util.spec_from_file_location('mk', os.path.join(os.path.dirname(HANDMADE), 'make_handmade_bundle.py'))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    assert open(HANDMADE + '.index', 'rb').read() == bytes(mk.index_file)
    assert open(HANDMADE + '.data-00000-of-00003', 'rb').read() == mk.shard0
    assert open(HANDMADE + '.data-00002-of-00003', 'rb').read() == mk.shard2
    assert mk.crc32c_bitwise(b'123456789') == 0xe3069283 == tf_bundle.crc32c_py(b'123456789')


def test_handmade_fixture_corruption_is_detected(tmp_path):
    import shutil
    for suffix in ('.index', '.data-00000-of-00003', '.data-00002-of-00003'):
        shutil.copy(HANDMADE + suffix, str(tmp_path / ('h.ckpt' + suffix)))
    p = str(tmp_path / 'h.ckpt')
    raw = bytearray(open(p + '.index', 'rb').read())
    raw[40] ^= 0x01                                           # inside data block 1
    open(p + '.index', 'wb').write(bytes(raw))
    with pytest.raises(tf_bundle.BundleError, match='CRC'):
        tf_bundle.read_index(p + '.index')
    raw[40] ^= 0x01
    raw[-1] ^= 0xff                                           # footer magic
    open(p + '.index', 'wb').write(bytes(raw))
    with pytest.raises(tf_bundle.BundleError, match='magic'):
        tf_bundle.read_index(p + '.index')
    raw[-1] ^= 0xff
    open(p + '.index', 'wb').write(bytes(raw))
    d = bytearray(open(p + '.data-00002-of-00003', 'rb').read())
    d[3] ^= 0x80
    open(p + '.data-00002-of-00003', 'wb').write(bytes(d))
    with pytest.raises(tf_bundle.BundleError, match='CRC'):
        tf_bundle.read_bundle(p, names=['b/vec_f32'], verify_data=True)


def test_native_and_python_crc_agree():
    rng = np.random.default_rng(1)
    for n in (0, 1, 7, 8, 9, 255, 256, 257, 4099, 100003):
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert tf_bundle.crc32c(b) == tf_bundle.crc32c_py(b), n
    b = rng.integers(0, 256, 5000, dtype=np.uint8).tobytes()
    assert tf_bundle.crc32c(b[1000:], tf_bundle.crc32c(b[:1000])) == tf_bundle.crc32c_py(b)


def test_product_writer_is_read_back_by_the_reader_and_matches_the_test_writer_semantics(tmp_path):
    """tf_bundle.write_bundle (used by tools/make_eval_model) and tests/tf_bundle_writer.py are independent
    implementations: both must give the reader the same tensors, across several index blocks."""
    rng = np.random.default_rng(6)
    vars_ = teacher_like_variables(rng, n_layers=12)
    vars_['misc/empty'] = np.zeros((0, 4), np.float32)
    vars_['misc/f16'] = rng.normal(0, 1, (5,)).astype(np.float16)
    a = tf_bundle.write_bundle(str(tmp_path / 'a.ckpt'), vars_, block_size=512)
    b = str(tmp_path / 'b.ckpt')
    write_bundle(b, vars_, num_shards=1, block_size=512)
    ga, gb = tf_bundle.read_bundle(a, verify_data=True), tf_bundle.read_bundle(b, verify_data=True)
    assert set(ga) == set(gb) == set(vars_)
    for k, v in vars_.items():
        assert ga[k].dtype == v.dtype and np.array_equal(ga[k], v) and np.array_equal(gb[k], v), k
    ha, ea = tf_bundle.read_index(a + '.index')
    hb, eb = tf_bundle.read_index(b + '.index')
    assert ha == hb and ea == eb                              # identical header and entry protos (offsets, CRCs)


def test_make_eval_model_strips_a_training_checkpoint_to_its_ema_shadows(tmp_path):
    """tools/make_eval_model.py of the reference (:8-34): EMA variables only, under their shadow names, state file,
    config json copied; the result loads through checkpoint.load_weights like the ns_wn-eval directories of the
    reference's Readme."""
    from nsynth_wavenet_b200.tools.make_eval_model import save_eval_model
    rng = np.random.default_rng(8)
    vars_ = teacher_like_variables(rng, n_layers=4)
    train = tmp_path / 'logs'
    train.mkdir()
    write_bundle(str(train / 'model.ckpt-1234'), vars_, num_shards=2, block_size=256)
    (train / 'checkpoint').write_text('model_checkpoint_path: "model.ckpt-1234"\n')
    (train / 'wavenet_mol.json').write_text('{"num_layers": 4}')
    out = tmp_path / 'eval'
    out.mkdir()
    (out / 'stale').write_text('x')                           # the tool starts from an empty directory (:9-11)
    prefix = save_eval_model(str(train), str(out))
    assert sorted(os.listdir(str(out))) == ['checkpoint', 'model.ckpt-1234.data-00000-of-00001',
                                            'model.ckpt-1234.index', 'wavenet_mol.json']
    assert open(str(out / 'checkpoint')).read() == 'model_checkpoint_path: "model.ckpt-1234"'
    _, entries = tf_bundle.read_index(prefix + '.index')
    assert entries and all('ExponentialMovingAverage' in k for k in entries)
    assert len(entries) == sum('ExponentialMovingAverage' in k for k in vars_)
    got = checkpoint.load_weights(str(out))
    for k, v in vars_.items():
        if k.endswith('/ExponentialMovingAverage'):
            assert np.array_equal(got[k[:-len('/ExponentialMovingAverage')]], v)
    assert tf_bundle.latest_checkpoint(str(out)) == prefix
