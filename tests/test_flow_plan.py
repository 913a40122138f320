"""Host-side check of the persistent IAF flow kernel's work split and publish / read plan (engine tc3).

nsw_flow_plan_host runs the SAME integer functions the CUDA kernel uses (ft_range_of, ft_published, ft_tap_source in
nsynth_wavenet_b200/csrc/nsw_iaf_flow_tc.cu) on the CPU.  The kernel has no grid barrier: a CTA reads another CTA's
rows only through tiles that CTA decided to publish, so "every foreign read has a publish" is the property that keeps
it from dead-locking (a reader polling a flag nobody raises) or from reading stale rows.
"""
import ctypes as C

import numpy as np
import pytest

BM = 128


def plan(T, nclips, num_sms, l0, l1, num_stages, fuse_head):
    from nsynth_wavenet_b200 import _lib
    lib = _lib.load()
    n = C.c_int64(0)
    grid = lib.nsw_flow_plan_host(T, nclips, num_sms, l0, l1, num_stages, fuse_head, None, 0, C.byref(n))
    assert grid > 0, _lib.load().nsw_last_error()
    buf = np.zeros(n.value, np.int32)
    grid2 = lib.nsw_flow_plan_host(T, nclips, num_sms, l0, l1, num_stages, fuse_head,
                                   buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(n))
    assert grid2 == grid and n.value == buf.size
    ctas, pos, nl = [], 0, l1 - l0
    for _ in range(grid):
        clip, tk0, K, n_c = (int(v) for v in buf[pos:pos + 4])
        pos += 4
        rec = buf[pos:pos + nl * K * 3].reshape(nl, K, 3)
        pos += nl * K * 3
        ctas.append(dict(clip=clip, tk0=tk0, K=K, n_c=n_c, rec=rec))
    assert pos == buf.size
    return ctas


SHAPES = [
    # T, clips, SMs, l0, l1, stages, fused head
    (7680, 8, 148, 0, 10, 10, 1),     # the benchmark shape, 10-layer flow
    (7680, 8, 148, 0, 30, 10, 1),     # 30-layer flow
    (7680, 8, 148, 3, 10, 10, 0),     # second launch of a flow split by a debug tap
    (4096, 1, 148, 0, 10, 10, 1),     # BASELINE config 1: 32 CTAs of one tile
    (512, 1, 148, 0, 10, 10, 1),
    (1024, 74, 148, 0, 10, 10, 1),    # as many clips as one launch takes
    (2560, 3, 148, 0, 30, 10, 1),
    (31744, 2, 148, 0, 30, 10, 1),    # 248 tiles per clip: 3 or 4 tiles per CTA
    (75776, 1, 148, 0, 10, 10, 0),    # the longest clip one launch can hold (592 tiles)
    (7680, 8, 132, 0, 10, 10, 1),     # another SM count
]


@pytest.mark.parametrize('T,nclips,sms,l0,l1,stages,fuse', SHAPES)
def test_ranges_partition_every_clip(T, nclips, sms, l0, l1, stages, fuse):
    ctas = plan(T, nclips, sms, l0, l1, stages, fuse)
    tiles = T // BM
    assert len(ctas) <= sms
    seen = np.zeros((nclips, tiles), np.int32)
    for c in ctas:
        assert 1 <= c['K'] <= 4, c
        assert 0 <= c['clip'] < nclips and c['tk0'] >= 0 and c['tk0'] + c['K'] <= tiles   # never straddles a clip
        seen[c['clip'], c['tk0']:c['tk0'] + c['K']] += 1
    assert np.all(seen == 1)                                                           # every tile exactly once
    for clip in range(nclips):                                                         # consecutive in time
        mine = sorted((c['tk0'], c['K']) for c in ctas if c['clip'] == clip)
        assert len(mine) == [c['n_c'] for c in ctas if c['clip'] == clip][0]
        t = 0
        for tk0, K in mine:
            assert tk0 == t
            t += K


@pytest.mark.parametrize('T,nclips,sms,l0,l1,stages,fuse', SHAPES)
def test_every_foreign_read_has_a_publish(T, nclips, sms, l0, l1, stages, fuse):
    ctas = plan(T, nclips, sms, l0, l1, stages, fuse)
    tiles, nl = T // BM, l1 - l0
    owner = {}
    for i, c in enumerate(ctas):
        for k in range(c['K']):
            owner[(c['clip'], c['tk0'] + k)] = (i, k)
    reach = max(1, 2 * (1 << (stages - 1)) // BM)
    n_foreign = 0
    for i, c in enumerate(ctas):
        for li in range(nl):
            d = 1 << ((l0 + li) % stages)
            for k in range(c['K']):
                for tap in (0, 1):
                    src = int(c['rec'][li, k, 1 + tap])
                    tk = c['tk0'] + k
                    want = tk - (2 - tap) * d // BM if 2 * d > BM else None
                    if src == -2:        # causal zeros: the window lies before the clip
                        assert (want is not None and want < 0) or (want is None and k == 0 and c['tk0'] == 0)
                        continue
                    if src == -1:        # own shared memory
                        if want is not None:
                            assert c['tk0'] <= want <= tk
                        else:
                            assert k >= 1
                        continue
                    # foreign tile of the same clip, strictly earlier in time
                    assert 0 <= src < c['tk0']
                    if want is not None:
                        assert src == want
                    else:
                        assert k == 0 and src == c['tk0'] - 1 and 2 * d <= BM
                    n_foreign += 1
                    j, kj = owner[(c['clip'], src)]
                    assert j != i
                    if li >= 1:          # layer li reads layer li-1's output; li == 0 reads the start conv rows
                        assert ctas[j]['rec'][li - 1, kj, 0] == 1, 'tile read by CTA %d is never published' % i
                    # the publisher only waits for the "consumed" counters of the CTAs owning the next `reach`
                    # tiles before it overwrites a published tile: every reader must be among them
                    last_tk = ctas[j]['tk0'] + ctas[j]['K'] - 1
                    assert c['tk0'] <= min(last_tk + reach, tiles - 1)
    if nclips * tiles > len(ctas) or any(c['tk0'] > 0 for c in ctas):
        assert n_foreign > 0
    # the last layer is published for the separate head kernel unless the head is fused
    for c in ctas:
        assert np.all(c['rec'][nl - 1, :, 0] == (0 if fuse else 1))


def test_plan_rejects_what_one_launch_cannot_hold():
    from nsynth_wavenet_b200 import _lib
    lib = _lib.load()
    n = C.c_int64(0)
    assert lib.nsw_flow_plan_host(75776 + 128, 1, 148, 0, 10, 10, 1, None, 0, C.byref(n)) < 0   # 593 tiles > 4 x 148
    assert lib.nsw_flow_plan_host(7680, 10, 148, 0, 10, 10, 1, None, 0, C.byref(n)) < 0         # 10 x 15 CTAs > 148
    assert lib.nsw_flow_plan_host(7680 + 64, 1, 148, 0, 10, 10, 1, None, 0, C.byref(n)) < 0      # not whole tiles


# ---------------------------------------------------------------------------------------------------------------------
# the CTA-pair kernel (nsw_iaf_flow_pair.cu): same properties on its own split (fp_range_of / fp_owner / fp_published)
def pair_plan(T, nclips, max_pairs, nl, num_stages):
    from nsynth_wavenet_b200 import _lib
    lib = _lib.load()
    n = C.c_int64(0)
    grid = lib.nsw_flow_pair_plan_host(T, nclips, max_pairs, nl, num_stages, None, 0, C.byref(n))
    if grid <= 0:
        return None
    buf = np.zeros(n.value, np.int32)
    assert lib.nsw_flow_pair_plan_host(T, nclips, max_pairs, nl, num_stages, buf.ctypes.data_as(C.c_void_p), buf.size,
                                       C.byref(n)) == grid
    ctas, pos = [], 0
    for _ in range(grid):
        clip, tk0, K, n_c, far = (int(v) for v in buf[pos:pos + 5])
        pos += 5
        rec = buf[pos:pos + nl * K * 3].reshape(nl, K, 3)
        pos += nl * K * 3
        ctas.append(dict(clip=clip, tk0=tk0, K=K, n_c=n_c, far=far, rec=rec))
    assert pos == buf.size
    return ctas


PAIR_SHAPES = [
    # T, clips, CTA pairs on the device, layers, stages
    (7680, 8, 74, 10, 10),     # the benchmark shape: 9 pairs per clip, 3 + 3 or 4 + 4 tiles
    (7680, 8, 72, 30, 10),
    (7680, 7, 74, 30, 10),     # the distillation shard
    (4096, 1, 74, 10, 10),     # one tile per CTA
    (3584, 8, 74, 10, 10),     # mixed 1 / 2 tiles per CTA
    (1024, 2, 74, 30, 10),
    (512, 1, 74, 10, 10),
    (7680, 8, 66, 10, 10),     # another device size: 8 pairs per clip, 4 + 4 and 3 + 3
]


@pytest.mark.parametrize('T,nclips,pairs,nl,stages', PAIR_SHAPES)
def test_pair_kernel_plan(T, nclips, pairs, nl, stages):
    ctas = pair_plan(T, nclips, pairs, nl, stages)
    assert ctas is not None
    tiles = T // BM
    assert len(ctas) <= 2 * pairs and len(ctas) % 2 == 0
    seen = np.zeros((nclips, tiles), np.int32)
    for i, c in enumerate(ctas):
        assert 1 <= c['K'] <= 4 and c['tk0'] + c['K'] <= tiles
        seen[c['clip'], c['tk0']:c['tk0'] + c['K']] += 1
        if i % 2 == 0:                                   # leader and peer: same clip, same tile count, consecutive tiles
            q = ctas[i + 1]
            assert q['clip'] == c['clip'] and q['K'] == c['K'] and q['tk0'] == c['tk0'] + c['K']
    assert np.all(seen == 1)
    owner = {}
    for i, c in enumerate(ctas):
        for k in range(c['K']):
            owner[(c['clip'], c['tk0'] + k)] = (i, k)
    first_of_clip = {}
    for i, c in enumerate(ctas):
        first_of_clip.setdefault(c['clip'], i)
    n_foreign = 0
    for i, c in enumerate(ctas):
        for li in range(nl):
            d = 1 << (li % stages)
            for k in range(c['K']):
                for tap in (0, 1):
                    src = int(c['rec'][li, k, 1 + tap])
                    if src < 0:
                        continue
                    assert 0 <= src < c['tk0']
                    if 2 * d > BM:
                        assert src == c['tk0'] + k - (2 - tap) * d // BM
                    else:
                        assert k == 0 and src == c['tk0'] - 1
                    n_foreign += 1
                    j, kj = owner[(c['clip'], src)]
                    assert j != i
                    if li >= 1:
                        assert ctas[j]['rec'][li - 1, kj, 0] == 1, 'tile read by CTA %d is never published' % i
                    # the publisher of CTA j polls the consumed counters of CTAs (j, far_j] of its clip: the reader is one
                    assert j < i <= first_of_clip[c['clip']] + ctas[j]['far']
    if len(ctas) > nclips:
        assert n_foreign > 0
    for c in ctas:                                       # the fused head consumes the last layer in place
        assert np.all(c['rec'][nl - 1, :, 0] == 0)


def test_pair_kernel_refuses_what_it_does_not_cover():
    assert pair_plan(7680 + 128, 8, 74, 10, 10) is None        # odd number of tiles per clip
    assert pair_plan(7680, 64, 74, 10, 10) is None             # 64 clips: one pair per clip would need 30 tiles per CTA
    assert pair_plan(154112, 1, 74, 10, 10) is None            # 1204 tiles: more than 8 per pair of ... 72 pairs
