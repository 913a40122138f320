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
