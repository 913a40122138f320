"""Kernel-level parity of the tcgen05 conv-GEMM (`nsw_conv_gemm_device`, include/nsw.h): the contraction every dense op of
the path runs on -- masked.conv1d as a GEMM (masked.py:160-232: taps t-2d, t-d, t with zero history), the 1x1
projections, and the two operand routes added for the teacher (a second source in front of K, an accumulate source
through identity k-blocks).  Reference: the same sums in NumPy float64.  Bar: 1e-5 relative to the largest output
(split fp16 x 3 products, fp32 accumulation), for both the CTA-pair kernel and the one-CTA kernel, on shapes that leave
ragged tiles everywhere: odd tile counts (a phantom m-tile in the last pair), rows % 128 != 0, N % 256 != 0."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def reference(x, w, bias, ntaps, a_off, stride, mclip, x2=None, a_off2=0, y=None):
    nclips, L, cin = x.shape
    N = w.shape[1]
    out = np.zeros((nclips, mclip, N), np.float64)
    for tap in range(ntaps):
        wt = w[tap * cin:(tap + 1) * cin].astype(np.float64)
        for m in range(mclip):
            f = m + a_off + tap * stride
            if 0 <= f < L:
                out[:, m] += x[:, f].astype(np.float64) @ wt
    if x2 is not None:
        w2 = w[ntaps * cin:].astype(np.float64)
        for m in range(mclip):
            f = m + a_off2
            if 0 <= f < x2.shape[1]:
                out[:, m] += x2[:, f].astype(np.float64) @ w2
    if bias is not None:
        out += bias.astype(np.float64)
    if y is not None:
        out += y.reshape(nclips, mclip, N).astype(np.float64)
    return out


def run(x, w, bias, ntaps, a_off, stride, mclip, x2=None, a_off2=0, y=None, flags=0):
    from nsynth_wavenet_b200 import _lib
    lib = _lib.load()
    dev = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()
    ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    dx, dw, db, dx2, dy = dev(x), dev(w), dev(bias), dev(x2), dev(y)
    nclips, L, cin = x.shape
    N = w.shape[1]
    out = torch.full((nclips * mclip, N), float('nan'), device='cuda')
    rc = lib.nsw_conv_gemm_device(ptr(dx), nclips, L, cin, ntaps, a_off, stride, mclip, ptr(dw), N, ptr(db), ptr(dx2),
                                  0 if x2 is None else x2.shape[1], 0 if x2 is None else x2.shape[2], a_off2, ptr(dy),
                                  flags, ptr(out), None)
    assert rc == 0, lib.nsw_last_error()
    torch.cuda.synchronize()
    return out.cpu().numpy().reshape(nclips, mclip, N)


CASES = [
    # nclips, L, cin, ntaps, a_off, stride, mclip, N, second (L2, cin2, a_off2) or None, accumulate
    dict(nclips=1, L=300, cin=64, ntaps=1, a_off=0, stride=1, mclip=300, N=320, second=None, acc=False),   # 3 m-tiles: phantom
    dict(nclips=3, L=200, cin=128, ntaps=3, a_off=-10, stride=5, mclip=200, N=256, second=None, acc=False),  # dilated, zero history
    dict(nclips=2, L=260, cin=64, ntaps=3, a_off=-4, stride=2, mclip=256, N=512, second=(300, 64, 30), acc=False),
    dict(nclips=2, L=256, cin=256, ntaps=1, a_off=0, stride=1, mclip=256, N=768, second=None, acc=True),    # teacher res+skip
    dict(nclips=1, L=384, cin=192, ntaps=2, a_off=-3, stride=3, mclip=384, N=576, second=(400, 128, 7), acc=True),
]


@pytest.mark.timeout(600)
@pytest.mark.parametrize('flags', [0, 1, 2], ids=['pair', 'pair-split-acc', 'one-cta'])
@pytest.mark.parametrize('case', range(len(CASES)))
def test_conv_gemm_matches_float64(case, flags):
    c = CASES[case]
    if flags == 2 and (c['second'] or c['acc']):
        pytest.skip('the one-CTA kernel has no second / accumulate source')
    rng = np.random.default_rng(100 + case)
    x = rng.normal(0, 1.0, (c['nclips'], c['L'], c['cin'])).astype(np.float32)
    K = c['ntaps'] * c['cin'] + (c['second'][1] if c['second'] else 0)
    w = rng.normal(0, 0.05, (K, c['N'])).astype(np.float32)
    bias = rng.normal(0, 0.3, (c['N'],)).astype(np.float32)
    x2 = rng.normal(0, 1.0, (c['nclips'], c['second'][0], c['second'][1])).astype(np.float32) if c['second'] else None
    a2 = c['second'][2] if c['second'] else 0
    y = rng.normal(0, 2.0, (c['nclips'] * c['mclip'], c['N'])).astype(np.float32) if c['acc'] else None
    ref = reference(x, w, bias, c['ntaps'], c['a_off'], c['stride'], c['mclip'], x2, a2, y)
    got = run(x, w, bias, c['ntaps'], c['a_off'], c['stride'], c['mclip'], x2, a2, y, flags)
    assert np.all(np.isfinite(got))
    err = np.abs(got - ref).max()
    print('case', case, 'flags', flags, 'max-abs err', err, 'max |ref|', np.abs(ref).max())
    assert err <= 1e-5 * max(1.0, np.abs(ref).max()), err


@pytest.mark.timeout(300)
def test_conv_gemm_small_n_takes_the_one_cta_kernel_and_rejects_bad_shapes():
    from nsynth_wavenet_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(7)
    x = rng.normal(0, 1, (2, 130, 64)).astype(np.float32)
    w = rng.normal(0, 0.05, (64, 128)).astype(np.float32)               # N = 128: no pair items
    ref = reference(x, w, None, 1, 0, 1, 130)
    got = run(x, w, None, 1, 0, 1, 130)
    assert np.abs(got - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())
    # N = 128 with an accumulate source needs the pair kernel: refused loudly, not computed some other way
    dx = torch.from_numpy(x).cuda(); dw = torch.from_numpy(w).cuda()
    dy = torch.zeros((260, 128), device='cuda'); out = torch.zeros((260, 128), device='cuda')
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = lib.nsw_conv_gemm_device(p(dx), 2, 130, 64, 1, 0, 1, 130, p(dw), 128, None, None, 0, 0, 0, p(dy), 0, p(out), None)
    assert rc != 0 and b'pair kernel' in lib.nsw_last_error()
    rc = lib.nsw_conv_gemm_device(p(dx), 2, 130, 64, 1, 0, 1, 130, p(dw), 100, None, None, 0, 0, 0, None, 0, p(out), None)
    assert rc != 0
