"""CPU check of the batched fastgen engine's create-time repacking (nsw_fastgen_gn_pack_host, a host-only test
hook): a NumPy emulation of the kernel's L+3 phase algorithm -- one exchange per layer via M_i = W2_i Wr_{i-1},
residual / skip slices owned per CTA, past taps from history, hoisted conditioning, out1 / out2 rows spread over
the CTAs -- driven by the PACKED blocks must reproduce the oracle's teacher-forced Fastgen outputs, for gate 512
(wavenet_mol.json) and for the double-gate mu-law CE model (wavenet_ce.json).  The grid barrier itself can only be
exercised on the GPU (tests/test_fastgen_gn_gpu.py)."""
import ctypes as C
from argparse import Namespace

import numpy as np
import pytest

from oracle import wavenet_oracle as O
from conftest import load_hparams

NC, W, S = 128, 512, 256


def layout(mh):
    ppc = mh // NC
    nd = 2 * ppc
    k1 = W + mh
    off_d = 0
    off_p = off_d + nd * k1
    off_l = off_p + nd * 2 * W
    off_s = off_l + 4 * mh
    off_c = off_s + 2 * W
    return dict(MH=mh, PPC=ppc, nD=nd, K1=k1, off_d=off_d, off_p=off_p, off_l=off_l, off_s=off_s, off_c=off_c,
                BF=off_c + 16)


def pack(hp, w):
    from nsynth_wavenet_b200 import _lib, engine
    lib = _lib.load()
    cfg = engine.wavenet_config(hp, engine='ffma')
    tensors, keep = _lib.make_tensors(w)
    sizes = (C.c_int64 * 6)()
    _lib.check(lib.nsw_fastgen_gn_pack_host(C.byref(cfg), tensors, len(tensors), None, 0, None, None, sizes))
    lay = layout(int(sizes[4]))
    assert (sizes[1], sizes[2], sizes[5]) == (lay['BF'], NC, hp.num_layers + 3)
    blocks = np.empty(sizes[0], np.float32)
    N = int(sizes[3])
    cond_w = np.empty((256, N), np.float32)
    cond_b = np.empty(N, np.float32)
    _lib.check(lib.nsw_fastgen_gn_pack_host(C.byref(cfg), tensors, len(tensors), blocks.ctypes.data, blocks.size,
                                            cond_w.ctypes.data, cond_b.ctypes.data, sizes))
    return (blocks.reshape(-1, NC, lay['BF']).astype(np.float64), cond_w.astype(np.float64),
            cond_b.astype(np.float64), lay)


def emulate(hp, w, blocks, cond_w, cond_b, lay, enc, fed):
    """fed[t] = the (input-encoded) sample fed at step t+1; returns out[T, O]."""
    L, O_ = hp.num_layers, O.teacher_out_width(hp)
    MH, PPC, nD, K1 = lay['MH'], lay['PPC'], lay['nD'], lay['K1']
    G = 2 * MH
    T = enc.shape[0]
    cond = enc @ cond_w + cond_b
    wcs = w['conv_start/W'][0, :, 0, :].astype(np.float64)
    bcs = w['conv_start/biases'].astype(np.float64)
    dil = [None] + [2 ** (i % hp.num_stages) for i in range(L)]
    hist = [dict() for _ in range(L + 1)]
    cidx = np.arange(NC)
    RO = (O_ + NC - 1) // NC
    outs = []
    x1 = x2 = xin = 0.0
    for t in range(T):
        l0 = wcs[2] * xin + wcs[1] * x1 + wcs[0] * x2 + bcs
        x2, x1 = x1, xin
        g = np.zeros(MH)
        lst = np.zeros((NC, 4))
        sst = np.zeros((NC, 2))
        for ph in range(1, L + 4):
            blk = blocks[ph - 1]
            cst = blk[:, lay['off_c']:lay['off_c'] + 16]
            Drows = blk[:, lay['off_d']:lay['off_d'] + nD * K1].reshape(NC, nD, K1)
            Srows = blk[:, lay['off_s']:lay['off_s'] + 2 * W].reshape(NC, 2, W)
            if ph <= L:
                v_l = l0 if ph == 1 else hist[ph - 1][t]           # l_{ph-2}
                v = np.concatenate([v_l, g])
                D = np.einsum('cjk,k->cj', Drows, v)
                d_ = dil[ph]
                hv = np.concatenate([hist[ph].get(t - 2 * d_, np.zeros(W)), hist[ph].get(t - d_, np.zeros(W))])
                Pv = np.einsum('cjk,k->cj', blk[:, lay['off_p']:lay['off_p'] + nD * 2 * W].reshape(NC, nD, 2 * W), hv)
                n0 = (ph - 1) * G + cidx * nD
                dd = D + Pv + np.stack([cond[t, n0 + j] for j in range(nD)], 1)
                gn = (O.sigmoid(dd[:, :PPC]) * np.tanh(dd[:, PPC:])).reshape(-1)   # pair c*PPC + j
                if ph == 1:
                    lst = l0.reshape(NC, 4).copy()
                    sst = np.einsum('cjk,k->cj', Srows, l0) + cst[:, 4:6]
                else:
                    Lrows = blk[:, lay['off_l']:lay['off_l'] + 4 * MH].reshape(NC, 4, MH)
                    lst = lst + np.einsum('cjk,k->cj', Lrows, g) + cst[:, :4]
                    sst = sst + np.einsum('cjk,k->cj', Srows[:, :, :MH], g) + cst[:, 4:6]
                hist[ph][t] = lst.reshape(-1).copy()
                g = gn
            elif ph == L + 1:
                sst = sst + np.einsum('cjk,k->cj', Srows[:, :, :MH], g) + cst[:, 4:6]
                s = np.maximum(sst.reshape(-1), 0)
            elif ph == L + 2:
                n0 = L * G + 2 * cidx
                h = np.maximum(np.einsum('cjk,k->cj', Drows[:, :2, :S], s) +
                               np.stack([cond[t, n0], cond[t, n0 + 1]], 1), 0).reshape(-1)
            else:
                o = (np.einsum('cjk,k->cj', Drows[:, :RO, :S], h) + cst[:, 6:6 + RO]).reshape(-1)[:O_]
        outs.append(o)
        xin = float(fed[t])
    return np.stack(outs)


def test_gn_packed_phase_algorithm_reproduces_oracle_gate512(teacher_hp):
    hp = teacher_hp
    w = O.init_teacher_weights(hp, seed=12345, bias_std=0.02)
    blocks, cond_w, cond_b, lay = pack(hp, w)
    assert lay['MH'] == 256
    rng = np.random.default_rng(3)
    T = 24
    enc = rng.uniform(-1, 1, (1, T, 256))
    wav = rng.uniform(-0.5, 0.5, (1, T))
    ref = O.fastgen_run(w, hp, enc, np.float64, teacher_force=wav)['out'][0]
    out = emulate(hp, w, blocks, cond_w, cond_b, lay, enc[0], wav[0])
    err = np.abs(out - ref).max()
    assert err < 2e-6, err   # only the fp32 rounding of M_i = W2 Wr and of the folded biases


@pytest.mark.timeout(600)
def test_gn_packed_phase_algorithm_reproduces_oracle_ce_double_gate():
    hp = load_hparams('wavenet_ce.json')
    hp = Namespace(**{**vars(hp), 'num_layers': 12})   # gate 1024, 256-way head, mu-law; fewer layers keep the CPU test short
    w = O.init_teacher_weights(hp, seed=5, bias_std=0.02)
    blocks, cond_w, cond_b, lay = pack(hp, w)
    assert lay['MH'] == 512 and O.teacher_out_width(hp) == 256
    rng = np.random.default_rng(4)
    T = 20
    enc = rng.uniform(-1, 1, (1, T, 256))
    codes = rng.integers(-128, 128, (1, T))
    wav = O.inv_mu_law(codes).astype(np.float64)       # what fastgen.synthesis feeds back (fastgen.py:163-164)
    ref = O.fastgen_run(w, hp, enc, np.float64, teacher_force=wav)['out'][0]
    fed = O.mu_law(wav[0]) / 128.0                     # wavenet.py:411-414
    out = emulate(hp, w, blocks, cond_w, cond_b, lay, enc[0], fed)
    err = np.abs(out - ref).max()
    assert err < 2e-6, err


def test_gn_pack_rejects_what_it_cannot_run(teacher_hp):
    from nsynth_wavenet_b200 import _lib, engine
    lib = _lib.load()
    w = {'conv_start/W': np.zeros((1, 3, 1, 512), np.float32)}
    tensors, keep = _lib.make_tensors(w)
    sizes = (C.c_int64 * 6)()
    ce16 = Namespace(**{**vars(teacher_hp), 'loss_type': 'ce', 'use_mu_law': False})   # 65536-way softmax
    rc = lib.nsw_fastgen_gn_pack_host(C.byref(engine.wavenet_config(ce16)), tensors, 1, None, 0, None, None, sizes)
    assert rc == -1 and b'65536-way' in lib.nsw_last_error()
    rc = lib.nsw_fastgen_gn_pack_host(C.byref(engine.wavenet_config(teacher_hp)), tensors, 1, None, 0, None, None, sizes)
    assert rc == -3 and b'missing weight tensor' in lib.nsw_last_error()
