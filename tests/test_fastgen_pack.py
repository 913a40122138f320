"""CPU check of the fastgen kernel's create-time repacking (nsw_fastgen_pack_host, a host-only
test hook): a NumPy emulation of the persistent kernel's phase algorithm -- one exchange per
layer via M_i = W2_i Wr_{i-1}, lazily updated residual/skip slices, past taps from history,
hoisted conditioning -- driven by the PACKED blocks must reproduce the oracle's teacher-forced
Fastgen outputs.  (Inter-CTA signalling itself can only be exercised on the GPU.)"""
import ctypes as C
import os

import numpy as np

from oracle import wavenet_oracle as O
from conftest import GOLDEN_DIR

OFF_D, OFF_L, OFF_S, OFF_C, OFF_P, BF = 0, 3072, 4096, 5120, 5128, 9224
NC = 128


def pack(hp, w):
    from nsynth_wavenet_b200 import _lib, engine
    lib = _lib.load()
    cfg = engine.wavenet_config(hp, engine='ffma')
    tensors, keep = _lib.make_tensors(w)
    sizes = (C.c_int64 * 4)()
    _lib.check(lib.nsw_fastgen_pack_host(C.byref(cfg), tensors, len(tensors), None, 0, None, None, sizes))
    assert (sizes[1], sizes[2]) == (BF, NC)
    blocks = np.empty(sizes[0], np.float32)
    N = sizes[3]
    cond_w = np.empty((256, N), np.float32)
    cond_b = np.empty(N, np.float32)
    _lib.check(lib.nsw_fastgen_pack_host(C.byref(cfg), tensors, len(tensors), blocks.ctypes.data,
                                         blocks.size, cond_w.ctypes.data, cond_b.ctypes.data, sizes))
    return blocks.reshape(-1, NC, BF).astype(np.float64), cond_w.astype(np.float64), cond_b.astype(np.float64)


def emulate(hp, w, blocks, cond_w, cond_b, enc, tf):
    L = hp.num_layers
    NPH = L + 2
    T = enc.shape[0]
    cond = enc @ cond_w + cond_b                     # hoisted, no centre trim
    wcs = w['conv_start/W'][0, :, 0, :].astype(np.float64)
    bcs = w['conv_start/biases'].astype(np.float64)
    wo2 = w['out2/W'][0, 0].astype(np.float64)
    bo2 = w['out2/biases'].astype(np.float64)
    dil = [None] + [2 ** (i % hp.num_stages) for i in range(L)]
    hist = [dict() for _ in range(L + 1)]
    cidx = np.arange(NC)
    outs = []
    x1 = x2 = 0.0
    xin = 0.0
    for t in range(T):
        l0 = wcs[2] * xin + wcs[1] * x1 + wcs[0] * x2 + bcs
        x2, x1 = x1, xin
        v_l = l0.copy()
        g = np.zeros(256)
        ls = l0.copy()
        sk = np.zeros(256)
        for ph in range(1, NPH + 1):
            blk = blocks[ph - 1]
            if 2 <= ph <= L:
                v_l = hist[ph - 1][t]                # l_{ph-2}, published in phase ph-1
            v = np.concatenate([v_l, g])
            D = np.einsum('cjk,k->cj', blk[:, OFF_D:OFF_D + 3072].reshape(NC, 4, 768), v)
            cst = blk[:, OFF_C:OFF_C + 8]
            if ph <= L:
                d_ = dil[ph]
                ls = ls + (np.einsum('cjk,k->cj', blk[:, OFF_L:OFF_L + 1024].reshape(NC, 4, 256), g)
                           + cst[:, :4]).reshape(-1)
                hv = np.concatenate([hist[ph].get(t - 2 * d_, np.zeros(512)),
                                     hist[ph].get(t - d_, np.zeros(512))])
                Pv = np.einsum('cjk,k->cj', blk[:, OFF_P:OFF_P + 4096].reshape(NC, 4, 1024), hv)
                n0 = (ph - 1) * 512 + 4 * cidx
                dd = D + Pv + np.stack([cond[t, n0 + j] for j in range(4)], 1)
                gn = (O.sigmoid(dd[:, :2]) * np.tanh(dd[:, 2:])).reshape(-1)   # channel 2c+j
                hist[ph][t] = ls.copy()
                S_ = blk[:, OFF_S:OFF_S + 1024].reshape(NC, 2, 512)
                if ph == 1:
                    sk = (np.einsum('cjk,k->cj', S_, v_l) + cst[:, 4:6]).reshape(-1)
                else:
                    sk = sk + (np.einsum('cjk,k->cj', S_[:, :, :256], g) + cst[:, 4:6]).reshape(-1)
                g = gn
            elif ph == L + 1:
                S_ = blk[:, OFF_S:OFF_S + 1024].reshape(NC, 2, 512)
                sk = np.maximum(sk + (np.einsum('cjk,k->cj', S_[:, :, :256], g) + cst[:, 4:6]).reshape(-1), 0)
                g = sk
            else:
                n0 = L * 512 + 2 * cidx
                h = np.maximum(D[:, :2] + np.stack([cond[t, n0], cond[t, n0 + 1]], 1), 0).reshape(-1)
                g = h
        outs.append(g @ wo2 + bo2)
        xin = float(tf[t])
    return np.stack(outs)


def test_packed_phase_algorithm_reproduces_oracle(teacher_hp):
    hp = teacher_hp
    w = O.init_teacher_weights(hp, seed=12345, bias_std=0.02)
    blocks, cond_w, cond_b = pack(hp, w)
    g = np.load(os.path.join(GOLDEN_DIR, 'fastgen_tf_1x96.npz'))
    T = 40   # covers dilations 1..16 twice over and the first uses of d = 32 history
    out = emulate(hp, w, blocks, cond_w, cond_b, g['enc'][0, :T].astype(np.float64), g['wav'][0, :T])
    err = np.abs(out - g['out'][0, :T]).max()
    assert err < 2e-5, err   # fp32-rounded M_i = W2 Wr and golden stored as fp32


def test_pack_rejects_unsupported_configs(teacher_hp):
    from argparse import Namespace
    from nsynth_wavenet_b200 import _lib, engine
    lib = _lib.load()
    w = {'conv_start/W': np.zeros((1, 3, 1, 512), np.float32)}
    tensors, keep = _lib.make_tensors(w)
    sizes = (C.c_int64 * 4)()
    ce = Namespace(**{**vars(teacher_hp), 'loss_type': 'ce', 'use_mu_law': True})
    rc = lib.nsw_fastgen_pack_host(C.byref(engine.wavenet_config(ce)), tensors, 1, None, 0, None, None, sizes)
    assert rc == -1 and b'mol / gauss' in lib.nsw_last_error()
    rc = lib.nsw_fastgen_pack_host(C.byref(engine.wavenet_config(teacher_hp)), tensors, 1, None, 0, None, None, sizes)
    assert rc == -3 and b'missing weight tensor' in lib.nsw_last_error()


def test_on_arrival_partition_and_butterfly_equal_the_row_dots(teacher_hp):
    """The poll group's on-arrival contraction (nsw_fastgen.cu, phases 2..L), restated on the packed block of one
    CTA: thread k holds entries {2k, 2k+1} of each third of the exchanged vector and the matching weights of the 8
    critical rows; a transposing butterfly leaves value (lane >> 2) & 7 in every lane; 4 warp partials are summed by
    the finalizer lanes.  Must equal the plain row dots (gate rows over [l | g], residual rows over g)."""
    hp = teacher_hp
    w = O.init_teacher_weights(hp, seed=12345, bias_std=0.02)
    blocks, _, _ = pack(hp, w)
    rng = np.random.default_rng(9)
    for ph, cta in ((2, 0), (7, 37), (hp.num_layers, 127)):
        blk = blocks[ph - 1][cta].astype(np.float64)
        v = rng.normal(0, 1, 768)
        k = np.arange(128)
        acc = np.zeros((128, 8))
        for j in range(4):
            for sgm in range(3):
                wj = np.stack([blk[OFF_D + j * 768 + 256 * sgm + 2 * k], blk[OFF_D + j * 768 + 256 * sgm + 2 * k + 1]], 1)
                e = np.stack([v[256 * sgm + 2 * k], v[256 * sgm + 2 * k + 1]], 1)
                acc[:, j] += (wj * e).sum(1)
            wr = np.stack([blk[OFF_L + j * 256 + 2 * k], blk[OFF_L + j * 256 + 2 * k + 1]], 1)
            acc[:, 4 + j] = (wr * np.stack([v[512 + 2 * k], v[512 + 2 * k + 1]], 1)).sum(1)
        part = np.zeros((4, 8))
        lane = np.arange(32)
        for wp in range(4):
            a8 = acc[32 * wp:32 * wp + 32]                                  # [lane, 8]
            h16, h8, h4 = (lane & 16) != 0, (lane & 8) != 0, (lane & 4) != 0
            b4 = np.zeros((32, 4))
            for i in range(4):
                mine = np.where(h16, a8[:, 4 + i], a8[:, i])
                theirs = np.where(h16, a8[:, i], a8[:, 4 + i])
                b4[:, i] = mine + theirs[lane ^ 16]
            b2 = np.zeros((32, 2))
            for i in range(2):
                mine = np.where(h8, b4[:, 2 + i], b4[:, i])
                theirs = np.where(h8, b4[:, i], b4[:, 2 + i])
                b2[:, i] = mine + theirs[lane ^ 8]
            mine = np.where(h4, b2[:, 1], b2[:, 0])
            theirs = np.where(h4, b2[:, 0], b2[:, 1])
            b1 = mine + theirs[lane ^ 4]
            b1 = b1 + b1[lane ^ 2]
            b1 = b1 + b1[lane ^ 1]
            for ln in range(0, 32, 4):
                part[wp, (ln >> 2) & 7] = b1[ln]
            assert all(np.allclose(b1[ln], b1[ln | 3]) for ln in range(0, 32, 4))   # replicated over 4 lanes
        tot = part.sum(0)
        rows_d = blk[OFF_D:OFF_D + 3072].reshape(4, 768) @ v
        rows_l = blk[OFF_L:OFF_L + 1024].reshape(4, 256) @ v[512:]
        assert np.abs(tot[:4] - rows_d).max() < 1e-10 and np.abs(tot[4:] - rows_l).max() < 1e-10
