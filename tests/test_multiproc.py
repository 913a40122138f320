"""N > 1 host logic on CPU: world_size-2 gloo processes (127.0.0.1 rendezvous)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nsynth_wavenet_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rng = np.random.default_rng(123)            # same shapes everywhere
    w = {'iaf_1/out1/W': rng.normal(size=(1, 1, 4, 4)).astype(np.float32),
         'iaf_1/out1/biases': rng.normal(size=4).astype(np.float32),
         'a/scalar': np.float32(3.0).reshape(())}
    if rank != 0:
        w = {k: np.zeros_like(v) for k, v in w.items()}   # must be overwritten by rank 0's
    got = parallel.broadcast_weights(w)
    clips = parallel.shard_clips(13, rank, world)
    slowest = parallel.max_over_ranks(10.0 + rank)
    q.put((rank, {k: v.copy() for k, v in got.items()}, clips, slowest))
    dist.barrier()
    dist.destroy_process_group()


def test_weight_broadcast_sharding_and_max_timing_two_ranks():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = np.random.default_rng(123)
    w0 = ref.normal(size=(1, 1, 4, 4)).astype(np.float32)
    for rank, w, clips, slowest in res:
        assert np.array_equal(w['iaf_1/out1/W'], w0)          # rank 1 received rank 0's values
        assert w['a/scalar'] == 3.0 and w['a/scalar'].shape == ()
        assert slowest == 11.0                                 # max over ranks
    all_clips = sorted(res[0][2] + res[1][2])
    assert all_clips == list(range(13)) and not set(res[0][2]) & set(res[1][2])


def test_flatten_roundtrip():
    w = {'b': np.arange(6, dtype=np.float32).reshape(2, 3), 'a': np.ones(2, np.float32)}
    flat, meta = parallel.flatten_weights(w)
    back = parallel.unflatten_weights(flat, meta)
    assert list(back) == ['a', 'b'] and np.array_equal(back['b'], w['b'])
    assert parallel.broadcast_weights(w) is w   # no process group: identity
