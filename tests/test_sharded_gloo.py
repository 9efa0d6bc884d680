"""CPU, world_size 2 over gloo: the multi-GPU path's host logic -- contiguous shards, no
collective on the data path, host gather in shard order.  The per-rank decode is the oracle
here (there is no GPU); on the GPU box each rank binds Code.decode of its own device."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from ldpc_ece535a.sharded import decode_sharded, shard_bounds


def test_shard_bounds():
    assert shard_bounds(10, 2) == [(0, 5), (5, 10)]
    assert shard_bounds(11, 4) == [(0, 3), (3, 6), (6, 9), (9, 11)]
    assert shard_bounds(1, 4) == [(0, 1), (1, 1), (1, 1), (1, 1)]
    assert shard_bounds(0, 8) == [(0, 0)] * 8
    for n in (0, 1, 7, 64, 1000):
        for w in (1, 2, 4, 8):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from oracle import oracle as O
    import util
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    codes = O.load_ref_codes()
    Hp, Lm, Um, _ = O.reorder_h(codes["shipped"]["H"])
    _, _, sym = util.synth_frames(Hp, Lm, Um, max(n_frames, 1), 3.0, seed=99)   # same on every rank
    sym = sym[:n_frames]

    def decode_fn(s):
        b, it, sy, _ = O.decode_frames(s, Hp, method=1, iterations=5, early_stop=True)
        return b, sy, it

    out = decode_sharded(decode_fn, sym.reshape(-1), n_frames, 64, dist=dist, dst=0)
    if rank == 0:
        if n_frames:
            whole = decode_fn(sym.reshape(-1))
            ok = all(np.array_equal(a, b) for a, b in zip(out, whole))
        else:
            ok = out is None
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [0, 1, 101, 256])
def test_two_rank_shards_gather_in_order(n_frames):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
