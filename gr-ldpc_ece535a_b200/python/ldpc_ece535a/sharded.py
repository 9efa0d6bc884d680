"""Independent-codeword sharding across the GPUs of one box: contiguous frame ranges per rank,
no data-path collective, outputs gathered on the host in shard order.

One process per GPU (torch.distributed, NCCL on the GPU box / gloo in CPU tests).  The only
communication is the final gather of the decoded bytes to rank 0, which is host traffic:
the kernels never exchange anything."""
import numpy as np


def shard_bounds(n_units, world):
    """Contiguous ranges of ceil(n/world) units: [(lo, hi)] * world (empty ranges at the tail)."""
    per = -(-n_units // world) if n_units else 0
    return [(min(r * per, n_units), min((r + 1) * per, n_units)) for r in range(world)]


def decode_sharded(decode_fn, sym, n_frames, N, dist=None, dst=0):
    """Decode frames [lo, hi) of `sym` (complex64, n_frames * N) on this rank with
    decode_fn(sym_shard) -> (bytes, synd, iters) and gather every rank's outputs on `dst`.

    Returns (bytes, synd, iters) for all n_frames on rank `dst`, None elsewhere.
    decode_fn is `Code.decode` bound to this rank's GPU on the GPU box."""
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    lo, hi = shard_bounds(n_frames, world)[rank]
    sym = np.asarray(sym, np.complex64).reshape(-1)
    part = decode_fn(sym[lo * N:hi * N]) if hi > lo else None
    if dist is None:
        return part
    parts = [None] * world if rank == dst else None
    dist.gather_object(part, parts, dst=dst)
    if rank != dst:
        return None
    parts = [p for p in parts if p is not None]
    if not parts:
        return None
    return tuple(np.concatenate([np.asarray(p[i]) for p in parts]) for i in range(3))
