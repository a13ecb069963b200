"""Multi-GPU sharding of the candidate sweep: one process per GPU, ``torch.distributed`` (NCCL over
NVLink/NVSwitch on the GPU box, gloo in the CPU tests) for the plumbing.

The path shards naturally: candidate rows are independent given the posterior state, so
each rank sweeps a contiguous block of rows with NO data-path collective.  The only exchanges are
  * after a refit: broadcast of the posterior state (X, alpha, L^-1, W^-1, hyper-parameters, fmin) from the
    fitting rank -- one NCCL broadcast per buffer, straight between the engines' own HBM buffers;
  * after a sweep: all-gather of the per-rank top-k records {score, global index} (k*16 bytes per rank) and an
    identical deterministic merge on every rank (ties -> lowest global index).
With a device-generated design (``acq_topk_uniform``) the candidates themselves never exist outside the GPUs: every
rank generates its own rows of ONE global counter-based stream, so the union over ranks is the same design for any
number of GPUs.
"""
import numpy as np


def shard_rows(N, rank, world):
    """Contiguous block [lo, hi) of N rows owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(int(N), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def merge_topk(records, k):
    """Merge per-rank (values, global_indices) records: k smallest values, ties by lowest global index.
    Independent of the order of `records`."""
    vals = np.concatenate([np.asarray(v, dtype=np.float64).reshape(-1) for v, _ in records])
    idx = np.concatenate([np.asarray(i, dtype=np.int64).reshape(-1) for _, i in records])
    keep = idx >= 0
    vals, idx = vals[keep], idx[keep]
    order = np.lexsort((idx, vals))[:k]
    return vals[order], idx[order]


def _dist():
    import torch.distributed as dist
    return dist


def allgather_topk(vals, idx, k, device=None, group=None):
    """All-gather fixed-size top-k records from every rank and merge them identically everywhere."""
    import torch
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return merge_topk([(vals, idx)], k)
    world = dist.get_world_size(group)
    rec = torch.full((k, 2), float('nan'), dtype=torch.float64)
    n = min(k, len(vals))
    rec[:n, 0] = torch.from_numpy(np.asarray(vals[:n], dtype=np.float64))
    # indices travel as float64 payload bit-cast from int64 so one collective carries the record
    ibuf = torch.full((k,), -1, dtype=torch.int64)
    ibuf[:n] = torch.from_numpy(np.asarray(idx[:n], dtype=np.int64))
    rec[:, 1] = ibuf.view(torch.float64)
    if device is not None:
        rec = rec.to(device)
    out = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(out, rec, group=group)
    records = []
    for o in out:
        o = o.cpu()
        records.append((o[:, 0].numpy(), o[:, 1].contiguous().view(torch.int64).numpy()))
    return merge_topk(records, k)


def allgather_topk_device(vals_dev, idx_dev, row_offset, k, group=None, negate=False):
    """Device-resident form of :func:`allgather_topk`: `vals_dev` (float64 [k]) and `idx_dev` (int64 [k], -1 = empty) are
    CUDA tensors as ``Engine.topk(..., out=...)`` leaves them.  The k records are packed on the device, exchanged with
    ONE collective (NCCL all_gather over NVLink) and read back with ONE device-to-host copy of world*k*16 bytes -- the
    only synchronisation of the step -- then merged identically on every rank.  ``negate``: the records hold f and the
    merge wants the score -f (smallest first)."""
    import torch
    dist = _dist()
    rec = torch.empty((k, 2), dtype=torch.float64, device=vals_dev.device)
    rec[:, 0] = -vals_dev if negate else vals_dev
    gidx = torch.where(idx_dev >= 0, idx_dev + int(row_offset), idx_dev)
    rec[:, 1] = gidx.view(torch.float64)   # int64 payload bit-cast so that one collective carries the record
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        out = torch.empty((dist.get_world_size(group) * k, 2), dtype=torch.float64, device=vals_dev.device)
        dist.all_gather_into_tensor(out, rec, group=group)     # rank r's records are rows [r k, (r + 1) k)
    else:
        out = rec
    host = out.cpu().reshape(-1, k, 2)
    records = [(host[r, :, 0].numpy(), host[r, :, 1].contiguous().view(torch.int64).numpy()) for r in range(host.shape[0])]
    return merge_topk(records, k)


def state_checksums(engine):
    """int64 wrap-around sum of every replicated state buffer (bit pattern), as a CUDA tensor: equal on two ranks iff
    (up to collisions) the replicas hold identical bits."""
    import torch
    sums = []
    for ptr, nbytes in engine.state_buffers():
        t = torch.as_tensor(_DevBuf(ptr, nbytes), device='cuda:%d' % engine.device)
        sums.append(t.view(torch.int64).sum())
    return torch.stack(sums)


class _DevBuf(object):
    """Expose a raw device allocation of the engine to torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {'shape': (nbytes // 8,), 'typestr': '<f8', 'data': (ptr, False), 'version': 3,
                                         'strides': None}


def broadcast_state(engine, src=0, group=None):
    """NCCL-broadcast every posterior-state buffer of `engine` from rank `src` (in place, HBM to HBM)."""
    import torch
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for ptr, nbytes in engine.state_buffers():
        t = torch.as_tensor(_DevBuf(ptr, nbytes), device='cuda:%d' % engine.device)
        dist.broadcast(t, src=src, group=group)
    torch.cuda.synchronize(engine.device)
    if dist.get_rank(group) != src:
        engine.state_commit()


class ShardedSweep(object):
    """Candidate sweep over all ranks of a process group.

    rank `src` owns the refit; every rank holds a replica of the posterior state and sweeps its own rows.
    """

    def __init__(self, engine, group=None):
        self.engine = engine
        self.group = group
        dist = _dist()
        self.distributed = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.world = dist.get_world_size(group) if self.distributed else 1

    def fit(self, kernel, X, Y, variance, lengthscale, noise, src=0):
        """Refit on `src`, then replicate the state.  All ranks must pass the same shapes."""
        X = np.asarray(X, dtype=np.float64)
        S = np.asarray(variance).reshape(-1).size
        if self.rank == src:
            self.engine.fit(kernel, X, Y, variance, lengthscale, noise)
        else:
            self.engine.alloc(kernel, X.shape[0], X.shape[1], S)
        broadcast_state(self.engine, src=src, group=self.group)

    def local_rows(self, N):
        return shard_rows(N, self.rank, self.world)

    def acq_topk(self, kind, par, Xs_local, row_offset, k, f_out=None):
        """Sweep this rank's candidates (`Xs_local`, whose first row has global index `row_offset`) and return
        the global k best (score = -f, smallest first), identical on every rank."""
        vals, idx = self.engine.acq_topk(kind, par, Xs_local, k, f_out=f_out)
        dev = 'cuda:%d' % self.engine.device if self.distributed and _dist().get_backend(self.group) == 'nccl' else None
        return allgather_topk(vals, idx + row_offset, k, device=dev, group=self.group)

    def acq_topk_uniform(self, kind, par, key, N, bounds, k):
        """Sweep a device-generated uniform design of N global rows (Engine.design_uniform's stream), sharded by rows,
        and return (scores [k], global row numbers [k], rows [k, d]) of the k best -- identical on every rank and
        independent of the number of ranks."""
        lo, hi = self.local_rows(N)
        if hi > lo:
            vals, idx, _ = self.engine.acq_topk_uniform(kind, par, key, hi - lo, bounds, min(k, hi - lo), row0=lo)
        else:
            vals, idx = np.empty(0), np.empty(0, dtype=np.int64)
        dev = 'cuda:%d' % self.engine.device if self.distributed and _dist().get_backend(self.group) == 'nccl' else None
        vals, gidx = allgather_topk(vals, idx + lo, k, device=dev, group=self.group)
        # the winners are regenerated from their global row numbers: no candidate rows travel between ranks
        rows = np.vstack([self.engine.design_uniform(key, 1, bounds, row0=int(g)) for g in gidx]) if len(gidx) else \
            np.empty((0, np.atleast_2d(bounds).shape[0]))
        return vals, gidx, rows
