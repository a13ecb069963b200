"""world_size-2 gloo worker for tests/test_host_logic.py: exercises the sharded-sweep protocol on CPU with a
NumPy stand-in scorer in place of the CUDA engine (which needs a B200)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gpyopt_b200 import parallel  # noqa: E402


def _records(scores, k, world):
    """what allgather_topk_device should have merged: rank 0 contributes 3 records, the others k"""
    out = []
    for r in range(world):
        lo, hi = parallel.shard_rows(scores.size, r, world)
        li = np.argsort(scores[lo:hi], kind='stable')[:k if r == 1 else 3]
        out.append((scores[lo:hi][li], li + lo))
    return out


def main():
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    assert world == 2
    N, d, k = 1001, 3, 7
    rng = np.random.default_rng(0)           # identical candidate set on both ranks
    Xs = rng.random((N, d))
    scores = np.sin(7 * Xs).sum(1)
    scores[5] = scores[900]                  # a tie that straddles the two shards
    lo, hi = parallel.shard_rows(N, rank, world)
    local = scores[lo:hi]
    li = np.argsort(local, kind='stable')[:k]
    vals, idx = parallel.allgather_topk(local[li], li + lo, k)
    ref = np.argsort(scores, kind='stable')[:k]
    assert np.array_equal(idx, ref), (idx, ref)
    assert np.array_equal(vals, scores[ref])
    # every rank must hold the identical merged record
    t = torch.from_numpy(idx.astype(np.int64).copy())
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    assert all(torch.equal(g, gathered[0]) for g in gathered)
    # the device-resident form (records packed in a tensor, ONE collective, one copy back): same merge
    tv = torch.full((k,), float('nan'), dtype=torch.float64)
    ti = torch.full((k,), -1, dtype=torch.int64)
    kk = k if rank == 1 else 3                     # rank 0 holds fewer records than k: padded with (NaN, -1)
    tv[:kk] = torch.from_numpy(-local[li][:kk].copy())   # the engine leaves f; the merge wants the score -f
    ti[:kk] = torch.from_numpy(li[:kk].astype(np.int64))
    vd, idd = parallel.allgather_topk_device(tv, ti, lo, k, negate=True)
    ref_d = parallel.merge_topk([(local0, idx0) for local0, idx0 in _records(scores, k, world)], k)
    assert np.array_equal(idd, ref_d[1]) and np.array_equal(vd, ref_d[0]), (idd, ref_d[1])
    # fewer local rows than k on one rank
    vals2, idx2 = parallel.allgather_topk(local[li][:2] if rank == 0 else local[li], (li + lo)[:2] if rank == 0 else li + lo, k)
    assert len(idx2) == k and np.all(np.diff(vals2) >= 0)
    # device-generated design, sharded by rows: same winners as ONE process sweeping the whole design
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from fake_engine import FakeEngine
    from oracle import gpyopt_numpy as og
    eng = FakeEngine()
    Xt = rng.random((25, 3)) * 4 - 2
    eng.fit('rbf', Xt, np.sin(Xt).sum(1), 1.0, np.full((1, 3), 0.7), 1e-3)
    bounds = np.array([[-2., 2.], [-2., 2.], [-1., 3.]])
    sw = parallel.ShardedSweep(eng)
    v, gi, rows = sw.acq_topk_uniform('EI', 0.01, [11, 22], 1003, bounds, 5)
    Xall = og.design_uniform([11, 22], 1003, bounds)
    v1, i1 = eng.acq_topk('EI', 0.01, Xall, 5)
    assert np.array_equal(gi, i1) and np.array_equal(v, v1) and np.array_equal(rows, Xall[i1]), (gi, i1)
    dist.barrier()
    if rank == 0:
        print('DIST_OK')
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
