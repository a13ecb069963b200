"""Multi-GPU self-check (run under torchrun, one rank per GPU, NCCL): the sharded sweeps must return exactly what ONE
engine returns for the whole candidate set -- host design and device-generated design.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_multigpu.py"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gpyopt_b200 import Engine
from gpyopt_b200.parallel import ShardedSweep


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    eng = Engine(local)
    sw = ShardedSweep(eng)
    rng = np.random.default_rng(0)                      # same data on every rank (only rank 0's fit is used)
    n, d, N, k = 700, 6, 200_003, 7
    X = 1 + 9 * rng.random((n, d)); Y = np.sin(X).sum(1); Y = (Y - Y.mean()) / Y.std()
    sw.fit('rbf', X, Y, [1.0], [[2.0]], [1e-2])
    bounds = np.array([[1., 10.]] * d)
    key = [77, 5]
    # (1) device-generated design, sharded: winners + rows identical to a single-engine sweep of the whole design
    v, gi, rows = sw.acq_topk_uniform('EI', 0.01, key, N, bounds, k)
    v1, i1, r1 = eng.acq_topk_uniform('EI', 0.01, key, N, bounds, k)
    assert np.array_equal(gi, i1) and np.array_equal(v, v1) and np.array_equal(rows, r1), (rank, gi, i1)
    # (2) host design, sharded by rows
    Xs = eng.design_uniform(key, N, bounds)
    lo, hi = sw.local_rows(N)
    v2, i2 = sw.acq_topk('EI', 0.01, Xs[lo:hi], lo, k)
    assert np.array_equal(i2, i1) and np.array_equal(v2, v1)
    # (3) replicated state is bit-identical to the fitting rank's: same predictions everywhere
    m, s = eng.predict(Xs[:1000])
    t = torch.from_numpy(np.concatenate([m[0], s[0]])).cuda()
    g = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(g, t)
    assert all(torch.equal(g[0], x) for x in g)
    dist.barrier()
    if rank == 0:
        print('MULTIGPU_OK world=%d winners=%s' % (world, gi.tolist()))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
