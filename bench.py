#!/usr/bin/env python
"""Benchmark of the hot path: EI + gradient candidate evaluations per second at n = 2048, d = 6 (BASELINE.json
config 3: alpine2, RBF, EI jitter 0.01), one process per GPU.

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on the host cores

A step is ONE pass of the hot path (AcquisitionEI._compute_acq_withGradients: K1 -> DMMA GEMM -> K3) over this rank's
batch of synthetic candidates.  `value` is device-resident throughput (inputs in HBM when the clock starts), `e2e` the
same call through the reference-facing plugin class with HOST buffers (H2D of the candidates and D2H of f, df inside
the timed region).  Scaling is weak: every rank sweeps its own `--batch` candidates; the only collectives are the
state broadcast after the refit (outside the timed region) and the all-gather of per-rank top-k records after each
sweep (inside it).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_OBS, DIM = 2048, 6
JITTER = 0.01
METRIC = "EI+grad candidate evals/sec (n=2048,d=6)"
UNIT = "candidates/s"


def synthetic_problem():
    """BASELINE config 3: X ~ U[1,10]^6 seed 0, Y = standardised alpine2, RBF(1, l=2), noise 1e-2."""
    rng = np.random.default_rng(0)
    X = 1 + 9 * rng.random((N_OBS, DIM))
    fval = -(np.prod(np.sqrt(X), axis=1) * np.prod(np.sin(X), axis=1))  # alpine2 (objective_examples/experimentsNd.py:59-67)
    Y = (fval - fval.mean()) / fval.std()
    return X, Y[:, None]


def flops_per_candidate(n, d):
    """Algorithmic FP64 flops of EI + gradient per candidate: 2 n^2 dense + elementwise terms."""
    return 2 * n * n + 2 * n + n * (3 * d + 8) + 4 * n * d


# ----------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        while not self._halt.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(',')
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith('active'):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------------------
# CPU baseline: the reference's path restated in NumPy (oracle/), timed on the host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_rate(sample, chunk=8192, repeats=1):
    """EI + gradient through the oracle (reference acquisition formulas on the NumPy restatement of GPy's exact GP),
    chunked at 8192 rows (un-chunked K* would need 164 GB at 1e7 candidates), fmin cached per refit.  BLAS runs on all
    host cores (torchrun exports OMP_NUM_THREADS=1, which would otherwise throttle the baseline)."""
    from oracle import gpy_numpy as G, gpyopt_numpy as og
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    X, Y = synthetic_problem()
    gp = G.GPRegression(X, Y, G.RBF(DIM, variance=1.0, lengthscale=2.0), noise_var=1e-2)
    fmin = og.gp_get_fmin(gp)
    Xs = 1 + 9 * np.random.default_rng(1).random((sample, DIM))
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        for lo in range(0, sample, chunk):
            m, s, dm, ds = og.gp_predict_withGradients(gp, Xs[lo:lo + chunk])
            og.acq_EI_withGradients(m, s, dm, ds, fmin, JITTER)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return sample / best, best


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    sample = args.cpu_sample
    cores = os.cpu_count()
    cpu_reference_rate(min(sample, 4096))  # warm-up (BLAS threads, page faults)
    for _ in range(max(0, args.warmup - 1)):
        cpu_reference_rate(min(sample, 4096))
    times = []
    for _ in range(args.steps):
        rate, dt = cpu_reference_rate(sample)
        times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = sample / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "alpine2 d=6, n=2048 observations, EI+gradient candidate sweep (BASELINE config 3)",
                   "kernel": "RBF variance=1 lengthscale=2", "noise_var": 1e-2, "jitter": JITTER,
                   "candidates_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d candidates per step, chunks of 8192 rows, NumPy/SciPy BLAS on all host cores; "
                                   "oracle port (GPy is not installable: reference path restated, oracle/)" % sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
def measure_fp64_peak(torch, device):
    """cuBLAS DGEMM 8192^3, best of 5 (the FP64 roofline denominator; MEASURED_PEAKS.json carries no FP64 entry)."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=device)
    b = torch.randn(n, n, dtype=torch.float64, device=device)
    torch.matmul(a, b)
    torch.cuda.synchronize(device)
    best = None
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize(device)
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / (best * 1e-3) * 1e-12


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from gpyopt_b200 import Engine, GPModel, AcquisitionEI, kern
    from gpyopt_b200.parallel import ShardedSweep, allgather_topk

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)

    X, Y = synthetic_problem()
    # the reference-facing model: fixed hyper-parameters (max_iters=0), state fitted on rank 0 and broadcast
    # posterior state: rank 0 factorises (K4), the other ranks receive L^-1, W^-1, alpha, ... over NCCL
    eng = Engine(local_rank)
    sweep = ShardedSweep(eng)
    sweep.fit('rbf', X, Y[:, 0], [1.0], [[2.0] * DIM], [1e-2], src=0)
    # the reference-facing model (fixed hyper-parameters: max_iters=0) adopts the resident state
    model = GPModel(kernel=kern.RBF(DIM, variance=1.0, lengthscale=2.0), noise_var=1e-2, max_iters=0, verbose=False,
                    device=local_rank, engine=eng, adopt_state=True)
    model.updateModel(X, Y, None, None)
    acq = AcquisitionEI(model, None, None, jitter=JITTER)

    N = args.batch
    g = torch.Generator(device=device)
    g.manual_seed(1000 + rank)
    Xs = 1 + 9 * torch.rand((N, DIM), dtype=torch.float64, device=device, generator=g)
    f = torch.empty(N, dtype=torch.float64, device=device)
    df = torch.empty((N, DIM), dtype=torch.float64, device=device)
    stream = torch.cuda.current_stream(device)
    K_TOP = 5

    def device_step():
        eng.acq('EI', JITTER, Xs, grad=True, out=(f, df), stream=stream.cuda_stream)
        vals, idx = eng.topk(f, K_TOP, largest=True, stream=stream.cuda_stream)
        lo = rank * N
        return allgather_topk(-vals, idx + lo, K_TOP, device=device if world > 1 else None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    for _ in range(args.warmup):
        device_step()
    peak_tf = measure_fp64_peak(torch, device) if rank == 0 else None

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    eng.launch_count(reset=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    best = None
    for _ in range(args.steps):
        best = device_step()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = eng.launch_count(reset=True)
    # dominant kernel, timed live with CUDA events on its own stream in one extra pass over the same batch; in this
    # pass the chunks are serialised on one stream slot so each event pair brackets exactly one GEMM launch
    eng.set_profiling(True)
    eng.get_profile()
    eng.acq('EI', JITTER, Xs, grad=True, out=(f, df), stream=stream.cuda_stream)
    torch.cuda.synchronize(device)
    gemm_ms, gemm_launches, gemm_cands = eng.get_profile()
    eng.set_profiling(False)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * N / (ms_step * 1e-3)

    # ---- end to end through the reference-facing plugin call, host buffers --------------------------------------
    host_x = torch.empty((N, DIM), dtype=torch.float64, pin_memory=True)
    host_x.copy_(Xs)
    x_np = host_x.numpy()
    f_h = df_h = None
    for _ in range(3):                      # warm-up at full size, holding the previous result like the timed loop does: the
        f_h, df_h = acq._compute_acq_withGradients(x_np)   # TWO page-locked result sets it alternates between enter torch's host cache
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        f_h, df_h = acq._compute_acq_withGradients(x_np)
        score = float(-f_h.max())           # device->host read of the step's result
    torch.cuda.synchronize(device)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * N / float(te.item())
    # parity spot-check of the timed outputs against the plugin path (same kernels, host vs device pointers)
    assert np.array_equal(f_h[:1000, 0], f[:1000].cpu().numpy())

    if rank == 0:
        info = eng.info()
        n_pad = info['n_pad']
        dense_flops_per_cand = 2.0 * n_pad * n_pad
        gemm_avg_ms = gemm_ms / max(1, gemm_launches)
        cands_per_launch = gemm_cands / max(1, gemm_launches)
        achieved_tf = dense_flops_per_cand * cands_per_launch / (gemm_avg_ms * 1e-3) * 1e-12 if gemm_launches else None
        traffic = None
        summary = os.path.join(ROOT, 'profiles', 'gemm_grad_summary.json')
        if os.path.exists(summary):
            try:
                traffic = json.load(open(summary)).get('dram_bytes_per_launch')
            except Exception:
                traffic = None
        cpu_rate, cpu_dt = cpu_reference_rate(args.cpu_sample)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "alpine2 d=6, n=2048 observations, EI+gradient candidate sweep (BASELINE config 3)",
                       "kernel": "RBF variance=1 lengthscale=2", "noise_var": 1e-2, "jitter": JITTER,
                       "candidates_per_gpu_per_step": N, "candidates_per_step": world * N, "chunk": info['chunk'],
                       "l2": "inputs_larger_than_l2 (48 B/candidate x %d + 16 KiB/candidate K* chunks >> 126 MB)" % N,
                       "strict_variance_pass": info['strict_variance'], "parallelism": "candidate rows sharded x%d" % world},
            "flops_per_candidate": flops_per_candidate(N_OBS, DIM),
            "tflops_algorithmic": value * flops_per_candidate(N_OBS, DIM) * 1e-12,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": N * DIM * 8, "d2h_bytes_per_step": N * (DIM + 1) * 8,
                    "api": "gpyopt_b200.AcquisitionEI._compute_acq_withGradients(numpy [N,6]) -> (f [N,1], df [N,6])"},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "gpo::dgemm_nt_kernel<EPI_GRAD> (FP64 DMMA m8n8k4 + TMA)",
                         "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": (achieved_tf / peak_tf) if achieved_tf and peak_tf else None, "traffic": traffic,
                         "peak_source": "cuBLAS DGEMM 8192^3 measured live in this run (best of 5); MEASURED_PEAKS.json has "
                                        "no FP64 entry; recorded pool value 36.06 TF in profiles/fp64_peak_r01.json",
                         "flops_per_launch": dense_flops_per_cand * cands_per_launch, "avg_launch_ms": gemm_avg_ms,
                         "launches_timed": gemm_launches,
                         "how": "CUDA events around every GEMM launch on its stream, one extra pass over the step's batch with "
                                "the chunk pipeline serialised; value/ms_per_step come from the normal overlapped pipeline",
                         "frac_whole_step": value / world * flops_per_candidate(N_OBS, DIM) * 1e-12 / peak_tf if peak_tf else None},
            "cpu_baseline": {"value": cpu_rate, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                             "sample": "%d candidates, chunks of 8192 rows, NumPy/SciPy BLAS on all host cores (%.1f s)"
                                       % (args.cpu_sample, cpu_dt)},
            "best_candidates": {"scores": [float(v) for v in best[0]], "global_indices": [int(i) for i in best[1]]},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=10_000_000, help='candidates per GPU per step')
    ap.add_argument('--cpu-sample', type=int, default=32768, help='candidates in the bounded CPU-baseline sample')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == '__main__':
    sys.exit(main())
