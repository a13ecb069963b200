#!/usr/bin/env python
"""Benchmark of the hot path: EI + gradient candidate evaluations per second at n = 2048, d = 6 (BASELINE.json
config 3: alpine2, RBF, EI jitter 0.01, 10 M candidates sharded across the GPUs), one process per GPU.

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on the host cores
    python bench.py --config c4 ...                          # secondary line: BASELINE config 4 (S=64, n=512, d=4)

A step is ONE pass of the hot path (AcquisitionEI._compute_acq_withGradients: K1 -> DMMA GEMMs -> K3) over the step's
candidates, followed by the selection of the 5 best (per-rank top-k, one all-gather, identical merge).  `value` is
device-resident throughput (candidates already in HBM when the clock starts), `e2e` the same call through the
reference-facing plugin class with a plain pageable NumPy array (H2D of the candidates and D2H of f, df inside the
timed region).

Scaling is STRONG by default, as BASELINE config 3 states it: `--batch` candidates IN TOTAL per step, row-sharded over the
ranks (every rank generates its own rows of one global counter-based design, so the candidate set does not depend on
the number of GPUs).  `--scaling weak` gives every rank `--batch` candidates of its own.  The only collectives are the
state broadcast after the refit (outside the timed region) and the all-gather of the per-rank top-k records after each
sweep (inside it).

After the timed region the run checks itself against the CPU oracle (`parity`: the first 4096 rows of the design the
`cpu_baseline` leg times are scored on the GPU as well) and, on more than one GPU, that the replicas hold identical
state and that the sharded selection equals a single-engine sweep (`multi_gpu_check`).
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

JITTER = 0.01
UNIT = "candidates/s"
DESIGN_KEY = [20261020, 3]      # Philox key of the candidate design (device and NumPy streams are bit-identical)
PARITY_ROWS = 4096
K_TOP = 5

CONFIGS = {
    'c3': dict(n=2048, d=6, S=1, acq='EI', bounds=(1.0, 10.0),
               metric="EI+grad candidate evals/sec (n=2048,d=6)",
               workload="alpine2 d=6, n=2048 observations, EI+gradient candidate sweep (BASELINE config 3)",
               kernel="RBF variance=1 lengthscale=2"),
    'c4': dict(n=512, d=4, S=64, acq='EI_MCMC', bounds=(0.0, 1.0),
               metric="EI_MCMC+grad candidate evals/sec (S=64,n=512,d=4)",
               workload="GP_MCMC with EI_MCMC, 64 hyper-parameter samples, d=4, n=512, candidate rows sharded (BASELINE config 4)",
               kernel="RBF, theta_s = (1.0, 0.5, 1e-2) * exp(0.1 N(0,1)), seed 2"),
}


def alpine2(X):
    """objective_examples/experimentsNd.py:59-67"""
    return -(np.prod(np.sqrt(X), axis=1) * np.prod(np.sin(X), axis=1))


def synthetic_problem(cfg):
    """SURVEY.md section 8d.  C3: X ~ U[1,10]^6 seed 0, Y = standardised alpine2, RBF(1, l=2), noise 1e-2.
    C4: X ~ U[0,1]^4, Y = standardised alpine2 on rescaled inputs, S=64 synthetic hyper-parameter samples, seed 2."""
    n, d = cfg['n'], cfg['d']
    if cfg['S'] == 1:
        rng = np.random.default_rng(0)
        X = 1 + 9 * rng.random((n, d))
        fval = alpine2(X)
        Y = (fval - fval.mean()) / fval.std()
        return X, Y[:, None], np.array([1.0]), np.full((1, d), 2.0), np.array([1e-2])
    rng = np.random.default_rng(2)
    X = rng.random((n, d))
    fval = alpine2(1 + 9 * X)
    Y = (fval - fval.mean()) / fval.std()
    theta = np.array([1.0, 0.5, 1e-2]) * np.exp(0.1 * rng.standard_normal((cfg['S'], 3)))
    return X, Y[:, None], theta[:, 0].copy(), np.repeat(theta[:, 1:2], d, axis=1), theta[:, 2].copy()


def flops_per_candidate(n, d, S=1):
    """Algorithmic FP64 flops of EI + gradient per candidate: 2 n^2 dense + elementwise terms."""
    return S * (2 * n * n + 2 * n + n * (3 * d + 8) + 4 * n * d)


def design_rows(cfg, row0, rows):
    """Rows [row0, row0 + rows) of the global candidate design, on the host (NumPy's Philox stream)."""
    lo, hi = cfg['bounds']
    d = cfg['d']
    bg = np.random.Philox(key=np.array(DESIGN_KEY, dtype=np.uint64))
    bg.advance((row0 * d) // 4)
    g = np.random.Generator(bg)
    skip = (row0 * d) % 4
    u = g.uniform(lo, hi, size=skip + rows * d)[skip:]
    return np.ascontiguousarray(u.reshape(rows, d))


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
def oracle_model(cfg):
    from oracle import gpy_numpy as G, gpyopt_numpy as og
    X, Y, var, ls, noise = synthetic_problem(cfg)
    gps = [G.GPRegression(X, Y, G.RBF(cfg['d'], variance=float(var[z]), lengthscale=float(ls[z, 0])), noise_var=float(noise[z]))
           for z in range(cfg['S'])]
    fmins = [og.gp_get_fmin(gp) for gp in gps]
    return gps, fmins


def oracle_eval(cfg, gps, fmins, Xs):
    """EI (or EI_MCMC) value and gradient through the oracle: the reference acquisition formulas on the NumPy
    restatement of GPy's exact GP; fmin cached per refit."""
    from oracle import gpyopt_numpy as og
    if cfg['S'] == 1:
        m, s, dm, ds = og.gp_predict_withGradients(gps[0], Xs)
        return og.acq_EI_withGradients(m, s, dm, ds, fmins[0], JITTER)
    preds = [og.gp_predict_withGradients(gp, Xs) for gp in gps]
    return og.acq_EI_MCMC_withGradients([p[0] for p in preds], [p[1] for p in preds], [p[2] for p in preds],
                                        [p[3] for p in preds], fmins, JITTER)


def set_blas_threads(n):
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:
        pass


def cpu_reference_rate(cfg, sample, chunk=8192, model=None, keep=0):
    """Time the oracle over `sample` rows of the global design, chunked at 8192 rows (un-chunked K* would need 164 GB at
    1e7 candidates).  Returns (rate, seconds, f[:keep], df[:keep])."""
    gps, fmins = model if model is not None else oracle_model(cfg)
    Xs = design_rows(cfg, 0, sample)
    fs, dfs = [], []
    t0 = time.perf_counter()
    for lo in range(0, sample, chunk):
        f, df = oracle_eval(cfg, gps, fmins, Xs[lo:lo + chunk])
        if lo < keep:
            fs.append(f)
            dfs.append(df)
    dt = time.perf_counter() - t0
    f_keep = np.vstack(fs)[:keep] if keep else None
    df_keep = np.vstack(dfs)[:keep] if keep else None
    return sample / dt, dt, f_keep, df_keep


def run_reference_arm(args, cfg):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    sample = args.cpu_sample
    cores = os.cpu_count()
    set_blas_threads(cores)   # (torchrun exports OMP_NUM_THREADS=1, which would otherwise throttle the baseline)
    model = oracle_model(cfg)
    for _ in range(max(1, args.warmup)):
        cpu_reference_rate(cfg, min(sample, 4096), model=model)   # warm-up (BLAS threads, page faults)
    times = []
    for _ in range(args.steps):
        rate, dt, _, _ = cpu_reference_rate(cfg, sample, model=model)
        times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = sample / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": cfg['metric'], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg['workload'], "kernel": cfg['kernel'], "noise_var": 1e-2, "jitter": JITTER,
                   "candidates_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d candidates per step (first rows of the design the GPU arm sweeps), chunks of 8192 rows, "
                                   "NumPy/SciPy BLAS on all host cores; oracle port (GPy is not installable: reference path "
                                   "restated, oracle/)" % sample},
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


def scaled_err(a, ref):
    den = float(np.max(np.abs(ref)))
    return float(np.max(np.abs(np.asarray(a) - np.asarray(ref))) / (den if den > 0 else 1.0))


def multi_gpu_check(torch, dist, eng, sweep, cfg, rank, world, device):
    """Outside the timed region, on the hardware the number was taken on: (1) every replica holds bit-identical state
    (incl. L^-1, L^-T, W^-1); (2) the row-sharded sweep of a shared 200 000-row design selects the same 5 candidates, with
    the same scores, as ONE engine sweeping all of it; (3) gradient sweeps on every replica (N > 256: the two triangular
    products that read L^-T) equal rank 0's bit for bit."""
    from gpyopt_b200.parallel import state_checksums
    out = {}
    cs = state_checksums(eng)
    all_cs = [torch.empty_like(cs) for _ in range(world)]
    dist.all_gather(all_cs, cs)
    out['state_identical'] = bool(all(torch.equal(all_cs[0], c) for c in all_cs))
    bounds = np.array([cfg['bounds']] * cfg['d'], dtype=np.float64)
    key = [777, 1]
    v_sh, i_sh, _ = sweep.acq_topk_uniform(cfg['acq'], JITTER, key, 200_000, bounds, K_TOP)
    v_1, i_1, _ = eng.acq_topk_uniform(cfg['acq'], JITTER, key, 200_000, bounds, K_TOP)
    same = torch.tensor([int(np.array_equal(i_sh, i_1) and np.array_equal(v_sh, v_1))], device=device)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    out['sharded_top5_equals_single_engine'] = bool(same.item())
    Xp = torch.from_numpy(design_rows(cfg, 0, 1024)).to(device)
    f, df = eng.acq(cfg['acq'], JITTER, Xp, grad=True)
    torch.cuda.synchronize(device)
    pack = torch.cat([f.reshape(-1), df.reshape(-1)]).view(torch.int64)
    ref = pack.clone()
    dist.broadcast(ref, src=0)
    eq = torch.tensor([int(torch.equal(pack, ref))], device=device)
    dist.all_reduce(eq, op=dist.ReduceOp.MIN)
    out['replica_gradients_bitwise_equal'] = bool(eq.item())
    out['status'] = 'ok' if all(v for v in out.values()) else 'FAILED'
    return out


def run_gpu_arm(args, cfg):
    import torch
    import torch.distributed as dist
    from gpyopt_b200 import Engine, GPModel, GPModel_MCMC, AcquisitionEI, AcquisitionEI_MCMC, kern
    from gpyopt_b200.parallel import ShardedSweep, allgather_topk_device, shard_rows

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

    n, d, S, kind = cfg['n'], cfg['d'], cfg['S'], cfg['acq']
    X, Y, var, ls, noise = synthetic_problem(cfg)
    # posterior state: rank 0 factorises (K4), the other ranks receive L^-1, L^-T, W^-1, alpha, ... over NCCL
    eng = Engine(local_rank)
    if args.strict is not None:
        eng.set_strict_variance(args.strict)
    sweep = ShardedSweep(eng)
    sweep.fit('rbf', X, Y[:, 0], var, ls, noise, src=0)
    # the reference-facing plugin objects on top of the resident state (fixed hyper-parameters)
    if S == 1:
        model = GPModel(kernel=kern.RBF(d, variance=float(var[0]), lengthscale=float(ls[0, 0])), noise_var=float(noise[0]),
                        max_iters=0, verbose=False, device=local_rank, engine=eng, adopt_state=True)
        model.updateModel(X, Y, None, None)
        acq = AcquisitionEI(model, None, None, jitter=JITTER)
    else:
        model = GPModel_MCMC(kernel=kern.RBF(d, variance=1.0, lengthscale=0.5), noise_var=1e-2, device=local_rank)
        model.adopt_samples(X, Y, np.column_stack([var, ls[:, 0], noise]), engine=eng)
        acq = AcquisitionEI_MCMC(model, None, None, jitter=JITTER)

    strong = args.scaling == 'strong'
    n_total = args.batch if strong else args.batch * world
    lo, hi = shard_rows(n_total, rank, world)
    N = hi - lo
    bounds = np.array([cfg['bounds']] * d, dtype=np.float64)
    Xs = torch.empty((N, d), dtype=torch.float64, device=device)
    eng.design_uniform(DESIGN_KEY, N, bounds, row0=lo, out=Xs)      # this rank's rows of the one global design
    f = torch.empty(N, dtype=torch.float64, device=device)
    df = torch.empty((N, d), dtype=torch.float64, device=device)
    tk_vals = torch.empty(K_TOP, dtype=torch.float64, device=device)
    tk_idx = torch.empty(K_TOP, dtype=torch.int64, device=device)
    stream = torch.cuda.current_stream(device)

    def device_step():
        eng.acq(kind, JITTER, Xs, grad=True, out=(f, df), stream=stream.cuda_stream)
        eng.topk(f, K_TOP, largest=True, stream=stream.cuda_stream, out=(tk_vals, tk_idx))
        return allgather_topk_device(tk_vals, tk_idx, lo, K_TOP, negate=True)   # one collective, one D2H (sync)

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
    # dominant kernel(s), timed live with CUDA events on their own stream in one extra pass over the same batch; in this
    # pass the chunks are serialised on one stream slot so each event pair brackets exactly the variance contraction of
    # one chunk
    eng.set_profiling(True)
    eng.get_profile()
    eng.acq(kind, JITTER, Xs, grad=True, out=(f, df), stream=stream.cuda_stream)
    torch.cuda.synchronize(device)
    gemm_ms, gemm_launches, gemm_cands = eng.get_profile()
    eng.set_profiling(False)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = n_total / (ms_step * 1e-3)

    # ---- end to end through the reference-facing plugin call: pageable NumPy in, NumPy out ----------------------
    x_np = np.array(Xs.cpu().numpy(), copy=True)           # plain malloc'ed memory, what GPyOpt hands the acquisition
    f_h = df_h = None
    for _ in range(3):                      # warm-up at full size, holding the previous result like the timed loop does
        f_h, df_h = acq._compute_acq_withGradients(x_np)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 8))
    for _ in range(e2e_steps):
        f_h, df_h = acq._compute_acq_withGradients(x_np)
        score = float(-f_h.max())           # the step's result is read on the host
    torch.cuda.synchronize(device)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = n_total / float(te.item())
    # the plugin path (host pointers) and the device-pointer path run the same kernels: bitwise equal
    assert np.array_equal(f_h[:1000, 0], f[:1000].cpu().numpy())

    mg = multi_gpu_check(torch, dist, eng, sweep, cfg, rank, world, device) if world > 1 else {"status": "n/a (1 GPU)"}

    rc = 0
    if rank == 0:
        info = eng.info()
        n_pad = info['n_pad']
        dense_flops_per_cand = 2.0 * n_pad * n_pad * S
        gemm_avg_ms = gemm_ms / max(1, gemm_launches)
        cands_per_launch = gemm_cands / max(1, gemm_launches)
        achieved_tf = dense_flops_per_cand * cands_per_launch / (gemm_avg_ms * 1e-3) * 1e-12 if gemm_launches else None
        traffic, traffic_src = None, None
        summary = os.path.join(ROOT, 'profiles', 'gemm_sweep_r02_summary.json')
        if os.path.exists(summary) and args.config == 'c3':
            try:
                js = json.load(open(summary))      # the variance-contraction launches of ONE chunk (two in the default form)
                want = 2 if info['strict_variance'] else 1
                ls = [l for l in js['launches'] if 'dgemm_nt_kernel' in l['kernel']][:want]
                if len(ls) == want:
                    stored_cands = float(js.get('candidates_per_launch', 9472))
                    traffic = float(sum(l['dram_bytes_total'] for l in ls)) * cands_per_launch / stored_cands
                    traffic_src = ("stored ncu value (profiles/gemm_sweep_r02_summary.json: dram__bytes_read.sum + dram__bytes_write.sum of "
                                   "`ncu --set full` captures of the same kernels on one %d-candidate chunk, scaled to the %d candidates "
                                   "per launch of this run), NOT measured in this run" % (int(stored_cands), int(round(cands_per_launch))))
            except Exception:
                traffic = None
        # CPU baseline + parity of what was timed: the oracle sweeps the first rows of the same design
        set_blas_threads(os.cpu_count())
        omodel = oracle_model(cfg)
        cpu_reference_rate(cfg, min(args.cpu_sample, 2048), model=omodel)          # warm-up
        cpu_rate, cpu_dt, f_o, df_o = cpu_reference_rate(cfg, args.cpu_sample, model=omodel, keep=min(PARITY_ROWS, args.cpu_sample, N))
        set_blas_threads(1)
        one_rate, _, _, _ = cpu_reference_rate(cfg, min(args.cpu_sample, 4096 if S == 1 else 512), model=omodel)
        set_blas_threads(os.cpu_count())
        P = f_o.shape[0]
        f_g = f[:P].cpu().numpy()
        df_g = df[:P].cpu().numpy()
        _, idx_g = eng.acq_topk(kind, JITTER, Xs[:P].contiguous(), K_TOP)
        idx_o = np.argsort(-f_o[:, 0], kind='stable')[:K_TOP]
        parity = {"rows": int(P), "f": scaled_err(f_g, f_o[:, 0]), "df": scaled_err(df_g, df_o),
                  "top5_equal": bool(np.array_equal(idx_g, idx_o)), "tolerance": 1e-10,
                  "metric": "max|a-ref|/max|ref| against the CPU oracle on the first rows of the timed design"}
        parity["status"] = "ok" if parity["f"] <= 1e-10 and parity["df"] <= 1e-10 and parity["top5_equal"] else "FAILED"
        line = {
            "metric": cfg['metric'], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": cfg['workload'], "kernel": cfg['kernel'], "noise_var": 1e-2, "jitter": JITTER,
                       "candidates_per_step": n_total, "candidates_per_gpu_per_step": N, "chunk": info['chunk'],
                       "l2": "inputs_larger_than_l2 (K* chunks of %d KiB/candidate x %d candidates per chunk >> 126 MB; fresh chunk "
                             "every GEMM)" % (8 * n_pad * S // 1024, info['chunk']),
                       "strict_variance_pass": info['strict_variance'],
                       "variance_form": "sigma_f^2 - |L^-1 k|^2 (reference form, two triangular products)" if info['strict_variance']
                       else "sigma_f^2 - k.W^-1.k (fast form, opt-in)",
                       "parallelism": "candidate rows sharded x%d (%s scaling), state replicated by NCCL broadcast, per-rank top-%d "
                                      "records all-gathered" % (world, args.scaling, K_TOP)},
            "flops_per_candidate": flops_per_candidate(n, d, S),
            "tflops_algorithmic": value * flops_per_candidate(n, d, S) * 1e-12,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n_total * d * 8, "d2h_bytes_per_step": n_total * (d + 1) * 8,
                    "steps": e2e_steps, "input": "pageable numpy.ndarray",
                    "api": "gpyopt_b200.Acquisition%s._compute_acq_withGradients(numpy [N,%d]) -> (f [N,1], df [N,%d])" % (kind, d, d)},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor",
                         "kernel": "gpo::dgemm_nt_kernel<EPI_SUMSQ|EPI_GRAD> (FP64 DMMA m8n8k4 + TMA): the variance contraction of one "
                                   "chunk = t = L^-1 K* (sum of squares) then b = L^-T t (gradient reductions)" if info['strict_variance']
                         else "gpo::dgemm_nt_kernel<EPI_GRAD> (FP64 DMMA m8n8k4 + TMA): b = W^-1 K*",
                         "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": (achieved_tf / peak_tf) if achieved_tf and peak_tf else None, "traffic": traffic,
                         "traffic_source": traffic_src,
                         "peak_source": "cuBLAS DGEMM 8192^3 measured live in this run (best of 5); MEASURED_PEAKS.json has "
                                        "no FP64 entry",
                         "flops_per_launch": dense_flops_per_cand * cands_per_launch, "avg_launch_ms": gemm_avg_ms,
                         "launches_timed": gemm_launches,
                         "how": "algorithmic 2 n^2 S flop per candidate; CUDA events around the variance contraction of every chunk "
                                "on its stream, one extra pass over the step's batch with the chunk pipeline serialised; "
                                "value/ms_per_step come from the normal overlapped pipeline",
                         "frac_whole_step": value / world * flops_per_candidate(n, d, S) * 1e-12 / peak_tf if peak_tf else None},
            "cpu_baseline": {"value": cpu_rate, "unit": UNIT, "cores": os.cpu_count(),
                             "threads_effective": round(cpu_rate / one_rate, 2), "one_thread_value": one_rate, "kind": "port",
                             "sample": "%d candidates (first rows of the timed design), chunks of 8192 rows, NumPy/SciPy BLAS on all "
                                       "host cores (%.1f s); threads_effective = all-core rate / one-thread rate" % (args.cpu_sample, cpu_dt)},
            "parity": parity,
            "multi_gpu_check": mg,
            "best_candidates": {"scores": [float(v) for v in best[0]], "global_indices": [int(i) for i in best[1]]},
        }
        print(json.dumps(line))
        if parity["status"] != "ok" or (world > 1 and mg.get("status") != "ok"):
            sys.stderr.write("bench.py: self-check FAILED: parity=%r multi_gpu_check=%r\n" % (parity, mg))
            rc = 1
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return rc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='c3', choices=sorted(CONFIGS))
    ap.add_argument('--scaling', default='strong', choices=['strong', 'weak'])
    ap.add_argument('--batch', type=int, default=None,
                    help='candidates per step: in total (strong scaling, default) or per GPU (--scaling weak)')
    ap.add_argument('--cpu-sample', type=int, default=None, help='candidates in the bounded CPU-baseline sample')
    ap.add_argument('--strict', type=int, default=None, choices=[-1, 0, 1],
                    help='variance form (gpo_set_strict_variance); default: the engine default (1, reference form)')
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.batch is None:
        args.batch = 10_000_000 if args.config == 'c3' else 500_000
    if args.cpu_sample is None:
        args.cpu_sample = 65536 if args.config == 'c3' else 4096
    if args.impl == 'reference':
        return run_reference_arm(args, cfg)
    return run_gpu_arm(args, cfg)


if __name__ == '__main__':
    sys.exit(main())
