#!/bin/bash
# Round-2 measurement pass on one B200: every artefact that profiles/ cites.  Run under gpurun; results land in gpurun_out/final/.
set -u
O=gpurun_out/final; mkdir -p $O
python bench.py --steps 5 --warmup 3 > $O/bench_r02_1gpu.json 2> $O/bench_r02_1gpu.err; echo "bench rc=$?"
python bench.py --steps 5 --warmup 3 --strict 0 --cpu-sample 4096 > $O/bench_r02_1gpu_fastform.json 2> $O/bench_fast.err; echo "bench fast rc=$?"
python bench.py --steps 3 --warmup 3 --config c4 > $O/bench_r02_c4_1gpu.json 2> $O/bench_c4.err; echo "bench c4 rc=$?"
python tests/manual_bench_configs.py > $O/configs_r02.jsonl 2> $O/configs.err; echo "configs rc=$?"
python tools/bench_kernels.py > $O/kernels_r02.jsonl 2>&1
python tools/time_fit.py 20 64 128 256 512 1024 2048 3072 4096 5120 6144 > $O/refit_r02.jsonl 2>&1
tools/microbench/cusolver_potrf 512 1024 2048 3072 4096 > $O/refit_cusolver_baseline_r02.jsonl 2>&1
python tools/fit_ab.py 256 1000 2048 3072 --no-check > $O/refit_ab_r02.jsonl 2>&1
python tools/time_single_point.py > $O/latency_r02.json 2>&1
python tools/time_lbfgs_rounds.py > $O/lbfgs_rounds_r02.jsonl 2>&1
python tools/desync_ab.py > $O/desync_ab_r02.jsonl 2>&1
python tools/stress_latency_path.py 4000 > $O/stress_latency_path_r02.jsonl 2>&1
python tools/stress_sweep_repeat.py > $O/stress_sweep_repeat_r02.jsonl 2>&1
python tools/fit_chain_limit.py 4096 4480 4608 5120 6144 > $O/refit_large_n_r02.jsonl 2>&1
for n in 2048 4096; do python tools/fit_trace.py $n 2> $O/fit_trace_${n}_r02.jsonl; done
# launch list of the bench command (small batch), and the same command's line without ncu
python bench.py --steps 2 --warmup 1 --batch 189440 --cpu-sample 2048 > $O/bench_r02_small_for_ncu.json 2> $O/bench_small.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_bench_r02.csv python bench.py --steps 2 --warmup 1 --batch 189440 --cpu-sample 2048 > $O/ncu_bench.log 2>&1; echo "ncu bench rc=$?"
# launch list of one refit at n = 2048 (second fit of the process: plain streams, the graph replay shows as the same kernels)
python tools/prof_fit.py 2048 > $O/pfit.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_fit2048_r02.csv python tools/prof_fit.py 2048 > $O/ncu_fit.log 2>&1; echo "ncu fit rc=$?"
# full captures of the dominant kernels
python tools/prof_sweep.py 18944 grad > $O/ps.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:dgemm_nt_kernel<\(int\)[12]," -s 2 -c 2 -f -o $O/gemm_sweep_r02 python tools/prof_sweep.py 18944 grad > $O/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"chol_chain_kernel|rank128_dmma_kernel|trsm_panel_kernel" -s 60 -c 6 -f -o $O/refit_chain_r02 python tools/prof_fit.py 2048 > $O/ncu_chain.log 2>&1; echo "ncu chain rc=$?"
python tools/prof_lbfgs.py 4096 8 > $O/plb.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"smalln_both|cov_small|acq_small" -s 9 -c 3 -f -o $O/latency_kernels_r02 python tools/prof_lbfgs.py 4096 8 > $O/ncu_lat.log 2>&1; echo "ncu latency rc=$?"
ls -la $O
