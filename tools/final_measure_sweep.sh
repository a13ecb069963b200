#!/bin/bash
set -u
O=gpurun_out/final; mkdir -p $O
python bench.py --steps 5 --warmup 3 > $O/bench_r02_1gpu.json 2> $O/bench_r02_1gpu.err; echo "bench rc=$?"
python bench.py --steps 5 --warmup 3 --strict 0 --cpu-sample 4096 > $O/bench_r02_1gpu_fastform.json 2> $O/bench_fast.err; echo "bench fast rc=$?"
python tools/bench_kernels.py > $O/kernels_r02.jsonl 2>&1
python bench.py --steps 2 --warmup 1 --batch 189440 --cpu-sample 2048 > $O/bench_r02_small_for_ncu.json 2> $O/bench_small.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_bench_r02.csv python bench.py --steps 2 --warmup 1 --batch 189440 --cpu-sample 2048 > $O/ncu_bench.log 2>&1; echo "ncu bench rc=$?"
python tools/prof_sweep.py 37888 grad > $O/ps.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:dgemm_nt_kernel<\(int\)[12]," -s 2 -c 2 -f -o $O/gemm_sweep_r02 python tools/prof_sweep.py 37888 grad > $O/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
python tests/manual_bench_configs.py > $O/configs_r02.jsonl 2> $O/configs.err; echo "configs rc=$?"
