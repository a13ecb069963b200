#!/bin/bash
# trace + timing of the refit at the given n:  tools/fit_trace_n.sh TAG n...
tag=$1; shift
for n in "$@"; do
  python tools/fit_trace.py $n 2> gpurun_out/${tag}_trace$n.jsonl
  python tools/show_trace.py gpurun_out/${tag}_trace$n.jsonl | tail -3
done
python tools/time_fit.py 256 512 1024 2048 3072 4096 | tee gpurun_out/${tag}_timefit.jsonl
