#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== tensor debug"; timeout 300 python tools/rq_tensor_debug.py > gpurun_out/tensor_debug.txt 2>&1; echo "rc=$?"; cat gpurun_out/tensor_debug.txt
echo "== pytest rq/kmeans"; timeout 900 python -m pytest tests/test_gpu_rq_encode.py tests/test_gpu_kmeans.py -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_rq.txt 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_rq.txt
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_quick.err; python -c "
import json; j=json.load(open('gpurun_out/bench_quick.json')); print({k:j[k] for k in ['value','ms_per_step','gpu_launches','prefilter_flagged_fraction','fast_vs_exact_kernel_mismatch_rows']}, j['roofline']['frac'], j['e2e']['value'], j['clocks'])"
