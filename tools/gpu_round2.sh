#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== probe"; timeout 180 tools/umma_probe > gpurun_out/probe2.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/probe2.txt
echo "== tensor debug"; timeout 300 python tools/rq_tensor_debug.py > gpurun_out/tensor_debug.txt 2>&1; echo "rc=$?"; cat gpurun_out/tensor_debug.txt
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu.txt
echo "== bench"; timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
