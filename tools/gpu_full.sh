#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 240 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.txt
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.txt
echo "== bench"; timeout 420 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
