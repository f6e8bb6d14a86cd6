#!/bin/bash
# First GPU pass: hardware probe, parity tests, smoke, bench, ncu launch list + full capture.
set -u
mkdir -p gpurun_out
nvidia-smi > gpurun_out/env.txt 2>&1
free -g >> gpurun_out/env.txt; nproc >> gpurun_out/env.txt
echo "== probe"; timeout 180 tools/umma_probe > gpurun_out/probe.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/probe.txt
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 -x -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.txt
echo "== bench"; timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --docs 2000000 --no-cpu-baseline > gpurun_out/bench_ncu.json 2> gpurun_out/bench_ncu.err; echo "ncu rc=$?"
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rq_encode_exact|rerank_kernel|kmeans_accumulate" -c 4 -o gpurun_out/prof_v0 python bench.py --steps 1 --warmup 3 --docs 1000000 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_full.err; echo "ncu full rc=$?"
ls -la gpurun_out
