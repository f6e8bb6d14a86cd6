#!/bin/bash
# correctness + timing + pipeline trace of the fifth-generation encode kernel
export MEVI_RQ_KERNEL=${MEVI_RQ_KERNEL:-5}
mkdir -p gpurun_out
timeout 200 python tools/rq_tensor_debug.py 2>&1 | tail -7
timeout 200 python tools/rq_ablate.py ${ABL:-0,4,6,2,0} 2>&1 | grep -E "debug="
MEVI_RQ_TRACE=gpurun_out/trace_v5.bin timeout 200 python tools/rq_trace.py run > gpurun_out/trace_v5.txt 2>&1
tail -2 gpurun_out/trace_v5.txt
