#!/bin/bash
# round 2, session L: score-histogram bar in the exact kernel (c5), four 40 KB stages in the large class (c2),
# directory-probe load flavours (prepare_kernel), chunk schedules of the host-buffer call
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/sweep.py --workload c3 --steps 5 --variants 0,0x4000,0x8000,0xC000 --check 0x4000,0x8000,0xC000 > gpurun_out/sweep_c3.log 2>&1; grep -E "variant|rror" gpurun_out/sweep_c3.log | tail -5
for wl in c5 c2; do
  timeout 400 python bench.py --workload $wl --steps 5 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.log; echo "$wl rc=$?"; python tools/show_bench.py gpurun_out/bench_$wl.json; grep -iE "error|Traceback" -A8 gpurun_out/bench_$wl.log | head -20
done
TRACE=0 timeout 300 python tools/trace_e2e.py 131072:0 131072:1 131072:2 131072:3 131072:4 131072:5 65536:1 65536:3 262144:1 262144:4 > gpurun_out/e2e_sched.log 2>&1; grep -E "chunk .*per call|rror" gpurun_out/e2e_sched.log
