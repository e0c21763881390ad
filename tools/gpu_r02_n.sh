#!/bin/bash
# round 2, session N: hybrid producers (TMA bulk copies + cp.async warps), one stage class of four 40 KB stages,
# dynamic row hand-out in the multi-pass exact kernel
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
V=0,0x2000000,0x3000000,0x4000000,0x5000000,0x6000000,0x7000000,0x1000000
timeout 400 python tools/sweep.py --workload c3 --steps 5 --variants $V --check ${V#0,} > gpurun_out/sweep_c3.log 2>&1; grep -E "variant|rror" gpurun_out/sweep_c3.log | tail -9
timeout 200 python tools/sweep.py --workload c2 --steps 5 --variants $V --check ${V#0,} > gpurun_out/sweep_c2.log 2>&1; grep -E "variant|rror" gpurun_out/sweep_c2.log | tail -9
for wl in c5; do
  timeout 400 python bench.py --workload $wl --steps 5 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.log; echo "$wl rc=$?"; python tools/show_bench.py gpurun_out/bench_$wl.json; grep -iE "error|Traceback" -A8 gpurun_out/bench_$wl.log | head -20
done
