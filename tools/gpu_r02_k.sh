#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/race_after_build.py 200 build > gpurun_out/race_build.log 2>&1; echo "stress rc=$?"; tail -4 gpurun_out/race_build.log
for i in 1 2; do
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/pytest_k$i.log 2>&1; echo "run $i rc=$?"; grep -E "differ|mismatch|passed|failed" gpurun_out/pytest_k$i.log | cut -c1-300 | head -5
done
timeout 120 python tools/sweep.py --workload c2 --steps 5 --variants 0,0x1000000 --check 0x1000000 > gpurun_out/sweep_c2.log 2>&1; grep -E "variant|rror" gpurun_out/sweep_c2.log | tail -3
timeout 200 python tools/sweep.py --workload c3 --steps 5 --variants 0,0x1000000,0x2000000,0x3000000 --check 0x1000000,0x2000000,0x3000000 > gpurun_out/sweep_c3.log 2>&1; grep -E "variant|rror" gpurun_out/sweep_c3.log | tail -5
