#!/bin/bash
# round 2, session F: v1 structure with rows ordered by the hash (cheap range search), resolver scan unroll 8
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
timeout 200 python tools/sweep.py --workload c3 --steps 5 --variants 0,0x2000,0x1000000,0x2000000,0x3000000,0x4000000,0x5000000,0x6000000,512,0x2000200 --check 0x2000,0x1000000,0x2000000,0x3000000,0x4000000,0x5000000,0x6000000 --out gpurun_out/sweep_c3.txt > gpurun_out/sweep_c3.log 2>&1
echo "sweep c3 rc=$?"; grep -E "variant|fpx dbg|rror" gpurun_out/sweep_c3.log | tail -40
timeout 120 python tools/sweep.py --workload c2 --steps 5 --variants 0,0x2000 --check 0x2000 --out gpurun_out/sweep_c2.txt > gpurun_out/sweep_c2.log 2>&1
echo "sweep c2 rc=$?"; grep -E "variant|rror" gpurun_out/sweep_c2.log | tail -5
