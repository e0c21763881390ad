#!/bin/bash
# round 2, session R: resolver variant FV 1 (pivots loaded while the counters count, per-warp finding lists)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
V=0,0x2000000,0x1000000,0x3000000,0,0x2000000
timeout 400 python tools/sweep.py --workload c3 --steps 8 --variants $V --check 0x2000000,0x1000000,0x3000000 > gpurun_out/sweep_c3.log 2>&1; grep -E "variant|rror" gpurun_out/sweep_c3.log | tail -9
timeout 200 python tools/sweep.py --workload c2 --steps 8 --variants $V --check 0x2000000,0x1000000,0x3000000 > gpurun_out/sweep_c2.log 2>&1; grep -E "variant|rror" gpurun_out/sweep_c2.log | tail -9
