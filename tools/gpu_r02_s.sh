#!/bin/bash
# round 2, session W: producer variant PV 1 (expect_tx + copies first, stage directory and a second arrival after)
mkdir -p gpurun_out
V=0,0x2000000,0x3000000,0x4000000,0x1000000,0,0x2000000
timeout 400 python tools/sweep.py --workload c3 --steps 8 --variants $V --check 0x2000000,0x3000000,0x4000000,0x1000000 > gpurun_out/sweep_c3.log 2>&1; grep -E "variant|rror" gpurun_out/sweep_c3.log | tail -9
timeout 200 python tools/sweep.py --workload c2 --steps 8 --variants $V --check 0x2000000,0x3000000,0x4000000,0x1000000 > gpurun_out/sweep_c2.log 2>&1; grep -E "variant|rror" gpurun_out/sweep_c2.log | tail -9
timeout 400 python -m pytest tests -m gpu -x -q -k "variants or saturation or c2_full" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
