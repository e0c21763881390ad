#!/bin/bash
# round 2, session O: 40 KB against 32 KB stages on c3 (same kernel otherwise)
mkdir -p gpurun_out
V=0,0x2000000,0x1000000,0x3000000,0,0x2000000
timeout 400 python tools/sweep.py --workload c3 --steps 8 --variants $V --check 0x2000000,0x1000000,0x3000000 > gpurun_out/sweep_c3.log 2>&1; grep -E "variant|rror" gpurun_out/sweep_c3.log | tail -9
