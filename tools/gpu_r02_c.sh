#!/bin/bash
# round 2, session C: phase timers of the hot kernel (short timeouts: a deadlock must not eat the budget)
mkdir -p gpurun_out
timeout 150 python tools/sweep.py --workload c3 --steps 5 --variants 0,512,0x3000000,0x3000200,0x1000000,0x7000000 --check 0x3000000,0x1000000,0x7000000 --out gpurun_out/sweep_c3.txt > gpurun_out/sweep_c3.log 2>&1
echo "sweep c3 rc=$?"; grep -E "variant|fpx dbg|rror" gpurun_out/sweep_c3.log | tail -40
