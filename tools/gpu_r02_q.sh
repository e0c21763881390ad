#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/sweep.py --workload c3 --steps 8 --variants 0,512,515 > gpurun_out/sweep_c3_timers.log 2>&1; grep -E "variant|fpx dbg|rror" gpurun_out/sweep_c3_timers.log | tail -12
timeout 300 python tools/sweep.py --workload c2 --steps 8 --variants 0,512 > gpurun_out/sweep_c2_timers.log 2>&1; grep -E "variant|fpx dbg|rror" gpurun_out/sweep_c2_timers.log | tail -12
