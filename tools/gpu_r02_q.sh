#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/sweep.py --workload c3 --steps 8 --variants 0,512,0 > gpurun_out/sweep_c3_timers.log 2>&1; grep -E "variant|fpx dbg|rror" gpurun_out/sweep_c3_timers.log | tail -8
