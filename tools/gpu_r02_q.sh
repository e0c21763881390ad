#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/sweep.py --workload c3 --steps 8 --variants 0,0x100000,0x200000,0 --check 0x100000,0x200000 > gpurun_out/sweep_c3_prep.log 2>&1; grep -E "variant|rror" gpurun_out/sweep_c3_prep.log | tail -8
