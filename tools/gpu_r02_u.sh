#!/bin/bash
# round 2, session U: is the gather bound by the number of 128-byte lines?  Skeleton of the hot kernel (no counting, no
# recount: debug bits 0+1) with the rows' sources as they are and rounded down to 128 bytes (bit 6; wrong data, same bytes)
mkdir -p gpurun_out
timeout 400 python tools/sweep.py --workload c3 --steps 8 --variants 0,3,67,3,67,64 > gpurun_out/sweep_c3_align.log 2>&1; grep -E "variant|rror" gpurun_out/sweep_c3_align.log | tail -9
timeout 400 python tools/sweep.py --workload c2 --steps 8 --variants 0,3,67,3,67 > gpurun_out/sweep_c2_align.log 2>&1; grep -E "variant|rror" gpurun_out/sweep_c2_align.log | tail -9
