#!/bin/bash
# how much of prepare_kernel's time is the directory's size?  c3's row lengths and queries on a quarter / half of its terms
mkdir -p gpurun_out
for wl in c3q c3h; do timeout 300 python tools/sweep.py --workload $wl --steps 8 --variants 0,0 > gpurun_out/sweep_$wl.log 2>&1; grep -E "variant|snapshot:|rror" gpurun_out/sweep_$wl.log | tail -3; done
