#!/bin/bash
# round 2, session T: smoke(), then both multi-GPU modes on the small four-segment workload with the final kernels
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
NG=$(nvidia-smi -L | wc -l) WLS=tiny4 MODES="replicated sharded" STEPS=3 bash tools/gpu_r02_mg.sh 2>&1 | grep -vE "^$" | tail -12
