#!/bin/bash
mkdir -p gpurun_out
for dbg in ${ABL:-1 2 11}; do
  FPX_DEBUG_ABLATE=$dbg timeout 300 python bench.py --workload c3 --steps 5 --no-cpu-baseline > gpurun_out/ablate_$dbg.json 2> gpurun_out/ablate_$dbg.log
  echo "ablate=$dbg"; python tools/show_bench.py gpurun_out/ablate_$dbg.json
done
