#!/bin/bash
mkdir -p gpurun_out
for dbg in ${ABL:-512 521}; do
  FPX_DEBUG_ABLATE=$dbg timeout 300 python bench.py --workload c3 --steps 5 --no-cpu-baseline > gpurun_out/var_$dbg.json 2> gpurun_out/var_$dbg.log
  echo "variant=$dbg"; python tools/show_bench.py gpurun_out/var_$dbg.json; grep "fpx dbg" gpurun_out/var_$dbg.log | tail -1
done
