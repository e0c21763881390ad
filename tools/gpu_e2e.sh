#!/bin/bash
mkdir -p gpurun_out
for ch in 8192 16384 32768 65536; do
  timeout 300 python bench.py --workload c3 --steps 5 --no-cpu-baseline --chunk $ch > gpurun_out/e2e_$ch.json 2> gpurun_out/e2e_$ch.log
  echo "chunk=$ch"; python -c "
import json;j=json.loads(open('gpurun_out/e2e_$ch.json').read().strip().splitlines()[-1]);e=j['e2e'];print('e2e %.2fM q/s  %.3f ms/step | h2d %.3f d2h %.3f kernels %.3f ms (event time, overlapped)'%(e['value']/1e6,e['ms_per_step'],e['h2d_ms_per_step'],e['d2h_ms_per_step'],e['kernel_ms_per_step']))"
done
