#!/bin/bash
mkdir -p gpurun_out
for ch in ${CHUNKS:-0}; do
  timeout 600 python bench.py --workload c3 --steps 10 --no-cpu-baseline --chunk $ch > gpurun_out/bench_q_$ch.json 2> gpurun_out/bench_q_$ch.log
  echo "chunk=$ch rc=$?"; python tools/show_bench.py gpurun_out/bench_q_$ch.json
done
