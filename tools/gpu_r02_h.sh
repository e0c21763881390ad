#!/bin/bash
# round 2, session H: bench.py plumbing (tiny4 both modes), the headline line (c3), and c4 on one GPU
mkdir -p gpurun_out
for m in replicated sharded; do
  timeout 120 python bench.py --workload tiny4 --mode $m --steps 3 --parity-queries 2000 > gpurun_out/bench_tiny4_$m.json 2> gpurun_out/bench_tiny4_$m.log; echo "tiny4 $m rc=$?"; tail -c 600 gpurun_out/bench_tiny4_$m.json; echo; grep -iE "error|Traceback" -A5 gpurun_out/bench_tiny4_$m.log | head -20
done
timeout 300 python bench.py --workload c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.log; echo "c3 rc=$?"; python tools/show_bench.py gpurun_out/bench_c3.json; grep -iE "error|Traceback" -A5 gpurun_out/bench_c3.log | head
timeout 600 python bench.py --workload c4 --steps 3 --no-e2e --parity-queries 10000 --no-cpu-baseline > gpurun_out/bench_c4_1gpu.json 2> gpurun_out/bench_c4_1gpu.log; echo "c4 rc=$?"; grep "\[bench\]" gpurun_out/bench_c4_1gpu.log | tail -12; python tools/show_bench.py gpurun_out/bench_c4_1gpu.json; grep -iE "error|Traceback" -A8 gpurun_out/bench_c4_1gpu.log | head -30
nvidia-smi --query-gpu=memory.used --format=csv | tail -1
