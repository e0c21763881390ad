#!/bin/bash
# round 2, session X: e2e with two host threads submitting alternate batches (default chunking, then one chunk per call)
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.log; echo "c3 rc=$?"; python tools/show_bench.py gpurun_out/bench_c3.json; grep -iE "error|Traceback" -A8 gpurun_out/bench_c3.log | head -20
timeout 300 python bench.py --chunk 1048576 --no-cpu-baseline > gpurun_out/bench_c3_onechunk.json 2> gpurun_out/bench_c3_onechunk.log; echo "c3 one chunk rc=$?"; python tools/show_bench.py gpurun_out/bench_c3_onechunk.json
timeout 300 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.log; python tools/show_bench.py gpurun_out/bench_c2.json
