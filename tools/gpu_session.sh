#!/bin/bash
# One GPU session: smoke, microbench, gpu tests, benches.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt; lscpu | head -20 >> gpurun_out/host.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== ubench"; timeout 120 ./tools/ubench_smem > gpurun_out/ubench.log 2>&1; cat gpurun_out/ubench.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== bench c2"; timeout 600 python bench.py --workload c2 --steps 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.log; echo "rc=$?"; tail -5 gpurun_out/bench_c2.log; cat gpurun_out/bench_c2.json
echo "== bench c3"; timeout 900 python bench.py --workload c3 --steps 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.log; echo "rc=$?"; tail -8 gpurun_out/bench_c3.log; cat gpurun_out/bench_c3.json
