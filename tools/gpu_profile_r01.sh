#!/bin/bash
# Round-end artefacts for profiles/r01: launch list, full ncu capture of the hot kernel, bench lines of both arms,
# other workloads, e2e timeline, GPU test log.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt; lscpu | grep "Model name" >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"prepare_kernel|prepare_long_kernel|search_sketch_kernel|search_smem_kernel|search_wide_kernel" -s 21 -c 28 --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3 --steps 4 --no-cpu-baseline > gpurun_out/launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"search_sketch_kernel" -s 3 -c 1 -f -o gpurun_out/prof_final python bench.py --workload c3 --steps 1 --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1; echo "ncu rc=$?"
timeout 900 python bench.py --workload c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.log; python tools/show_bench.py gpurun_out/bench_c3.json
timeout 900 python bench.py --workload c3 --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c3_reference.json 2> gpurun_out/bench_c3_reference.log; tail -c 400 gpurun_out/bench_c3_reference.json; echo
for wl in c2 c5; do timeout 900 python bench.py --workload $wl --steps 5 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.log; python tools/show_bench.py gpurun_out/bench_$wl.json; done
timeout 300 python tools/trace_e2e.py 32768 > gpurun_out/trace_e2e.log 2>&1; grep -E "trace" gpurun_out/trace_e2e.log
