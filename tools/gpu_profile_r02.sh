#!/bin/bash
# Round-2 artefacts for profiles/r02 (tools/make_profiles_r02.py turns them into the tracked summaries): GPU tests, launch
# list, full ncu captures of the hot kernel and of prepare_kernel, bench lines of both arms, the other workloads, e2e timeline.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt; lscpu | grep "Model name" >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --workload c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.log; python tools/show_bench.py gpurun_out/bench_c3.json
timeout 900 python bench.py --workload c3 --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c3_reference.json 2> gpurun_out/bench_c3_reference.log; tail -c 300 gpurun_out/bench_c3_reference.json; echo
for wl in c2 c5; do timeout 900 python bench.py --workload $wl --steps 5 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.log; python tools/show_bench.py gpurun_out/bench_$wl.json; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"prepare_kernel|prepare_long_kernel|search_find_kernel|search_smem_kernel|search_wide_kernel" -s 18 -c 28 --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3 --steps 4 --no-cpu-baseline --no-e2e > gpurun_out/launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"search_find_kernel" -s 3 -c 1 -f -o gpurun_out/prof_find_final python bench.py --workload c3 --steps 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_final.log 2>&1; echo "ncu find rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"prepare_kernel" -s 3 -c 1 -f -o gpurun_out/prof_prep_final python bench.py --workload c3 --steps 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_prep.log 2>&1; echo "ncu prepare rc=$?"
timeout 300 python tools/sweep.py --workload c3 --steps 5 --variants 0,512 > gpurun_out/sweep_c3_timers.log 2>&1; grep -E "variant|fpx dbg" gpurun_out/sweep_c3_timers.log
timeout 300 python tools/trace_e2e.py 131072 > gpurun_out/trace_e2e.log 2>&1; grep -E "trace|per call" gpurun_out/trace_e2e.log
