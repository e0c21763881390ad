#!/bin/bash
# Round-end profiling artefacts for profiles/: launch list, full ncu capture of the hot kernel, bench lines.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"prepare_kernel|prepare_long_kernel|search_sketch_kernel|search_smem_kernel|search_wide_kernel" -s 21 -c 28 --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3 --steps 4 --no-cpu-baseline > gpurun_out/launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"search_sketch_kernel" -s 3 -c 1 -o gpurun_out/prof_final python bench.py --workload c3 --steps 1 --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1; echo "ncu rc=$?"
timeout 900 python bench.py --workload c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.log; python tools/show_bench.py gpurun_out/bench_c3.json
timeout 900 python bench.py --workload c3 --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c3_reference.json 2> gpurun_out/bench_c3_reference.log; tail -c 600 gpurun_out/bench_c3_reference.json
