#!/bin/bash
# compute-sanitizer over the final code of round 2: the hot kernel on queries that almost fill a stage (tools/race_large.py)
# under racecheck / synccheck / memcheck, then GPU tests that reach the exact, wide, long-query and build kernels
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 500 $S --tool racecheck python tools/race_large.py 300 1 > gpurun_out/san_race_large.log 2>&1; echo "racecheck race_large rc=$?"; grep -E "RACECHECK SUMMARY|sketch_queries|MISMATCH" gpurun_out/san_race_large.log | tail -3
timeout 200 $S --tool synccheck python tools/race_large.py 300 1 > gpurun_out/san_sync_large.log 2>&1; echo "synccheck race_large rc=$?"; grep -E "ERROR SUMMARY|sketch_queries" gpurun_out/san_sync_large.log | tail -2
timeout 200 $S --tool memcheck python tools/race_large.py 300 1 > gpurun_out/san_mem_large.log 2>&1; echo "memcheck race_large rc=$?"; grep -E "ERROR SUMMARY|sketch_queries" gpurun_out/san_mem_large.log | tail -2
K="c1_single or kats or saturation or count_overflow or long_queries or random_multi"
timeout 500 $S --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$K" > gpurun_out/san_mem_tests.log 2>&1; echo "memcheck tests rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_mem_tests.log | tail -2
timeout 500 $S --tool synccheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$K or zipf" > gpurun_out/san_sync_tests.log 2>&1; echo "synccheck tests rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_sync_tests.log | tail -2
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
