#!/bin/bash
# racecheck on the exact / wide kernels through the GPU tests that reach them
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $S --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "zipf or count_overflow or kats or random_multi or c1_single" > gpurun_out/san_race_tests.log 2>&1; echo "racecheck tests rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/san_race_tests.log | tail -3; grep -E "hazard detected|Race reported" gpurun_out/san_race_tests.log | sort | uniq -c | head
