#!/bin/bash
# memcheck over every GPU test but the full-size ones (those decode 10^8..10^9 postings)
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 1000 $S --tool memcheck python -m pytest tests -m gpu -q -k "not c3_full and not 50k and not c2_full" > gpurun_out/san_mem_all.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_mem_all.log | tail -3; grep -E "Invalid|out of bounds|Misaligned" gpurun_out/san_mem_all.log | sort | uniq -c | head
