#!/bin/bash
mkdir -p gpurun_out
WL=c3 timeout 300 python tools/two_callers.py 100000 50000 25000 > gpurun_out/two_callers_c3.log 2>&1; grep -E "nq |rror" gpurun_out/two_callers_c3.log | tail
WL=c2 timeout 300 python tools/two_callers.py 10000 5000 > gpurun_out/two_callers_c2.log 2>&1; grep -E "nq |rror" gpurun_out/two_callers_c2.log | tail
