#!/bin/bash
# round 2, session V: ablations of the hot kernel's skeleton (wrong results, timing only): 3 = no counting, no recount;
# 11 = also no copies (hand-overs only); 19 = skeleton without the sketch clear; 27 = hand-overs only, no clear; 8 = no copies but counting
mkdir -p gpurun_out
timeout 400 python tools/sweep.py --workload c3 --steps 8 --variants 0,3,11,19,27,8,2,1 > gpurun_out/sweep_c3_skel.log 2>&1; grep -E "variant|rror" gpurun_out/sweep_c3_skel.log | tail -9
