#!/bin/bash
# last check of the round on the committed build: GPU tests, smoke(), both bench arms
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.log; echo "bench rc=$?"; python tools/show_bench.py gpurun_out/bench_c3.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c3_reference.json 2> gpurun_out/bench_c3_reference.log; echo "reference rc=$?"; python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_c3_reference.json").read().strip().splitlines()[-1]); print("reference arm: %.0f q/s x%d" % (j["value"], j["cpu_baseline"]["cores"]))
PY
timeout 300 python bench.py --workload c5 --steps 5 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.log; python tools/show_bench.py gpurun_out/bench_c5.json
