#!/bin/bash
# round 2, session I: GPU tests + the headline bench line (e2e through the packed API, new chunk schedule)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --workload c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.log; echo "c3 rc=$?"; python tools/show_bench.py gpurun_out/bench_c3.json; grep -iE "error|Traceback" -A5 gpurun_out/bench_c3.log | head
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_c3.json').read().strip().splitlines()[-1])
print(json.dumps(j.get("e2e")))
PY
