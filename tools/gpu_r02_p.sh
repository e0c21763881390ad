#!/bin/bash
# round 2, session P: GPU tests on the committed build, the headline line, then two ranks (NUMA binding, packed e2e)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.log; echo "c3 rc=$?"; python tools/show_bench.py gpurun_out/bench_c3.json
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG > gpurun_out/bench_c3_${NG}gpu.json 2> gpurun_out/bench_c3_${NG}gpu.log; echo "x$NG rc=$?"; python tools/show_bench.py gpurun_out/bench_c3_${NG}gpu.json
python - <<PY
import json
j=json.loads(open("gpurun_out/bench_c3_${NG}gpu.json").read().strip().splitlines()[-1])
print({k: j.get(k) for k in ("value","ms_per_step","n_gpus","host_affinity")}); print("e2e", j.get("e2e"))
PY
fi
nproc; lscpu | grep -E "NUMA|Model name|Socket" | head -8
