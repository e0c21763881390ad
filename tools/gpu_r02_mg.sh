#!/bin/bash
# multi-GPU session: N = $NG ranks.  WL / MODES / STEPS select what runs.
mkdir -p gpurun_out
NG=${NG:-2}
for wl in ${WLS:-tiny4}; do
for m in ${MODES:-replicated sharded}; do
  out=gpurun_out/bench_${wl}_${m}_${NG}gpu
  timeout ${TMO:-600} python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --workload $wl --mode $m --steps ${STEPS:-3} --parity-queries ${PQ:-2000} ${EXTRA} > $out.json 2> $out.log
  echo "$wl $m x$NG rc=$?"; python tools/show_bench.py $out.json; grep -iE "error|Traceback" -A8 $out.log | head -30
  python - <<PY
import json
try:
    j=json.loads(open("$out.json").read().strip().splitlines()[-1])
    print({k: j.get(k) for k in ("value","ms_per_step","scaling","mode","parallelism","parity_checked_queries","parity_bit_exact","exchange_bytes_per_rank_per_step")})
    print("e2e", j.get("e2e"))
except Exception as e:
    print("no json", e)
PY
done; done
nvidia-smi --query-gpu=index,name --format=csv | head -10; nproc; free -g | head -2
