#!/usr/bin/env python
import json, sys
for path in sys.argv[1:]:
    try:
        j = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:
        print(path, "unreadable", e); continue
    r = j.get("roofline", {})
    print("%s: value %.2fM q/s  e2e %.2fM  ms/step %.3f | sketch %.3f exact %.3f prepare %.3f wide %.3f ms | frac %.3f (%.0f GB/s) | cpu %.0f q/s x%d | parity %s | clocks %s" % (
        path, j["value"] / 1e6, j.get("e2e", {}).get("value", 0) / 1e6, j["ms_per_step"], r.get("sketch_ms_per_step", 0),
        r.get("exact_ms_per_step", 0), r.get("prepare_ms_per_step", 0), r.get("wide_ms_per_step", 0), r.get("frac", 0),
        r.get("achieved", 0), j.get("cpu_baseline", {}).get("value", 0), j.get("cpu_baseline", {}).get("cores", 0),
        j.get("parity_bit_exact"), j.get("clocks", {}).get("sm_mhz")))
