#!/usr/bin/env python
"""Turn the outputs of tools/gpu_profile_r01.sh (gpurun_out/) into the tracked artefacts under profiles/r01/."""
import collections, csv, json, os, shutil, subprocess, sys
R = "profiles/r01"
os.makedirs(R, exist_ok=True)
shutil.copy("gpurun_out/launches_c3.csv", R + "/launches_c3.csv")
rows = [r for r in csv.reader(open("gpurun_out/launches_c3.csv")) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(r[4].split("::")[-1].split("(")[0], [0, 0.0]); a[0] += 1; a[1] += float(r[14])
tot = sum(a[1] for a in agg.values())
with open(R + "/launch_list_c3.txt", "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --workload c3 --steps 4 (launches of the timed region; cold-cache, serialised: compare SHARES)\n")
    for k, (n, t) in agg.items():
        f.write("%-34s launches %3d  total %12.1f ns  share %5.1f%%\n" % (k, n, t, 100 * t / tot))
out = subprocess.run(["ncu", "-i", "gpurun_out/prof_final.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rs = list(csv.reader(out.splitlines())); h, u, v = rs[0], rs[1], rs[-1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__shared_mem_per_block_dynamic', 'lts__t_sector_hit_rate.pct', 'smsp__warps_eligible.avg.per_cycle_active']
vals, lines = {}, []
for w in want:
    if w in h:
        i = h.index(w); vals[w] = (v[i], u[i]); lines.append("%-72s %s %s" % (w, v[i], u[i]))
open(R + "/ncu_sketch_kernel_final_summary.txt", "w").write(
    "ncu --set full --clock-control none --import-source on -k regex:search_sketch_kernel -s 3 -c 1 ; bench.py --workload c3 --steps 1 (10 M x 120, 100 K queries x 100 terms)\n"
    "kernel: search_sketch_kernel<14,3,6>  grid 148 x 1024 threads, 196 KB dynamic smem, 1 CTA/SM (the default hot kernel at round end)\n\n" + "\n".join(lines) + "\n")
mul = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}
tr = float(vals['dram__bytes_read.sum'][0]) * mul[vals['dram__bytes_read.sum'][1]] + float(vals['dram__bytes_write.sum'][0]) * mul[vals['dram__bytes_write.sum'][1]]
json.dump({"kernel": "search_sketch_kernel", "workload": "c3", "dram_bytes_per_launch": tr,
           "source": "profiles/r01/ncu_sketch_kernel_final_summary.txt (ncu --set full, 100 K-query launch)"}, open("profiles/traffic_c3.json", "w"))
top = subprocess.run([sys.executable, "tools/ncu_lines.py", "gpurun_out/prof_final.ncu-rep", "30"], capture_output=True, text=True).stdout
# per-role budget: the line ranges are located in the source, so they follow the code
src = open("acoustid-index_b200/csrc/fpx_kernels.cu").read().split("\n")
def line_of(needle, start=0):
    for i in range(start, len(src)):
        if needle in src[i]:
            return i + 1
    raise SystemExit("marker not found: " + needle)
k0 = line_of("search_sketch_kernel(BatchArgs a) {")
lp, lr = line_of("if (warp >= kSkFirstProducer) {", k0), line_of("if (warp >= kSkFirstResolver) {", k0)
lc, ll, le = line_of("// ===== counters", k0), line_of("if (!(a.debug & 1u)) {", k0), line_of("struct S2Shared", k0)
lh, lm = line_of("// ranking helpers"), line_of("// mbarrier / TMA bulk-copy primitives")
regions = ["init:%d-%d" % (k0, lp - 1), "producer:%d-%d" % (lp, lr - 1), "resolver:%d-%d" % (lr, lc - 1),
           "counter_wait:%d-%d" % (lc, ll - 1), "counter_loop:%d-%d" % (ll, le - 1), "helpers_rank:%d-%d" % (lh, lm - 1),
           "mbar_fns:%d-%d" % (lm, k0 - 1)]
reg = subprocess.run([sys.executable, "tools/ncu_regions.py", "gpurun_out/prof_final.ncu-rep"] + regions, capture_output=True, text=True).stdout
smem = subprocess.run([sys.executable, "tools/ncu_smem_lines.py", "gpurun_out/prof_final.ncu-rep", "12"], capture_output=True, text=True).stdout
open(R + "/ncu_sketch_kernel_final_top_lines.txt", "w").write(
    top + "\nper role (source line ranges " + " ".join(regions) + "):\n" + reg +
    "\nshared-memory wavefronts per source line (source-page attribution; inlined atomics appear on two lines):\n" + smem)
for n in (2, 4, 8):
    if os.path.exists("gpurun_out/bench_%dgpu.json" % n):
        shutil.copy("gpurun_out/bench_%dgpu.json" % n, R + "/bench_c3_%dgpu.json" % n)
for a, b in (("bench_c3.json", "bench_c3_final.json"), ("bench_c3_reference.json", "bench_c3_reference_arm.json"), ("bench_c2.json", "bench_c2_final.json"),
             ("bench_c5.json", "bench_c5_final.json"), ("gpu.txt", "gpu_box.txt"), ("pytest_gpu.log", "pytest_gpu.log")):
    shutil.copy("gpurun_out/" + a, R + "/" + b)
open(R + "/e2e_timeline_c3.txt", "w").write("".join(l for l in open("gpurun_out/trace_e2e.log") if "trace" in l))
print("\n".join(lines)); print(open(R + "/launch_list_c3.txt").read())
