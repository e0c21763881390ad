#!/usr/bin/env python
"""Turn the outputs of tools/gpu_profile_r02.sh (gpurun_out/) into the tracked artefacts under profiles/r02/."""
import collections, csv, json, os, shutil, subprocess, sys
R = "profiles/r02"
os.makedirs(R, exist_ok=True)
SRC = "acoustid-index_b200/csrc/fpx_kernels.cu"

# ---- launch list: shares of the step
shutil.copy("gpurun_out/launches_c3.csv", R + "/launches_c3.csv")
rows = [r for r in csv.reader(open("gpurun_out/launches_c3.csv")) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(r[4].split("::")[-1].split("(")[0], [0, 0.0]); a[0] += 1; a[1] += float(r[14])
tot = sum(a[1] for a in agg.values())
with open(R + "/launch_list_c3.txt", "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --workload c3 --steps 4 (launches of the timed region; "
            "cold-cache, serialised: compare SHARES)\n")
    for k, (n, t) in agg.items():
        f.write("%-40s launches %3d  total %12.1f ns  share %5.1f%%\n" % (k, n, t, 100 * t / tot))

# ---- full captures
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static', 'lts__t_sector_hit_rate.pct',
        'smsp__warps_eligible.avg.per_cycle_active', 'lts__t_sectors_srcunit_tex_op_read.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'dram__sectors_read.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio']
MUL = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}


def summary(rep, out, head):
    o = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rs = list(csv.reader(o.splitlines())); h, u, v = rs[0], rs[1], rs[-1]
    vals, lines = {}, []
    for w in WANT:
        if w in h:
            i = h.index(w); vals[w] = (v[i], u[i]); lines.append("%-88s %s %s" % (w, v[i], u[i]))
    stalls = sorted(((float(v[i] or 0), n) for i, n in enumerate(h) if "issue_stalled" in n and n.endswith("per_issue_active.ratio")), reverse=True)
    lines.append("")
    lines.append("warp stall reasons (warps per issue-active cycle), largest first:")
    for x, n in stalls[:8]:
        lines.append("  %-86s %.2f" % (n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), x))
    open(out, "w").write(head + "kernel: " + v[h.index("Kernel Name")] + "\n\n" + "\n".join(lines) + "\n")
    tr = sum(float(vals[k][0]) * MUL[vals[k][1]] for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
    return tr, vals


tr, vals = summary("gpurun_out/prof_find_final.ncu-rep", R + "/ncu_find_kernel_summary.txt",
                   "ncu --set full --clock-control none --import-source on -k regex:search_find_kernel -s 3 -c 1 ; bench.py --workload c3 --steps 1 "
                   "(10 M x 120, 100 K queries x 100 terms)\n")
json.dump({"kernel": "search_find_kernel", "workload": "c3", "dram_bytes_per_launch": tr,
           "source": "profiles/r02/ncu_find_kernel_summary.txt (ncu --set full, 100 K-query launch)"}, open("profiles/traffic_c3.json", "w"))
summary("gpurun_out/prof_prep_final.ncu-rep", R + "/ncu_prepare_kernel_summary.txt",
        "ncu --set full --clock-control none --import-source on -k regex:prepare_kernel -s 3 -c 1 ; bench.py --workload c3 --steps 1 (100 K queries x 100 terms: "
        "10 M directory probes)\n")
if os.path.exists("gpurun_out/prof_c5.ncu-rep"):
    summary("gpurun_out/prof_c5.ncu-rep", R + "/ncu_exact_kernel_c5_summary.txt",
            "ncu --set full --clock-control none --import-source on, search_smem_kernel<15,1024> ; bench.py --workload c5 --steps 1 (Zipf, 100 K queries; the capture was "
            "taken before rows were handed out dynamically: 62.7 ms, now 58.1)\n")

# ---- per source line / per role
top = subprocess.run([sys.executable, "tools/ncu_lines.py", "gpurun_out/prof_find_final.ncu-rep", "30"], capture_output=True, text=True).stdout
src = open(SRC).read().split("\n")


def line_of(needle, start=0):
    for i in range(start, len(src)):
        if needle in src[i]:
            return i + 1
    raise SystemExit("marker not found: " + needle)


k0 = line_of("search_find_kernel(BatchArgs a, uint32_t cls) {")
lp, lr = line_of("if (warp >= kFirstProducer) {", k0), line_of("if (warp >= kFirstResolver) {", k0)
lc, ll, lb, le = line_of("// ===== counters", k0), line_of("if (!(a.debug & 1u)) {", k0), line_of("const uint32_t b2 = ", k0), line_of("// exact shared-memory path", k0)
lh, lm = line_of("// ranking helpers"), line_of("// mbarrier / TMA bulk-copy primitives")
regions = ["init:%d-%d" % (k0, lp - 1), "producer:%d-%d" % (lp, lr - 1), "resolver:%d-%d" % (lr, lc - 1),
           "counter_wait:%d-%d" % (lc, ll - 1), "counter_count:%d-%d" % (ll, lb - 1), "counter_readback:%d-%d" % (lb, le - 1),
           "helpers_rank_barriers:%d-%d" % (lh, lm - 1), "mbar_tma_fns:%d-%d" % (lm, k0 - 1)]
reg = subprocess.run([sys.executable, "tools/ncu_regions.py", "gpurun_out/prof_find_final.ncu-rep"] + regions, capture_output=True, text=True).stdout
smem = subprocess.run([sys.executable, "tools/ncu_smem_lines.py", "gpurun_out/prof_find_final.ncu-rep", "12"], capture_output=True, text=True).stdout
open(R + "/ncu_find_kernel_top_lines.txt", "w").write(
    top + "\nper role (source line ranges " + " ".join(regions) + "; a bar.sync blocks at the next memory instruction, so barrier waits show\n"
    "up on the lines after the barrier):\n" + reg +
    "\nshared-memory wavefronts per source line (source-page attribution; inlined atomics appear on two lines):\n" + smem)
top = subprocess.run([sys.executable, "tools/ncu_lines.py", "gpurun_out/prof_prep_final.ncu-rep", "16"], capture_output=True, text=True).stdout
open(R + "/ncu_prepare_kernel_summary.txt", "a").write("\nstall samples per source line:\n" + top)

# ---- bench lines, logs
for a, b in (("bench_c3.json", "bench_c3_final.json"), ("bench_c3_reference.json", "bench_c3_reference_arm.json"), ("bench_c2.json", "bench_c2_final.json"),
             ("bench_c5.json", "bench_c5_final.json"), ("gpu.txt", "gpu_box.txt"), ("pytest_gpu.log", "pytest_gpu.log")):
    shutil.copy("gpurun_out/" + a, R + "/" + b)
open(R + "/e2e_timeline_c3.txt", "w").write("".join(l for l in open("gpurun_out/trace_e2e.log") if "trace" in l or "per call" in l))
open(R + "/phase_timers_c3.txt", "w").write(
    "tools/sweep.py --workload c3 --variants 0,512: variant 512 runs the instance of the hot kernel with phase timers (CTA 0, clock cycles per own\n"
    "query from the start of the role's iteration; counter and resolver groups take every second query; the timed instance is ~10 % slower)\n" +
    "".join(l for l in open("gpurun_out/sweep_c3_timers.log") if "variant" in l or "fpx dbg" in l))
subprocess.run([sys.executable, "tools/sass_excerpt.py", R + "/sass_hot_kernel.txt"], check=True)
print(open(R + "/launch_list_c3.txt").read()); print(open(R + "/ncu_find_kernel_summary.txt").read())
