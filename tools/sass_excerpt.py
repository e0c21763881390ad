#!/usr/bin/env python
"""SASS evidence for the hot kernel: cuobjdump -sass of libfpx.so, the default instance of search_find_kernel.
Writes the mnemonic histogram and the lines that show the TMA bulk copies (UBLKCP), the mbarrier (SYNCS), the
fire-and-forget shared atomics (ATOMS with RZ destination), the named barriers (BAR) and the byte-sum (IDP.4A).
   python tools/sass_excerpt.py [out.txt]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "acoustid-index_b200", "libfpx.so")
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02", "sass_hot_kernel.txt")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout.split("\n")
funcs, cur = collections.OrderedDict(), None
for ln in sass:
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1); funcs[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
        funcs[cur].append(ln.rstrip())
def demangled(n):
    return subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
want = [k for k in funcs if "search_find_kernelILi8ELi2ELi8ELi4ELj2552ELi15ELb0E" in k]
assert want, "default hot-kernel instance not found"
lines = funcs[want[0]]
ops = collections.Counter()
for ln in lines:
    m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m:
        ops[m.group(1)] += 1
fam = collections.Counter()
for k, v in ops.items():
    fam[k.split(".")[0]] += v
with open(out, "w") as f:
    f.write("cuobjdump -sass acoustid-index_b200/libfpx.so (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a)\n")
    f.write("kernel: %s\n%d SASS instructions\n\n" % (demangled(want[0]), len(lines)))
    f.write("instruction families (static count): " + ", ".join("%s %d" % kv for kv in fam.most_common(28)) + "\n\n")
    for title, pat in (("TMA bulk copy global -> shared, completion on an mbarrier (producers)", r"UBLKCP"),
                       ("mbarrier operations (init / expect_tx arrive / try_wait)", r"SYNCS"),
                       ("shared-memory atomics: the count (destination RZ: the result is never read), the hot-counter and findings lists", r"ATOMS"),
                       ("named barriers (role hand-overs; 0xN = barrier id register / immediate)", r"\bBAR\."),
                       ("byte sum of the sketch read-back", r"IDP"),
                       ("128-bit shared loads of the staged postings / the sketch", r"LDS\.128")):
        hits = [ln for ln in lines if re.search(pat, ln)]
        f.write("---- %s: %d\n" % (title, len(hits)))
        for ln in hits[:14]:
            f.write(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/$", "", ln) + "\n")
        if len(hits) > 14:
            f.write("        ... %d more\n" % (len(hits) - 14))
        f.write("\n")
    f.write("other kernels of the library (static SASS instruction counts; tensor-map TMA (UTMALDG) and tensor-core\n"
            "instructions are not expected: ragged 1-D rows and integer counting):\n")
    for k, v in funcs.items():
        if "cub" in k[:40] or "3cub" in k:
            continue  # CUB's kernels (snapshot build, off the hot path)
        c = collections.Counter(re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln).group(1) for ln in v if re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln))
        f.write("  %-90s %6d instr, UBLKCP %d, ATOMS %d, ATOMG %d, REDG %d\n" % (demangled(k)[:90], len(v), c["UBLKCP"], c["ATOMS"], c["ATOMG"] + c["ATOM"], c["REDG"] + c["RED"]))
print("wrote", out)
