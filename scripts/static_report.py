"""Static evidence gathered WITHOUT a GPU (nvcc cross-compiles sm_100a here): ptxas -v resource usage of every kernel and the
SASS instruction mix of the hot loop of selected kernels (cuobjdump -sass).  Writes profiles/<tag>_static_ptxas.txt.

    python scripts/static_report.py r01d
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "ddcmd_b200", "csrc", "api.cu")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-ffp-contract=off"]


def main(tag):
    obj = "/tmp/_static_api.o"
    r = subprocess.run(["nvcc"] + FLAGS + ["-Xptxas", "-v", "-c", SRC, "-o", obj], capture_output=True, text=True, check=True)
    rows = []
    name = None
    props = ""
    for line in r.stderr.splitlines():
        m = re.search(r"Compiling entry function '([^']+)'", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            continue
        if "bytes stack frame" in line:
            props = line.strip()
        m = re.search(r"Used (\d+) registers(.*)", line)
        if m and name:
            sm = re.search(r"(\d+) bytes smem", line)
            sp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", props)
            rows.append((name, int(m.group(1)), int(sm.group(1)) if sm else 0, int(sp.group(1)), int(sp.group(2)), int(sp.group(3))))
            name = None
    out = ["# static resource usage, nvcc %s (cross-compiled, no GPU): registers / static smem / stack / spill bytes" % " ".join(FLAGS[:2]),
           "%-42s %5s %7s %6s %7s %7s" % ("kernel", "regs", "smem_B", "stack", "spill_st", "spill_ld")]
    for n, regs, smem, stack, s1, s2 in sorted(rows):
        out.append("%-42s %5d %7d %6d %7d %7d" % (n[:42], regs, smem, stack, s1, s2))
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    for kern, note in (("k_nbr_cell", "one-pass list build: whole kernel (the 4-way unrolled candidate loop dominates)"),
                       ("k_pairILb0", "pair kernel, force-only variant: whole kernel")):
        m = re.search(r"Function : (_Z\d+%s\S*)(.*?)(?=Function : |\Z)" % kern, sass, re.S)
        if not m:
            continue
        ops = collections.Counter(re.findall(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", m.group(2), re.M))
        total = sum(ops.values())
        out += ["", "# SASS instruction mix of %s (%s): %d instructions" % (kern, note, total),
                "  " + "  ".join("%s %d" % kv for kv in ops.most_common(18))]
        fp64 = sum(v for k, v in ops.items() if k in ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX"))
        ld = ops.get("LDG", 0)
        out.append("  fp64 pipe instructions %d (%.0f%%), LDG %d, LDS/STS %d/%d" % (fp64, 100.0 * fp64 / total, ld, ops.get("LDS", 0), ops.get("STS", 0)))
    path = os.path.join(ROOT, "profiles", "%s_static_ptxas.txt" % tag)
    open(path, "w").write("\n".join(out) + "\n")
    print(path)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "static")
