"""SASS op-count summary of libldconv_b200.so per kernel (cuobjdump -sass): the tcgen05 / TMEM / TMA mnemonics that prove which
kernels use the Blackwell units (B200_PROFILING.md): UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA load), UTMASTG (TMA
store), UTMAPF (TMA L2 prefetch), plus the ELECT / BRA.U.ANY single-lane loops the control warps no longer contain.
    python scripts/sass_summary.py > profiles/r2_sass_opcounts.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "experiment_yolo_b200", "libldconv_b200.so")
OPS = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "SYNCS", "BRA.U.ANY", "FFMA2", "MUFU.TANH", "REDG", "ATOMG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name)
            counts[name] = collections.Counter()
            continue
        if name is None:
            continue
        for op in OPS:
            if re.search(r"\b" + re.escape(op), line):
                counts[name][op] += 1
        if re.search(r"/\*[0-9a-f]{4}\*/", line):
            counts[name]["instructions"] += 1
    import hashlib
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    sha = bench.lib_sha16()
    print(f"library source sha256[:16] (bench.lib_sha16) = {sha}; SASS op counts per kernel (static instruction counts, not executions)")
    print("%-96s %7s " % ("kernel", "instr") + " ".join("%9s" % o for o in OPS))
    agg = collections.OrderedDict()
    for k, c in counts.items():
        base = re.sub(r"<.*", "", k).replace("void ", "").replace("ldc::", "")
        a = agg.setdefault(base, [0, collections.Counter()])
        a[0] += 1
        a[1].update(c)
    for base, (n, c) in agg.items():
        if not any(c[o] for o in OPS[:5]) and "kernel" not in base:
            continue
        print("%-96s %7d " % (f"{base} (x{n} instances)", c["instructions"]) + " ".join("%9d" % c[o] for o in OPS))


if __name__ == "__main__":
    main()
