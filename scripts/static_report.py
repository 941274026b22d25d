"""Static evidence for kernels that have not run on the device yet: resource usage (cuobjdump --dump-resource-usage) and the
SASS opcode mix (cuobjdump -sass) of the kernels named by a regular expression, written as a markdown table.
    python scripts/static_report.py 'sbem_|_m2p_kernel' > profiles/static_r01_unrun_kernels.md
Nothing here is a measurement; it is what `-Xptxas -v` / `cuobjdump -sass` show before spending GPU time."""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else ".")
KEEP = ["DFMA", "DMUL", "DADD", "MUFU", "LDG", "STG", "LDS", "STS", "LDL", "STL", "LDC", "IMAD", "IADD3", "BRA", "BAR", "WARPSYNC"]


def demangle(name):
    return subprocess.check_output(["c++filt", "-p", name]).decode().strip().split("(anonymous namespace)::")[-1]


print("# Static report: kernels written after the round's GPU minutes were spent\n")
print("`cuobjdump --dump-resource-usage` and `cuobjdump -sass` of the objects built by `csrc/Makefile` "
      "(sm_100a, -O3, -lineinfo).  Opcode counts are static instruction counts of the whole kernel, not executed counts.\n")
print("| kernel | object | registers | shared B | local (stack) B | " + " | ".join(KEEP) + " | total |")
print("|---|---|---:|---:|---:|" + "---:|" * (len(KEEP) + 1))
for obj in sorted(glob.glob(os.path.join(ROOT, "fmm_bem_relaxed_b200", "csrc", "*.o"))):
    res = subprocess.check_output(["cuobjdump", "--dump-resource-usage", obj], stderr=subprocess.DEVNULL).decode()
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+).*?STACK:(\d+).*?SHARED:(\d+)", line)
        if m and cur:
            usage[cur] = (int(m.group(1)), int(m.group(3)), int(m.group(2)))
    sass = subprocess.check_output(["cuobjdump", "-sass", obj], stderr=subprocess.DEVNULL).decode()
    mix = collections.defaultdict(collections.Counter)
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?\S+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            mix[cur][m.group(1)] += 1
    for fn in sorted(mix):
        if not pat.search(fn):
            continue
        r = usage.get(fn, (0, 0, 0))
        c = mix[fn]
        print("| `%s` | %s | %d | %d | %d | %s | %d |" % (demangle(fn), os.path.basename(obj)[:-2] + ".cu", r[0], r[1], r[2],
                                                        " | ".join(str(c.get(k, 0)) for k in KEEP), sum(c.values())))
