"""Per-kernel SASS evidence for the shipped library: `cuobjdump -sass prmf_b200/libprmf_b200.so`, counted.

    python tools/sass_summary.py > profiles/sass_summary.txt

For every kernel: instruction count and the mnemonics that prove which hardware path it takes -- UBLKCP (1-D TMA bulk
copy), UTMALDG (tensor-map TMA), SYNCS (mbarrier), UTCHMMA / UTCQMMA... (tcgen05.mma), LDTM (tcgen05.ld from TMEM),
UTCBAR (tcgen05.commit), DFMA / DMMA (fp64), LDGSTS (cp.async), RED/ATOM, MEMBAR, and register / spill figures from
`cuobjdump -res-usage`."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "prmf_b200", "libprmf_b200.so")
KEYS = ["UBLKCP", "UTMALDG", "SYNCS", "UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "LDTM", "STTM", "UTCBAR", "DFMA", "DMMA",
        "DMUL", "DADD", "FFMA", "HMMA", "LDGSTS", "LDG", "STG", "LDS", "STS", "ATOM", "RED", "MEMBAR", "BAR", "MUFU"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.split("\n"):
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = line.strip()
            cur = None
    counts = collections.OrderedDict()
    cur = None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for key in KEYS:
                if op == key or op.startswith(key + "."):
                    counts[cur][key] += 1
                    break
            else:
                for key in KEYS:
                    if op.startswith(key):
                        counts[cur][key] += 1
                        break
    names = demangle(list(counts))
    print("SASS summary of prmf_b200/libprmf_b200.so (cuobjdump -sass, sm_100a); counts are static instructions\n")
    for fn, c in sorted(counts.items(), key=lambda kv: names[kv[0]]):
        shown = "  ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k])
        print(names[fn][:170])
        print("    instr=%d  %s" % (c["_total"], shown))
        if fn in usage:
            print("    " + usage[fn])
    print("\nkernels: %d" % len(counts))


if __name__ == "__main__":
    main()
