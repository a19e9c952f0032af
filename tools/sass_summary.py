#!/usr/bin/env python
"""Per-kernel SASS evidence of the Blackwell-native path: counts of tcgen05 MMAs (UTCHMMA / UTCQMMA ...), TMEM loads
and stores (LDTM / STTM), TMA tensor loads / stores (UTMALDG / UTMASTG), TMEM allocation (UTCATOMSWS / UTCBAR) and
mbarrier waits (SYNCS) in libspb200.so.   python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pytorch_superpoint_b200", "libspb200.so")
PATTERNS = ["UTCHMMA", "UTCHMMA.WS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "SYNCS", "MUFU", "HMMA", "FFMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    demangle = {}
    names = re.findall(r"Function : (\S+)", sass)
    if names:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(names, out))
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["instructions"] += 1
            for p in PATTERNS:
                if op == p or op.startswith(p + "."):
                    counts[cur][p] += 1
            if op.startswith("UTCHMMA.WS"):
                counts[cur]["UTCHMMA.WS"] += 0  # (already counted by the prefix rule)
    arch = re.findall(r"arch = (sm_\w+)", sass)
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (arch: {sorted(set(arch))}); mnemonics per kernel")
    cols = ["instructions", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "FFMA", "MUFU"]
    print(f"{'kernel':100s} " + " ".join(f"{c:>12s}" for c in cols))
    tot = collections.Counter()
    for fn, c in counts.items():
        name = demangle.get(fn, fn)
        name = re.sub(r"\(.*", "", name)[:100]
        print(f"{name:100s} " + " ".join(f"{c[k]:12d}" for k in cols))
        tot.update(c)
    print(f"{'TOTAL':100s} " + " ".join(f"{tot[k]:12d}" for k in cols))


if __name__ == "__main__":
    sys.exit(main())
