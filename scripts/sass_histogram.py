#!/usr/bin/env python
"""Per-kernel SASS instruction histogram of libenspara_b200.so (cuobjdump -sass): the
mnemonics that prove which hardware paths the shipped binary uses -- UTCHMMA (tcgen05.mma),
LDTM/STTM (tcgen05.ld/st, TMEM), UTMALDG (TMA tensor loads), UBLKCP (cp.async.bulk), SYNCS
(mbarrier), DFMA (FP64 accumulate), HMMA/IMMA (legacy mma.sync; expected 0).
Runs without a GPU.  Usage: python scripts/sass_histogram.py > profiles/r2_sass_histogram.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "enspara_b200", "libenspara_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS",
        "DFMA", "FFMA", "HMMA", "IMMA", "LDG", "STG", "LDS", "STS", "ATOM", "RED", "SHFL"]


def _strip_params(name):
    """Drop the trailing parameter list of a demangled name, keep the template arguments."""
    if not name.endswith(")"):
        return name
    depth = 0
    for i in range(len(name) - 1, -1, -1):
        depth += name[i] == ")"
        depth -= name[i] == "("
        if depth == 0:
            return name[:i]
    return name


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True,
                         check=True).stdout
    demangle = subprocess.run(["cu++filt"], input="\n".join(
        re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
    names = iter(demangle)
    hist, order, cur = {}, [], None
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = next(names, m.group(1))
            cur = _strip_params(cur).replace("void eb::", "")
            if cur not in hist:
                hist[cur] = collections.Counter()
                order.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1).split(".")[0]
            hist[cur][op] += 1
            hist[cur]["_total"] += 1
    tot = collections.Counter()
    for h in hist.values():
        tot.update(h)
    print("# SASS instruction histogram of enspara_b200/libenspara_b200.so (sm_100a)\n")
    print("`cuobjdump -sass` of the shipped library, counted per kernel by "
          "`scripts/sass_histogram.py` (static counts, not executed counts).\n")
    print("Whole library: " + ", ".join("%s %d" % (k, tot[k]) for k in KEYS if tot[k]) +
          " (of %d instructions in %d kernels)\n" % (tot["_total"], len(order)))
    print("| kernel | total | " + " | ".join(KEYS) + " |")
    print("|---|---|" + "---|" * len(KEYS))
    for k in sorted(order):
        h = hist[k]
        print("| `%s` | %d | " % (k[:110], h["_total"]) +
              " | ".join(str(h[x]) if h[x] else "" for x in KEYS) + " |")


if __name__ == "__main__":
    sys.exit(main())
