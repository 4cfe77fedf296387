#!/usr/bin/env python
"""Per-kernel SASS evidence of the shipped library (no GPU needed): counts of the Blackwell-native mnemonics
(/opt/skills/guides/B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, bulk copies -> UBLKCP / UTMALDG,
mbarrier -> SYNCS, tcgen05.commit -> UTCBAR, tcgen05.alloc -> UTCATOMSWS), instruction count, registers.

    python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "turboae_b200", "lib", "libturboae_b200.so")
PATS = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTCATOMSWS", "SYNCS", "MUFU.EX2", "MUFU.TANH",
        "HMMA", "STS", "LDS", "ELECT", "STL", "LDL"]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except Exception:
        return name


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True)
    regs = dict(re.findall(r"Function (\S+):\s*\n\s*REG:(\d+)", res.stdout + res.stderr))
    cur, counts, n_ins = None, collections.OrderedDict(), collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if cur and m:
            ins = m.group(1)
            n_ins[cur] += 1
            for p in PATS:
                if re.search(r"(^|\s|@!?U?P\d\s+)" + re.escape(p) + r"(\.|\s|$)", ins):
                    if p == "UTCHMMA" and "UTCHMMA.2CTA" in ins:
                        continue
                    counts[cur][p] += 1
    print("# SASS summary of %s (cuobjdump -sass; sm_100a)" % os.path.relpath(LIB, ROOT))
    print("# columns: instructions, registers, then counts of " + ", ".join(PATS))
    for fn, c in counts.items():
        short = demangle(fn).replace("(anonymous namespace)::", "").replace("void ", "").replace("tae::", "")
        short = re.sub(r"\((?!.*\().*$", "", short) if short.count("(") == 1 else short.split("(")[0]
        print("%-44s ins %5d regs %3s  %s" % (short[:44], n_ins[fn], regs.get(fn, "?"),
                                              "  ".join("%s=%d" % (p, c[p]) for p in PATS if c[p])))


if __name__ == "__main__":
    main()
