#!/usr/bin/env python
"""Regenerates profiles/sass/*.txt: for every kernel object of the tensor path, the Blackwell-specific SASS it contains
(tcgen05 MMA = UTCHMMA, TMEM loads = LDTM, TMEM alloc = UTCATOMSWS/UTCALLOC, TMA bulk copies = UBLKCP / UTMALDG, commit
barriers = UTCBAR, mbarrier waits = SYNCS, packed fp32x2 FMA = FFMA2), with counts per kernel and the first occurrences in
context -- so that the use of the sm_100a instructions can be checked from the repository without rebuilding.
    python scripts/dump_sass.py        (needs cuobjdump; run after `make -C adaptive_mcmc_b200/csrc`)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "adaptive_mcmc_b200", "csrc")
OUT = os.path.join(ROOT, "profiles", "sass")
OBJECTS = ["diamonds_tc_adapt.o", "diamonds_tc.o", "mmd_tc.o", "tc_selftest.o"]
PAT = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCBAR|UTCATOMSWS|UTCALLOC|LDTM|STTM|UBLKCP|UTMALDG|UTMASTG|SYNCS|FFMA2|MUFU\.EX2|SETMAXREG|USETMAXREG|ELECT)\b[.\w]*")


def main():
    os.makedirs(OUT, exist_ok=True)
    for obj in OBJECTS:
        path = os.path.join(CSRC, obj)
        if not os.path.exists(path):
            print("missing", path)
            continue
        txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
        kernels = re.split(r"\n\s*Function : ", txt)[1:]
        lines = [f"# cuobjdump -sass {obj} (sm_100a) -- Blackwell-specific instructions per kernel\n"]
        for k in kernels:
            name = k.split("\n", 1)[0].strip()
            try:
                dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
            except FileNotFoundError:
                dem = name
            body = k.split("\n")
            total = sum(1 for l in body if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l))
            cnt = collections.Counter(m.group(0) for l in body for m in [PAT.search(l)] if m and "/*" in l)
            if not cnt:
                continue
            lines.append(f"\n## {dem[:200]}\n   {total} SASS instructions; " + ", ".join(f"{k} x{v}" for k, v in sorted(cnt.items())) + "\n")
            shown = collections.Counter()
            for l in body:
                m = PAT.search(l)
                if m and "/*" in l and shown[m.group(1)] < 2:
                    shown[m.group(1)] += 1
                    lines.append("   " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l.rstrip()).strip() + "\n")
        with open(os.path.join(OUT, obj.replace(".o", ".txt")), "w") as f:
            f.writelines(lines)
        print("wrote", os.path.join(OUT, obj.replace(".o", ".txt")), len(lines), "lines")


if __name__ == "__main__":
    sys.exit(main())
