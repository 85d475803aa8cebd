#!/usr/bin/env python
"""Executed warp instructions and stall samples per SOURCE LINE of a kernel in an `ncu --set full --import-source on` capture.

    python tools/ncu_lines.py <file.ncu-rep> <cubin> <mangled-name-substring> [units] [top]

The capture's source page lists SASS addresses with their execution counts; `nvdisasm --print-line-info` of a cubin compiled
from the same source with the same flags (-lineinfo) maps the addresses to lines.  `units`: divide the counts by this number
(e.g. pixel steps of the launch).  Run here on the CPU box."""
import collections
import csv
import re
import subprocess
import sys


def main():
    rep, cubin, name = sys.argv[1:4]
    units = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.split("\n")
    line, amap, infn = None, {}, False
    for l in txt:
        if l.startswith(".text."):
            infn = name in l
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            line = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            amap[int(m.group(1), 16)] = (line, m.group(2).strip())
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    si, ei, pi, ai = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Address")
    base = int(rows[2][ai], 16)
    by = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
    tot = totp = 0.0
    mism = 0
    for r in rows[2:]:
        try:
            a = int(r[ai], 16) - base
        except ValueError:
            continue
        e, p = float(r[ei] or 0), float(r[pi] or 0)
        tot += e
        totp += p
        ln, t = amap.get(a, (None, None))
        op = re.sub(r"^@!?U?P\d+\s+", "", r[si]).split()[0]
        if t is None or re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] != op.split(".")[0]:
            mism += 1
        x = by[ln]
        x[0] += e
        x[1] += p
        x[2][op.split(".")[0]] += e
    print(rows[0][1][:100])
    print(f"{len(amap)} instructions mapped, {mism} do not match the capture's SASS (must be 0); executed {tot:.0f} = {tot / units:.2f} per unit")
    files = {}
    for ln, (e, p, c) in sorted(by.items(), key=lambda kv: -kv[1][0])[:top]:
        src = ""
        if ln:
            if ln[0] not in files:
                try:
                    files[ln[0]] = open(f"introtocomputervision_b200/csrc/{ln[0]}").read().split("\n")
                except OSError:
                    files[ln[0]] = []
            f = files[ln[0]]
            src = f[ln[1] - 1].strip()[:64] if 0 < ln[1] <= len(f) else ""
        mix = " ".join(f"{k}:{v / units:.2f}" for k, v in c.most_common(3))
        print(f"{(ln[0] + ':' + str(ln[1])) if ln else '?':24s} {e / units:7.2f}/unit {100 * e / tot:5.1f}% instr {100 * p / max(totp, 1):5.1f}% samples  {mix:40s} | {src}")


if __name__ == "__main__":
    main()
