#!/usr/bin/env python
"""Where a kernel's time goes OUTSIDE its inner loop: stall samples against executed instructions per block of SASS.

    python tools/ncu_hotspots.py <file.ncu-rep> [block-bytes-hex [lo-hex hi-hex]]

Reads the source page of an `ncu --set full --import-source on` capture (run here, on the CPU box).  Prints, per block of
SASS (default 0x400 bytes = 64 instructions), its share of executed warp instructions and of warp stall samples; blocks whose
sample share is well above their instruction share are the ones to look at (a row epilogue made of divergence regions, a
producer path redoing integer divisions).  With lo/hi: the instructions of that address range with their counts, and the 40
instructions holding the most samples with their two leading stall reasons."""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    si, ei, pi, ai = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Address")
    stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_")]
    base = int(rows[2][ai], 16)
    data = []
    for r in rows[2:]:
        try:
            a = int(r[ai], 16) - base
        except ValueError:
            continue
        data.append((a, r[si].strip(), float(r[ei] or 0), float(r[pi] or 0), r))
    tot, toti = sum(d[3] for d in data), sum(d[2] for d in data)
    print(rows[0][1][:110])
    print(f"warp instructions executed {toti:.0f}, stall samples {tot:.0f}")
    block = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0x400
    b = collections.OrderedDict()
    for a, s, e, p, _ in data:
        x = b.setdefault(a // block, [0, 0, 0])
        x[0] += e; x[1] += p; x[2] += 1
    for k, (e, p, n) in b.items():
        if p / tot > 0.004:
            print(f"{k * block:#08x}  instr {e / toti * 100:5.1f}%  samples {p / tot * 100:5.1f}%  ratio {(p / tot) / (e / toti + 1e-9):4.1f}  executions per instruction {e / n / 1e6:6.2f}M")
    if len(sys.argv) > 4:
        lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
        for a, s, e, p, _ in data:
            if lo <= a <= hi:
                print(f"{a:#08x} {s[:84]:84s} {int(e) // 1000:8d}k {int(p):6d}")
    print("== the 40 instructions holding the most samples")
    for a, s, e, p, r in sorted(sorted(data, key=lambda d: -d[3])[:40], key=lambda d: d[0]):
        st = sorted(((float(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        print(f"{a:#08x} {s[:70]:70s} {int(e) // 1000:8d}k {int(p):6d} {100 * p / tot:4.1f}%  {st}")


if __name__ == "__main__":
    main()
