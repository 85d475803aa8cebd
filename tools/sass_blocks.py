"""Opcode mix of the largest straight-line blocks of a SASS dump (development helper):
    python tools/sass_blocks.py /tmp/dev/sass_ncc.txt [n_blocks] [units_per_block]"""
import collections
import re
import sys

path = sys.argv[1]
nblk = int(sys.argv[2]) if len(sys.argv) > 2 else 2
units = float(sys.argv[3]) if len(sys.argv) > 3 else 96.0
ins = []
for line in open(path):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
# block boundaries: branch targets and branch instructions
targets = set()
for a, t in ins:
    m = re.search(r"\b(BRA|BSSY|CALL)\S*\s+.*?(0x[0-9a-f]+)", t)
    if m:
        targets.add(int(m.group(2), 16))
blocks, cur = [], []
for a, t in ins:
    if a in targets and cur:
        blocks.append(cur); cur = []
    cur.append((a, t))
    if re.match(r"(@!?U?P\d+\s+)?(BRA|EXIT|RET|BSYNC|WARPSYNC)", t):
        blocks.append(cur); cur = []
if cur:
    blocks.append(cur)
blocks.sort(key=len, reverse=True)
for b in blocks[:nblk]:
    ops = collections.Counter()
    for a, t in b:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", t)
        ops[m.group(2)] += 1
    print(f"block @{b[0][0]:#x}: {len(b)} instrs = {len(b) / units:.2f} per unit")
    print("   " + "  ".join(f"{op}:{c / units:.2f}" for op, c in ops.most_common(18)))
