#!/usr/bin/env python
"""Summarises ncu captures (read on the CPU box with `ncu -i`):

    python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/<name>.summary.txt
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep [--traffic workload cost directions_per_launch] > profiles/<name>.txt

`launches`: per-kernel count, mean device time and share of the total.  `full`: the raw-page metrics that matter for
this path, the stall reasons, and the SASS opcode mix with sampled stall shares from the source page; with --traffic
the DRAM bytes of the first launch are recorded in profiles/roofline_traffic.json (bench.py's roofline.traffic)."""
import csv
import io
import json
import re
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.strip()


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    agg = defaultdict(list)
    for r in rows[1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        agg[short(r[ki])].append(v)
    total = sum(sum(v) for v in agg.values())
    for k, v in agg.items():
        print(f"{k[:52]:52s} n={len(v):4d} mean={sum(v) / len(v):10.2f} us  share={100 * sum(v) / total:5.1f}%")


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def full(rep, traffic=None):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    first = None
    for n, r in enumerate(rows[2:]):
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        if first is None:
            first = d
        print(f"== launch {n}: {d.get('Kernel Name', '?')}")
        for m in METRICS:
            if m in d:
                print(f"  {m:88s} {d[m]:>16s} {u.get(m, '')}")
        stalls = {k: float(v.replace(",", "")) for k, v in d.items()
                  if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v}
        if not stalls:
            stalls = {k: float(v.replace(",", "")) for k, v in d.items()
                      if k.startswith("smsp__average_warp") and "issue_stalled" in k and k.endswith(".ratio") and v}
        print("  stall reasons (warps per issue-active cycle):")
        for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:9]:
            nm = re.sub(r"smsp__average_warps?_(latency_)?issue_stalled_|_per_issue_active\.ratio|\.ratio", "", k)
            print(f"    {nm:24s} {v:.3f}")
    try:
        src = ncu_csv(rep, "source")
    except subprocess.CalledProcessError:
        src = None
    if src:
        # the source page lists one table per launch; keep the first
        hdr = None
        ex, smp = defaultdict(float), defaultdict(float)
        for r in src:
            if "Source" in r and any("Instructions Executed" in c for c in r):
                if hdr is not None:
                    break
                hdr = r
                si = hdr.index("Source")
                ei = next(i for i, c in enumerate(hdr) if c.strip() == "# Instructions Executed" or c.strip() == "Instructions Executed")
                pi = next((i for i, c in enumerate(hdr) if c.strip() in ("# Samples", "Warp Stall Sampling (All Samples)", "Warp Stall Sampling (All Cycles)")), None)
                continue
            if hdr is None or len(r) <= max(si, ei):
                continue
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[si])
            if not m:
                continue
            op = m.group(2)
            op = ".".join(op.split(".")[:2]) if op.startswith(("IDP", "VIMNMX", "CREDUX", "LDS", "IMAD", "RED", "STS", "SHFL")) else op.split(".")[0]
            try:
                ex[op] += float(r[ei].replace(",", "") or 0)
                if pi is not None:
                    smp[op] += float(r[pi].replace(",", "") or 0)
            except ValueError:
                pass
        tot, stot = sum(ex.values()), sum(smp.values()) or 1.0
        print(f"== SASS mix of the first launch (warp instructions executed: {int(tot)})")
        for op, v in sorted(ex.items(), key=lambda kv: -kv[1])[:22]:
            print(f"  {op:22s} {int(v):12d} {100 * v / tot:5.1f}%   pc samples {100 * smp[op] / stot:5.1f}%")
    if traffic and first:
        wl, cost, dpl = traffic
        p = ROOT / "profiles" / "roofline_traffic.json"
        data = json.loads(p.read_text()) if p.exists() else {"captures": []}
        u = dict(zip(rows[0], rows[1]))

        def b(name):
            v = float(first[name].replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[name]]
        ent = {"workload": wl, "cost": cost, "directions_per_launch": float(dpl), "kernel": short(first.get("Kernel Name", "")),
               "dram_bytes_read": b("dram__bytes_read.sum"), "dram_bytes_write": b("dram__bytes_write.sum"), "source": Path(rep).name}
        data["captures"] = [e for e in data["captures"] if not (e["workload"] == wl and e["cost"] == cost and e["directions_per_launch"] == float(dpl))] + [ent]
        p.write_text(json.dumps(data, indent=1) + "\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        t = None
        if "--traffic" in sys.argv:
            i = sys.argv.index("--traffic")
            t = (sys.argv[i + 1], sys.argv[i + 2], sys.argv[i + 3])
        full(sys.argv[2], t)
