"""Summarise an `ncu --set full` report (.ncu-rep) into the few numbers the roofline report cites:
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/<name>.txt]
Runs `ncu -i ... --page raw --csv` and `--page source --csv` (no GPU needed)."""
import collections
import csv
import io
import re
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows = ncu(rep, "raw")
    hdr, units = rows[0], rows[1]
    for k, vals in enumerate(rows[2:]):
        d = dict(zip(hdr, zip(units, vals)))
        print(f"== launch {k}: {d.get('Kernel Name', ('', '?'))[1]}")
        for m in RAW:
            if m in d:
                print(f"  {m:82s} {d[m][1]:>16s} {d[m][0]}")
        st = {h: float(v[1]) for h, v in d.items() if h.startswith("smsp__average_warps_issue_stalled_") and v[1]}
        print("  stall reasons (warps per issue-active cycle):")
        for h, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]:
            print(f"    {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):24s} {v:.3f}")
    src = ncu(rep, "source")
    if len(src) > 2:
        hdr = src[1]
        ix = {h: i for i, h in enumerate(hdr)}
        by_op, samples = collections.Counter(), collections.Counter()
        for r in src[2:]:
            if len(r) < len(hdr):
                continue
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
            op = ".".join((m.group(2) if m else "?").split(".")[:2])
            by_op[op] += int(r[ix["Instructions Executed"]] or 0)
            samples[op] += int(r[ix["# Samples"]] or 0)
        tot, ts = sum(by_op.values()), max(1, sum(samples.values()))
        print(f"== SASS mix of the first launch (warp instructions executed: {tot})")
        for op, n in by_op.most_common(16):
            print(f"  {op:20s} {n:12d} {100 * n / tot:5.1f}%   pc samples {100 * samples[op] / ts:5.1f}%")


if __name__ == "__main__":
    main()
