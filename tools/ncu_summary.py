#!/usr/bin/env python
"""Turns the raw ncu artefacts a gpurun call brings back (gpurun_out/, scratch) into the small,
tracked summaries under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
  python tools/ncu_summary.py kernel   gpurun_out/prof_vote_r1.ncu-rep profiles/r1_vote_count_full.md

`launches` aggregates the `--metrics gpu__time_duration.sum` launch list per kernel (count, total,
mean, share of all GPU time).  `kernel` extracts the metrics DESIGN.md / bench.py quote from an
`ncu --set full` report (first profiled launch): duration, DRAM bytes, pipe utilisation, issue
slots, occupancy, registers, plus the SASS instruction mix of the hot loop when source counters
are present.
"""
import collections
import csv
import io
import subprocess
import sys


def launches(src, dst):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            ns = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        key = (row["Kernel Name"].split("(")[0].replace("void ", ""), row["Grid Size"], row["Block Size"])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(v[1] for v in agg.values())
    out = ["# ncu launch list summary (`%s`)" % src, "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none` over one bench.py run; per-launch times are",
           "cold-cache and serialised, so only each kernel's SHARE of the GPU time is comparable with bench.py.", "",
           "| kernel | grid | block | launches | total us | mean us | share |", "|---|---|---|---|---|---|---|"]
    for (k, g, b), (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %s | %s | %d | %.1f | %.2f | %.3f |" % (k, g, b, n, ns / 1e3, ns / 1e3 / n, ns / total))
    out.append("")
    out.append("total GPU time %.1f us over %d launches" % (total / 1e3, sum(v[0] for v in agg.values())))
    open(dst, "w").write("\n".join(out) + "\n")


WANT = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_active", "sm__instruction_throughput.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__maximum_warps_per_active_cycle_pct",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_static",
    "launch__shared_mem_per_block_dynamic",
    "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
]


def kernel(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = ["# ncu --set full summary (`%s`)" % src, "",
           "`ncu --set full --clock-control none --import-source on`; numbers under the profiler are for analysis,",
           "never bench values.", ""]
    for li, row in enumerate(data):
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        out.append("## launch %d: `%s` grid %s block %s" % (li, d.get("Kernel Name", "?"), d.get("Grid Size", "?"),
                                                          d.get("Block Size", "?")))
        out.append("")
        out.append("| metric | value | unit |")
        out.append("|---|---|---|")
        for m in WANT:
            if m in d and d[m] != "":
                out.append("| `%s` | %s | %s |" % (m, d[m], u.get(m, "")))
        # stall reasons, top 6
        stalls = [(k, float(v.replace(",", ""))) for k, v in d.items()
                  if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v]
        stalls.sort(key=lambda kv: -kv[1])
        for k, v in stalls[:6]:
            out.append("| `%s` | %.3f | warps/issue |" % (k, v))
        out.append("")
        if li >= 1:
            break
    # instruction mix of the whole kernel from the source page
    src_csv = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv", "--print-source", "sass"],
                             capture_output=True, text=True).stdout
    try:
        body = src_csv[src_csv.index('"Address"'):]
        nxt = body.find('"Kernel Name"')          # a second profiled launch follows: keep launch 0 only
        if nxt > 0:
            body = body[:nxt]
        srows = list(csv.DictReader(io.StringIO(body)))
        mix = collections.Counter()
        col = next((c for c in srows[0].keys() if c.strip().startswith("# Warp Instructions Executed")
                    or c.strip() == "Instructions Executed"), None)
        if col:
            for r in srows:
                ins = (r.get("Source") or "").strip()
                if not ins:
                    continue
                parts = ins.split()
                op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
                try:
                    mix[op.split(".")[0]] += int(float((r[col] or "0").replace(",", "") or 0))
                except (ValueError, AttributeError):
                    pass
            tot = sum(mix.values())
            if tot:
                out.append("## executed warp-instruction mix by SASS opcode (launch 0)")
                out.append("")
                out.append("| opcode | warp instructions | share |")
                out.append("|---|---|---|")
                for op, n in mix.most_common(24):
                    out.append("| %s | %d | %.3f |" % (op, n, n / tot))
                out.append("")
    except (ValueError, IndexError, StopIteration):
        pass
    open(dst, "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
