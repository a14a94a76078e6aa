"""Developer tool: summarise an Nsight Compute report (.ncu-rep) of the accumulation kernel
into a markdown file for profiles/.  Usage: ncu_summary.py report.ncu-rep out.md "title" """
import collections
import csv
import io
import re
import subprocess
import sys

rep, out_md, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")


def page(name):
    txt = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(txt)))


raw = page("raw")
hdr, units, vals = raw[0], raw[1], raw[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
want = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
lines = [f"# {title}", "", f"Source: `{rep}` (ncu --set full --clock-control none --import-source on, 1 launch).", "",
         "| metric | unit | value |", "|---|---|---|"]
for k in want:
    if k in d:
        lines.append(f"| `{k}` | {d[k][0]} | {d[k][1]} |")
lines += ["", "## Warp stall reasons (warps stalled per issue-active cycle)", "", "| reason | ratio |", "|---|---|"]
for k in sorted(d):
    m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", k)
    if m and float(d[k][1] or 0) > 0.005:
        lines.append(f"| {m.group(1)} | {float(d[k][1]):.3f} |")

src = page("source")
if len(src) > 2:
    h = src[1]
    ix = {name: i for i, name in enumerate(h)}
    rows = [r for r in src[2:] if len(r) == len(h)]
    tot = sum(int(r[ix["# Samples"]]) for r in rows) or 1
    agg = collections.defaultdict(lambda: [0, 0])
    for r in rows:
        s = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]].strip())
        op = s.split()[0].split(".")[0] if s else "?"
        agg[op][0] += int(r[ix["# Samples"]])
        agg[op][1] += int(r[ix["Instructions Executed"]])
    lines += ["", "## SASS opcode mix (whole kernel)", "", "| opcode | warp instructions executed | share of stall samples |",
              "|---|---|---|"]
    for op, (n, ex) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
        lines.append(f"| {op} | {ex:.4g} | {100 * n / tot:.1f} % |")
    lines += ["", "## Hottest instructions (stall samples)", "", "| share | executed | instruction |", "|---|---|---|"]
    for r in sorted(rows, key=lambda r: -int(r[ix["# Samples"]]))[:15]:
        lines.append(f"| {100 * int(r[ix['# Samples']]) / tot:.2f} % | {int(r[ix['Instructions Executed']]):.4g} | `{r[ix['Source']].strip()[:100]}` |")
open(out_md, "w").write("\n".join(lines) + "\n")
print("wrote", out_md)
