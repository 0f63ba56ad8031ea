"""Summarise an ncu report (read on the CPU box, no GPU needed) into a small markdown file for profiles/.
Usage: python tools/ncu_summary.py gpurun_out/solve_r1.ncu-rep profiles/solve_r1.md "title" """
import collections
import csv
import io
import subprocess
import sys

rep, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else sys.argv[1])
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__instruction_throughput.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
lines = [f"# {title}", "", f"source: `{rep}` (ncu --set full --clock-control none), read with `ncu -i ... --page raw --csv`", ""]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    lines += [f"## {name}", "", "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f"| `{k}` | {r[i]} | {units[i]} |")
    lines.append("")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
h = None
ops, tot = collections.Counter(), 0
stalls = collections.Counter()
for r in srows:
    if "Source" in r and "Instructions Executed" in r:
        h = {k: i for i, k in enumerate(r)}
        continue
    if h is None or len(r) < len(h):
        continue
    toks = r[h["Source"]].split()
    if not toks:
        continue
    op = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).split(".")[0]
    n = int(r[h["Instructions Executed"]] or 0)
    ops[op] += n
    tot += n
    for k, i in h.items():
        if k.startswith("stall_") and "Not Issued" not in k:
            stalls[k] += int(r[i] or 0)
if tot:
    lines += ["## SASS instruction mix (warp instructions executed, last kernel in the report)", "",
              "| opcode | executed | share |", "|---|---|---|"]
    for op, n in ops.most_common(16):
        lines.append(f"| {op} | {n} | {100 * n / tot:.2f}% |")
    lines += ["", f"total warp instructions: {tot}", "", "## warp stall samples", "", "| reason | share |", "|---|---|"]
    st = sum(stalls.values()) or 1
    for k, v in stalls.most_common(8):
        lines.append(f"| {k} | {100 * v / st:.1f}% |")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
