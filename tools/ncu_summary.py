"""Condense an ncu capture of the bulk daily-integration kernel into what profiles/ keeps.

  python tools/ncu_summary.py <raw.csv> <source.csv> <cells> <days> <tag>
writes profiles/<tag>_bulk_metrics.txt, profiles/<tag>_bulk_opcodes.txt and profiles/fp64_work.json
(raw.csv / source.csv come from `ncu -i X.ncu-rep --page raw|source --csv`)."""
import collections
import csv
import json
import os
import re
import sys

raw_csv, src_csv, cells, days, tag = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cd = cells * days
rows = list(csv.reader(open(raw_csv)))
d = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
f = lambda k: float(d[k][1])
keep = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__sass_average_branch_targets_threads_uniform.pct",
] + sorted(k for k in d if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"))
with open(os.path.join(ROOT, "profiles", f"{tag}_bulk_metrics.txt"), "w") as out:
    out.write(f"# ncu --set full --clock-control none, k_run_bulk<bulk>, {cells} cells x {days} days (tools/profile_case.py bulk)\n")
    for k in keep:
        if k in d:
            out.write(f"{k}\t{d[k][1]}\t{d[k][0]}\n")
dfma, dmul, dadd = (f("smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % o) for o in ("dfma", "dmul", "dadd"))
unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
dram = f("dram__bytes_read.sum") * unit[d["dram__bytes_read.sum"][0]] + f("dram__bytes_write.sum") * unit[d["dram__bytes_write.sum"][0]]
work = {
    "source": f"ncu capture profiles/{tag}_bulk_metrics.txt ({cells} cells x {days} days, k_run_bulk<bulk>)",
    "fp64_inst_per_cell_day": (dfma + dmul + dadd) / cd, "dfma_per_cell_day": dfma / cd, "dmul_per_cell_day": dmul / cd,
    "dadd_per_cell_day": dadd / cd, "flops_per_cell_day": (2 * dfma + dmul + dadd) / cd,
    "thread_inst_per_cell_day": f("smsp__thread_inst_executed.sum") / cd, "dram_bytes_per_cell_day": dram / cd,
    "kernel_ms": f("gpu__time_duration.sum"), "cell_days": cd,
}
json.dump(work, open(os.path.join(ROOT, "profiles", "fp64_work.json"), "w"), indent=1)
# opcode mix from the source page
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, ie, it, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
ops, thr, samp = collections.Counter(), collections.Counter(), collections.Counter()
for r in rows[2:]:
    if len(r) <= it:
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia])
    if not m:
        continue
    op = m.group(2).split(".")[0]
    ops[op] += int(r[ie]); thr[op] += int(r[it]); samp[op] += int(r[isamp])
tot, ts = sum(ops.values()), sum(samp.values())
with open(os.path.join(ROOT, "profiles", f"{tag}_bulk_opcodes.txt"), "w") as out:
    out.write("# opcode\tshare of warp instructions\tthread-instructions per cell-day\tavg active lanes\tshare of stall samples\n")
    for op, n in ops.most_common(24):
        out.write(f"{op}\t{n / tot * 100:.1f}%\t{thr[op] / cd:.1f}\t{thr[op] / max(n, 1):.1f}\t{samp[op] / ts * 100:.1f}%\n")
print(json.dumps(work, indent=1))
