"""ncu --page raw --csv of ONE kernel launch -> a small JSON summary for profiles/ (and the per-pairing DRAM figure
bench.py's roofline.traffic uses).  usage: python tools/ncu_summary.py <raw.csv> <kernel> <n_pairings> <out.json>"""
import csv
import json
import sys

raw, kernel, n, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
KEYS = ["gpu__time_duration.sum", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
KEYS += [k for k in d if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
m = {k: {"value": d[k][0], "unit": d[k][1]} for k in KEYS if k in d}


def to_bytes(v, u):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


traffic = to_bytes(*d["dram__bytes_read.sum"]) + to_bytes(*d["dram__bytes_write.sum"])
json.dump({"kernel": kernel, "n_pairings": n, "dram_bytes_per_pairing": traffic / n, "metrics": m}, open(out, "w"), indent=1)
print(out, "dram B/pairing %.0f" % (traffic / n))
