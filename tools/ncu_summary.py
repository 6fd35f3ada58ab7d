"""Condense an .ncu-rep (ncu --set full) into a small csv: one row per kernel launch with the metrics the roofline
discussion needs.  usage: ncu_summary.py <rep> <out.csv>"""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(raw))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.sum.per_cycle_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel"] + ["%s [%s]" % (c, units[ix[c]]) for c in cols if c in ix] + ["top_stalls"])
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].replace("void ctb::", "").split("(")[0]
        top = sorted(((float(r[ix[s]]), s.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                      for s in stalls), reverse=True)[:4]
        w.writerow([name] + [r[ix[c]] for c in cols if c in ix] + [" ".join("%s=%.2f" % (n, v) for v, n in top)])
print("wrote", out, len(rows) - 2, "launches")
