"""Summarise gpurun_out/bench.log (per-op table) and an ncu launch-list csv (per-kernel mean of the last rep)."""
import collections
import csv
import json
import re
import sys

IDEAL_PEAK = 6555.5


def bench(path):
    lines = [x for x in open(path) if x.startswith("{")]
    if not lines:
        print(open(path).read()[-2000:])
        return
    d = json.loads(lines[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "algorithmic_gbs", "frac_of_hbm_peak", "gpu_launches")},
          "e2e", d["e2e"]["value"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    ops = d["ops"]
    tot = collections.Counter()
    for c in ("a2d", "a3d", "b2d", "b3d", "c2d", "c3d"):
        row = [o for o in ops if o["class"] == c]
        print(c, " ".join("%s=%.3f(%2.0f%%)" % (o["op"][:9], o["ms"], 100 * o["gbs"] / IDEAL_PEAK) for o in row),
              "| sum %.3f ideal %.3f" % (sum(o["ms"] for o in row), sum(o["bytes"] for o in row) / IDEAL_PEAK / 1e6))
        for o in row:
            tot[o["op"]] += o["ms"]
    print("per-op totals:", {k: round(v, 3) for k, v in tot.items()}, "all", round(sum(tot.values()), 3))


def launches(path):
    lines = open(path).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')]
    if not start:
        print("no csv in", path)
        return
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines[start[0]:]):
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        key = (r["ID"], name)
        agg.setdefault(key, {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    for (i, name), m in agg.items():
        t, u = m.get("gpu__time_duration.sum", (0, "us"))
        t = t / 1e3 if u in ("ns", "nsecond") else (t * 1e3 if u in ("ms", "msecond") else t)
        inst = m.get("smsp__inst_executed.sum", (0, ""))[0]
        rd, ru = m.get("dram__bytes_read.sum", (0, ""))
        wr, wu = m.get("dram__bytes_write.sum", (0, ""))
        sc = {"Mbyte": 1, "Gbyte": 1e3, "Kbyte": 1e-3, "byte": 1e-6}
        print("%3s %-38s %8.1f us  inst %7.1fM  dram r %7.1f w %7.1f MB" % (i, name[:38], t, inst / 1e6,
                                                                        rd * sc.get(ru, 1), wr * sc.get(wu, 1)))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        (bench if p.endswith(".log") else launches)(p)
