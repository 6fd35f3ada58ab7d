"""Join an ncu SASS-level source page (csv) with nvdisasm line info: instructions executed per SOURCE line.
usage: hotlines.py <ncu-rep> <launch-index> <mangled-substring> [top]"""
import csv
import collections
import re
import subprocess
import sys

rep, launch, sub = sys.argv[1], int(sys.argv[2]), sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 22
dis = open("/tmp/sass/all.disasm").read().splitlines()
# locate function
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and sub in l][0]
line_of = {}
cur = None
for l in dis[start + 1:]:
    if l.startswith("//------") or l.startswith(".text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"],
                     capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
h = [i for i, r in enumerate(rows) if "# Samples" in r][0]
hdr = rows[h]
ix = {x: i for i, x in enumerate(hdr)}
data = [r for r in rows[h + 1:] if len(r) > ix["Instructions Executed"] and r[ix["Instructions Executed"]].isdigit()]
# the page may repeat; keep first occurrence of each address
seen, uniq = set(), []
for r in data:
    if r[ix["Address"]] in seen:
        continue
    seen.add(r[ix["Address"]])
    uniq.append(r)
base = min(int(r[ix["Address"]], 16) for r in uniq)
inst = collections.Counter()
smp = collections.Counter()
for r in uniq:
    off = int(r[ix["Address"]], 16) - base
    key = line_of.get(off)
    inst[key] += int(r[ix["Instructions Executed"]])
    smp[key] += int(r[ix["# Samples"]])
ti, ts = sum(inst.values()), sum(smp.values())
print(rows[0][1][:80] if len(rows[0]) > 1 else "", "total inst %.1fM samples %d" % (ti / 1e6, ts))
src = {}
for k, _ in inst.most_common(top):
    if k and k[0] not in src:
        try:
            src[k[0]] = open("/root/repo/cloud_transformers_b200/csrc/" + k[0]).read().splitlines()
        except Exception:
            src[k[0]] = []
    text = src[k[0]][k[1] - 1].strip()[:90] if k and len(src.get(k[0], [])) >= k[1] else ""
    print("%5.1f%% inst %5.1f%% smp  %s:%s  %s" % (100 * inst[k] / ti, 100 * smp[k] / max(ts, 1), k[0] if k else None,
                                                 k[1] if k else None, text))
