"""SASS evidence for libctb200.so: per kernel, the counts of the mnemonics that matter (TMA bulk copies UBLKCP,
mbarrier SYNCS, shared-memory atomics ATOMS.*, L2 atomics REDG / ATOMG, packed fp32 FFMA2 / FMUL2 / FADD2, programmatic
dependent launch PREEXIT / ACQBULK, MATCH for the plan's radix sort) and registers / spill stack / shared memory from
cuobjdump --dump-resource-usage.    usage: python tools/sass_evidence.py > profiles/rNN_sass_evidence.txt"""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cloud_transformers_b200", "libctb200.so")
KEEP = ("UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "ATOMS", "REDG", "ATOMG", "RED.", "FFMA2", "FMUL2", "FADD2", "PREEXIT",
        "ACQBULK", "MATCH", "LDS", "STS", "LDG", "STG", "BAR", "F2I", "SHFL", "UTC", "LDTM", "HMMA")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sha = hashlib.sha256(open(LIB, "rb").read()).hexdigest()[:16]
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = " ".join(re.findall(r"(?:REG|STACK|SHARED|LOCAL):\d+", line))
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, total, fn = collections.defaultdict(collections.Counter), collections.Counter(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            op = m.group(1)
            total[fn] += 1
            for k in KEEP:
                if op.startswith(k):
                    key = op if k == "ATOMS" else k
                    if k == "ATOMS":
                        key = ".".join(op.split(".")[:2])
                    counts[fn][key] += 1
                    break
    names = demangle(sorted(total))
    print("# SASS evidence for cloud_transformers_b200/libctb200.so (sm_100a), sha256[:16] = %s" % sha)
    print("# UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier, ATOMS = native shared-memory atomics, REDG / ATOMG = L2")
    print("# atomics, FFMA2 / FMUL2 / FADD2 = packed fp32 (sm_100), PREEXIT = griddepcontrol.launch_dependents, ACQBULK =")
    print("# griddepcontrol.wait, MATCH = __match_any_sync (plan radix sort).  No tensor-core mnemonics (UTC*MMA / HMMA): the")
    print("# path is byte / atomic bound by design (north star).  Columns: mnemonic counts | total SASS | resources")
    for fn in sorted(total, key=lambda f: names[f]):
        short = re.sub(r"\(.*", "", names[fn]).replace("void ", "").replace("ctb::", "")
        short = short.replace("(bool)", "").replace("(int)", "")
        c = counts[fn]
        print("%-58s %s total=%d | %s" % (short[:58], " ".join("%s=%d" % kv for kv in sorted(c.items())), total[fn],
                                          usage.get(fn, "")))


if __name__ == "__main__":
    main()
