"""Aggregate an ncu report per CUDA source line: python scripts/ncu_lines.py report.ncu-rep [top]"""
import collections
import csv
import subprocess
import sys

def num(x):
    try:
        return int(x)
    except (ValueError, TypeError):
        return 0


rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, ix = None, None, None
agg = collections.defaultdict(lambda: collections.Counter())
src = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {}
        for i, h in enumerate(hdr):
            ix.setdefault(h, i)
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] != "":          # a source-line row (aggregate of its SASS rows)
        key = (cur_file, int(r[0]))
        src[key] = r[1]
        a = agg[key]
        a["samples"] += num(r[ix["# Samples"]])
        a["inst"] += num(r[ix["Instructions Executed"]])
        for h in hdr:
            if h.startswith("stall_") and "Not Issued" not in h:
                a[h] += num(r[ix[h]])
tot = sum(a["samples"] for a in agg.values())
toti = sum(a["inst"] for a in agg.values())
print("total samples", tot, "total warp instructions", toti)
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(((v, k) for k, v in a.items() if k.startswith("stall_")), reverse=True)[:3]
    print("%5.1f%% smp %5.1f%% ins  %s:%d  %s   [%s]" % (100 * a["samples"] / max(tot, 1), 100 * a["inst"] / max(toti, 1),
          key[0], key[1], src[key].strip()[:70], ", ".join("%s %d" % (k[6:], v) for v, k in st)))

print("---- by instruction share ----")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["inst"])[:top]:
    print("%5.1f%% ins %5.1f%% smp  %s:%d  %s" % (100 * a["inst"] / max(toti, 1), 100 * a["samples"] / max(tot, 1), key[0], key[1],
          src[key].strip()[:90]))
