"""CPU box: SASS / resource summary of the built library per kernel -> markdown on stdout.

    python scripts/sass_summary.py [path/to/libcsplotch_b200.so] > profiles/r02_sass_summary.md

Per kernel: registers, stack frame, static shared memory (cuobjdump -res-usage) and counts of the mnemonics that show
what the code is made of: UBLKCP (1-D TMA bulk copies), SYNCS (mbarrier), UCGABAR / CGABAR (cluster barrier), DFMA / DADD /
DMUL (fp64 pipe), MUFU, SHFL, ATOMS (shared atomics), LDS / STS, LDG / STG / LD / ST (generic), LDL / STL (local = spills),
BAR (CTA barriers).  No tensor-core or tensor-map instructions are expected (1-D streams, fp64 path)."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "csplotch_b200/libcsplotch_b200.so"
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
name = None
for ln in res.splitlines():
    m = re.search(r"Function (\S+):", ln)
    if m:
        name = m.group(1)
        continue
    if name and "REG:" in ln:
        usage[name] = dict(re.findall(r"(\w+):(\d+)", ln))
        name = None
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts = collections.defaultdict(collections.Counter)
cur = None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if cur and m:
        counts[cur][m.group(1)] += 1
demangle = subprocess.run(["c++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
short = {k: re.sub(r"\(.*", "", d) for k, d in zip(counts, demangle)}
cols = ["UBLKCP", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT", "DFMA", "DADD", "DMUL", "DMMA", "MUFU", "SHFL", "ATOMS", "LDS", "STS", "LDG", "STG",
        "LD", "ST", "LDL", "STL", "BAR", "HMMA", "UTMALDG", "UTCHMMA"]
print("| kernel | regs | stack B | total instr | " + " | ".join(cols) + " |")
print("|---|---|---|---|" + "---|" * len(cols))
for k in sorted(counts, key=lambda x: short[x]):
    if "k_nuts_advance" not in short[k] and "k_leapfrog_fixed" not in short[k] and "k_lpgrad" not in short[k] and sum(counts[k].values()) < 3000:
        continue
    u = usage.get(k, {})
    print("| `%s` | %s | %s | %d | %s |" % (short[k].replace("void ", ""), u.get("REG", "?"), u.get("STACK", "?"), sum(counts[k].values()),
                                          " | ".join(str(counts[k].get(c, 0)) for c in cols)))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("\nwhole library:", ", ".join("%s %d" % (c, tot.get(c, 0)) for c in cols))
