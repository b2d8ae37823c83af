"""profiles/traffic_<shape>.json from one `ncu --set full` capture of k_nuts_advance taken by
scripts/ncu_capture.sh <tag> k_nuts_advance scripts/profile_nuts.py <shape> <G> <iters>
(gpurun_out/raw_<tag>.csv + gpurun_out/plain_<tag>.log).  usage: traffic_from_ncu.py <tag> <shape> [launch index]"""
import csv, json, sys

tag, shape = sys.argv[1], sys.argv[2]
idx = int(sys.argv[3]) if len(sys.argv) > 3 else 1        # index into the launches list of the launch ncu profiled (NCU_SKIP)
rows = [r for r in csv.reader(l for l in open("gpurun_out/raw_%s.csv" % tag) if not l.startswith("==")) if r]
hdr, unit, val = rows[0], rows[1], rows[2]


def metric(name):
    i = hdr.index(name)
    x = float(val[i].replace(",", ""))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3,
             "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}
    return x * scale.get(unit[i], 1.0)


info = json.loads([l for l in open("gpurun_out/plain_%s.log" % tag) if l.startswith("{")][-1])
L = info["launches"][idx]
N, D = info["N"], info["D"]
rd, wr = metric("dram__bytes_read.sum"), metric("dram__bytes_write.sum")
alg = (L["leapfrogs"] * (56 * D + 4 * N) + (L["grads"] - L["leapfrogs"]) * (16 * D + 4 * N)) / L["grads"]
out = {
    "source": "ncu --clock-control none (dram__bytes_read/write.sum), k_nuts_advance, one gene-chain per CTA, "
              "scripts/profile_nuts.py %s %d (launch %d: %d gradient evaluations, %d leapfrogs)" % (shape, info["G"], idx, L["grads"], L["leapfrogs"]),
    "dram_bytes_read": rd, "dram_bytes_write": wr, "grad_evals": L["grads"], "leapfrogs": L["leapfrogs"],
    "dram_bytes_per_grad": (rd + wr) / L["grads"], "algorithmic_bytes_per_grad": alg,
    "duration_s_under_ncu": metric("gpu__time_duration.sum"),
    "issue_active_pct": float(val[hdr.index("smsp__issue_active.avg.pct_of_peak_sustained_active")]),
    "registers_per_thread": int(float(val[hdr.index("launch__registers_per_thread")])),
    "note": "Stan dimension D used for the algorithmic bytes (SURVEY 8d); the kernel streams the padded internal vectors",
}
json.dump(out, open("profiles/traffic_%s.json" % shape, "w"), indent=1)
print(json.dumps(out))
