#!/bin/bash
# TEST INFRASTRUCTURE.  What the reference's wrappers do around the CmdStan executable, restated for boxes where
# /root/reference is absent (the GPU box): /root/reference/bin/splotch:122-143 (MODE=plain) and
# /root/reference/bin/splotch_vinit:129-206 (MODE=vinit).  One process per chain, then header + draws of all chains
# with every line that starts with '#' or 'l' removed.
# usage: run_wrapper.sh MODE GENE DATA_DIR OUT_DIR BINARY NUM_SAMPLES NUM_CHAINS
set -e
MODE=$1; G=$2; DATA=$3; OUT=$4; BIN=$5; NS=$6; NC=$7
SUB=$((G / 100))
mkdir -p "$OUT/$SUB"
IN="$DATA/$SUB/data_$G.R"
if [ "$MODE" = vinit ]; then
  PF="$OUT/$SUB/output_$G.csv"
  "$BIN" pathfinder num_draws=1000 num_paths="$NC" data file="$IN" output file="$PF"
  python3 - "$PF" "$OUT/$SUB/init_$G" "$NC" <<'PY'
import json, sys
import numpy as np
pf, stem, nc = sys.argv[1], sys.argv[2], int(sys.argv[3])
rows = [l for l in open(pf) if not l.startswith("#")]
header, body = rows[0].split(","), rows[1:]
for i, r in enumerate(sorted(np.random.choice(len(body), size=nc, replace=False))):
    vals = body[r].split(",")
    init = {k: float(v) for k, v in zip(header, vals) if k not in ("lp_approx__", "lp__", "path__")}
    json.dump(init, open("%s_%d.json" % (stem, i + 1), "w"), indent=2)
PY
  rm "$PF"
fi
pids=()
for C in $(seq 1 "$NC"); do
  if [ "$MODE" = vinit ]; then
    "$BIN" sample num_samples="$NS" num_warmup=150 random id="$C" init="$OUT/$SUB/init_${G}_$C.json" \
        data file="$IN" output file="$OUT/$SUB/output_${G}_$C.csv" refresh=10 &
  else
    "$BIN" sample num_samples="$NS" num_warmup="$NS" random id="$C" \
        data file="$IN" output file="$OUT/$SUB/output_${G}_$C.csv" refresh=10 &
  fi
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
grep lp__ "$OUT/$SUB/output_${G}_1.csv" > "$OUT/$SUB/combined_$G.csv"
sed '/^[#l]/d' "$OUT/$SUB"/output_"$G"_*.csv >> "$OUT/$SUB/combined_$G.csv"
rm -f "$OUT/$SUB"/output_"$G"_*.csv
