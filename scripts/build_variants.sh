#!/bin/bash
# Build development variants of the library side by side (CPU box, parallel): scripts/build_variants.sh name1:"-DX -DY" name2:"" ...
# -> csplotch_b200/libcsplotch_b200_<name>.so (K in {1, 10} only; load with CSPLOTCH_LIB)
cd "$(dirname "$0")/.."
pids=()
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  ( CSPLOTCH_DEV_DEFS="$defs" python -m csplotch_b200.build --dev --out "$PWD/csplotch_b200/libcsplotch_b200_$name.so" > /tmp/build_$name.log 2>&1 \
      && echo "built $name [$defs]" || { echo "FAILED $name"; tail -20 /tmp/build_$name.log; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
