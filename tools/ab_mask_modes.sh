#!/bin/bash
# A/B of the three bracket-search modes of k_topo_mask (B200_MASK_MODE) on a B200: layer hashes must agree, then the
# kernel times on the bench terrain (no layover) and on the rough terrain (fold-over on every line).
cd "$(dirname "$0")/.."
for m in "0 0 default" "0 0 nopf" "2 0 default"; do
  set -- $m
  export B200_MASK_MODE=$1 B200_MASK_ELEV_KB=$2
  if [ "$3" = default ]; then unset B200GEOM_LIB; else export B200GEOM_LIB="$PWD/variants/$3.so"; [ -f "$B200GEOM_LIB" ] || continue; fi
  echo "== B200_MASK_MODE=$1 B200_MASK_ELEV_KB=$2 lib=$3"
  python tools/gpu_layer_hash.py 2>&1 | tail -1 | cut -c1-400
  python tools/gpu_perf.py --no-parity --rough 2>&1 | tail -2 | cut -c1-500
done
