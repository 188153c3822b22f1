#!/bin/bash
# A/B of the three bracket-search modes of k_topo_mask (B200_MASK_MODE) on a B200: layer hashes must agree, then the
# kernel times on the bench terrain (no layover) and on the rough terrain (fold-over on every line).
cd "$(dirname "$0")/.."
for m in "0 160" "0 0" "2 160"; do
  set -- $m
  export B200_MASK_MODE=$1 B200_MASK_ELEV_KB=$2
  echo "== B200_MASK_MODE=$1 B200_MASK_ELEV_KB=$2"
  python tools/gpu_layer_hash.py 2>&1 | tail -1 | cut -c1-400
  python tools/gpu_perf.py --no-parity --rough 2>&1 | tail -2 | cut -c1-500
done
