#!/bin/bash
# A/B of the three bracket-search modes of k_topo_mask (B200_MASK_MODE) on a B200: layer hashes must agree, then the
# kernel times on the bench terrain (no layover) and on the rough terrain (fold-over on every line).
cd "$(dirname "$0")/.."
for m in 0 2; do
  export B200_MASK_MODE=$m
  echo "== B200_MASK_MODE=$m"
  python tools/gpu_layer_hash.py 2>&1 | tail -1 | cut -c1-400
  python tools/gpu_perf.py --no-parity --rough 2>&1 | tail -2 | cut -c1-500
done
