#!/bin/bash
# Build the experimental variants of libb200geom listed below (same ABI, one -D each) into variants/, then A/B them on a
# B200 in ONE gpurun call:
#   tools/sweep_variants.sh build
#   gpurun --timeout 600 -- 'tools/sweep_variants.sh run 2>&1 | tee gpurun_out/sweep.log'
# Every variant must print the same layer hashes as the default build (tools/gpu_layer_hash.py) before its timings count.
VARIANTS=(
  "s6|-DB2_SOLVE_MINBLOCKS=6"      # solve kernel at 84 registers (6 CTAs / SM)
  "seg2k|-DB2_SEG_MAX=2048"        # longer runs per warp: fewer idle tails, fewer warps per line
  "seg512|-DB2_SEG_MAX=512"
  "ring128|-DB2_RING=128"          # deeper look-ahead of the per-pixel constants
  "knots1k|-DB2_MASK_KNOTS=1024"   # finer piecewise-linear search guess in the layover pass
  "m512|-DB2_MASK_BLOCK=512"       # two lines per SM in the layover pass (slower on fold-over heavy terrain: check --rough)
  "ppt8|-DB2_GEO_PPT=8"            # geo2rdr: eight pixels per thread behind one exposed first load
  "ppt2|-DB2_GEO_PPT=2"
)
ROOT=$(cd "$(dirname "$0")/.." && pwd)
case "$1" in
  build)
    for v in "${VARIANTS[@]}"; do "$ROOT/tools/build_variant.sh" "${v%%|*}" "${v#*|}" & 
      while [ "$(jobs -r | wc -l)" -ge 3 ]; do sleep 2; done
    done; wait ;;
  run)
    cd "$ROOT"
    for name in default $(for v in "${VARIANTS[@]}"; do echo "${v%%|*}"; done); do
      if [ "$name" = default ]; then unset B200GEOM_LIB; else export B200GEOM_LIB="$ROOT/variants/$name.so"; [ -f "$B200GEOM_LIB" ] || continue; fi
      python tools/gpu_layer_hash.py 2>&1 | tail -1 | cut -c1-420
      python tools/gpu_perf.py --no-parity --rough 2>&1 | tail -2 | cut -c1-420
    done ;;
  *) echo "usage: $0 build|run"; exit 1 ;;
esac
