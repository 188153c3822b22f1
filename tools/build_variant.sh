#!/bin/bash
# Build an experimental libb200geom (same ABI) into variants/NAME.so:  tools/build_variant.sh NAME "-DB2_FINAL_MINBLOCKS=5" [git-rev]
# A/B it on the GPU box with B200GEOM_LIB=/root/repo/variants/NAME.so python tools/gpu_perf.py
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; EXTRA=$2; REV=$3
W=/tmp/b200_variant_$NAME
rm -rf $W && mkdir -p $W/isce2_b200 $W/include
if [ -n "$REV" ]; then
  git -C $ROOT archive $REV isce2_b200/csrc include | tar -x -C $W
else
  cp -r $ROOT/isce2_b200/csrc $W/isce2_b200/ && cp $ROOT/include/*.h $W/include/
fi
rm -f $W/isce2_b200/csrc/*.o
env -u CC -u CXX make -s -C $W/isce2_b200/csrc -j4 EXTRA="$EXTRA" > $W/build.log 2>&1
mkdir -p $ROOT/variants && cp $W/isce2_b200/libb200geom.so $ROOT/variants/$NAME.so
echo "built variants/$NAME.so"
