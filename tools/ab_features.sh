#!/bin/bash
# Build alternative libraries with parts of the one-launch kernels compiled out (HERE, no GPU),
# e.g.  tools/ab_features.sh build nosd "-DCARLE_FEAT_SD=0"   ->  carle_b200/lib/ab/libcarle_nosd.so
# and measure all of them on the GPU box:  tools/ab_features.sh run  (appends to gpurun_out/ab_features.jsonl)
set -u
cd "$(dirname "$0")/.."
case "${1:-}" in
  build)
    mkdir -p carle_b200/lib/ab
    CARLE_NVCC_EXTRA="$3" python -m carle_b200.build --out=carle_b200/lib/ab/libcarle_$2.so ;;
  run)
    mkdir -p gpurun_out
    python tools/ab_headline.py default >> gpurun_out/ab_features.jsonl
    for lib in carle_b200/lib/ab/libcarle_*.so; do
      [ -e "$lib" ] || continue
      CARLE_B200_LIB=$PWD/$lib python tools/ab_headline.py "$(basename $lib .so)" >> gpurun_out/ab_features.jsonl
    done
    cat gpurun_out/ab_features.jsonl ;;
  *) echo "usage: $0 build <tag> <nvcc-extra> | run" ;;
esac
