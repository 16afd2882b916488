#!/bin/bash
# GPU visit: full GPU test-suite, A/B table (run-time switches on the default library + any library
# variants under carle_b200/lib/ab), host-side API profile, the 65536^2 torus.
set -u
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/ab_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/ab_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $OUT/ab_pytest.log
rm -f $OUT/ab_features.jsonl
python tools/ab_headline.py default >> $OUT/ab_features.jsonl
CARLE_REVERSE=0 python tools/ab_headline.py no_reverse >> $OUT/ab_features.jsonl
CARLE_OBS_WRITE_BACK=1 python tools/ab_headline.py obs_write_back >> $OUT/ab_features.jsonl
for lib in carle_b200/lib/ab/libcarle_*.so; do
  [ -e "$lib" ] || continue
  CARLE_B200_LIB=$PWD/$lib python tools/ab_headline.py "$(basename $lib .so)" >> $OUT/ab_features.jsonl
done
cat $OUT/ab_features.jsonl
python tools/api_profile.py > $OUT/api_profile.txt 2>&1; grep "===" $OUT/api_profile.txt
python tools/biggrid.py 16384 65536 > $OUT/tile_65536.txt 2>&1; cat $OUT/tile_65536.txt
