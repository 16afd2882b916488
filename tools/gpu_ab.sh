#!/bin/bash
# GPU visit: full GPU test-suite, repeated SpeedDetector-tail tests (memory ordering), A/B table
# (run-time switches on the default library + any library variants under carle_b200/lib/ab), the 65536^2 torus.
set -u
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/ab_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/ab_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $OUT/ab_pytest.log
for i in 1 2 3 4 5 6; do
  timeout 600 python -m pytest tests/test_round2_gpu.py -m gpu -x -q -k "speed_detector or rollout_plan or sharded" >> $OUT/ab_pytest_sd_repeat.log 2>&1 || echo "SD repeat $i FAILED"
done
tail -2 $OUT/ab_pytest_sd_repeat.log
rm -f $OUT/ab_features.jsonl
python tools/ab_headline.py default >> $OUT/ab_features.jsonl
python tools/ab_headline.py default_again >> $OUT/ab_features.jsonl
for lib in carle_b200/lib/ab/libcarle_*.so; do
  [ -e "$lib" ] || continue
  CARLE_B200_LIB=$PWD/$lib python tools/ab_headline.py "$(basename $lib .so)" >> $OUT/ab_features.jsonl
done
cat $OUT/ab_features.jsonl
python tools/biggrid.py 16384 65536 > $OUT/tile_65536.txt 2>&1; cat $OUT/tile_65536.txt
