#!/bin/bash
# Short GPU visit: smoke, a test subset, the headline A/B line (twice), the 65536^2 torus.
set -u
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/q_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_wrappers_gpu.py -m gpu -x -q "$@" > $OUT/q_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $OUT/q_pytest.log
rm -f $OUT/ab_features.jsonl
python tools/ab_headline.py default >> $OUT/ab_features.jsonl
CARLE_REVERSE=0 python tools/ab_headline.py no_reverse >> $OUT/ab_features.jsonl
for lib in carle_b200/lib/ab/libcarle_*.so; do
  [ -e "$lib" ] || continue
  CARLE_B200_LIB=$PWD/$lib python tools/ab_headline.py "$(basename $lib .so)" >> $OUT/ab_features.jsonl
done
cat $OUT/ab_features.jsonl
python tools/biggrid.py 65536 > $OUT/tile_65536.txt 2>&1; cat $OUT/tile_65536.txt
python tools/random_agent_bench.py > $OUT/random_agent.txt 2>&1; CARLE_RANDOM_IMPL=direct python tools/random_agent_bench.py >> $OUT/random_agent.txt 2>&1; grep RESULT $OUT/random_agent.txt
