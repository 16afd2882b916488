#!/bin/bash
# Last GPU-box visit of a round: the default bench line first, then smoke and the whole GPU suite.
#   tools/gpu_final.sh [tag]
set -u
TAG=${1:-r2i}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1; echo "build rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench rc=$?"
tail -c 1200 $OUT/${TAG}_bench_n1.err
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
timeout ${PYTEST_LIMIT:-330} python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 $OUT/${TAG}_pytest.log
