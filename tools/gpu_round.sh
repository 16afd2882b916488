#!/bin/bash
# One GPU-box visit: GPU test-suite, the default bench line, the ncu launch list of the bench
# command and one `--set full` capture of its dominant kernel.  Run through gpurun; everything
# lands in gpurun_out/.
#   tools/gpu_round.sh [tag] [pytest-args...]
set -u
TAG=${1:-r2}
shift || true
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q "$@" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $OUT/${TAG}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench rc=$?"
tail -c 600 $OUT/${TAG}_bench_n1.err
BENCH_ARGS="--steps 20 --warmup 5 --no-extras --no-e2e --no-cpu-baseline --repeats 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py $BENCH_ARGS > $OUT/${TAG}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_strip_kernel -s 30 -c 1 \
    -f -o $OUT/${TAG}_strip_cfg3 python bench.py $BENCH_ARGS > $OUT/${TAG}_ncu_strip.log 2>&1
echo "ncu strip rc=$?"
ls -la $OUT | tail -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_stream_kernel -s 300 -c 1 \
    -f -o $OUT/${TAG}_stream_cfg2 python tools/cfg2_profile.py > $OUT/${TAG}_ncu_stream.log 2>&1
echo "ncu stream rc=$?"
