#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): the multi-GPU tests, then bench.py at N ranks (config 4 sharded and
# config 5 banded ride in its extras).  Everything under a timeout: a wedged neighbour-flag wait must not
# hold the box.    tools/gpu_multi.sh <N> [tag]
set -u
N=${1:-2}
TAG=${2:-r2_n$N}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $OUT/${TAG}_smi.txt 2>&1
nvidia-smi topo -m >> $OUT/${TAG}_smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "banded_multi" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $OUT/${TAG}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
tail -c 1500 $OUT/${TAG}_bench.err
python tools/show_bench.py $OUT/${TAG}_bench.json
