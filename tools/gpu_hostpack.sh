#!/bin/bash
# Short GPU-box visit for the host-side packing path: the packer's thread sweep on the box's cores, the
# tests that feed host actions, the default bench line.   tools/gpu_hostpack.sh [tag]
set -u
TAG=${1:-r2h}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1; echo "build rc=$?"
timeout 120 python tools/host_pack_bench.py > $OUT/${TAG}_host_pack.json 2> $OUT/${TAG}_host_pack.err; echo "host_pack rc=$?"
cat $OUT/${TAG}_host_pack.err
timeout 300 python -m pytest tests -m gpu -x -q -k "host or smoke or staged" > $OUT/${TAG}_pytest_host.log 2>&1; echo "pytest rc=$?"
tail -3 $OUT/${TAG}_pytest_host.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench rc=$?"
tail -c 1500 $OUT/${TAG}_bench_n1.err
