#!/bin/bash
# One GPU-box visit with the reference checkout pushed as scratch (gpurun_scratch/ is git-ignored; it
# is created HERE before the call and removed afterwards -- nothing of it is committed):
#   mkdir -p gpurun_scratch && cp -r /root/reference gpurun_scratch/reference
#   gpurun -- bash tools/gpu_reference_visit.sh
# (1) the reference's own wrappers and test body on carle_b200.CARLE (tests/test_reference_wrappers_gpu.py)
# (2) the real reference's CPU step timed beside oracle/torch_port.py on this box's host cores.
set -u
OUT=gpurun_out
mkdir -p $OUT
export CARLE_REFERENCE_PATH=$PWD/gpurun_scratch/reference
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
timeout 900 python -m pytest tests/test_reference_wrappers_gpu.py -m gpu -v -rxs > $OUT/r2_reference_wrappers_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 $OUT/r2_reference_wrappers_pytest.log
timeout 600 python tests/port_vs_reference_cpu.py > $OUT/r2_port_vs_reference_cpu.json 2> $OUT/r2_port_vs_reference_cpu.err
echo "cpu compare rc=$?"; cat $OUT/r2_port_vs_reference_cpu.json
