#!/bin/bash
# The two bench arms exactly as the driver launches them at N = 1 (reference arm first), with wall times.
set -u
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
t0=$(date +%s)
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/refarm_reference.json 2> $OUT/refarm_reference.err; echo "reference rc=$? wall=$(( $(date +%s) - t0 )) s"
t0=$(date +%s)
python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/refarm_ours.json 2> $OUT/refarm_ours.err; echo "ours rc=$? wall=$(( $(date +%s) - t0 )) s"
python - <<'PY'
import json
r=json.loads([l for l in open('gpurun_out/refarm_reference.json') if l.startswith('{')][-1])
o=json.loads([l for l in open('gpurun_out/refarm_ours.json') if l.startswith('{')][-1])
print("reference", r['value'], r['cpu_baseline']['sample'][:160])
print("ours value", o['value'], "e2e", o['e2e']['value'], "ratio", o['value']/r['value'], "e2e ratio", o['e2e']['value']/r['value'], "same config", r['config']==o['config'])
PY
