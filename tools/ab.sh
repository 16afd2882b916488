set -x
B="python bench.py --no-extras --no-cpu-baseline --no-e2e --steps 40"
for impl in direct tma; do
  for cfg in "--instances 4096 --size 128 --window 32" "--instances 16384 --size 256 --window 64 --rule B368/S245 --fused-reductions --pool-mib 1024" "--instances 131072 --size 64 --window 32 --pool-mib 1024" "--instances 1048576 --size 64 --window 32 --pool-mib 8192 --steps 10"; do
    CARLE_FUSED_IMPL=$impl $B $cfg 2>&1 | python -c "
import sys,json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('RESULT $impl', d['config']['instances_per_gpu'], d['config']['grid'], 'us/step', round(d['ms_per_step']*1e3,2), 'cells/s %.3e'%d['value'], 'GB/s', round(d['roofline']['achieved']), 'frac', round(d['roofline']['frac'],3))
"
  done
done
