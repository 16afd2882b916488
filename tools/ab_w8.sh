B="python bench.py --no-extras --no-cpu-baseline --no-e2e --steps 40 --instances 16384 --size 256 --window 64 --pool-mib 1024"
for q in 1 0; do
  CARLE_QUAD=$q $B --rule B368/S245 --fused-reductions 2>/dev/null | python -c "
import sys,json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('RESULT morley+sums quad=$q', 'us/step', round(d['ms_per_step']*1e3,2), 'cells/s %.3e'%d['value'], 'frac', round(d['roofline']['frac'],3))
"
  CARLE_QUAD=$q $B --rule B3/S23 2>/dev/null | python -c "
import sys,json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('RESULT life-nosums quad=$q', 'us/step', round(d['ms_per_step']*1e3,2), 'cells/s %.3e'%d['value'], 'frac', round(d['roofline']['frac'],3))
"
done
