B="python bench.py --no-extras --no-cpu-baseline --no-e2e --steps 40 --instances 16384 --size 256 --window 64 --rule B368/S245 --fused-reductions --pool-mib 1024"
for v in "" tools/variants/ctas3.so tools/variants/ctas4.so; do
  CARLE_B200_LIB=$v $B 2>/dev/null | python -c "
import sys,json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('RESULT variant=[$v]', 'us/step', round(d['ms_per_step']*1e3,2), 'cells/s %.3e'%d['value'], 'frac', round(d['roofline']['frac'],3))
"
done
B2="python bench.py --no-extras --no-cpu-baseline --no-e2e --steps 40 --instances 16384 --size 256 --window 64 --rule B3/S23 --pool-mib 1024"
for v in "" tools/variants/ctas3.so tools/variants/ctas4.so; do
  CARLE_B200_LIB=$v $B2 2>/dev/null | python -c "
import sys,json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('RESULT life-nosums variant=[$v]', 'us/step', round(d['ms_per_step']*1e3,2), 'cells/s %.3e'%d['value'], 'frac', round(d['roofline']['frac'],3))
"
done
