B="python bench.py --no-extras --no-cpu-baseline --no-e2e --steps 40"
run() {  # label, env assignments, bench args
  local label="$1"; shift; local envs="$1"; shift
  env $envs $B "$@" 2>/dev/null | python -c "
import sys,json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('RESULT %-28s' % '$label', d['config']['instances_per_gpu'], d['config']['grid'], d['config']['rule'], 'us/step', round(d['ms_per_step']*1e3,2), 'cells/s %.3e'%d['value'], 'frac', round(d['roofline']['frac'],3))
"
}
C2="--instances 4096 --size 128 --window 32"
C2B="--instances 32768 --size 128 --window 32 --pool-mib 1024"
C3="--instances 16384 --size 256 --window 64 --pool-mib 1024"
C4="--instances 131072 --size 64 --window 32 --pool-mib 1024"
run cfg2-life       "X=1" $C2
run cfg2-dynamic    "X=1" $C2 --rule B36/S125
run 32k-life        "X=1" $C2B
run 32k-dynamic     "X=1" $C2B --rule B36/S125
run cfg3-life       "X=1" $C3
run cfg3-dynamic    "X=1" $C3 --rule B36/S125
run cfg3-dyn-sums   "X=1" $C3 --rule B36/S125 --fused-reductions
run cfg4-life       "X=1" $C4
run cfg4-dynamic    "X=1" $C4 --rule B36/S125
