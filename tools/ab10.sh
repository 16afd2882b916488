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
C3="--instances 16384 --size 256 --window 64 --pool-mib 1024"
run cfg3-strip2-morley-sums "CARLE_STRIP_R=2" $C3 --rule B368/S245 --fused-reductions
run cfg3-strip4-morley-sums "CARLE_STRIP_R=4" $C3 --rule B368/S245 --fused-reductions
run cfg3-strip2-life        "CARLE_STRIP_R=2" $C3
run cfg3-strip4-life        "CARLE_STRIP_R=4" $C3
run cfg3-strip4-jit-sums    "CARLE_STRIP_R=4" $C3 --rule B36/S125 --fused-reductions
