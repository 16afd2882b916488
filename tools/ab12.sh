# A/B of the strip kernel's sum hand-over: atomics with a return value (libcarle_prev.so, built from
# the previous commit) vs fire-and-forget reductions read back one trip later.
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
PREV="CARLE_B200_LIB=$PWD/carle_b200/lib/libcarle_prev.so"
C3="--instances 16384 --size 256 --window 64 --pool-mib 1024"
run cfg3-prev-morley-sums "$PREV" $C3 --rule B368/S245 --fused-reductions
run cfg3-new-morley-sums  "X=1"   $C3 --rule B368/S245 --fused-reductions
run cfg3-prev-life-sums   "$PREV" $C3 --fused-reductions
run cfg3-new-life-sums    "X=1"   $C3 --fused-reductions
run cfg3-new-life         "X=1"   $C3
run cfg3-new-morley-sums-r2 "CARLE_STRIP_R=2" $C3 --rule B368/S245 --fused-reductions
run cfg3-prev-morley-sums-r2 "$PREV CARLE_STRIP_R=2" $C3 --rule B368/S245 --fused-reductions
