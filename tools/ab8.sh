B="python bench.py --no-extras --no-cpu-baseline --no-e2e --steps 40"
run() {  # label, env assignments, bench args
  local label="$1"; shift; local envs="$1"; shift
  env $envs $B "$@" 2>/dev/null | python -c "
import sys,json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('RESULT %-28s' % '$label', d['config']['instances_per_gpu'], d['config']['grid'], 'us/step', round(d['ms_per_step']*1e3,2), 'cells/s %.3e'%d['value'], 'frac', round(d['roofline']['frac'],3))
"
}
W16="CARLE_B200_LIB=$PWD/carle_b200/lib/libcarle_w16.so"
W4="CARLE_B200_LIB=$PWD/carle_b200/lib/libcarle_w4.so"
C2="--instances 4096 --size 128 --window 32"
C4="--instances 131072 --size 64 --window 32 --pool-mib 1024"
C1="--instances 1 --size 64 --window 32"
run cfg2-w8   "X=1"  $C2
run cfg2-w16  "$W16" $C2
run cfg2-w4   "$W4"  $C2
run cfg4-w8   "X=1"  $C4
run cfg4-w16  "$W16" $C4
run cfg4-w4   "$W4"  $C4
run cfg1-w8   "X=1"  $C1
run cfg1-w4   "$W4"  $C1
