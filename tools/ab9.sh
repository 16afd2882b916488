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
run cfg2      "X=1" --instances 4096 --size 128 --window 32
run cfg2-u8   "X=1" --instances 4096 --size 128 --window 32
run 32k-128   "X=1" --instances 32768 --size 128 --window 32 --pool-mib 1024
run cfg4      "X=1" --instances 131072 --size 64 --window 32 --pool-mib 1024
run cfg1      "X=1" --instances 1 --size 64 --window 32
