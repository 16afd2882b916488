B="python bench.py --no-extras --no-cpu-baseline --no-e2e --steps 40"
run() {  # label, env assignments, bench args
  local label="$1"; shift; local envs="$1"; shift
  env $envs $B "$@" 2>/dev/null | python -c "
import sys,json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('RESULT %-28s' % '$label', d['config']['instances_per_gpu'], d['config']['grid'], 'us/step', round(d['ms_per_step']*1e3,2), 'cells/s %.3e'%d['value'], 'GB/s', round(d['roofline']['achieved']), 'frac', round(d['roofline']['frac'],3), 'clk', d['clocks']['sm_mhz'])
"
}
C1="--instances 1 --size 64 --window 32"
C2="--instances 4096 --size 128 --window 32"
C4="--instances 131072 --size 64 --window 32 --pool-mib 1024"
run cfg2-tma          "CARLE_FUSED_IMPL=tma"                          $C2
run cfg2-tma-early    "CARLE_FUSED_IMPL=tma CARLE_EARLY_ACTIONS=1"    $C2
run cfg2-tma-nopdl    "CARLE_FUSED_IMPL=tma CARLE_PDL=0"              $C2
run cfg4-tma          "CARLE_FUSED_IMPL=tma"                          $C4
run cfg4-tma-early    "CARLE_FUSED_IMPL=tma CARLE_EARLY_ACTIONS=1"    $C4
run cfg1-tma          "CARLE_FUSED_IMPL=tma"                          $C1
run cfg1-tma-early    "CARLE_FUSED_IMPL=tma CARLE_EARLY_ACTIONS=1"    $C1
