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
L3="CARLE_B200_LIB=$PWD/carle_b200/lib/libcarle_c3.so"
L4="CARLE_B200_LIB=$PWD/carle_b200/lib/libcarle_c4.so"
C2="--instances 4096 --size 128 --window 32"
C2B="--instances 32768 --size 128 --window 32 --pool-mib 1024"
run cfg2-ctas2        "CARLE_FUSED_IMPL=tma"        $C2
run cfg2-ctas3        "CARLE_FUSED_IMPL=tma $L3"    $C2
run cfg2-ctas4        "CARLE_FUSED_IMPL=tma $L4"    $C2
run cfg2-ctas4-rankb  "CARLE_FUSED_IMPL=tma $L4 CARLE_RANK=b"    $C2
run 32k-ctas2         "CARLE_FUSED_IMPL=tma"        $C2B
run 32k-ctas3         "CARLE_FUSED_IMPL=tma $L3"    $C2B
run 32k-ctas4         "CARLE_FUSED_IMPL=tma $L4"    $C2B
