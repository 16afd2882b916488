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
D1="CARLE_B200_LIB=$PWD/carle_b200/lib/libcarle_d1.so"
C1="--instances 1 --size 64 --window 32"
C2="--instances 4096 --size 128 --window 32"
C3="--instances 16384 --size 256 --window 64 --rule B368/S245 --fused-reductions --pool-mib 1024"
C3L="--instances 16384 --size 256 --window 64 --pool-mib 1024"
C4="--instances 131072 --size 64 --window 32 --pool-mib 1024"
run cfg2-tma-d2       "CARLE_FUSED_IMPL=tma"                 $C2
run cfg2-tma-d1       "CARLE_FUSED_IMPL=tma $D1"             $C2
run cfg2-tma-d2-nopdl "CARLE_FUSED_IMPL=tma CARLE_PDL=0"     $C2
run cfg2-strip        "CARLE_FUSED_IMPL=strip"               $C2
run cfg2-strip-b      "CARLE_FUSED_IMPL=strip CARLE_RANK=b"  $C2
run cfg3-strip2       "CARLE_STRIP_R=2"                      $C3
run cfg3-strip4       "CARLE_STRIP_R=4"                      $C3
run cfg3life-strip2   "CARLE_STRIP_R=2"                      $C3L
run cfg3life-strip4   "CARLE_STRIP_R=4"                      $C3L
run cfg4-tma-d2       "CARLE_FUSED_IMPL=tma"                 $C4
run cfg4-tma-d1       "CARLE_FUSED_IMPL=tma $D1"             $C4
run cfg1-tma          "CARLE_FUSED_IMPL=tma"                 $C1
