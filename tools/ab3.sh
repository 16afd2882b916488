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
C2="--instances 4096 --size 128 --window 32"
C3="--instances 16384 --size 256 --window 64 --rule B368/S245 --fused-reductions --pool-mib 1024"
C3L="--instances 16384 --size 256 --window 64 --pool-mib 1024"
C3LS="--instances 16384 --size 256 --window 64 --pool-mib 1024 --fused-reductions"
C3M="--instances 16384 --size 256 --window 64 --pool-mib 1024 --rule B368/S245"
C4="--instances 131072 --size 64 --window 32 --pool-mib 1024"
run cfg2-tma-inter    "CARLE_FUSED_IMPL=tma CARLE_RANK=i"    $C2
run cfg2-tma-blocked  "CARLE_FUSED_IMPL=tma CARLE_RANK=b"    $C2
run cfg3-strip2-b     "CARLE_STRIP_R=2 CARLE_RANK=b"         $C3
run cfg3-strip2-i     "CARLE_STRIP_R=2 CARLE_RANK=i"         $C3
run cfg3-strip4-b     "CARLE_STRIP_R=4 CARLE_RANK=b"         $C3
run cfg3-strip4-i     "CARLE_STRIP_R=4 CARLE_RANK=i"         $C3
run cfg3life-strip2   "CARLE_STRIP_R=2"                      $C3L
run cfg3life+sums-s2  "CARLE_STRIP_R=2"                      $C3LS
run cfg3morley-s2     "CARLE_STRIP_R=2"                      $C3M
run cfg3life-strip4   "CARLE_STRIP_R=4"                      $C3L
run cfg4-tma-blocked  "CARLE_FUSED_IMPL=tma CARLE_RANK=b"    $C4
run cfg4-tma-inter    "CARLE_FUSED_IMPL=tma CARLE_RANK=i"    $C4
