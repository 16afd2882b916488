# targeted GPU tests + config-3 bench with an alternative build of the library (CARLE_B200_LIB)
export CARLE_B200_LIB=$PWD/carle_b200/lib/libcarle_alt.so
python -m pytest tests -m gpu -x -q -k "rollout_digests or sweep or wrappers or every_step_kernel_variant or fused_step_matches or banded_single or fused_random" > gpurun_out/t_alt2.log 2>&1; tail -n 2 gpurun_out/t_alt2.log
python bench.py --no-extras --no-cpu-baseline --no-e2e --steps 40 --instances 16384 --size 256 --window 64 --pool-mib 1024 --rule B368/S245 --fused-reductions 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('cfg3 morley sums us/step', round(d['ms_per_step']*1e3,2), 'cells/s %.4e'%d['value'], 'frac', round(d['roofline']['frac'],3))
"
