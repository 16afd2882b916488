# GPU test suite + bench with an alternative build of the library (CARLE_B200_LIB), then the bench
# with the default library for comparison
export ALT=$PWD/carle_b200/lib/libcarle_flags.so
CARLE_B200_LIB=$ALT python -m pytest tests -m gpu -x -q > gpurun_out/t_alt.log 2>&1; tail -n 2 gpurun_out/t_alt.log
for lib in $ALT ""; do
  CARLE_B200_LIB=$lib python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
  python - <<'PY'
import json, os
d = json.loads([l for l in open("gpurun_out/bench_ab.json") if l.startswith("{")][0])
print("LIB", os.environ.get("CARLE_B200_LIB") or "default", d["value"], d["ms_per_step"], d["roofline"]["frac"])
for k in ("cfg3_morley_speed", "cfg3_shape_life_no_sums", "cfg4_shard_131072x64x64", "device_random_agent_fused"):
    v = d["extras"][k]; print("   ", k, v.get("cell_updates_per_sec"), v.get("ms_per_step") or v.get("us_per_step"))
PY
done
