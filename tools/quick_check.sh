# full GPU test suite + the default bench line with extras (no CPU baseline / e2e), printed compactly
python -m pytest tests -m gpu -x -q > gpurun_out/t_quick.log 2>&1; tail -n 2 gpurun_out/t_quick.log
python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_quick.json") if l.startswith("{")][0])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"])
for k, v in d["extras"].items():
    print(k, v.get("cell_updates_per_sec"), v.get("ms_per_step") or v.get("us_per_step") or v.get("us_per_generation"))
PY
