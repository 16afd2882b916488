#!/usr/bin/env python
"""Print the numbers of a bench.py JSON line as a compact table (tools/show_bench.py FILE)."""
import json
import sys


def main(path):
    # (under torchrun a library banner may precede the line on stdout)
    d = json.loads([ln for ln in open(path) if ln.lstrip().startswith("{")][-1])
    r, e = d["roofline"], d["e2e"]
    print(f"value {d['value']:.4e}  ms/step {d['ms_per_step']:.5f}  n_gpus {d['n_gpus']}  steps {d['steps']}")
    print(f"roofline {r['achieved']:.0f} GB/s frac {r['frac']:.3f} us/launch {r['us_per_launch']:.2f} "
          f"share {r.get('share_of_step')} traffic {r['traffic']}")
    if e:
        print(f"e2e {e['value']:.4e} ms/step {e['ms_per_step']:.4f} h2d bytes {e['h2d_bytes_per_step']} "
              f"host input GB/s {e.get('host_input_gbs', e.get('h2d_gbs_lower_bound', 0)):.1f}")
    for k, v in (d.get("e2e_variants") or {}).items():
        print(f"  e2e.{k:22s} {v['value']:.4e}  ms/step {v['ms_per_step']:.4f}")
    if d.get("cpu_baseline"):
        print(f"cpu_baseline {d['cpu_baseline']['value']:.4e} cores {d['cpu_baseline']['cores']}")
    for k, v in (d.get("extras") or {}).items():
        if "error" in v:
            print(f"  {k:36s} ERROR {v['error'][:150]}")
            continue
        rate = v.get("cell_updates_per_sec")
        rest = {a: b for a, b in v.items() if a not in ("note", "cell_updates_per_sec") and not isinstance(b, dict)}
        print(f"  {k:36s} {rate if rate is None else format(rate, '.4e')}  " +
              " ".join(f"{a}={b:.4g}" if isinstance(b, float) else f"{a}={b}" for a, b in rest.items()))
        for a, b in v.items():
            if isinstance(b, dict):
                print(f"    .{a}: {b.get('cell_updates_per_sec', 0):.4e} ms/step {b.get('ms_per_step')}")


if __name__ == "__main__":
    main(sys.argv[1])
