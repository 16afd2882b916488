#!/usr/bin/env python
"""Summarise an `ncu --set full` capture for profiles/ (run HERE, on the .ncu-rep brought back in
gpurun_out/):

    python tools/ncu_summary.py gpurun_out/X.ncu-rep profiles/r2_X.summary.txt [--traffic KERNEL WORKLOAD]

Writes the counters the judge looks at (duration, DRAM bytes, registers, grid, pipe utilisation,
stall breakdown), a stall-sample profile over consecutive SASS regions with their opcode mix, and --
with --traffic -- records dram__bytes_read/write per launch in profiles/r2_kernel_traffic.json,
which bench.py reports as roofline.traffic for that kernel / workload."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active",
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True)
    if out.returncode:
        raise SystemExit(out.stderr)
    return list(csv.reader(io.StringIO(out.stdout)))


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return int(round(float(value) * scale))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    rows = ncu_csv(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    lines = [f"== {os.path.basename(rep)} {vals[col['Kernel Name']]}"]
    for k in KEYS:
        if k in col:
            lines.append(f"   {k} = {vals[col[k]]} {units[col[k]]}")
    stalls = {h[len("smsp__pcsamp_warps_issue_stalled_"):]: int(vals[i]) for h, i in col.items()
              if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
    total = sum(stalls.values()) or 1
    lines.append("   stalls: " + ", ".join(f"{k} {100 * v / total:.0f}%" for k, v in
                                           sorted(stalls.items(), key=lambda kv: -kv[1]) if v * 100 >= total))
    # stall samples / instructions over consecutive SASS regions
    sass = ncu_csv(rep, "source", ("--print-source", "sass"))
    shdr = sass[1]
    ix = {h: i for i, h in enumerate(shdr)}
    data = [r for r in sass[2:] if len(r) == len(shdr)]
    num = lambda r, k: int(r[ix[k]] or 0)
    tot_s = sum(num(r, "# Samples") for r in data) or 1
    tot_i = sum(num(r, "Instructions Executed") for r in data) or 1
    lines.append(f"   SASS: {len(data)} instructions, {tot_i} warp-instructions executed, {tot_s} stall samples")
    lines.append("   region (SASS index)  samples  instr   opcode mix (warp-instructions, thousands) | top stall reasons")
    chunk = 160
    for s in range(0, len(data), chunk):
        seg = data[s:s + chunk]
        smp = sum(num(r, "# Samples") for r in seg)
        ins = sum(num(r, "Instructions Executed") for r in seg)
        if smp * 200 < tot_s and ins * 200 < tot_i:
            continue
        ops = collections.Counter()
        for r in seg:
            words = [w for w in r[ix["Source"]].split() if not w.startswith("@")]
            ops[words[0].split(".")[0] if words else "?"] += num(r, "Instructions Executed")
        reasons = collections.Counter({k[6:]: sum(num(r, k) for r in seg) for k in shdr
                                       if k.startswith("stall_") and "Not Issued" not in k})
        lines.append(f"   {s:5d}-{s + len(seg):5d}  {100 * smp / tot_s:5.1f}%  {100 * ins / tot_i:5.1f}%   " +
                     " ".join(f"{k}:{v // 1000}" for k, v in ops.most_common(6)) + " | " +
                     " ".join(f"{k}:{v}" for k, v in reasons.most_common(4)))
    with open(dst, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:24]))
    if "--traffic" in sys.argv:
        at = sys.argv.index("--traffic")
        kernel, workload = sys.argv[at + 1], sys.argv[at + 2]
        path = os.path.join(ROOT, "profiles", "r2_kernel_traffic.json")
        table = json.load(open(path)) if os.path.exists(path) else {"kernels": {}}
        rd = to_bytes(vals[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(vals[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        table["kernels"].setdefault(kernel, {})[workload] = {
            "dram_bytes_read": rd, "dram_bytes_write": wr,
            "ncu_duration_us": float(vals[col["gpu__time_duration.sum"]]),
            "kernel_name": vals[col["Kernel Name"]],
            "source": f"profiles/{os.path.basename(dst)} (ncu --set full --clock-control none, one launch)"}
        with open(path, "w") as f:
            json.dump(table, f, indent=1, sort_keys=True)
        print("traffic:", rd + wr, "bytes per launch ->", path)


if __name__ == "__main__":
    main()
