#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into the few numbers DESIGN.md / profiles/ quote.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [n_lattice_updates]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
nlu = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, d = rows[0], rows[1], rows[2]
g = lambda k: d[hdr.index(k)]  # noqa: E731
print("kernel:", g("Kernel Name"), "grid", g("Grid Size"), "block", g("Block Size"))
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.max", "smsp__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_cbu.sum",
        "sm__inst_executed_pipe_adu.sum", "sm__inst_executed_pipe_lsu.sum"]
for k in keys:
    if k in hdr:
        print(f"{k:80s} {g(k):>18s} {units[hdr.index(k)]}")
if nlu:
    def val(k):
        v, u = float(g(k)), units[hdr.index(k)]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}.get(u, 1)
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    t = float(g("gpu__time_duration.sum")) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}.get(units[hdr.index("gpu__time_duration.sum")], 1)
    print(f"DRAM bytes/LU: read {rd / nlu:.2f} write {wr / nlu:.2f} total {(rd + wr) / nlu:.2f};  DRAM GB/s {(rd + wr) / t / 1e9:.1f};"
          f"  MLUPS(under ncu) {nlu / t / 1e6:.0f};  warp-inst/LU*32 {float(g('smsp__inst_executed.sum')) * 32 / nlu:.1f}")
stalls = []
for k in hdr:
    if "smsp__average_warps_issue_stalled" in k and k.endswith("_per_issue_active.ratio"):
        try:
            stalls.append((float(g(k)), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        except ValueError:
            pass
print("stalls (warps per issue-active cycle):", ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:9]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]
ix = {k: i for i, k in enumerate(h2)}
data, seen = [], set()
for r in rows[2:]:
    if len(r) != len(h2) or not r[ix["# Samples"]].isdigit():
        continue
    if r[ix["Address"]] in seen:
        break
    seen.add(r[ix["Address"]])
    data.append(r)
tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
print(f"source page: {len(data)} SASS instructions, {tot} samples; top stall sites:")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 14]:
    s = int(r[ix["# Samples"]])
    reasons = {k: int(r[ix[k]]) for k in h2 if k.startswith("stall_") and "Not Issued" not in k and r[ix[k]].isdigit() and int(r[ix[k]]) > 0}
    main = sorted(reasons.items(), key=lambda kv: -kv[1])[:2]
    print(f"  {100 * s / tot:5.2f}%  {r[ix['Source']].strip()[:60]:60s} {main}")
