#!/bin/bash
# round 2, GPU call 29 (1 GPU): the other workloads through the final tree (e2e = host loop): C4 complete, thermal, tau = 0.9
mkdir -p gpurun_out; O=gpurun_out
timeout 300 python bench.py --workload thermal_moving --steps 200 --warmup 5 --no-cpu-baseline --no-parity > $O/bench_final_c4.json 2> $O/bench_final_c4.err
timeout 300 python bench.py --workload thermal --steps 200 --warmup 5 --no-cpu-baseline --no-parity > $O/bench_final_thermal.json 2> $O/bench_final_thermal.err
timeout 300 python bench.py --tau 0.9 --steps 100 --warmup 5 --no-cpu-baseline --no-parity > $O/bench_final_tau09.json 2> $O/bench_final_tau09.err
python - <<'PY'
import json
for f in ("bench_final_c4", "bench_final_thermal", "bench_final_tau09"):
    try:
        l = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, l["value"], l.get("ms_per_step"), (l.get("e2e") or {}).get("value"), l["roofline"]["frac"], l["clocks"]["sm_mhz"], l["clocks"]["reasons"])
    except Exception as e:
        print(f, "unreadable:", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
