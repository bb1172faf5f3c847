#!/bin/bash
# round 2, GPU call 28 (1 GPU): mass-log pieces on the band stage's own stream (race fix) -- repeated, then every GPU test
mkdir -p gpurun_out; O=gpurun_out
( for i in 1 2 3 4 5 6; do timeout 300 python -m pytest tests/test_gpu_host_loop.py -m gpu -x -q -k mass 2>&1 | tail -1; done ) > $O/pytest_mass_repeat.log 2>&1; cat $O/pytest_mass_repeat.log
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_final.log 2>&1; tail -4 $O/pytest_final.log | head -2
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_final_20c.json 2> $O/bench_final_20c.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/bench_final_20c.json").read().strip().splitlines()[-1])
print(l["value"], l.get("ms_per_step"), (l.get("e2e") or {}).get("value"), l["roofline"]["frac"], l["roofline"]["dram_frac"], l.get("parity_vs_1gpu"), l["clocks"])
PY
