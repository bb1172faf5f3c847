#!/bin/bash
# round 2, GPU call 27 (1 GPU): final-tree validation -- every GPU test, smoke, the bench lines, ncu launch list of the bench
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_final.log 2>&1; tail -4 $O/pytest_final.log
( timeout 300 python __graft_entry__.py smoke ) > $O/smoke_final.log 2>&1; tail -2 $O/smoke_final.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_final_20b.json 2> $O/bench_final_20b.err
timeout 600 python bench.py > $O/bench_final_defaultb.json 2> $O/bench_final_defaultb.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_final_referenceb.json 2> /dev/null
python - <<'PY'
import json
for f in ("bench_final_20b", "bench_final_defaultb", "bench_final_referenceb"):
    l = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, l["value"], l.get("ms_per_step"), (l.get("e2e") or {}).get("value"), (l.get("roofline") or {}).get("frac"), (l.get("roofline") or {}).get("dram_frac"),
          l.get("parity_vs_1gpu"), (l.get("clocks") or {}).get("sm_mhz"), (l.get("clocks") or {}).get("reasons"), l.get("gpu_launches"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_final.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-parity > $O/bench_under_ncu.log 2>&1
python tools/launch_list.py $O/launches_final.csv > $O/launch_list_final.txt 2>&1 || true; head -40 $O/launch_list_final.txt
