#!/bin/bash
# round 2, GPU call 16 (1 GPU): in-loop mass log, time_loop as one library call; whole-job lines at 20 / 200 steps
mkdir -p gpurun_out; O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_host_loop.py tests/test_gpu_step.py tests/test_gpu_baseline_sizes.py -m gpu -x -q ) > $O/pytest_call16.log 2>&1; tail -5 $O/pytest_call16.log
for rep in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_20_call16_$rep.json 2> $O/bench_20_call16.err
done
timeout 600 python bench.py --no-cpu-baseline --no-parity > $O/bench_200_call16.json 2>> $O/bench_20_call16.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*_call16*.json")):
    l = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, l["value"], l["ms_per_step"], l["e2e"]["value"], round(l["e2e"]["value"] / l["value"], 3), l["clocks"]["sm_mhz"], l["clocks"]["reasons"])
PY
SWALBE_HOST_TRACE=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity 2> $O/trace16.txt > /dev/null; tail -75 $O/trace16.txt
