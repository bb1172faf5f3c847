#!/bin/bash
# round 2, GPU call 15 (1 GPU): asynchronous mass read-back; whole-job lines at the driver's 20 steps and at 200
mkdir -p gpurun_out; O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_host_loop.py tests/test_gpu_step.py -m gpu -x -q ) > $O/pytest_call15.log 2>&1; tail -3 $O/pytest_call15.log
for rep in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_20_call15_$rep.json 2> $O/bench_20_call15.err
SWALBE_HOST_STREAM=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity > $O/bench_20_call15_plain_$rep.json 2>> $O/bench_20_call15.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_20_call15*.json")):
    l = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, l["value"], l["ms_per_step"], l["e2e"]["value"], round(l["e2e"]["value"] / l["value"], 3), l["clocks"]["sm_mhz"], l["clocks"]["reasons"])
PY
SWALBE_HOST_TRACE=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity 2> $O/trace15.txt > /dev/null; tail -120 $O/trace15.txt
