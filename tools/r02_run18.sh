#!/bin/bash
# round 2, GPU call 18 (1 GPU): two compute streams inside the sweeps of the host loop, A/B against one stream
mkdir -p gpurun_out; O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_host_loop.py -m gpu -x -q ) > $O/pytest_call18.log 2>&1; tail -5 $O/pytest_call18.log
for rep in 1 2; do
for ns in 2 1; do
SWALBE_HOST_STREAMS=$ns timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity > $O/bench_20_call18_s${ns}_$rep.json 2>> $O/bench_20_call18.err
done; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_20_call18*.json")):
    l = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, l["value"], l["ms_per_step"], l["e2e"]["value"], round(l["e2e"]["value"] / l["value"], 3), l["clocks"]["sm_mhz"], l["clocks"]["reasons"])
PY
SWALBE_HOST_TRACE=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity 2> $O/trace18.txt > /dev/null; tail -48 $O/trace18.txt
for b in 512 2048; do SWALBE_BAND_ROWS=$b timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('band $b', l['value'], l['e2e']['value'])"; done
