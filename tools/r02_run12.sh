#!/bin/bash
# round 2, GPU call 12 (1 GPU): swalbe_time_loop_host -- parity tests, band / sweep-length sweep of the whole job, bench lines
mkdir -p gpurun_out; O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_host_loop.py -m gpu -x -q ) > $O/pytest_host_loop.log 2>&1; tail -5 $O/pytest_host_loop.log
( timeout 600 python tools/e2e_probe.py --steps 20; timeout 300 python tools/e2e_probe.py --steps 200 --bands 0,1024 --kmax 0 ) > $O/probes12.txt 2>&1; cat $O/probes12.txt
timeout 600 python bench.py --steps 20 --warmup 3 --check-e2e --no-cpu-baseline > $O/bench_e2e_check.json 2> $O/bench_e2e_check.err; tail -c 1500 $O/bench_e2e_check.json
timeout 600 python bench.py > $O/bench_default_call12.json 2> $O/bench_default_call12.err; python - <<'PY'
import json
l = json.loads(open("gpurun_out/bench_default_call12.json").read().strip().splitlines()[-1])
print({k: l[k] for k in ("value", "ms_per_step", "e2e", "clocks", "materialise_step_ms", "parity_vs_1gpu") if k in l})
PY
