#!/bin/bash
# round 2, GPU call 31 (1 GPU): host-plane layout checks against the callers (drivers test, a small bench line)
mkdir -p gpurun_out; O=gpurun_out
( timeout 120 python -m pytest tests/test_gpu_host_loop.py -m gpu -x -q -k "drivers or prints or 150-200-50-5" ) 2>&1 | tail -2
timeout 120 python bench.py --L 2048 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['value'], l['e2e']['value'], l['parity_vs_1gpu'])"
