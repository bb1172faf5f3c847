#!/bin/bash
# round 2, GPU call 24 (2 GPUs): the slab runtime's host loop on two ranks (parity), bench N = 2 at the driver's 20 steps
mkdir -p gpurun_out; O=gpurun_out
( timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q ) > $O/pytest_call24.log 2>&1; tail -6 $O/pytest_call24.log
TR="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
$TR 29821 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_n2_20_call24.json 2> $O/bench_n2_call24.err
SWALBE_HOST_STREAM=0 $TR 29822 bench.py --gpus 2 --steps 20 --warmup 3 --no-parity > $O/bench_n2_20_call24_plain.json 2>> $O/bench_n2_call24.err
$TR 29823 bench.py --gpus 2 --steps 20 --warmup 3 --workload thermal_moving > $O/bench_n2_20_call24_c4.json 2>> $O/bench_n2_call24.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_n2_20_call24*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, l["value"], l["ms_per_step"], l.get("halo_transport"), l.get("parity_vs_1gpu"), (l.get("e2e") or {}).get("value"), l["clocks"]["sm_mhz"])
PY
tail -3 $O/bench_n2_call24.err
