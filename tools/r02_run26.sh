#!/bin/bash
# round 2, GPU call 26 (4 GPUs): the driver's 20-step line with the slab host loop as e2e (distinct up / down neighbours)
mkdir -p gpurun_out; O=gpurun_out
TR="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port"
$TR 29831 bench.py --gpus 4 --steps 20 --warmup 3 > $O/bench_n4_20_call26.json 2> $O/bench_n4_call26.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/bench_n4_20_call26.json").read().strip().splitlines()[-1])
print(l["value"], l["ms_per_step"], l.get("halo_transport"), l.get("parity_vs_1gpu"), (l.get("e2e") or {}).get("value"), l["clocks"]["sm_mhz"])
PY
tail -3 $O/bench_n4_call26.err
