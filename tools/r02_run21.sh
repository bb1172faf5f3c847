#!/bin/bash
# round 2, GPU call 21 (8 GPUs): peer-memory halos at 8 ranks: the driver's 20-step line with the bitwise leg, and 200 steps
mkdir -p gpurun_out; O=gpurun_out
TR="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
$TR 29811 bench.py --gpus 8 --steps 20 --warmup 3 > $O/bench_n8_20_p2p1.json 2> $O/bench_n8_call21.err
$TR 29812 bench.py --gpus 8 --steps 200 --warmup 5 --no-e2e --no-parity > $O/bench_n8_200_p2p1.json 2>> $O/bench_n8_call21.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_n8_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, l["value"], l["ms_per_step"], l.get("halo_transport"), l.get("dist_loop_ms_per_step_rank0"), l.get("parity_vs_1gpu"), (l.get("e2e") or {}).get("value"), l["clocks"]["sm_mhz"])
PY
tail -3 $O/bench_n8_call21.err
