#!/bin/bash
# round 2, GPU call 20 (N GPUs): peer-memory halos at N > 2 (distinct up / down neighbours): bench line with the bitwise leg
N=${1:-4}
mkdir -p gpurun_out; O=gpurun_out
TR="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
SWALBE_DIST_P2P=1 $TR 29801 bench.py --gpus $N --steps 20 --warmup 3 > $O/bench_n${N}_20_p2p1.json 2> $O/bench_n${N}_call20.err
SWALBE_DIST_P2P=1 $TR 29802 bench.py --gpus $N --steps 200 --warmup 5 --no-e2e --no-parity > $O/bench_n${N}_200_p2p1.json 2>> $O/bench_n${N}_call20.err
SWALBE_DIST_P2P=0 $TR 29803 bench.py --gpus $N --steps 200 --warmup 5 --no-e2e --no-parity > $O/bench_n${N}_200_p2p0.json 2>> $O/bench_n${N}_call20.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_n${N}_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, l["value"], l["ms_per_step"], l.get("halo_transport"), l.get("dist_loop_ms_per_step_rank0"), l.get("parity_vs_1gpu"), (l.get("e2e") or {}).get("value"), l["clocks"]["sm_mhz"])
PY
tail -3 $O/bench_n${N}_call20.err
