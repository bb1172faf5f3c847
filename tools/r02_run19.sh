#!/bin/bash
# round 2, GPU call 19 (2 GPUs): peer-memory halo exchange (k_halo_push / k_halo_wait) against the NCCL exchange
mkdir -p gpurun_out; O=gpurun_out
( SWALBE_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q ) > $O/pytest_call19.log 2>&1; tail -5 $O/pytest_call19.log
TR="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
for rep in 1 2; do
for p2p in 1 0; do
SWALBE_DIST_P2P=$p2p $TR $((29700 + rep * 10 + p2p)) bench.py --gpus 2 --steps 200 --warmup 5 --no-e2e > $O/bench_n2_p2p${p2p}_$rep.json 2>> $O/bench_n2_call19.err
done; done
SWALBE_DIST_P2P=1 $TR 29790 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_n2_20_p2p1.json 2>> $O/bench_n2_call19.err
SWALBE_DIST_P2P=1 $TR 29791 bench.py --gpus 2 --steps 200 --warmup 5 --workload thermal_moving --no-e2e > $O/bench_n2_c4_p2p1.json 2>> $O/bench_n2_call19.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_n2_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, l["value"], l["ms_per_step"], l.get("halo_transport"), l.get("dist_loop_ms_per_step_rank0"), l.get("parity_vs_1gpu"), (l.get("e2e") or {}).get("value"), l["clocks"]["sm_mhz"])
PY
tail -5 $O/bench_n2_call19.err
