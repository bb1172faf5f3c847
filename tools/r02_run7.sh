#!/bin/bash
# round 2, GPU call 7 (8 GPUs): BASELINE config 5 -- strong 16384^2, weak 8192^2 / 32768x4096 per GPU -- and C4 at
# 8 GPUs, each line with clocks, the N-rank-vs-1-GPU bitwise leg and the rank-0 loop time (overlap evidence).
# Phase A: N = 8 runs one after the other; phases B/C: N = 4, 2, 1 runs side by side on disjoint GPUs of the same box.
mkdir -p gpurun_out; O=gpurun_out
run() {  # run <gpus list> <nproc> <port> <outfile> <bench args...>
  local devs=$1 n=$2 port=$3 out=$4; shift 4
  if [ "$n" = 1 ]; then CUDA_VISIBLE_DEVICES=$devs python bench.py --gpus 1 "$@" > $O/$out.json 2> $O/$out.err
  else CUDA_VISIBLE_DEVICES=$devs python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
         bench.py --gpus $n "$@" > $O/$out.json 2> $O/$out.err; fi
}
ALL=0,1,2,3,4,5,6,7
# ---- phase A: 8 GPUs
run $ALL 8 29801 c5_weak8192_n8 --steps 200 --warmup 10 --no-cpu-baseline
run $ALL 8 29802 c5_strong16384_n8 --L 16384 --scaling strong --steps 200 --warmup 10 --no-cpu-baseline --no-parity --no-e2e
run $ALL 8 29803 c5_weak32768x4096_n8 --L 32768 --rows 4096 --steps 100 --warmup 5 --no-cpu-baseline --no-parity --no-e2e
run $ALL 8 29805 c4_thermal_moving_n8 --workload thermal_moving --steps 200 --warmup 10 --no-cpu-baseline
run $ALL 8 29806 c4_thermal_n8 --workload thermal --steps 200 --warmup 10 --no-cpu-baseline --no-parity --no-e2e
# ---- phase B: strong 16384^2 at N = 4, 2 and weak 8192^2 at N = 1 (x2: two GPUs) side by side
run 0,1,2,3 4 29811 c5_strong16384_n4 --L 16384 --scaling strong --steps 200 --warmup 10 --no-cpu-baseline --no-parity --no-e2e &
run 4,5 2 29812 c5_strong16384_n2 --L 16384 --scaling strong --steps 200 --warmup 10 --no-cpu-baseline --no-parity --no-e2e &
run 6 1 0 c5_weak8192_n1 --steps 200 --warmup 10 --no-cpu-baseline --no-parity --no-e2e &
run 7 1 0 c5_strong16384_n1 --L 16384 --steps 100 --warmup 5 --no-cpu-baseline --no-parity --no-e2e &
wait
# ---- phase C: weak 32768x4096 per GPU at N = 4, 2, 1 and weak 8192^2 at N = 4 ... side by side
run 0,1,2,3 4 29821 c5_weak32768x4096_n4 --L 32768 --rows 4096 --steps 100 --warmup 5 --no-cpu-baseline --no-parity --no-e2e &
run 4,5 2 29822 c5_weak32768x4096_n2 --L 32768 --rows 4096 --steps 100 --warmup 5 --no-cpu-baseline --no-parity --no-e2e &
run 6 1 0 c5_weak32768x4096_n1 --L 32768 --rows 4096 --steps 100 --warmup 5 --no-cpu-baseline --no-parity --no-e2e &
run 7 1 0 c4_thermal_moving_n1 --workload thermal_moving --steps 200 --warmup 10 --no-cpu-baseline --no-parity --no-e2e &
wait
# ---- phase D: weak 8192^2 at N = 4 and N = 2 (the driver's SCALE config) side by side
run 0,1,2,3 4 29831 c5_weak8192_n4 --steps 200 --warmup 10 --no-cpu-baseline --no-e2e &
run 4,5 2 29832 c5_weak8192_n2 --steps 200 --warmup 10 --no-cpu-baseline --no-e2e &
wait
python - <<'PY'
import glob, json, os
for f in sorted(glob.glob("gpurun_out/c[45]_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), "N", d["n_gpus"], d["value"], "MLUPS", d["ms_per_step"], "ms/step frac", d["roofline"]["frac"],
              "clk", d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"), "loop/step", d.get("dist_loop_ms_per_step_rank0"),
              "parity", d.get("parity_vs_1gpu"), "e2e", (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(os.path.basename(f), "FAILED", e, open(f.replace(".json", ".err")).read()[-600:])
PY
