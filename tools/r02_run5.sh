#!/bin/bash
# round 2, GPU call 5 (2 GPUs): 2-rank NCCL parity tests, bench --gpus 2 (parity leg vs 1 GPU), reference arm under torchrun
mkdir -p gpurun_out; O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( time timeout 900 python -m pytest tests/test_gpu_dist.py -x -q ) > $O/pytest_dist2.log 2>&1; tail -4 $O/pytest_dist2.log
$TR --master-port 29701 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2_20.json 2> $O/bench_n2_20.err; tail -c 700 $O/bench_n2_20.json
$TR --master-port 29702 bench.py --gpus 2 --steps 200 --warmup 10 > $O/bench_n2_200.json 2> $O/bench_n2_200.err
$TR --master-port 29703 bench.py --gpus 2 --workload thermal_moving --steps 200 --warmup 10 > $O/bench_c4_n2.json 2> $O/bench_c4_n2.err
$TR --master-port 29704 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; cat $O/bench_ref_n2.json | cut -c1-400
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n1_samebox.json 2> $O/bench_n1_samebox.err
ls -la $O | tail -8
