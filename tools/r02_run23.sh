#!/bin/bash
# round 2, GPU call 23 (1 GPU): staged from-moments kernels (FMS) against FM at tau = 0.9; the slab host loop on one rank
mkdir -p gpurun_out; O=gpurun_out
( timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_step.py tests/test_gpu_baseline_sizes.py -m gpu -x -q -k "single_rank or tau or general or moments or nccl_path or special" ) > $O/pytest_call23.log 2>&1; tail -6 $O/pytest_call23.log
{
for rep in 1 2; do
for st in 1 0; do
  SWALBE_FM_STAGE=$st python tools/rate_probe.py --tau 0.9 --steps 100 --label fm_stage$st
done; done
for nt in 128 160 192 224; do SWALBE_FM_STAGE=1 SWALBE_NT=$nt python tools/rate_probe.py --tau 0.9 --steps 100 --label fms_nt$nt; done
SWALBE_FM_STAGE=1 python tools/rate_probe.py --tau 0.9 --steps 100 --L 4096 --label fms_4096
SWALBE_FM_STAGE=0 python tools/rate_probe.py --tau 0.9 --steps 100 --L 4096 --label fm_4096
SWALBE_FM_STAGE=1 python tools/rate_probe.py --tau 0.9 --n 3 --m 2 --steps 100 --label fms_32
SWALBE_FM_STAGE=0 python tools/rate_probe.py --tau 0.9 --n 3 --m 2 --steps 100 --label fm_32
} > $O/probes23.txt 2>&1
sed 's/theta_field=\(True\|False\) //; s/(144 B.*//' $O/probes23.txt
