#!/bin/bash
# round 2, GPU call 11 (1 GPU): A/B of the steady-state loop unrolled by the ring period (libswalbe_b200_u8.so, -DSW_UNROLL8)
# against the shipped library on the same box; CTA-width sweep of the complete C4 kernel
mkdir -p gpurun_out; O=gpurun_out
ALT=$PWD/swalbe.jl_b200/libswalbe_b200_u8.so
( SWALBE_B200_SO=$ALT timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_baseline_sizes.py -m gpu -x -q ) > $O/pytest_u8.log 2>&1; tail -3 $O/pytest_u8.log
{
for rep in 1 2; do
for so in main alt; do
  if [ $so = alt ]; then export SWALBE_B200_SO=$ALT; else unset SWALBE_B200_SO; fi
  python tools/rate_probe.py --steps 200 --label film8192_$so
  python tools/rate_probe.py --lazy --steps 200 --label lazy8192_$so
  python tools/rate_probe.py --thermal --steps 200 --label thermal8192_$so
  python tools/rate_probe.py --L 4096 --steps 400 --label film4096_$so
  python tools/rate_probe.py --L 2048 --steps 800 --label film2048_$so
  SWALBE_GRAPH=0 python tools/rate_probe.py --L 1024 --steps 2000 --label film1024_$so
  SWALBE_GRAPH=0 python tools/rate_probe.py --L 1024 --n 3 --m 2 --steps 2000 --lazy --label lazy1024_$so
done; done
unset SWALBE_B200_SO
for nt in 0 128 160 192 224; do SWALBE_NT=$nt python tools/rate_probe.py --thermal --theta-field --n 3 --m 2 --label c4like_nt$nt; done
} > $O/probes11.txt 2>&1
cat $O/probes11.txt | sed 's/theta_field=\(True\|False\) //; s/(144 B.*//'
