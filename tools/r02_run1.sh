#!/bin/bash
# round 2, GPU call 1: parity suite, bench lines, general-tau FM sweep, ncu of the FM kernel, 512^2 theta-field tile A/B
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
tail -5 $O/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/bench_default_20.json 2> $O/bench_default_20.err; tail -c 600 $O/bench_default_20.json
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $O/bench_default_200.json 2>> $O/bench_default_20.err
python bench.py --steps 100 --warmup 10 --tau 0.9 --no-cpu-baseline --no-parity > $O/bench_tau09.json 2> $O/bench_tau09.err
{
for pf in 0 2 3 5 8; do SWALBE_FM_PREFETCH=$pf python tools/rate_probe.py --tau 0.9 --label fm_pf$pf; done
SWALBE_FM=0 python tools/rate_probe.py --tau 0.9 --label fm_off
for nt in 128 160 192 224 256; do SWALBE_NT=$nt python tools/rate_probe.py --tau 0.9 --label fm_nt$nt; done
for r in 64 256 512; do SWALBE_RMAX=$r python tools/rate_probe.py --tau 0.9 --label fm_rmax$r; done
python tools/rate_probe.py --tau 0.9 --n 3 --m 2 --label fm_32
python tools/rate_probe.py --tau 0.9 --L 4096 --label fm_4096
python tools/rate_probe.py --tau 0.9 --L 2048 --steps 400 --label fm_2048
# 512^2 moving-wettability size: theta-field through the tile kernel or the marching kernel (5 calls of 98 steps)
for tt in 0 1; do SWALBE_TILE_THETA=$tt python tools/rate_probe.py --L 512 --n 3 --m 2 --theta-field --steps 980 --calls 10 --label tile_theta$tt; done
for tt in 0 1; do SWALBE_TILE_THETA=$tt python tools/rate_probe.py --L 256 --n 3 --m 2 --theta-field --steps 980 --calls 10 --label tile_theta$tt; done
python tools/rate_probe.py --label film_8192
python tools/rate_probe.py --lazy --label film_8192_lazy
python tools/rate_probe.py --thermal --label thermal_8192
} > $O/probes.txt 2>&1
cat $O/probes.txt
# ncu: launch list of the tau = 0.9 bench + full capture of the FM kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02_launches_tau09.csv \
    python bench.py --steps 6 --warmup 3 --tau 0.9 --no-cpu-baseline --no-e2e --no-parity > $O/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused_step -s 12 -c 1 -o $O/r02_fm_tau09 \
    python tools/rate_probe.py --tau 0.9 --steps 10 > $O/ncu_fm.log 2>&1
ls -la $O
