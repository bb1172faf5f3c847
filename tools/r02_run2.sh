#!/bin/bash
# round 2, GPU call 2: thermal noise v3 (statistics tests + rates + ncu), FM residency hints A/B, small theta-field lattices
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu2.log 2>&1
tail -5 $O/pytest_gpu2.log
{
for nt in 0 128 160 192 224; do SWALBE_NT=$nt python tools/rate_probe.py --thermal --label thermal_nt$nt; done
SWALBE_NT=224 python tools/rate_probe.py --thermal --n 3 --m 2 --label thermal32_nt224
python tools/rate_probe.py --thermal --theta-field --n 3 --m 2 --label c4_like
for h in 0 1 2 3; do SWALBE_FM_HINTS=$h python tools/rate_probe.py --tau 0.9 --label fm_hints$h; done
SWALBE_NT=128 python tools/rate_probe.py --tau 0.9 --label fm_nt128_hints3
SWALBE_NT=224 python tools/rate_probe.py --tau 0.9 --label fm_nt224_hints3
for L in 128 256 384 512; do for tt in 0 1; do SWALBE_DEBUG=1 SWALBE_TILE_THETA=$tt python tools/rate_probe.py --L $L --n 3 --m 2 --theta-field --steps 980 --calls 10 --label tile_theta$tt; done; done
for L in 100 256 512; do python tools/rate_probe.py --L $L --steps 2000 --calls 10 --label small_film; done
python tools/rate_probe.py --L 1024 --steps 1000 --calls 5 --label mid_film
python tools/rate_probe.py --L 2048 --steps 400 --label mid_film
python tools/rate_probe.py --L 4096 --steps 200 --label mid_film
} > $O/probes2.txt 2>&1
cat $O/probes2.txt | grep -v "^\[swalbe\]" 
grep "^\[swalbe\]" $O/probes2.txt | sort | uniq -c | sort -rn | head -30
python bench.py --workload thermal --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_thermal.json 2> $O/bench_thermal.err
python bench.py --workload thermal_moving --steps 200 --warmup 10 --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err
python bench.py --tau 0.9 --steps 100 --warmup 10 --no-cpu-baseline --no-parity > $O/bench_tau09_b.json 2> $O/bench_tau09_b.err
ncu --set full --clock-control none --import-source on -k regex:k_fused_step -s 12 -c 1 -o $O/r02_thermal_v3 \
    python tools/rate_probe.py --thermal --steps 10 > $O/ncu_thermal.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused_step -s 12 -c 1 -o $O/r02_fm_v2 \
    python tools/rate_probe.py --tau 0.9 --steps 10 > $O/ncu_fm2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r02_launches_theta256.csv \
    python tools/rate_probe.py --L 256 --n 3 --m 2 --theta-field --steps 20 --calls 1 > $O/ncu_t256.log 2>&1
ls -la $O | tail -20
