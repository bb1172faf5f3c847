#!/bin/bash
# round 2, GPU call 3: neighbour-sync flavour A/B (film, thermal; large and mid lattices), FM prefetch/hint combinations
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu3.log 2>&1
tail -5 $O/pytest_gpu3.log
{
for ns in 0 1; do SWALBE_NSYNC=$ns python tools/rate_probe.py --steps 200 --label film_ns$ns; done
for nt in 128 160 192 224 256; do SWALBE_NSYNC=1 SWALBE_NT=$nt python tools/rate_probe.py --label film_ns1_nt$nt; done
SWALBE_NSYNC=1 SWALBE_BULK=0 python tools/rate_probe.py --label film_ns1_ldgsts
SWALBE_NSYNC=0 SWALBE_BULK=0 python tools/rate_probe.py --label film_ns0_ldgsts
for ns in 0 1; do SWALBE_NSYNC=$ns python tools/rate_probe.py --lazy --label film_lazy_ns$ns; done
for ns in 0 1; do SWALBE_NSYNC=$ns python tools/rate_probe.py --thermal --label thermal_ns$ns; done
for L in 4096 2048 1024; do for ns in 0 1; do SWALBE_NSYNC=$ns python tools/rate_probe.py --L $L --steps 400 --calls 2 --label mid_ns$ns; done; done
for ns in 0 1; do SWALBE_NSYNC=$ns SWALBE_TILE_MAX=0 python tools/rate_probe.py --L 512 --steps 980 --calls 10 --label march512_ns$ns; done
SWALBE_FM_PREFETCH=3 SWALBE_FM_HINTS=5 python tools/rate_probe.py --tau 0.9 --label fm_pf3_h5
SWALBE_FM_PREFETCH=2 SWALBE_FM_HINTS=5 python tools/rate_probe.py --tau 0.9 --label fm_pf2_h5
SWALBE_FM_PREFETCH=4 SWALBE_FM_HINTS=7 python tools/rate_probe.py --tau 0.9 --label fm_pf4_h7
SWALBE_FM_PREFETCH=3 SWALBE_FM_HINTS=1 python tools/rate_probe.py --tau 0.9 --label fm_pf3_h1
SWALBE_FM_PREFETCH=0 SWALBE_FM_HINTS=1 python tools/rate_probe.py --tau 0.9 --label fm_pf0_h1
SWALBE_FM_PREFETCH=0 SWALBE_FM_HINTS=1 SWALBE_RMAX=128 python tools/rate_probe.py --tau 0.9 --label fm_pf0_h1_rmax128
} > $O/probes3.txt 2>&1
cat $O/probes3.txt
ncu --set full --clock-control none --import-source on -k regex:k_fused_step -s 12 -c 1 -o $O/r02_film_ns \
    env SWALBE_NSYNC=1 python tools/rate_probe.py --steps 10 > $O/ncu_ns.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused_step -s 12 -c 1 -o $O/r02_film_base \
    env SWALBE_NSYNC=0 python tools/rate_probe.py --steps 10 > $O/ncu_base.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused_step -s 12 -c 1 -o $O/r02_fm_v3 \
    env SWALBE_FM_PREFETCH=3 SWALBE_FM_HINTS=5 python tools/rate_probe.py --tau 0.9 --steps 10 > $O/ncu_fm3.log 2>&1
ls -la $O | tail -8
