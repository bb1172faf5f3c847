#!/bin/bash
# round 2, GPU call 10 (1 GPU): ncu captures of the kernels the review found missing from profiles/: tile kernel, edge strips
mkdir -p gpurun_out; O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_tile_step -s 20 -c 1 -o $O/r02_tile \
    env SWALBE_GRAPH=0 python tools/small_probe.py 256 100 nologs > $O/ncu_tile.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tile_step -s 20 -c 1 -o $O/r02_tile_theta \
    env SWALBE_GRAPH=0 python tools/rate_probe.py --L 512 --n 3 --m 2 --theta-field --steps 30 > $O/ncu_tile_theta.log 2>&1
# slab runtime, one rank: per step two 3-row edge-strip launches and the interior launch
ncu --set full --clock-control none --import-source on -k regex:k_fused_step -s 30 -c 3 -o $O/r02_edge_strips \
    python tools/dist_probe.py film > $O/ncu_edges.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_fused_step -s 30 -c 30 --csv --log-file $O/r02_launches_dist.csv \
    python tools/dist_probe.py film > $O/ncu_dist_list.log 2>&1
ls -la $O | tail -6
