#!/bin/bash
# round 2, GPU call 9 (1 GPU): validation of the final tree -- parity suite, smoke, default bench lines, reference arm -- and
# the cluster kernel with incremental site indices
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu9.log 2>&1; tail -5 $O/pytest_gpu9.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke9.log 2>&1; tail -2 $O/smoke9.log
{
for L in 32 64 100 128; do
  for lg in nologs logs; do
    SWALBE_CLUSTER_MAX=100000 SWALBE_CLUSTER_MAX_LOGS=100000 python tools/small_probe.py $L 100 $lg
  done
done
SWALBE_CLUSTER_SIZE=8 SWALBE_CLUSTER_MAX=100000 python tools/small_probe.py 100 100 nologs
} > $O/probes9.txt 2>&1
cat $O/probes9.txt
python bench.py --steps 20 --warmup 5 > $O/bench_final_20.json 2> $O/bench_final_20.err; tail -c 500 $O/bench_final_20.json
python bench.py > $O/bench_final_default.json 2> $O/bench_final_default.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_final_reference.json 2> $O/bench_final_reference.err; cut -c1-300 $O/bench_final_reference.json
ncu --set full --clock-control none --import-source on -k regex:k_cluster_steps -c 1 -o $O/r02_cluster_v2 \
    env SWALBE_CLUSTER_MAX=100000 python tools/small_probe.py 100 100 nologs > $O/ncu_cluster2.log 2>&1
ls -la $O | tail -5
