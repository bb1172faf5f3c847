#!/bin/bash
# round 2, GPU call 4: persistent cluster kernel (parity + latency), README example wall time, default bench + ncu launch list
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu4.log 2>&1
tail -5 $O/pytest_gpu4.log
{
for L in 32 64 100 128; do
  for lg in nologs logs; do
    python tools/small_probe.py $L 100 $lg
    SWALBE_CLUSTER_SIZE=8 python tools/small_probe.py $L 100 $lg
    SWALBE_CLUSTER=0 python tools/small_probe.py $L 100 $lg
  done
done
python tools/small_probe.py 100 1000 logs
SWALBE_CLUSTER=0 python tools/small_probe.py 100 1000 logs
python tools/small_probe.py 100 10 nologs
SWALBE_CLUSTER=0 python tools/small_probe.py 100 10 nologs
SWALBE_DEBUG=1 python tools/small_probe.py 100 100 nologs 2>&1 | grep -m3 "cluster kernel"
python - <<'PY'
import time, sys
sys.path.insert(0, ".")
import swalbe_b200 as sw, torch
for rep in range(3):
    sysc = sw.SysConst(Lx=100, Ly=100, param=sw.Taumucs(g=-0.001, γ=0.0005, Tmax=1000))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h, diff = sw.run_rayleightaylor(sysc, "GPU", h0=1.0, ϵ=0.01, verbos=False)
    torch.cuda.synchronize(); print(f"run_rayleightaylor 100^2 Tmax=1000 (README example): {1e3 * (time.perf_counter() - t0):.2f} ms wall, {len(diff)} log entries", flush=True)
PY
} > $O/probes4.txt 2>&1
cat $O/probes4.txt
python bench.py --steps 20 --warmup 5 > $O/bench_default_20_b.json 2> $O/bench_default_20_b.err; tail -c 300 $O/bench_default_20_b.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_cluster_steps -c 1 -o $O/r02_cluster \
    python tools/small_probe.py 100 100 logs > $O/ncu_cluster.log 2>&1
ls -la $O | tail -6
