#!/bin/bash
# round 2, GPU call 6 (1 GPU): full parity suite (1-D family, cluster kernel), small-lattice latencies, C5 single-GPU baselines
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu6.log 2>&1
tail -6 $O/pytest_gpu6.log
{
for L in 32 64 100; do
  for lg in nologs logs; do
    SWALBE_CLUSTER_MAX=100000 SWALBE_CLUSTER_MAX_LOGS=100000 python tools/small_probe.py $L 100 $lg
    SWALBE_CLUSTER=0 python tools/small_probe.py $L 100 $lg
  done
done
python tools/small_probe.py 100 1000 logs
python - <<'PY'
import time, sys
sys.path.insert(0, ".")
import swalbe_b200 as sw, torch, numpy as np
for rep in range(3):
    sysc = sw.SysConst(Lx=100, Ly=100, param=sw.Taumucs(g=-0.001, γ=0.0005, Tmax=1000))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h, diff = sw.run_rayleightaylor(sysc, "GPU", h0=1.0, ϵ=0.01, verbos=False)
    torch.cuda.synchronize(); print(f"run_rayleightaylor 100^2 Tmax=1000 (README example): {1e3 * (time.perf_counter() - t0):.2f} ms wall", flush=True)
# 1-D: L = 1024, 10000 steps (persistent single-CTA loop)
s1 = sw.SysConst_1D(L=1024, param=sw.Taumucs(Tmax=10000, tdump=5000))
for rep in range(3):
    st = sw.Sys(s1); st.height.set(1.0 + 0.1 * np.random.default_rng(1).standard_normal(1024))
    torch.cuda.synchronize(); t0 = time.perf_counter(); sw.time_loop(s1, st); torch.cuda.synchronize()
    print(f"1-D time_loop L=1024 Tmax=10000: {1e3 * (time.perf_counter() - t0):.2f} ms wall = {1e2 * (time.perf_counter() - t0):.3f} us/step", flush=True)
PY
} > $O/probes6.txt 2>&1
cat $O/probes6.txt
python bench.py --L 16384 --steps 100 --warmup 5 --no-cpu-baseline --no-parity > $O/bench_c5_16384_n1.json 2> $O/bench_c5_16384_n1.err; tail -c 400 $O/bench_c5_16384_n1.json
python bench.py --L 32768 --rows 4096 --steps 100 --warmup 5 --no-cpu-baseline --no-parity > $O/bench_c5_32768x4096_n1.json 2> $O/bench_c5_32768x4096_n1.err; tail -c 400 $O/bench_c5_32768x4096_n1.json
python bench.py --L 16384 --rows 2048 --steps 100 --warmup 5 --no-cpu-baseline --no-parity > $O/bench_c5_16384x2048_n1.json 2> $O/bench_c5_16384x2048_n1.err
python bench.py --L 4096 --workload spinodal --steps 200 --warmup 10 --no-cpu-baseline --no-parity > $O/bench_c3.json 2> $O/bench_c3.err
python bench.py --L 1024 --workload droplet --steps 2000 --warmup 10 --no-cpu-baseline --no-parity > $O/bench_c2.json 2> $O/bench_c2.err
ls -la $O | tail -6
