#!/bin/bash
# round 2, GPU call 8 (1 GPU): compute-sanitizer over every kernel flavour, slab-runtime vs fused-loop probe, 1-D latency,
# ncu traffic of the default FM kernel
mkdir -p gpurun_out; O=gpurun_out
( time timeout 900 compute-sanitizer --tool memcheck python tools/sanitize.py ) > $O/sanitize_memcheck.log 2>&1; tail -4 $O/sanitize_memcheck.log
( time timeout 1200 compute-sanitizer --tool racecheck python tools/sanitize.py ) > $O/sanitize_racecheck.log 2>&1; tail -4 $O/sanitize_racecheck.log
{
SWALBE_DEBUG=1 python tools/dist_probe.py thermal_moving 2>&1 | grep -v "cluster kernel"
SWALBE_DEBUG=1 python tools/dist_probe.py thermal 2>&1
python tools/dist_probe.py film
python - <<'PY'
import time, sys
sys.path.insert(0, ".")
import swalbe_b200 as sw, torch, numpy as np
for L in (256, 1024, 4096):
    s1 = sw.SysConst_1D(L=L, param=sw.Taumucs(Tmax=10000, tdump=5000))
    for rep in range(3):
        st = sw.Sys(s1); st.height.set(1.0 + 0.1 * np.random.default_rng(1).standard_normal(L))
        torch.cuda.synchronize(); t0 = time.perf_counter(); sw.time_loop(s1, st); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"1-D time_loop L={L} Tmax=10000: {1e3 * dt:.2f} ms wall = {1e2 * dt:.3f} us/step", flush=True)
PY
} > $O/probes8.txt 2>&1
cat $O/probes8.txt
ncu --set full --clock-control none --import-source on -k regex:k_fused_step -s 12 -c 1 -o $O/r02_fm_default \
    python tools/rate_probe.py --tau 0.9 --steps 10 > $O/ncu_fm_default.log 2>&1
ls -la $O | tail -5
