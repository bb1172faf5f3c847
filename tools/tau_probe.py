"""Throughput of the general-tau path (old populations are read: 192 B/LU) next to tau = 1."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import swalbe_b200 as sw
import bench

L = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
for tau, g in ((1.0, 0.0), (1.0, -0.001), (0.9, 0.0)):
    sysc = sw.SysConst(Lx=L, Ly=L, param=sw.Taumucs(τ=tau, g=g))
    st = sw.Sys(sysc, "GPU")
    st.height.set(bench.initial_height(L))
    sw.equilibrium(st, sysc)
    st.ftemp.t.copy_(st.feq.t)
    sw.fused_steps(st, sysc, 10)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sw.fused_steps(st, sysc, 100); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 100
    bytes_lu = 120 if tau == 1.0 else 192
    print(f"tau={tau} g={g}: {L*L/ms/1e3:9.1f} MLUPS  {ms:.3f} ms/step  moved {bytes_lu} B/LU -> {L*L*bytes_lu/ms/1e6:.0f} GB/s", flush=True)
    del st; torch.cuda.empty_cache()
