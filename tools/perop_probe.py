"""Throughput of a user-written loop that calls the seven operators one by one (src/simulate.jl:15-22 as scripts do),
each through its own C-ABI kernel, next to the fused loop."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import swalbe_b200 as sw
import bench

for L in (1024, 4096):
    sysc = sw.SysConst(Lx=L, Ly=L, param=sw.Taumucs())
    st = sw.Sys(sysc, "GPU")
    st.height.set(bench.initial_height(L))
    def step():
        sw.filmpressure(st, sysc); sw.hgradp(st); sw.slippage(st, sysc); sw.update(st)
        sw.equilibrium(st, sysc); sw.BGKandStream(st, sysc); sw.moments(st)
    for _ in range(5): step()
    n = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    sw.fused_steps(st, sysc, 5)
    e0.record(); sw.fused_steps(st, sysc, n); e1.record(); torch.cuda.synchronize()
    msf = e0.elapsed_time(e1) / n
    print(f"L={L}: operator-by-operator {L*L/ms/1e3:8.1f} MLUPS ({ms:.3f} ms/step, 8 launches)   fused {L*L/msf/1e3:8.1f} MLUPS ({msf:.3f} ms/step)")
