"""Throughput with a contact-angle FIELD (moving-wettability scripts), slip variants and inclination."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import swalbe_b200 as sw
import bench

L = 8192
sysc = sw.SysConst(Lx=L, Ly=L, param=sw.Taumucs(n=3, m=2, hmin=0.07))
st = sw.Sys(sysc, "GPU")
i = torch.arange(L, device="cuda", dtype=torch.float64)
theta = sw.Field(L, L)
theta.t.copy_(1 / 9 + 1 / 36 * torch.sin(4 * np.pi * i / L)[None, :] * torch.sin(4 * np.pi * i / L)[:, None])
ct = sw.Field(L, L); ct.t.copy_(torch.cos(np.pi * theta.t))
def timeit(tag, **kw):
    st.height.set(bench.initial_height(L)); st.velx.t.zero_(); st.vely.t.zero_()
    sw.fused_steps(st, sysc, 10, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sw.fused_steps(st, sysc, 100, **kw); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 100
    print(f"{tag:34s} {L*L/ms/1e3:9.1f} MLUPS  {ms:.3f} ms/step", flush=True)
timeit("scalar theta")
timeit("theta field", θ=theta)
timeit("theta field + ring_riv slip", θ=theta, slip_variant=2)
timeit("scalar + inclination", incl=([1e-5, 0.0], 1.0))
timeit("theta field, logs on (full kernel)", θ=theta, log_minmax=True)
