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
st.height.set(bench.initial_height(L))
sw.fused_steps(st, sysc, 10, θ=theta)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); sw.fused_steps(st, sysc, 60, θ=theta); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 60
print(f"NT={os.environ.get('SWALBE_NT','auto')} theta field {L*L/ms/1e3:9.1f} MLUPS  {ms:.3f} ms/step", flush=True)
