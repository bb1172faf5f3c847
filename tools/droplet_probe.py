"""Where does the droplet workload lose time?  Times lazy (compute-bound) steps for variations of the initial state."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import swalbe_b200 as sw
import bench

L = 4096
def run(tag, h0, u_noise=0.0, n=3, m=2, hmin=0.07, steps=100, pre=0):
    sysc = sw.SysConst(Lx=L, Ly=L, param=sw.Taumucs(n=n, m=m, hmin=hmin))
    st = sw.Sys(sysc, "GPU")
    st.height.set(h0)
    if u_noise:
        st.velx.t.copy_(u_noise * torch.randn((L, L), device="cuda", dtype=torch.float64))
        st.vely.t.copy_(u_noise * torch.randn((L, L), device="cuda", dtype=torch.float64))
    if pre:
        sw.fused_steps(st, sysc, pre, lazy_populations=True)
    sw.fused_steps(st, sysc, 5, lazy_populations=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sw.fused_steps(st, sysc, steps, lazy_populations=True); e1.record(); torch.cuda.synchronize()
    z = float(((st.velx.t == 0) & (st.vely.t == 0)).double().mean())
    fin = bool(torch.isfinite(st.height.t).all())
    print(f"{tag:40s} {L*L*steps/e0.elapsed_time(e1)/1e3:9.1f} MLUPS  zero-u frac {z:.3f} finite {fin}", flush=True)
    del st; torch.cuda.empty_cache()

film = bench.initial_height(L, workload="film")
drop = bench.initial_height(L, workload="droplet")
run("film", film)
run("flat h=1 (u == 0 everywhere)", np.asfortranarray(np.ones((L, L))))
run("flat h=0.05", np.asfortranarray(np.full((L, L), 0.05)))
run("droplet", drop)
run("droplet + u noise 1e-12", drop, u_noise=1e-12)
run("droplet after 1500 steps", drop, pre=1500)
run("droplet cap only region test: h=50 flat", np.asfortranarray(np.full((L, L), 50.0)))
