"""compute-sanitizer coverage of what the second half of round 2 added: the host loop (banded sweeps on two compute streams,
seam strips, in-loop mass log), the slab runtime's host loop on one rank, the expanded 1-D kinds.  Same conventions as
tools/sanitize.py (which covers everything older)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import swalbe_b200 as sw
from swalbe_b200.dist import DistSim

rng = np.random.default_rng(0)
# ---- later in round 2 ----------------------------------------------------------------------------------------------
# host loop: banded sweeps on two compute streams, seam strips, mass log (rows summed behind the launches), plain fallback
os.environ["SWALBE_HOST_MIN_SITES"], os.environ["SWALBE_BAND_ROWS"] = "1", "40"
for kw, kind, seed in ((dict(g=-0.001), "simple", None), (dict(kbt=1e-6), "thermal", 9)):
    sysc = sw.SysConst(Lx=192, Ly=170, param=sw.Taumucs(**kw))
    st = sw.Sys(sysc, "GPU", kind=kind)
    hin = torch.from_numpy(np.ascontiguousarray((np.abs(1.0 + 0.2 * rng.standard_normal((192, 170))) + 0.06).transpose())).pin_memory()
    hout = torch.empty_like(hin).pin_memory()
    mass = torch.full((8,), float("nan"), dtype=torch.float64).pin_memory()
    for n in (3, 17):
        sw.fused_steps(st, sysc, n, host_in=hin, host_out=hout, thermal_seed=seed, mass_log=(1, 3, mass), log_minmax=True)
    sw.fused_steps(st, sysc, 7, mass_log=(0, 2, mass))
    os.environ["SWALBE_HOST_STREAMS"] = "1"
    sw.fused_steps(st, sysc, 9, host_in=hin, host_out=hout, lazy_populations=kind == "simple", thermal_seed=seed)
    del os.environ["SWALBE_HOST_STREAMS"]
torch.cuda.synchronize()
# the slab runtime's host loop on one rank (self-neighbour): bands, boundary strips, exchanges
sysc = sw.SysConst(Lx=130, Ly=72, param=sw.Taumucs(g=-0.001))
sim = DistSim(sysc, 0, 1, None)
os.environ["SWALBE_BAND_ROWS"] = "16"
hin = torch.ones(72 * 130, dtype=torch.float64).pin_memory()
hout = torch.empty_like(hin).pin_memory()
sim.time_loop_host(9, host_in=hin, host_out=hout)
sim.time_loop(2)
sim.time_loop_host(4, host_out=hout)
torch.cuda.synchronize()
sim.close()
del os.environ["SWALBE_HOST_MIN_SITES"], os.environ["SWALBE_BAND_ROWS"]
# the expanded 1-D kinds: operators, fused gamma / Marangoni / inclination loop, bounce-back loop
L = 700
s1 = sw.SysConst_1D(L=L, param=sw.Taumucs(Tmax=6, tdump=3, kbt=1e-6))
sg = sw.Sys(s1, kind="gamma")
sg.basestate.height.set(np.abs(1.0 + 0.2 * rng.standard_normal(L)) + 0.06)
sg.γ.set(0.01 * (1.0 + 0.1 * rng.random(L)))
sw.gradgamma(sg); sw.gradgamma(sg, s1)
sw.filmpressure(sg, s1, γ=sg.γ)
sw.update(sg)
sw.one_d.fused_steps(sg, s1, 5, gamma_field=True, marangoni=True)
sw.one_d.fused_steps(sw.Sys(s1), s1, 4, incl=(1e-4, 1.0))
stt = sw.Sys(s1, kind="thermal")
sw.thermal(stt, s1, seed=1, step=2); sw.update(stt); sw.inclination(1e-4, stt)
rho = sw.Field(L).set(0.1 * rng.random(L))
sw.update_rho(rho, sw.Field(L), sg.basestate.height, None, sw.Field(L, 4))
sw.filmpressure(sw.Field(L), sg.basestate.height, None, rho, 0.01, 1 / 9, 3, 2, 0.07, 0.05, Gamma=0.3)
obs = np.zeros(L); obs[:4] = 1; obs[-4:] = 1
sb = sw.SysConstWithBound_1D(L=L, param=sw.Taumucs(Tmax=5, tdump=2, τ=0.9), obs=obs)
sw.obslist(sb)
sw.time_loop(sb, sw.Sys(sb, kind="gamma_bound"))
torch.cuda.synchronize()
print("sanitize run complete")
