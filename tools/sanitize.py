"""Small fused-loop runs for compute-sanitizer (memcheck / racecheck / initcheck): all kernel flavours, wrap + ghost rows."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

import swalbe_b200 as sw
from swalbe_b200.dist import DistSim

rng = np.random.default_rng(0)
for (Lx, Ly, kw) in [(25, 26, {}), (300, 40, dict(n=3, m=2, hmin=0.07)), (130, 33, dict(τ=0.8)), (64, 20, dict(n=4, m=2))]:
    sysc = sw.SysConst(Lx=Lx, Ly=Ly, param=sw.Taumucs(g=-0.001, **kw))
    st = sw.Sys(sysc, "GPU")
    st.height.set(np.asfortranarray(np.abs(1.0 + 0.2 * rng.standard_normal((Lx, Ly))) + 0.06))
    sw.fused_steps(st, sysc, 4, log_minmax=True, log_wetted=True)
    sw.fused_steps(st, sysc, 3)
    th = sw.Field(Lx, Ly).set(np.full((Lx, Ly), 1 / 9))
    sw.fused_steps(st, sysc, 2, θ=th, slip_variant=2, incl=([1e-4, 0.0], 1.0))
    print("ok", Lx, Ly, kw, float(st.height.t.sum()))
# small-lattice flavours: marching lean kernels forced, then the tile kernel with CUDA-graph capture + replay
os.environ["SWALBE_TILE_MAX"] = "0"
sysc = sw.SysConst(Lx=70, Ly=44, param=sw.Taumucs())
st = sw.Sys(sysc, "GPU")
st.height.set(np.asfortranarray(np.abs(1.0 + 0.2 * rng.standard_normal((70, 44))) + 0.06))
for _ in range(3):
    sw.fused_steps(st, sysc, 9)
del os.environ["SWALBE_TILE_MAX"]
for _ in range(4):
    sw.fused_steps(st, sysc, 9)
    sw.fused_steps(st, sysc, 9)
print("ok small-lattice flavours", float(st.height.t.sum()))
sysc = sw.SysConst(Lx=200, Ly=24, param=sw.Taumucs(kbt=1e-6))
st = sw.Sys(sysc, "GPU", kind="thermal")
sw.fused_steps(st, sysc, 3, thermal_seed=5)
sim = DistSim(sysc, 0, 1, None, thermal_seed=5)
h = sw.Field(200, 24, fill=1.0)
z = sw.Field(200, 24)
sim.set_state(h, z, z)
sim.time_loop(3)
sim.get_state(h)
sim.close()
for op in ("filmpressure", "hgradp", "slippage", "update", "equilibrium", "BGKandStream", "moments"):
    f = getattr(sw, op)
    f(st, sysc) if op not in ("hgradp", "update", "moments") else f(st)
sw.thermal(st, sysc, seed=3, step=1)  # per-operator thermal kernel (shared-memory noise tables)
# on-device initial conditions (slab offsets, noisy variants with the shared-memory tables), circshift, cospi
for jb, rows in ((0, 24), (7, 10)):
    f = sw.Field(200, rows)
    sw.singledroplet(f, 40, 1 / 6, (100, 12), j_begin=jb)
    sw.torus(f, 8, 30, 1 / 9, (100, 12), noise=0.01, seed=2, j_begin=jb)
    sw.rivulet(f, 10, 1 / 9, "x", 12, noise=0.01, seed=2, j_begin=jb)
    sw.sinewave2d(f, 1.0, 0.01, 3, 2, j_begin=jb, Ly=24)
    sw.randinterface(f, 1.0, 0.01, seed=4, j_begin=jb)
g = sw.Field(200, 24)
sw.circshift(g, h, (3, -5))
sw.cospi_field(g)
print("stats", sw.field_stats(st.height))
import torch
torch.cuda.synchronize()
print("sanitize run complete")
