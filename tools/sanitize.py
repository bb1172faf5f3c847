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
# ---- round 2 flavours ----------------------------------------------------------------------------------------------
# tau != 1 from-moments kernels (first step on the planes, FM steps, chunk that vouches for its moments), hints on / off
for hints in ("1", "7"):
    os.environ["SWALBE_FM_HINTS"] = hints
    os.environ["SWALBE_FM_PREFETCH"] = "3" if hints == "7" else "0"
    sysc = sw.SysConst(Lx=300, Ly=40, param=sw.Taumucs(τ=0.8, n=3, m=2, hmin=0.07))
    st = sw.Sys(sysc, "GPU")
    st.height.set(np.asfortranarray(np.abs(1.0 + 0.2 * rng.standard_normal((300, 40))) + 0.06))
    sw.fused_steps(st, sysc, 5, skip_aux=True)
    sw.fused_steps(st, sysc, 4, moments_consistent=True)
del os.environ["SWALBE_FM_HINTS"], os.environ["SWALBE_FM_PREFETCH"]
# persistent cluster kernel: every cluster size, logs, lazy populations, in place; then the tile kernel with a theta field
for csize in ("0", "1", "4", "16"):
    os.environ["SWALBE_CLUSTER_SIZE"] = csize
    sysc = sw.SysConst(Lx=64, Ly=64, param=sw.Taumucs(g=-0.001))
    st = sw.Sys(sysc, "GPU")
    st.height.set(np.asfortranarray(np.abs(1.0 + 0.2 * rng.standard_normal((64, 64))) + 0.06))
    sw.fused_steps(st, sysc, 6, log_minmax=True, log_wetted=True)
    sw.fused_steps(st, sysc, 5, lazy_populations=True, skip_aux=True)
del os.environ["SWALBE_CLUSTER_SIZE"]
sysc = sw.SysConst(Lx=130, Ly=64, param=sw.Taumucs(n=3, m=2, hmin=0.07))
st = sw.Sys(sysc, "GPU")
th = sw.Field(130, 64).set(1 / 9 + rng.random((130, 64)) / 36)
sw.fused_steps(st, sysc, 4, θ=th, skip_aux=True)
# neighbour-sync kernels: LDGSTS rows and per-warp TMA rows, film and thermal
os.environ["SWALBE_NSYNC"] = "1"; os.environ["SWALBE_TILE_MAX"] = "0"; os.environ["SWALBE_CLUSTER"] = "0"
for bulk in ("0", "2"):
    os.environ["SWALBE_BULK"] = bulk
    sysc = sw.SysConst(Lx=512, Ly=48, param=sw.Taumucs())
    st = sw.Sys(sysc, "GPU")
    st.height.set(np.asfortranarray(np.abs(1.0 + 0.2 * rng.standard_normal((512, 48))) + 0.06))
    sw.fused_steps(st, sysc, 5, skip_aux=True)
sysc = sw.SysConst(Lx=200, Ly=24, param=sw.Taumucs(kbt=1e-6))
st = sw.Sys(sysc, "GPU", kind="thermal")
sw.fused_steps(st, sysc, 3, thermal_seed=5)
for k in ("SWALBE_NSYNC", "SWALBE_TILE_MAX", "SWALBE_CLUSTER", "SWALBE_BULK"):
    del os.environ[k]
# 1-D family: persistent loop (tau == 1 and tau != 1, logs), step-by-step kernel (too long for one CTA), operators
for L, kw in ((300, dict()), (1000, dict(τ=0.8)), (20000, dict(g=0.001))):
    s1 = sw.SysConst_1D(L=L, param=sw.Taumucs(Tmax=9, tdump=4, **kw))
    s = sw.Sys(s1)
    s.height.set(np.abs(1.0 + 0.2 * rng.standard_normal(L)) + 0.06)
    sw.time_loop(s1, s, [])
    for op in ("filmpressure", "hgradp", "slippage", "update", "equilibrium", "BGKandStream", "moments"):
        f = getattr(sw, op)
        f(s, s1) if op not in ("hgradp", "update", "moments") else f(s)
import torch
torch.cuda.synchronize()
print("sanitize run complete")
