"""The slab runtime with ONE rank (self-exchange by device copies) against the single-GPU fused loop on the same lattice:
isolates what the ghost-row launch pattern (two edge strips + interior per step, three streams) costs by itself.
   python tools/dist_probe.py [film|thermal|thermal_moving] [L]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse as ap
import numpy as np, torch
import swalbe_b200 as sw
import bench
from swalbe_b200 import _lib
from swalbe_b200.dist import DistSim

wl = sys.argv[1] if len(sys.argv) > 1 else "film"
L = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
K = 98
prm = sw.Taumucs(**bench.workload_params(ap.Namespace(workload=wl), K))
sysc = sw.SysConst(Lx=L, Ly=L, param=prm)
seed = 1234 if wl != "film" else None
sim = DistSim(sysc, 0, 1, None, thermal_seed=seed)
h = sw.Field(L, L).set(bench.initial_height(L, workload=wl)); z = sw.Field(L, L)
sim.set_state(h, z, z)
th = None
if wl == "thermal_moving":
    th = sw.Field(L, L).set(bench.theta_pattern(L))
    sim.set_theta(sw.cospi_field(th))
sim.time_loop(10)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(); sim.time_loop(K, 10); e1.record(); torch.cuda.synchronize()
print(f"[{wl}] slab runtime, 1 rank: {e0.elapsed_time(e1) / K:.4f} ms/step (loop on its own streams: {sim.last_loop_ms() / K:.4f})", flush=True)
sim.close(); del sim
st = sw.Sys(sysc, "GPU", kind="thermal" if seed is not None else "simple")
st.height.set(h)
kw = dict(thermal_seed=seed, θ=th, skip_aux=True, pressure_variant=_lib.PRESSURE_POWER_BROAD)
sw.fused_steps(st, sysc, 10, **kw)
torch.cuda.synchronize(); e0.record(); sw.fused_steps(st, sysc, K, step0=10, **kw); e1.record(); torch.cuda.synchronize()
print(f"[{wl}] single-GPU fused loop: {e0.elapsed_time(e1) / K:.4f} ms/step", flush=True)
