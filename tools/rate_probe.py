"""Device-timed rate of swalbe_time_loop for one configuration (lean steps only: skip_aux), for A/B sweeps over the
library's environment knobs (SWALBE_FM, SWALBE_FM_PREFETCH, SWALBE_NT, SWALBE_RMAX, SWALBE_TILE_THETA ...).
  python tools/rate_probe.py [--L 8192] [--tau 1.0] [--steps 100] [--theta-field] [--thermal] [--n 9 --m 3] [--label x]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import swalbe_b200 as sw
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=8192); ap.add_argument("--Ly", type=int, default=0)
ap.add_argument("--tau", type=float, default=1.0); ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--calls", type=int, default=1, help="split the timed steps into this many swalbe_time_loop calls")
ap.add_argument("--theta-field", action="store_true"); ap.add_argument("--thermal", action="store_true")
ap.add_argument("--n", type=int, default=9); ap.add_argument("--m", type=int, default=3)
ap.add_argument("--g", type=float, default=0.0); ap.add_argument("--lazy", action="store_true")
ap.add_argument("--label", default="")
a = ap.parse_args()
L, Ly = a.L, a.Ly or a.L
kw = dict(τ=a.tau, g=a.g, n=a.n, m=a.m)
if (a.n, a.m) == (3, 2):
    kw.update(hmin=0.07)
if a.thermal:
    kw.update(kbt=1e-7)
sysc = sw.SysConst(Lx=L, Ly=Ly, param=sw.Taumucs(**kw))
st = sw.Sys(sysc, "GPU", kind="thermal" if a.thermal else "simple")
st.height.set(bench.initial_height(L, Ly))
th = sw.Field(L, Ly).set(bench.theta_pattern(L, Ly)) if a.theta_field else None
if a.tau != 1.0:
    sw.equilibrium(st, sysc)
    st.ftemp.t.copy_(st.feq.t)
from swalbe_b200 import _lib
opt = dict(θ=th, thermal_seed=1234 if a.thermal else None, skip_aux=True, lazy_populations=a.lazy,
           pressure_variant=_lib.PRESSURE_POWER_BROAD)
sw.fused_steps(st, sysc, 10, **opt)
per = max(1, a.steps // a.calls)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for c in range(a.calls):
    sw.fused_steps(st, sysc, per, step0=10 + c * per, moments_consistent=True, **opt)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / (per * a.calls)
env = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("SWALBE_"))
print(f"[{a.label or 'probe'}] {L}x{Ly} tau={a.tau} n,m={a.n},{a.m} theta_field={a.theta_field} thermal={a.thermal} lazy={a.lazy} {env}: "
      f"{L*Ly/ms/1e3:9.1f} MLUPS  {ms*1e3:.2f} us/step  (144 B/LU -> {L*Ly*144/ms/1e6:.0f} GB/s)", flush=True)
