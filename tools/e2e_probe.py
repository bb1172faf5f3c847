"""Whole-job timing of run_host (swalbe_time_loop_host) at one lattice size: band height / sweep length sweep, against the
plain sequence copy -> equilibrium! -> time_loop -> copy.  Wall clock around the job, best of 3, bitwise check.
  python tools/e2e_probe.py [--L 8192] [--steps 20]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import swalbe_b200 as sw  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=8192)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--bands", default="0,256,512,1024,2048,4096")
ap.add_argument("--kmax", default="0,6,16")
ap.add_argument("--trace", action="store_true", help="one traced job (device timeline on stderr) per configuration")
args = ap.parse_args()
L, K = args.L, args.steps
i = np.arange(L, dtype=np.float64)[:, None]
j = np.arange(L, dtype=np.float64)[None, :]
h0 = 1.0 + 1e-3 * np.sin(2 * np.pi * i / L) * np.sin(2 * np.pi * j / L)
h_host = torch.from_numpy(np.ascontiguousarray(h0.transpose())).pin_memory()
out_host = torch.empty_like(h_host).pin_memory()
sysc = sw.SysConst(Lx=L, Ly=L, param=sw.Taumucs(Tmax=K))
st = sw.Sys(sysc, "GPU")


def plain():
    st.height.t.copy_(h_host, non_blocking=True)
    st.velx.t.zero_(); st.vely.t.zero_()
    sw.equilibrium(st, sysc)
    sw.time_loop(sysc, st)
    out_host.copy_(st.height.t, non_blocking=True)
    torch.cuda.synchronize()


def streamed():
    st.velx.t.zero_(); st.vely.t.zero_()
    sw.run_host(sysc, h_host, out_host, state=st)


def best(fn, n=3):
    fn()
    dt = 1e9
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        dt = min(dt, time.perf_counter() - t0)
    return dt


def h2d():
    st.height.t.copy_(h_host, non_blocking=True)
    torch.cuda.synchronize()


def d2h():
    out_host.copy_(st.height.t, non_blocking=True)
    torch.cuda.synchronize()


def loop_only():
    sw.time_loop(sysc, st)
    torch.cuda.synchronize()


for name, fn in (("H2D of one plane", h2d), ("D2H of one plane", d2h), ("time_loop alone", loop_only)):
    dt = best(fn)
    print(f"{name}: {dt * 1e3:8.2f} ms" + (f"  ({L * L * 8 / dt / 1e9:.1f} GB/s)" if "2" in name else ""), flush=True)
dt = best(plain)
want = out_host.clone()
print(f"L={L} steps={K} plain sequence: {dt * 1e3:8.2f} ms  {L * L * K / dt / 1e6:9.1f} MLUPS", flush=True)
for kmax in args.kmax.split(","):
    for band in args.bands.split(","):
        os.environ["SWALBE_BAND_ROWS"], os.environ["SWALBE_HOST_KMAX"] = band, kmax
        out_host.zero_()
        dt = best(streamed)
        ok = torch.equal(out_host, want)
        print(f"L={L} steps={K} band_rows={band:>5} kmax={kmax:>3}: {dt * 1e3:8.2f} ms  {L * L * K / dt / 1e6:9.1f} MLUPS  "
              f"bitwise={'yes' if ok else 'NO'}", flush=True)
        if args.trace:
            for nocopy in ("0", "1"):
                os.environ["SWALBE_HOST_TRACE"], os.environ["SWALBE_HOST_NOCOPY"] = "1", nocopy
                sys.stderr.write(f"--- trace band_rows={band} kmax={kmax} nocopy={nocopy}\n"); sys.stderr.flush()
                streamed()
            os.environ["SWALBE_HOST_TRACE"], os.environ["SWALBE_HOST_NOCOPY"] = "0", "0"
