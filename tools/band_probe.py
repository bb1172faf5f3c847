"""Device time of the banded sweeps of swalbe_time_loop_host without the copies (SWALBE_HOST_NOCOPY) for forced launch
geometries, against the whole-lattice loop: what a band launch costs.  python tools/band_probe.py [--L 8192] [--steps 12]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import swalbe_b200 as sw  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=8192)
ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--configs", default="0:0,128:0,160:0,192:0,224:0,224:64,224:128,224:256,160:128,160:256,128:128,128:256")
ap.add_argument("--bands", default="1024,2048")
args = ap.parse_args()
L, K = args.L, args.steps
h_host = torch.ones(L * L, dtype=torch.float64).pin_memory()
out_host = torch.empty_like(h_host).pin_memory()
sysc = sw.SysConst(Lx=L, Ly=L, param=sw.Taumucs(Tmax=K))
st = sw.Sys(sysc, "GPU")


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


os.environ["SWALBE_HOST_NOCOPY"] = "1"
for lazy in (True, False):
    t = timed(lambda: sw.fused_steps(st, sysc, K, skip_aux=True, lazy_populations=lazy))
    print(f"whole lattice, {K} steps, lazy={lazy}: {t:.3f} ms ({t / K:.3f} per step)", flush=True)
    for band in args.bands.split(","):
        os.environ["SWALBE_BAND_ROWS"] = band
        for cfg in args.configs.split(","):
            nt, rows = cfg.split(":")
            # a fresh state per configuration: launch geometries are cached per plan
            s2 = sw.Sys(sysc, "GPU")
            os.environ["SWALBE_NT"], os.environ["SWALBE_ROWS"] = nt, rows
            if nt == "0" and rows == "0":
                os.environ["SWALBE_DEBUG"] = "1"
            try:
                t = timed(lambda: sw.fused_steps(s2, sysc, K, skip_aux=True, lazy_populations=lazy, host_in=h_host))
                print(f"  bands of {band}, NT={nt} rows/CTA={rows}: {t:.3f} ms ({t / K:.3f} per step)", flush=True)
            except Exception as e:  # a forced geometry may not exist
                print(f"  bands of {band}, NT={nt} rows/CTA={rows}: {e}", flush=True)
            os.environ["SWALBE_DEBUG"] = "0"
            del s2
        os.environ["SWALBE_NT"], os.environ["SWALBE_ROWS"] = "0", "0"
