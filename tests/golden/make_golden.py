#!/usr/bin/env python
"""Generates tests/golden/*.npz -- small committed fixtures.

The reference is Julia and cannot be run in the build image, so these fixtures are produced by the ORACLE (the C
restatement, cross-checked bit for bit against the NumPy restatement and pinned to the reference's known-answer tests in
tests/test_oracle_golden.py).  They freeze the oracle's trajectories at commit time: a later change to the oracle, the
compiler flags or the kernels that moves a single bit shows up against these files, on the CPU (oracle) and on the GPU.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import oracle_c as oc  # noqa: E402
from oracle import oracle_np as onp  # noqa: E402

CASES = {
    # name: (Lx, Ly, nsteps, Params kwargs, seed)
    "default_25x26": (25, 26, 7, dict(g=-0.001, gamma=0.0005), 11),
    "n3m2_33x20": (33, 20, 5, dict(n=3, m=2, hmin=0.07, gamma=0.01, delta=2.0), 12),
    "tau075_20x31": (20, 31, 6, dict(tau=0.75, g=0.002), 13),
}
FIELDS = ("height", "velx", "vely", "pressure", "Fx", "Fy", "fout", "feq")


def initial_state(Lx, Ly, seed):
    rng = np.random.default_rng(seed)
    st = onp.State(Lx, Ly)
    st.height[...] = np.abs(1.0 + 0.2 * rng.standard_normal((Lx, Ly))) + 0.06
    st.velx[...] = 0.01 * rng.standard_normal((Lx, Ly))
    st.vely[...] = 0.01 * rng.standard_normal((Lx, Ly))
    st.ftemp[...] = 0.1 + 0.01 * rng.random((Lx, Ly, 9))
    return st


def main():
    out = {}
    for name, (Lx, Ly, nsteps, kw, seed) in CASES.items():
        st = initial_state(Lx, Ly, seed)
        for f in ("height", "velx", "vely", "ftemp"):
            out[f"{name}/in/{f}"] = getattr(st, f).copy()
        oc.time_loop(st, onp.Params(**kw), nsteps=nsteps)
        for f in FIELDS:
            out[f"{name}/out/{f}"] = getattr(st, f).copy()
    # BASELINE config 1 (README Rayleigh-Taylor): Lx=Ly=100, g=-0.001, gamma=0.0005, Tmax=1000, h0=1, eps=0.01
    p = onp.Params(Tmax=1000, g=-0.001, gamma=0.0005)
    st = onp.State(100, 100)
    st.height[...] = onp.rayleightaylor_ic(100, 100, kx=15, ky=18, eps=0.01)
    dh, _ = oc.time_loop(st, p, log_dh=True)
    out["readme_rt/height_after_1000"] = st.height.copy()
    out["readme_rt/dh"] = dh
    np.savez_compressed(os.path.join(HERE, "oracle_trajectories.npz"), **out)
    print("wrote", os.path.join(HERE, "oracle_trajectories.npz"), f"{len(out)} arrays")


if __name__ == "__main__":
    main()
