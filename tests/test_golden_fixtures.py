"""Committed fixtures (tests/golden/oracle_trajectories.npz, made by tests/golden/make_golden.py): the oracle must still
reproduce them bit for bit on the CPU, and the fused CUDA loop must reproduce them on the GPU."""
import os

import numpy as np
import pytest

from oracle import oracle_c as oc
from oracle import oracle_np as onp
from tests.golden.make_golden import CASES, FIELDS

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_trajectories.npz"))


def _state_from_fixture(name, Lx, Ly):
    st = onp.State(Lx, Ly)
    for f in ("height", "velx", "vely", "ftemp"):
        getattr(st, f)[...] = GOLD[f"{name}/in/{f}"]
    return st


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("impl", ["c", "numpy"])
def test_oracle_reproduces_fixture(name, impl):
    Lx, Ly, nsteps, kw, _ = CASES[name]
    st = _state_from_fixture(name, Lx, Ly)
    (oc if impl == "c" else onp).time_loop(st, onp.Params(**kw), nsteps=nsteps)
    for f in FIELDS:
        assert np.array_equal(getattr(st, f), GOLD[f"{name}/out/{f}"]), f


def test_oracle_reproduces_readme_rayleigh_taylor():
    p = onp.Params(Tmax=1000, g=-0.001, gamma=0.0005)
    st = onp.State(100, 100)
    st.height[...] = onp.rayleightaylor_ic(100, 100, kx=15, ky=18, eps=0.01)
    dh, _ = oc.time_loop(st, p, log_dh=True)
    assert np.array_equal(st.height, GOLD["readme_rt/height_after_1000"])
    assert np.array_equal(dh, GOLD["readme_rt/dh"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_reproduces_fixture(name):
    import swalbe_b200 as sw

    Lx, Ly, nsteps, kw, _ = CASES[name]
    jl = {"gamma": "γ", "delta": "δ", "tau": "τ"}
    sysc = sw.SysConst(Lx=Lx, Ly=Ly, param=sw.Taumucs(**{jl.get(k, k): v for k, v in kw.items()}))
    st = sw.Sys(sysc, "GPU")
    for f in ("height", "velx", "vely", "ftemp"):
        getattr(st, f).set(GOLD[f"{name}/in/{f}"])
    sw.fused_steps(st, sysc, nsteps)
    for f in FIELDS:
        assert np.array_equal(getattr(st, f).numpy(), GOLD[f"{name}/out/{f}"]), f


@pytest.mark.gpu
def test_gpu_reproduces_readme_rayleigh_taylor():
    import swalbe_b200 as sw

    sysc = sw.SysConst(Lx=100, Ly=100, param=sw.Taumucs(Tmax=1000, g=-0.001, γ=0.0005))
    h, diff = sw.run_rayleightaylor(sysc, "GPU", ϵ=0.01, verbos=False)
    assert np.array_equal(h.numpy(), GOLD["readme_rt/height_after_1000"])
    assert np.array_equal(np.asarray(diff), GOLD["readme_rt/dh"])
    rel = np.abs(h.numpy() - GOLD["readme_rt/height_after_1000"]).max() / np.abs(GOLD["readme_rt/height_after_1000"]).max()
    assert rel <= 1e-12  # north_star tolerance (met with 0 ulp)
