"""Direct oracle parity on the BASELINE.json configurations AT THEIR NAMED GRID SIZES: the fused B200 loop
through the C ABI against the low-memory restatement of the reference (oracle_time_loop_lowmem, pinned bit for bit to
the pass-structured oracle by tests/test_oracle_golden.py), every value of height / velx / vely / pressure / fout
compared BITWISE (tolerance 0; north_star allows 1e-12 relative).

  C3  spinodal dewetting, 4096^2, n=3 m=2 hmin=0.07 γ=0.01, h = 1 + 0.01 N(0,1) seed 20261017, time_loop(sys, state, 1/9)
      (src/simulate.jl:26-45, scripts/Moving_wet_stab.jl:109), 10 steps
  C4  deterministic part: 4096^2, contact-angle pattern moved by (1,1) (scripts/Moving_wettability_structs.jl:139-152),
      μ = 1/12, 5 + 5 steps around a move
  C5  flat film + perturbation at 8192^2 (the bench's grid), Taumucs defaults, 3 steps
"""
import os

import numpy as np
import pytest

from oracle import oracle_c as oc
from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu


def _threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def _same_plane(field, k, want, name):
    """one Lx x Ly plane of a device Field against the oracle's (Fortran-ordered) array, without a second host copy"""
    t = field.t if k is None else field.t[k]
    got = t.cpu().numpy()  # (Ly, Lx), C order == the Fortran-ordered (Lx, Ly) plane transposed
    w = want.T
    if not np.array_equal(got, w):
        bad = got != w
        raise AssertionError(f"{name}{'' if k is None else f'[{k}]'}: {int(bad.sum())} of {bad.size} values differ, "
                             f"max abs {np.abs(got - w)[bad].max():.3e}")


def _compare(st, h, ux, uy, f, pr):
    _same_plane(st.height, None, h, "height")
    _same_plane(st.velx, None, ux, "velx")
    _same_plane(st.vely, None, uy, "vely")
    _same_plane(st.pressure, None, pr, "pressure")
    for k in range(9):
        _same_plane(st.fout, k, f[:, :, k], "fout")
        _same_plane(st.ftemp, k, f[:, :, k], "ftemp")


def test_c3_spinodal_4096_bitwise():
    import swalbe_b200 as sw

    L, nsteps = 4096, 10
    rng = np.random.default_rng(20261017)
    h0 = np.asfortranarray(1.0 * (1.0 + 0.01 * rng.standard_normal((L, L))))  # randinterface!  src/initialvalues.jl:23-33
    prm = dict(n=3, m=2, hmin=0.07)
    sysc = sw.SysConst(Lx=L, Ly=L, param=sw.Taumucs(γ=0.01, **prm))
    st = sw.Sys(sysc, "GPU")
    st.height.set(h0)
    sw.equilibrium(st, sysc)
    sw.fused_steps(st, sysc, nsteps, θ=1 / 9)  # the loop body of time_loop(sys, state, θ)
    h, ux, uy = h0.copy(order="F"), np.zeros((L, L), order="F"), np.zeros((L, L), order="F")
    f = np.zeros((L, L, 9), order="F")  # the drivers never initialise ftemp (src/simulate.jl:349-356)
    pr = oc.time_loop_lowmem(h, ux, uy, f, onp.Params(gamma=0.01, **prm), nsteps, threads=_threads())
    _compare(st, h, ux, uy, f, pr)
    assert abs(h.sum() - h0.sum()) < 1e-12 * h0.sum()


def test_c4_moving_contact_angle_pattern_4096_bitwise():
    import swalbe_b200 as sw

    L = 4096
    i = np.arange(L, dtype=np.float64)[:, None]
    j = np.arange(L, dtype=np.float64)[None, :]
    h0 = np.asfortranarray(1.0 + 0.1 * np.sin(2 * np.pi * i / L) * np.sin(2 * np.pi * j / L))
    theta = np.asfortranarray(1 / 9 + (1 / 36) * np.sin(2 * np.pi * 2 * i / L) * np.sin(2 * np.pi * 2 * j / L))
    prm = dict(n=3, m=2, hmin=0.07)
    sysc = sw.SysConst(Lx=L, Ly=L, param=sw.Taumucs(γ=0.01, δ=1.0, μ=1 / 12, **prm))
    st = sw.Sys(sysc, "GPU")
    st.height.set(h0)
    th, inp = sw.Field(L, L).set(theta), sw.Field(L, L).set(theta)
    ct0 = sw.cospi_field(th).numpy()  # the device's cospi.(θ), handed to the oracle as data
    sw.fused_steps(st, sysc, 5, θ=th, skip_aux=True)
    sw.move_substrate(th, inp, 98, 98)  # t % tmove == 0: circshift!(θ, input, (1, 1)); input .= θ
    ct1 = sw.cospi_field(th).numpy()
    assert np.array_equal(ct1, onp.circshift(ct0, (1, 1)))
    sw.fused_steps(st, sysc, 5, θ=th)
    h, ux, uy = h0.copy(order="F"), np.zeros((L, L), order="F"), np.zeros((L, L), order="F")
    f = np.zeros((L, L, 9), order="F")
    p = onp.Params(gamma=0.01, delta=1.0, mu=1 / 12, **prm)
    oc.time_loop_lowmem(h, ux, uy, f, p, 5, cospi_theta=ct0, threads=_threads())
    pr = oc.time_loop_lowmem(h, ux, uy, f, p, 5, cospi_theta=ct1, threads=_threads())
    _compare(st, h, ux, uy, f, pr)


def test_c5_film_8192_bitwise():
    import swalbe_b200 as sw

    L, nsteps = 8192, 3
    i = np.arange(L, dtype=np.float64)[:, None]
    j = np.arange(L, dtype=np.float64)[None, :]
    h0 = np.asfortranarray(1.0 + 1e-3 * np.sin(2 * np.pi * i / L) * np.sin(2 * np.pi * j / L))
    sysc = sw.SysConst(Lx=L, Ly=L, param=sw.Taumucs())
    st = sw.Sys(sysc, "GPU")
    st.height.set(h0)
    sw.fused_steps(st, sysc, nsteps)  # (8192^2 runs the bulk-copy lean kernel for steps 1-2, the full kernel for step 3)
    h, ux, uy = h0.copy(order="F"), np.zeros((L, L), order="F"), np.zeros((L, L), order="F")
    del h0
    f = np.zeros((L, L, 9), order="F")
    pr = oc.time_loop_lowmem(h, ux, uy, f, onp.Params(), nsteps, threads=_threads())
    _compare(st, h, ux, uy, f, pr)
