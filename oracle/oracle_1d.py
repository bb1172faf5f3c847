"""NumPy restatement of Swalbe.jl's 1-D (D1Q3) thin-film LBM step  --  TEST INFRASTRUCTURE ONLY.

SURVEY.md 8f4: the `State_1D` family (the reference runs it on the CPU only; there is no device string in its 1-D
allocator or drivers).  Same conventions and the same pinning status as oracle_np.py: every expression keeps Julia's
association, NumPy rounds each operation separately (no FMA), `circshift(a, s)` == `np.roll(a, s)`; pinned against the
reference's own 1-D known answers (tests/test_oracle_1d.py transcribes test/collide.jl:141-226,
test/equilibrium.jl:87-153, test/moments.jl, test/pressure.jl:56-104, test/differences.jl:79-138,
test/forcing.jl:71-102, 125-138, test/simulate.jl:8-30), not against a live Julia run.
"""
from __future__ import annotations

import numpy as np

from .oracle_np import cospi, fast_32, fast_93, power_broad


def viewdists_1D(f):
    """src/collide.jl:303-309 -- three length-L views of f[L, 3]."""
    return f[:, 0], f[:, 1], f[:, 2]


def equilibrium(feq, height, velocity, gravity):
    """equilibrium!(feq, height, velocity, gravity)   src/equilibrium.jl:169-181"""
    f0, f1, f2 = viewdists_1D(feq)
    h, v, g = height, velocity, gravity
    f0[...] = h * ((1 - (0.5 * g) * h) - v * v)
    f1[...] = h * (((0.25 * g) * h + 0.5 * v) + 0.5 * (v * v))
    f2[...] = h * (((0.25 * g) * h - 0.5 * v) + 0.5 * (v * v))


def BGKandStream(fout, feq, ftemp, F, tau):
    """BGKandStream!(fout, feq, ftemp, F::Vector, τ)   src/collide.jl:179-201"""
    fe0, fe1, fe2 = viewdists_1D(feq)
    ft0, ft1, ft2 = viewdists_1D(ftemp)
    fo0, fo1, fo2 = viewdists_1D(fout)
    omeg = 1 - 1 / tau
    it = 1 / tau
    fo0[...] = omeg * ft0 + it * fe0
    fo1[...] = (omeg * ft1 + it * fe1) + 1 / 2 * F
    fo2[...] = (omeg * ft2 + it * fe2) - 1 / 2 * F
    ft0[...] = fo0
    ft1[...] = np.roll(fo1, 1)
    ft2[...] = np.roll(fo2, -1)
    fout[...] = ftemp


def moments(height, vel, fout):
    """moments!(height::Vector, vel, fout)   src/moments.jl:54-62 (sum! folds the three columns in order onto 0)"""
    f0, f1, f2 = viewdists_1D(fout)
    height[...] = ((0.0 + f0) + f1) + f2
    with np.errstate(divide="ignore", invalid="ignore"):
        vel[...] = (f1 - f2) / height


def _kappa(cospi_theta, n, m, hmin):
    return (1 - cospi_theta) * (n - 1) * (m - 1) / ((n - m) * hmin)


def filmpressure(output, f, dgrad, gamma, cospi_theta, n, m, hmin, hcrit, variant="fast"):
    """filmpressure!(output::Vector, f, dgrad, γ, θ, n, m, hmin, hcrit)   src/pressure.jl:196-227 (fast_93 / fast_32)
    filmpressure!(state::LBM_state_1D, sys; ...)                          src/pressure.jl:230-256 (power_broad)
    cospi_theta: cospi(θ) as a number or a length-L array (evaluated once by the caller)."""
    hip, him = np.roll(f, 1), np.roll(f, -1)
    if dgrad is not None:
        dgrad[:, 0], dgrad[:, 1] = hip, him
    with np.errstate(divide="ignore", invalid="ignore"):
        x = hmin / (f + hcrit)
        if variant == "fast":
            if (n, m) == (9, 3):
                pw = fast_93(x)
            elif (n, m) == (3, 2):
                pw = fast_32(x)
            else:
                raise ValueError(f"DomainError({(n, m)})")
        else:
            pw = power_broad(x, n) - power_broad(x, m)
        output[...] = -gamma * (_kappa(cospi_theta, n, m, hmin) * pw)
    output[...] = output - gamma * (hip - 2 * f + him)


def lap(output, f, dgrad=None):
    """∇²f!(output, f::Vector, dgrad)   src/differences.jl:77-85"""
    hip, him = np.roll(f, 1), np.roll(f, -1)
    output[...] = hip - 2 * f + him


def grad(output, f, a=None):
    """∇f!(output::Vector, f, dgrad, a) | ∇f!(output, f::Vector, dgrad)   src/differences.jl:208-230"""
    fip, fim = np.roll(f, 1), np.roll(f, -1)
    output[...] = (-0.5 * (fip - fim)) if a is None else (a * -0.5 * (fip - fim))


def hgradp(out, pressure, height):
    """h∇p!(state::LBM_state_1D)   src/forcing.jl:189-198"""
    grad(out, pressure, height)


def slippage(slip, height, vel, delta, mu):
    """slippage!(slip, height, vel, δ, μ)   src/forcing.jl:68-71"""
    with np.errstate(divide="ignore", invalid="ignore"):
        slip[...] = (6 * mu * height * vel) / (2 * (height * height) + 6 * delta * height + 3 * (delta * delta))


class State1D:
    """State_1D as built by Sys(sysc::Consts_1D)   src/initialize.jl:587-598 (height = 1, everything else 0)."""

    def __init__(self, L):
        self.L = L
        z = lambda *s: np.zeros(s, order="F")  # noqa: E731
        self.fout, self.ftemp, self.feq = z(L, 3), z(L, 3), z(L, 3)
        self.height = np.ones(L)
        self.vel, self.pressure, self.F, self.slip, self.hgradp = z(L), z(L), z(L), z(L), z(L)
        self.dgrad = z(L, 2)


def step(st: State1D, p, cospi_theta=None, pvariant="power_broad"):
    """one iteration of time_loop(sys::SysConst_1D, state::State_1D[, θ])   src/simulate.jl:98-136"""
    ct = cospi(p.theta) if cospi_theta is None else cospi_theta
    filmpressure(st.pressure, st.height, st.dgrad, p.gamma, ct, p.n, p.m, p.hmin, p.hcrit, variant=pvariant)
    hgradp(st.hgradp, st.pressure, st.height)
    slippage(st.slip, st.height, st.vel, p.delta, p.mu)
    st.F[...] = -st.hgradp - st.slip
    equilibrium(st.feq, st.height, st.vel, p.g)
    BGKandStream(st.fout, st.feq, st.ftemp, st.F, p.tau)
    moments(st.height, st.vel, st.fout)


def time_loop(st: State1D, p, nsteps=None, **kw):
    """returns the per-step max - min log of the Δh variant (src/simulate.jl:138-157)"""
    dh = []
    for _ in range(p.Tmax if nsteps is None else nsteps):
        dh.append(float(st.height.max() - st.height.min()))
        step(st, p, **kw)
    return dh
