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


# ---- the expanded 1-D kinds: State_thermal_1D, State_gamma_1D, StateWithBound_1D (src/initialize.jl:304-341, :587-616) -----
# Pinned against the reference's known answers: test/collide.jl:141-150 (bounce-back state at tau = 1 without forces ==
# plain streaming), test/forcing.jl:165-204 (inclination, surface-tension gradient, constant-field rho update),
# test/pressure.jl:56-131 (State_gamma_1D pressure, active-matter pressure).


def inclination(F, height, alpha, t=1000, tstart=0, tsmooth=1):
    """inclination!(α::Float64, state::State_1D; t, tstart, tsmooth)   src/forcing.jl:379-389"""
    F[...] = F + height * alpha * (0.5 + 0.5 * np.tanh((t - tstart) / tsmooth))


def gradgamma(out, gamma, height=None, delta=None):
    """∇γ!(state)  src/forcing.jl:423-432  |  ∇γ!(state, sys)  :434-447"""
    fip, fim = np.roll(gamma, 1), np.roll(gamma, -1)
    if height is None:
        out[...] = -3 / 2 * ((fip - fim) / 2.0)
    else:
        h = height
        out[...] = (2 * (h * h) + 6 * delta * h + 3 * (delta * delta)) / (6 * h) * h / 2 * ((fip - fim) / 2.0)


def filmpressure_gamma(output, f, gamma, cospi_theta, n, m, hmin, hcrit, rho=None, Gamma=0.0, ftemp=None):
    """filmpressure!(state::State_gamma_1D, sys; γ)   src/pressure.jl:284-315 (γ scalar or length-L array, ftemp columns 1, 2
    receive the two contributions) | filmpressure!(state::Expanded_1D, sys)  :258-282 (ftemp=None) |
    filmpressure!(output::Vector, f, dgrad, rho, γ, θ, n, m, hmin, hcrit; Gamma)  :318-338 (rho given)"""
    hip, him = np.roll(f, 1), np.roll(f, -1)
    g = gamma if rho is None else gamma + Gamma * rho
    with np.errstate(divide="ignore", invalid="ignore"):
        x = hmin / (f + hcrit)
        disj = -g * (_kappa(cospi_theta, n, m, hmin) * (power_broad(x, n) - power_broad(x, m)))
    lap = hip - 2 * f + him
    if ftemp is not None:
        ftemp[:, 1] = disj
        ftemp[:, 2] = -g * lap
    output[...] = disj - g * lap


def obslist1D(obs):
    """obslist1D(obs)   src/obstacle.jl:6-35 -> interior, [obsright, obsleft]"""
    L = len(obs)
    interior, obsleft, obsright = np.zeros(L), np.zeros(L), np.zeros(L)
    for i in range(L):
        ip, im = (i + 1) % L, (i - 1) % L
        interior[i] = 1 if (obs[i] == 1 and obs[ip] == 1 and obs[im] == 1) else 0
        obsleft[i] = 1 if (obs[i] == 1 and obs[im] == 1) else 0
        obsright[i] = 1 if (obs[i] == 1 and obs[ip] == 1) else 0
    return interior, [obsright, obsleft]


def BGKandStream_bound(fout, feq, ftemp, fbound, F, border, tau):
    """BGKandStream!(state::StateWithBound_1D, sys::SysConstWithBound_1D)   src/collide.jl:214-249"""
    fe0, fe1, fe2 = viewdists_1D(feq)
    ft0, ft1, ft2 = viewdists_1D(ftemp)
    fo0, fo1, fo2 = viewdists_1D(fout)
    _, fb1, fb2 = viewdists_1D(fbound)
    omeg, it = 1 - 1 / tau, 1 / tau
    fo0[...] = omeg * ft0 + it * fe0
    fo1[...] = (omeg * ft1 + it * fe1) + 1 / 2 * F
    fo2[...] = (omeg * ft2 + it * fe2) - 1 / 2 * F
    fb1[...] = fo1 * border[0]
    fb2[...] = fo2 * border[1]
    fo1[...] = fo1 - fb1
    fo2[...] = fo2 - fb2
    ft0[...] = fo0
    ft1[...] = np.roll(fo1, 1)
    ft2[...] = np.roll(fo2, -1)
    ft1[...] = ft1 + fb2
    ft2[...] = ft2 + fb1
    fout[...] = ftemp


def update_rho(rho, rho_int, height, differentials, D=1.0, M=0.0):
    """update_rho!(rho, rho_int, height, dgrad, differentials; D, M)   src/forcing.jl:399-417"""
    lap_rho, grad_rho, lap_h, grad_h = (differentials[:, k] for k in range(4))
    lap(lap_rho, rho)
    grad(grad_rho, rho)
    lap(lap_h, height)
    grad(grad_h, height)
    with np.errstate(divide="ignore", invalid="ignore"):
        rho_int[...] = (D * lap_rho - M * (grad_rho * grad_rho + rho * lap_rho)
                        - D * (grad_rho * grad_h / height + rho * (lap_h / height - (grad_h / height) * (grad_h / height))))
    rho[...] = rho + rho_int


def thermal_amplitude(height, kbt, mu, delta):
    """the factor of the unit normals in thermal!(fluc, height, kᵦT, μ, δ)   src/forcing.jl:322-333"""
    return np.sqrt(2 * kbt * mu * 6 * height / (2 * height * height + 6 * height * delta + 3 * delta * delta))


def step_gamma(st: State1D, p, gamma, dgamma, cospi_theta=None, alpha=None, incl_factor=None):
    """one iteration of the loop of run_gamma (src/simulate.jl:541-547); gamma / dgamma None: the plain loop; alpha: the
    callback slot of time_loop(sys, state, inclination!, α)  :159-179"""
    ct = cospi(p.theta) if cospi_theta is None else cospi_theta
    if gamma is None:
        filmpressure(st.pressure, st.height, st.dgrad, p.gamma, ct, p.n, p.m, p.hmin, p.hcrit, variant="power_broad")
    else:
        filmpressure_gamma(st.pressure, st.height, gamma, ct, p.n, p.m, p.hmin, p.hcrit, ftemp=st.ftemp)
    hgradp(st.hgradp, st.pressure, st.height)
    slippage(st.slip, st.height, st.vel, p.delta, p.mu)
    st.F[...] = (-st.hgradp - st.slip) if dgamma is None else (-st.hgradp - st.slip - dgamma)
    if alpha is not None:
        st.F[...] = st.F + st.height * alpha * incl_factor
    equilibrium(st.feq, st.height, st.vel, p.g)
    BGKandStream(st.fout, st.feq, st.ftemp, st.F, p.tau)
    moments(st.height, st.vel, st.fout)


def step_bound(st: State1D, fbound, p, border, cospi_theta=None):
    """one iteration of time_loop(sys::SysConstWithBound_1D, state::StateWithBound_1D)   src/simulate.jl:181-204"""
    ct = cospi(p.theta) if cospi_theta is None else cospi_theta
    filmpressure_gamma(st.pressure, st.height, p.gamma, ct, p.n, p.m, p.hmin, p.hcrit)
    hgradp(st.hgradp, st.pressure, st.height)
    slippage(st.slip, st.height, st.vel, p.delta, p.mu)
    st.F[...] = -st.hgradp - st.slip
    equilibrium(st.feq, st.height, st.vel, p.g)
    BGKandStream_bound(st.fout, st.feq, st.ftemp, fbound, st.F, border, p.tau)
    moments(st.height, st.vel, st.fout)
