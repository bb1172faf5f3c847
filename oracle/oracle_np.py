"""NumPy restatement of Swalbe.jl's 2-D (D2Q9) thin-film LBM step  --  TEST INFRASTRUCTURE ONLY.

This file is the *oracle*: a CPU restatement of the reference's Julia algorithm, operation for
operation, used only by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs as the checker.  Nothing under ``swalbe.jl_b200/`` imports it and the
product path never falls back to it.

Pinning status.  Julia is not installed in the build image and the reference has no non-Julia
implementation, so the oracle cannot be run against the live reference.  It IS pinned against every
known-answer vector the reference's own tests and doctests hold for this path
(tests/test_oracle_golden.py, transcribed from /root/reference/test/{collide,equilibrium,moments,
pressure,differences,forcing,simulate}.jl).  What those vectors do not pin (bit-level multi-step
trajectories, slippage2!/slippage_ring_riv!, theta fields, fast_93-vs-power_broad rounding) rests on
the Julia evaluation-order semantics documented in SURVEY.md Appendix A: for those cases parity is
"unpinned by a live reference" and DESIGN.md says so.

Conventions.  Arrays are NumPy float64 in **Fortran order** with Julia's shapes -- ``h[Lx,Ly]``,
``f[Lx,Ly,9]`` -- so memory is byte-identical to the Julia arrays (x = first index = contiguous) and
indices are Julia's minus one.  ``circshift(a,(sx,sy))`` == ``np.roll(a,(sx,sy),axis=(0,1))``.
Every expression below keeps Julia's association: ``a*b*c`` folds left, ``x^2`` is ``x*x``, ``-1/3``
is ``(-1)/3``, literal coefficients such as ``6mu`` are single products.  NumPy ufuncs round each
operation separately (no FMA contraction), which is what Julia does on the CPU.
"""
from __future__ import annotations

import math
import numpy as np

# ------------------------------------------------------------------------------------------------
# helpers


def zeros(*shape):
    return np.zeros(shape, dtype=np.float64, order="F")


def ones(*shape):
    return np.ones(shape, dtype=np.float64, order="F")


def circshift(a, shift):
    """Base.circshift(a,(sx,sy)): dest[i,j] = src[i-sx, j-sy] (periodic)."""
    return np.roll(a, shift, axis=(0, 1))


def viewdists(f):
    """src/collide.jl:270-282 -- nine Lx x Ly plane views of f[Lx,Ly,9]."""
    return [f[:, :, k] for k in range(9)]


def viewneighbors(d):
    """src/differences.jl:251-262 -- eight plane views of dgrad[Lx,Ly,8]."""
    return [d[:, :, k] for k in range(8)]


def cospi(x: float) -> float:
    """cos(pi*x) with exact range reduction (Julia's Base.cospi is <1 ulp; so is this).

    The value is computed ONCE on the host and handed as a number to both the oracle and the CUDA
    library, so its last-bit rounding never enters a parity comparison.
    """
    x = abs(float(x))
    x = math.fmod(x, 2.0)  # exact
    if x > 1.0:
        x = 2.0 - x  # exact for doubles in [1,2]
    # now x in [0,1]; reduce around 1/2 so the argument of sin/cos is small and exact
    if x == 0.5:
        return 0.0
    if x <= 0.25:
        return math.cos(math.pi * x)
    if x < 0.75:
        return math.sin(math.pi * (0.5 - x))
    return -math.cos(math.pi * (1.0 - x))


def power_broad(arg, n: int):
    """src/pressure.jl:363-369 -- temp = 1.0; temp *= arg, n times (left fold)."""
    temp = np.ones_like(arg) if isinstance(arg, np.ndarray) else 1.0
    for _ in range(n):
        temp = temp * arg
    return temp


def power_2(arg):  # src/pressure.jl:392-394
    return arg * arg


def power_3(arg):  # src/pressure.jl:401-403
    return arg * arg * arg


def fast_93(arg):  # src/pressure.jl:410-413
    temp = power_3(arg)
    return power_3(temp) - temp


def fast_32(arg):  # src/pressure.jl:420-422
    return power_3(arg) - power_2(arg)


# ------------------------------------------------------------------------------------------------
# operators (array forms)


def equilibrium(feq, height, velocityx, velocityy, vsquare, gravity):
    """src/equilibrium.jl:63-116."""
    f0, f1, f2, f3, f4, f5, f6, f7, f8 = viewdists(feq)
    g0 = 1.5 * gravity
    w1 = 1 / 9
    w5 = 1 / 36
    vsquare[...] = velocityx * velocityx + velocityy * velocityy
    f0[...] = height * (1 - 5 / 6 * gravity * height - 2 / 3 * vsquare)
    f1[...] = w1 * height * (g0 * height + 3 * velocityx + 4.5 * (velocityx * velocityx) - 1.5 * vsquare)
    f2[...] = w1 * height * (g0 * height + 3 * velocityy + 4.5 * (velocityy * velocityy) - 1.5 * vsquare)
    f3[...] = w1 * height * (g0 * height - 3 * velocityx + 4.5 * (velocityx * velocityx) - 1.5 * vsquare)
    f4[...] = w1 * height * (g0 * height - 3 * velocityy + 4.5 * (velocityy * velocityy) - 1.5 * vsquare)
    s = velocityx + velocityy
    f5[...] = w5 * height * (g0 * height + 3 * s + 4.5 * (s * s) - 1.5 * vsquare)
    d = velocityy - velocityx
    f6[...] = w5 * height * (g0 * height + 3 * d + 4.5 * (d * d) - 1.5 * vsquare)
    f7[...] = w5 * height * (g0 * height - 3 * s + 4.5 * (s * s) - 1.5 * vsquare)
    e = velocityx - velocityy
    f8[...] = w5 * height * (g0 * height + 3 * e + 4.5 * (e * e) - 1.5 * vsquare)


def BGKandStream(fout, feq, ftemp, Fx, Fy, tau):
    """src/collide.jl:70-105.  After the call fout == ftemp == streamed post-collision populations."""
    fe = viewdists(feq)
    ft = viewdists(ftemp)
    fo = viewdists(fout)
    omeg = 1 - 1 / tau
    it = 1 / tau
    fo[0][...] = omeg * ft[0] + it * fe[0]
    fo[1][...] = omeg * ft[1] + it * fe[1] + 1 / 3 * Fx
    fo[2][...] = omeg * ft[2] + it * fe[2] + 1 / 3 * Fy
    fo[3][...] = omeg * ft[3] + it * fe[3] - 1 / 3 * Fx
    fo[4][...] = omeg * ft[4] + it * fe[4] - 1 / 3 * Fy
    fo[5][...] = omeg * ft[5] + it * fe[5] + 1 / 24 * (Fx + Fy)
    fo[6][...] = omeg * ft[6] + it * fe[6] + 1 / 24 * (Fy - Fx)
    fo[7][...] = omeg * ft[7] + it * fe[7] - 1 / 24 * (Fx + Fy)
    fo[8][...] = omeg * ft[8] + it * fe[8] + 1 / 24 * (Fx - Fy)
    shifts = [(0, 0), (1, 0), (0, 1), (-1, 0), (0, -1), (1, 1), (-1, 1), (-1, -1), (1, -1)]
    for k in range(9):
        ft[k][...] = circshift(fo[k], shifts[k])
    fout[...] = ftemp


def moments(height, velx, vely, fout):
    """src/moments.jl:43-52.  sum! accumulates planes k=1..9 in order onto a zero-initialised height."""
    f0, f1, f2, f3, f4, f5, f6, f7, f8 = viewdists(fout)
    acc = np.zeros_like(height)
    for k in range(9):
        acc = acc + fout[:, :, k]
    height[...] = acc
    with np.errstate(divide="ignore", invalid="ignore"):
        velx[...] = (f1 - f3 + f5 - f6 - f7 + f8) / height
        vely[...] = (f2 - f4 + f5 + f6 - f7 - f8) / height


def _shift8(dgrad, f):
    """The eight circshift! calls shared by filmpressure!/h∇p!/∇f! (src/pressure.jl:131-139)."""
    hip, hjp, him, hjm, hipjp, himjp, himjm, hipjm = viewneighbors(dgrad)
    hip[...] = circshift(f, (1, 0))
    hjp[...] = circshift(f, (0, 1))
    him[...] = circshift(f, (-1, 0))
    hjm[...] = circshift(f, (0, -1))
    hipjp[...] = circshift(f, (1, 1))
    himjp[...] = circshift(f, (-1, 1))
    himjm[...] = circshift(f, (-1, -1))
    hipjm[...] = circshift(f, (1, -1))
    return hip, hjp, him, hjm, hipjp, himjp, himjm, hipjm


def filmpressure(output, f, dgrad, gamma, cospi_theta, n, m, hmin, hcrit, variant="fast"):
    """Film pressure.  ``variant='fast'``: array form src/pressure.jl:72-115 (fast_93/fast_32, DomainError
    -> ValueError otherwise).  ``variant='power_broad'``: state form src/pressure.jl:119-155.
    ``cospi_theta`` is cospi(theta), scalar or an Lx x Ly field, evaluated by the caller."""
    hip, hjp, him, hjm, hipjp, himjp, himjm, hipjm = _shift8(dgrad, f)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        x = hmin / (f + hcrit)
        if variant == "fast":
            if n == 9 and m == 3:
                powers = fast_93(x)
            elif n == 3 and m == 2:
                powers = fast_32(x)
            else:
                raise ValueError(f"DomainError({(n, m)}): exponents not supported by the array form")
        elif variant == "power_broad":
            powers = power_broad(x, n) - power_broad(x, m)
        else:
            raise ValueError(variant)
        output[...] = -gamma * ((1 - cospi_theta) * (n - 1) * (m - 1) / ((n - m) * hmin) * powers)
        output[...] = output - gamma * (
            2 / 3 * (hjp + hip + him + hjm) + 1 / 6 * (hipjp + himjp + himjm + hipjm) - 10 / 3 * f
        )


def lap9(output, f, gamma):
    """∇²f!  src/differences.jl:57-75."""
    d = zeros(*f.shape, 8)
    hip, hjp, him, hjm, hipjp, himjp, himjm, hipjm = _shift8(d, f)
    output[...] = gamma * (
        2 / 3 * (hjp + hip + him + hjm) + 1 / 6 * (hipjp + himjp + himjm + hipjm) - 10 / 3 * f
    )


def grad9(outputx, outputy, f, a=None, dgrad=None):
    """∇f! 3-, 4- and 5-argument forms  src/differences.jl:153-206 (a=None -> no multiplier)."""
    d = dgrad if dgrad is not None else zeros(*f.shape, 8)
    fip, fjp, fim, fjm, fipjp, fimjp, fimjm, fipjm = _shift8(d, f)
    gx = -1 / 3 * (fip - fim) - 1 / 12 * (fipjp - fimjp - fimjm + fipjm)
    gy = -1 / 3 * (fjp - fjm) - 1 / 12 * (fipjp + fimjp - fimjm - fipjm)
    if a is None:
        outputx[...] = gx
        outputy[...] = gy
    else:
        outputx[...] = a * gx
        outputy[...] = a * gy


def hgradp(hgpx, hgpy, pressure, height, dgrad):
    """h∇p!  src/forcing.jl:168-187 (identical arithmetic to the 5-arg ∇f!)."""
    fip, fjp, fim, fjm, fipjp, fimjp, fimjm, fipjm = _shift8(dgrad, pressure)
    hgpx[...] = height * (-1 / 3 * (fip - fim) - 1 / 12 * (fipjp - fimjp - fimjm + fipjm))
    hgpy[...] = height * (-1 / 3 * (fjp - fjm) - 1 / 12 * (fipjp + fimjp - fimjm - fipjm))


def slippage(slipx, slipy, height, velx, vely, delta, mu, hcrit=0.0, variant=0):
    """variant 0: slippage! src/forcing.jl:42-46; 1: slippage2! :85-99; 2: slippage_ring_riv! :107-111."""
    with np.errstate(divide="ignore", invalid="ignore"):
        if variant == 0:
            den = 2 * (height * height) + 6 * delta * height + 3 * (delta * delta)
            slipx[...] = (6 * mu * height * velx) / den
            slipy[...] = (6 * mu * height * vely) / den
        elif variant == 1:
            hh = height + hcrit
            den = 2 * (hh * hh) + 6 * delta * hh + 3 * (delta * delta)
            slipx[...] = (6 * mu * hh * velx) / den
            slipy[...] = (6 * mu * hh * vely) / den
        elif variant == 2:
            den = 2 * (height * height) + 6 * delta * (height + hcrit)
            slipx[...] = (6 * mu * height * velx) / den
            slipy[...] = (6 * mu * height * vely) / den
        else:
            raise ValueError(variant)


def thermal_amplitude(height, kbt, mu, delta):
    """Deterministic part of thermal!  src/forcing.jl:300-304: sqrt(2 kbt mu 6 h / (2hh + 6hδ + 3δδ))."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.sqrt(2 * kbt * mu * 6 * height / (2 * height * height + 6 * height * delta + 3 * delta * delta))


def thermal(kx, ky, height, kbt, mu, delta, normal_x, normal_y):
    """thermal!  src/forcing.jl:297-311 with the N(0,1) draws supplied by the caller (Julia's randn!
    stream is not reproducible outside Julia; thermal configs are compared statistically)."""
    amp = thermal_amplitude(height, kbt, mu, delta)
    kx[...] = normal_x * amp
    ky[...] = normal_y * amp


def force_sum(Fx, Fy, hgpx, hgpy, slipx, slipy, kbtx=None, kbty=None):
    """The inline 'update!'  src/simulate.jl:18-19 (+ thermal: scripts/Rivulet_stability.jl:123-124)."""
    if kbtx is None:
        Fx[...] = -hgpx - slipx
        Fy[...] = -hgpy - slipy
    else:
        Fx[...] = -hgpx - slipx - kbtx
        Fy[...] = -hgpy - slipy - kbty


def inclination(Fx, Fy, height, alpha, factor):
    """inclination!  src/forcing.jl:363-368.  ``factor`` = 0.5 + 0.5*tanh((t-tstart)/tsmooth), host-evaluated."""
    Fx[...] = Fx + height * alpha[0] * factor
    Fy[...] = Fy + height * alpha[1] * factor


# ------------------------------------------------------------------------------------------------
# state + drivers


class Params:
    """Taumucs  src/initialize.jl:43-60 (same defaults, including mu = cs^2*(tau-0.5))."""

    def __init__(self, Tmax=1000, tdump=None, tau=1.0, cs=None, mu=None, delta=1.0, kbt=0.0, gamma=0.01,
                 n=9, m=3, hmin=0.1, hcrit=0.05, theta=1 / 9, g=0.0):
        self.Tmax = int(Tmax)
        self.tdump = int(tdump) if tdump is not None else self.Tmax // 10
        self.tau = float(tau)
        self.cs = float(cs) if cs is not None else 1 / math.sqrt(3.0)
        self.mu = float(mu) if mu is not None else self.cs * self.cs * (self.tau - 0.5)
        self.delta, self.kbt, self.gamma = float(delta), float(kbt), float(gamma)
        self.n, self.m = int(n), int(m)
        self.hmin, self.hcrit, self.theta, self.g = float(hmin), float(hcrit), float(theta), float(g)


class State:
    """State  src/initialize.jl:149-168 as allocated by Sys  :491-508 (height=1, everything else 0)."""

    def __init__(self, Lx, Ly, thermal=False):
        self.Lx, self.Ly = Lx, Ly
        self.fout, self.ftemp, self.feq = zeros(Lx, Ly, 9), zeros(Lx, Ly, 9), zeros(Lx, Ly, 9)
        self.height = ones(Lx, Ly)
        for name in ("velx", "vely", "vsq", "pressure", "Fx", "Fy", "slipx", "slipy", "hgradpx", "hgradpy"):
            setattr(self, name, zeros(Lx, Ly))
        self.dgrad = zeros(Lx, Ly, 8)
        if thermal:
            self.kbtx, self.kbty = zeros(Lx, Ly), zeros(Lx, Ly)


def step(st: State, p: Params, cospi_theta=None, pvariant="power_broad", slip_variant=0,
         normals=None, incl=None):
    """One iteration of time_loop  src/simulate.jl:15-22 (state-form operators)."""
    ct = cospi(p.theta) if cospi_theta is None else cospi_theta
    filmpressure(st.pressure, st.height, st.dgrad, p.gamma, ct, p.n, p.m, p.hmin, p.hcrit, variant=pvariant)
    hgradp(st.hgradpx, st.hgradpy, st.pressure, st.height, st.dgrad)
    slippage(st.slipx, st.slipy, st.height, st.velx, st.vely, p.delta, p.mu, p.hcrit, slip_variant)
    if normals is not None:
        thermal(st.kbtx, st.kbty, st.height, p.kbt, p.mu, p.delta, normals[0], normals[1])
        force_sum(st.Fx, st.Fy, st.hgradpx, st.hgradpy, st.slipx, st.slipy, st.kbtx, st.kbty)
    else:
        force_sum(st.Fx, st.Fy, st.hgradpx, st.hgradpy, st.slipx, st.slipy)
    if incl is not None:
        inclination(st.Fx, st.Fy, st.height, incl[0], incl[1])
    equilibrium(st.feq, st.height, st.velx, st.vely, st.vsq, p.g)
    BGKandStream(st.fout, st.feq, st.ftemp, st.Fx, st.Fy, p.tau)
    moments(st.height, st.velx, st.vely, st.fout)


def time_loop(st: State, p: Params, nsteps=None, **kw):
    """time_loop  src/simulate.jl:6-25; returns the per-step max-min log of the Δh variant (:56)."""
    dh = []
    for _ in range(p.Tmax if nsteps is None else nsteps):
        dh.append(float(st.height.max() - st.height.min()))
        step(st, p, **kw)
    return dh


# initial conditions used by the configs (host-side, one-off) ----------------------------------


def rayleightaylor_ic(Lx, Ly, kx=15, ky=18, h0=1.0, eps=0.001):
    """src/simulate.jl:350-353 (divides by Lx-1 / Ly-1; replicate, do not fix)."""
    i = np.arange(1, Lx + 1, dtype=np.float64)[:, None]
    j = np.arange(1, Ly + 1, dtype=np.float64)[None, :]
    h = h0 * (1 + eps * np.sin(2 * np.pi * kx * i / (Lx - 1)) * np.sin(2 * np.pi * ky * j / (Ly - 1)))
    return np.asfortranarray(h)


def singledroplet(Lx, Ly, radius, theta, center):
    """src/initialvalues.jl:203-224 (precursor 0.05 hard-coded)."""
    i = np.arange(1, Lx + 1, dtype=np.float64)[:, None]
    j = np.arange(1, Ly + 1, dtype=np.float64)[None, :]
    circ = np.sqrt((i - center[0]) ** 2 + (j - center[1]) ** 2)
    inside = circ <= radius
    with np.errstate(invalid="ignore"):
        cap = (np.cos(np.arcsin(np.where(inside, circ / radius, 0.0))) - cospi(theta)) * radius
    h = np.where(inside, cap, 0.05)
    h = np.where(h < 0, 0.05, h)
    return np.asfortranarray(h)


def randinterface(Lx, Ly, h0, eps, rng):
    """src/initialvalues.jl:23-33 with a NumPy generator in place of Julia's unseeded randn!."""
    return np.asfortranarray(h0 * (1.0 + eps * rng.standard_normal((Lx, Ly))))


def torus(lx, ly, r1, R2, theta, center, hmin=0.05):
    """src/initialvalues.jl:144-168 (noise = 0)."""
    i = np.arange(1, lx + 1, dtype=np.float64)[:, None]
    j = np.arange(1, ly + 1, dtype=np.float64)[None, :]
    coord = np.sqrt((i - center[0]) ** 2 + (j - center[1]) ** 2)
    half = r1 ** 2 - (coord - R2) ** 2
    h = np.where(half <= 0.0, hmin, np.sqrt(np.where(half <= 0.0, 0.0, half)))
    corr = h - r1 * cospi(theta)
    return np.asfortranarray(np.where(corr < hmin, hmin, corr))


def rivulet(Lx, Ly, radius, theta, orientation, center, hmin=0.05):
    """src/initialvalues.jl:69-104 (noise = 0); orientation "y" -> profile in i, "x" -> profile in j."""
    i = np.arange(1, Lx + 1, dtype=np.float64)[:, None] + np.zeros((1, Ly))
    j = np.arange(1, Ly + 1, dtype=np.float64)[None, :] + np.zeros((Lx, 1))
    circ = np.sqrt(((i if orientation == "y" else j) - center) ** 2)
    inside = circ <= radius
    cap = (np.cos(np.arcsin(np.where(inside, circ / radius, 0.0))) - cospi(theta)) * radius
    h = np.where(inside, cap, hmin)
    return np.asfortranarray(np.where(h <= hmin, hmin, h))

