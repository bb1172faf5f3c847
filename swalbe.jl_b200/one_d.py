"""The 1-D (D1Q3) family of Swalbe.jl on the device: `SysConst_1D`, `State_1D` and the operators /
drivers that take them, mirrored with the reference's names (src/initialize.jl:83-98, 587-598; src/simulate.jl:98-157,
247-256, 296-304).  Upstream the 1-D family is CPU-only -- `Sys(sysc::Consts_1D)` has no device argument -- so the
device state built here (`CuState_1D`) has no upstream twin; everything else (field names, argument order, defaults,
DomainError) follows the reference.  A 1-D `Field` has shape (L,) or (L, 3): three contiguous columns, like Julia's.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class SysConst_1D:
    """Base.@kwdef struct SysConst_1D  src/initialize.jl:94-98."""

    def __init__(self, L=256, param=None):
        if param is None:
            raise TypeError("SysConst_1D: keyword argument param not assigned")
        self.L, self.param = int(L), param


class SysConstWithBound_1D:
    """Base.@kwdef struct SysConstWithBound_1D  src/initialize.jl:112-120: lattice with wall nodes `obs` (1 = solid);
    `interior` and `border = [obsright, obsleft]` are filled by obslist! (host NumPy; device copies travel with them)."""

    def __init__(self, L=256, param=None, obs=None):
        if param is None:
            raise TypeError("SysConstWithBound_1D: keyword argument param not assigned")
        self.L, self.param = int(L), param
        self.obs = np.zeros(self.L) if obs is None else np.asarray(obs, dtype=np.float64)
        self.interior = np.zeros(self.L)
        self.border = [np.zeros(self.L), np.zeros(self.L)]
        self._border_dev = None


def obslist1D(obs, verbose=False):
    """obslist1D(obs)   src/obstacle.jl:6-35 -> interior, [obsright, obsleft]  (host logic, like upstream)"""
    if verbose:
        print("WARNING: Always make your obstacles four nodes thick or the algorithm will crash.")
    obs = np.asarray(obs)
    on = obs == 1
    left, right = np.roll(on, 1), np.roll(on, -1)  # obs[i-1], obs[i+1] (periodic, mod1)
    return (on & right & left).astype(np.float64), [(on & right).astype(np.float64), (on & left).astype(np.float64)]


def obslist(sys_: SysConstWithBound_1D, verbose=False):
    """obslist!(sys::SysConstWithBound_1D)   src/obstacle.jl:41-53"""
    from . import Field

    interior, border = obslist1D(sys_.obs, verbose=verbose)
    sys_.interior[...] = interior
    sys_.border[0][...] = border[0]
    sys_.border[1][...] = border[1]
    sys_._border_dev = (Field(sys_.L).set(sys_.border[0]), Field(sys_.L).set(sys_.border[1]), Field(sys_.L).set(sys_.interior))


class LBM_state_1D:
    """abstract type LBM_state_1D  src/initialize.jl:8 (marker base of every 1-D device state)"""


class CuState_1D(LBM_state_1D):
    """State_1D  src/initialize.jl:587-598 on the device (height = 1, everything else 0)."""

    def __init__(self, L):
        from . import Field

        self.L = L
        self.fout, self.ftemp, self.feq = Field(L, 3), Field(L, 3), Field(L, 3)
        self.height = Field(L, fill=1.0)
        self.vel, self.pressure, self.F, self.slip, self.hgradp = Field(L), Field(L), Field(L), Field(L), Field(L)
        self.dgrad = Field(L, 2)

    def __getattr__(self, name):  # Julia field spelling
        if name == "h∇p":
            return self.hgradp
        raise AttributeError(name)

    def _c_state(self, extra=None):
        s = _lib.CState1D()
        for name, _ in _lib.CState1D._fields_:
            f = self.__dict__.get(name)
            if f is None and extra is not None:
                f = extra.__dict__.get({"gamma": "γ", "dgamma": "dγ"}.get(name, name))
            setattr(s, name, f.ptr if f is not None else None)
        return s


class Expanded_1D(LBM_state_1D):
    """abstract type Expanded_1D  src/initialize.jl:10: a `basestate::State_1D` plus the fields of the kind"""

    def __init__(self, L):
        self.L, self.basestate = L, CuState_1D(L)

    def _c_state(self):
        return self.basestate._c_state(extra=self)


class CuState_thermal_1D(Expanded_1D):
    """State_thermal_1D  src/initialize.jl:338-341"""

    def __init__(self, L):
        from . import Field

        super().__init__(L)
        self.kbt = Field(L)


class CuState_gamma_1D(Expanded_1D):
    """State_gamma_1D  src/initialize.jl:304-308 (γ, ∇γ; the Julia field name ∇γ is also reachable as `dγ`)"""

    def __init__(self, L):
        from . import Field

        super().__init__(L)
        self.γ, self.dγ = Field(L), Field(L)

    def __getattr__(self, name):
        if name == "∇γ":
            return self.dγ
        raise AttributeError(name)


class CuStateWithBound_1D(CuState_gamma_1D):
    """StateWithBound_1D  src/initialize.jl:317-324 (γ, ∇γ, fbound)"""

    def __init__(self, L):
        from . import Field

        super().__init__(L)
        self.fbound = Field(L, 3)


def Sys_1D(sysc, kind="simple"):
    """Sys(sysc::Consts_1D; T, kind)   src/initialize.jl:587-616"""
    if kind == "simple":
        return CuState_1D(sysc.L)
    if kind == "thermal":
        return CuState_thermal_1D(sysc.L)
    if kind == "gamma":
        return CuState_gamma_1D(sysc.L)
    if kind == "gamma_bound":
        return CuStateWithBound_1D(sysc.L)
    return None  # (the reference leaves `dyn` undefined for an unknown kind)


def base(st):
    """the State_1D of any 1-D state (Expanded_1D methods forward to state.basestate, e.g. src/collide.jl:205-211)"""
    return st.basestate if isinstance(st, Expanded_1D) else st


def _stream():
    from . import _stream as s

    return s()


def _theta(θ):
    from . import Field, _theta_args

    if isinstance(θ, Field):
        return _theta_args(θ)
    from . import cospi

    return cospi(θ), None


def equilibrium(feq, height, vel, g):
    """equilibrium!(feq, height, velocity, gravity)   src/equilibrium.jl:169-181"""
    _lib.call("swalbe_equilibrium_d1q3", feq.ptr, height.ptr, vel.ptr, float(g), height.shape[0], _stream())


def BGKandStream(fout, feq, ftemp, F, τ):
    """BGKandStream!(fout, feq, ftemp, F::Vector, τ)   src/collide.jl:179-201"""
    _lib.call("swalbe_bgk_stream_d1q3", fout.ptr, feq.ptr, ftemp.ptr, F.ptr, float(τ), F.shape[0], _stream())


def moments(height, vel, fout):
    """moments!(height::Vector, vel, fout)   src/moments.jl:54-62"""
    _lib.call("swalbe_moments_d1q3", height.ptr, vel.ptr, fout.ptr, height.shape[0], _stream())


def filmpressure(output, f, dgrad, γ, θ, n, m, hmin, hcrit, variant=_lib.PRESSURE_FAST):
    """filmpressure!(output::Vector, f, dgrad, γ, θ, n, m, hmin, hcrit)   src/pressure.jl:196-227"""
    ct, ctf = _theta(θ)
    _lib.call("swalbe_filmpressure_1d", output.ptr, f.ptr, dgrad.ptr if dgrad is not None else None, float(γ), ct, ctf, int(n),
              int(m), float(hmin), float(hcrit), variant, f.shape[0], _stream())


def gradf(output, f, dgrad=None, a=None):
    """∇f!(output::Vector, f, dgrad, a) | ∇f!(output, f::Vector, dgrad)   src/differences.jl:208-230"""
    _lib.call("swalbe_grad_1d", output.ptr, f.ptr, a.ptr if a is not None else None, f.shape[0], _stream())


def laplacianf(output, f, dgrad=None):
    """∇²f!(output, f::Vector, dgrad)   src/differences.jl:77-85"""
    _lib.call("swalbe_lap_1d", output.ptr, f.ptr, f.shape[0], _stream())


def slippage(slip, height, vel, δ, μ):
    """slippage!(slip, height, vel, δ, μ)   src/forcing.jl:68-71"""
    _lib.call("swalbe_slippage_1d", slip.ptr, height.ptr, vel.ptr, float(δ), float(μ), height.shape[0], _stream())


def hgradp(st):
    """h∇p!(state::LBM_state_1D)   src/forcing.jl:189-198 | h∇p!(state::Expanded_1D)  :223-232"""
    st = base(st)
    gradf(st.hgradp, st.pressure, st.dgrad, st.height)


def update(st, extra=None):
    """state.F .= -state.h∇p .- state.slip   src/simulate.jl:110;  with `extra`: ... .- extra  (run_gamma :544: ∇γ;
    thermal loops: kbt).  For a State_gamma_1D / State_thermal_1D the third term defaults to the state's own ∇γ / kbt."""
    b = base(st)
    if extra is None and isinstance(st, CuState_gamma_1D):
        extra = st.dγ
    if extra is None and isinstance(st, CuState_thermal_1D):
        extra = st.kbt
    if extra is None:
        _lib.call("swalbe_force_sum_1d", b.F.ptr, b.hgradp.ptr, b.slip.ptr, b.L, _stream())
    else:
        _lib.call("swalbe_force_sum3_1d", b.F.ptr, b.hgradp.ptr, b.slip.ptr, extra.ptr, b.L, _stream())


def thermal(fluc, height, kbt, μ, δ, seed, step):
    """thermal!(fluc, height, kᵦT, μ, δ)   src/forcing.jl:322-333"""
    _lib.call("swalbe_thermal_1d", fluc.ptr, height.ptr, float(kbt), float(μ), float(δ), int(seed), int(step), height.shape[0],
              _stream())


def inclination(α, st, t=1000, tstart=0, tsmooth=1):
    """inclination!(α::Float64, state::State_1D | Expanded_1D; t, tstart, tsmooth)   src/forcing.jl:379-389"""
    import math

    b = base(st)
    _lib.call("swalbe_inclination_1d", b.F.ptr, b.height.ptr, float(α), 0.5 + 0.5 * math.tanh((t - tstart) / tsmooth), b.L,
              _stream())


def gradgamma(st, sys_=None):
    """∇γ!(state)  src/forcing.jl:423-432 | ∇γ!(state, sys)  :434-447"""
    b = base(st)
    _lib.call("swalbe_gradgamma_1d", st.dγ.ptr, st.γ.ptr, b.height.ptr if sys_ is not None else None,
              float(sys_.param.delta) if sys_ is not None else 0.0, b.L, _stream())


def filmpressure_expanded(st, sys_, θ=None, n=None, m=None, hmin=None, hcrit=None, γ=None):
    """filmpressure!(state::Expanded_1D, sys; ...)  src/pressure.jl:258-282 | (state::State_gamma_1D, sys; ...)  :284-315
    (γ: scalar or a length-L Field, as run_gamma passes it; the gamma kind also parks the two contributions in ftemp)"""
    from . import Field

    p, b = sys_.param, base(st)
    ct, ctf = _theta(p.theta if θ is None else θ)
    γ = p.gamma if γ is None else γ
    gfield = γ.ptr if isinstance(γ, Field) else None
    park = type(st) is CuState_gamma_1D  # (StateWithBound_1D is a Boundary_1D <: Expanded_1D: the generic method)
    _lib.call("swalbe_filmpressure_gamma_1d", b.pressure.ptr, b.height.ptr, 0.0 if gfield else float(γ), gfield, None, 0.0, ct,
              ctf, int(p.n if n is None else n), int(p.m if m is None else m), float(p.hmin if hmin is None else hmin),
              float(p.hcrit if hcrit is None else hcrit), b.ftemp.ptr if park else None, b.L, _stream())


def filmpressure_rho(output, f, dgrad, rho, γ, θ, n, m, hmin, hcrit, Gamma=0.0):
    """filmpressure!(output::Vector, f, dgrad, rho, γ, θ, n, m, hmin, hcrit; Gamma)   src/pressure.jl:318-338"""
    ct, ctf = _theta(θ)
    _lib.call("swalbe_filmpressure_gamma_1d", output.ptr, f.ptr, float(γ), None, rho.ptr, float(Gamma), ct, ctf, int(n), int(m),
              float(hmin), float(hcrit), None, f.shape[0], _stream())


def BGKandStream_bound(st: "CuStateWithBound_1D", sys_: SysConstWithBound_1D):
    """BGKandStream!(state::StateWithBound_1D, sys::SysConstWithBound_1D)   src/collide.jl:214-249"""
    if sys_._border_dev is None:
        obslist(sys_)
    b = st.basestate
    _lib.call("swalbe_bgk_stream_bound_d1q3", b.fout.ptr, b.feq.ptr, b.ftemp.ptr, st.fbound.ptr, b.F.ptr,
              sys_._border_dev[0].ptr, sys_._border_dev[1].ptr, float(sys_.param.tau), b.L, _stream())


def update_rho(rho, rho_int, height, dgrad, differentials, D=1.0, M=0.0):
    """update_rho!(rho, rho_int, height, dgrad, differentials; D, M)   src/forcing.jl:399-417"""
    _lib.call("swalbe_update_rho_1d", rho.ptr, rho_int.ptr, height.ptr, differentials.ptr, float(D), float(M), height.shape[0],
              _stream())


def fused_steps(st, sys_: SysConst_1D, nsteps: int, θ=None, log_minmax=False, skip_aux=False,
                pressure_variant=_lib.PRESSURE_POWER_BROAD, incl=None, gamma_field=False, marangoni=False):
    """nsteps iterations of the loop body src/simulate.jl:107-114 through swalbe_time_loop_1d.
    incl = (α, factor): the inclination! callback slot (:159-179); gamma_field / marangoni (State_gamma_1D): the loop body
    of run_gamma (:541-547) -- pressure with the per-site tension state.γ, F = -h∇p - slip - state.∇γ."""
    import torch

    from . import _c_params

    q = _c_params(sys_.param, θ, pressure_variant=pressure_variant)
    if incl is not None:
        q.use_inclination, q.incl_ax, q.incl_factor = 1, float(incl[0]), float(incl[1])
    cs = st._c_state()
    skip_aux = (_lib.LOOP_SKIP_AUX if skip_aux else 0) | (_lib.LOOP_GAMMA_FIELD if gamma_field else 0) | \
               (_lib.LOOP_MARANGONI if marangoni else 0)
    st = base(st)
    logs = _lib.CLogs()
    mn = mx = None
    if log_minmax:
        mn = torch.empty(nsteps, dtype=torch.float64, device="cuda")
        mx = torch.empty(nsteps, dtype=torch.float64, device="cuda")
        logs.hmin, logs.hmax = mn.data_ptr(), mx.data_ptr()
    _lib.call("swalbe_time_loop_1d", C.byref(cs), C.byref(q), st.L, int(nsteps), skip_aux,
              C.byref(logs) if log_minmax else None, _stream())
    return mn, mx


def time_loop(sys_: SysConst_1D, st: CuState_1D, *extra, verbose=False):
    """time_loop(sys::SysConst_1D, state::State_1D)         src/simulate.jl:98-116
    time_loop(sys, state, θ)                               :118-136  (θ scalar or length-L Field)
    time_loop(sys, state, Δh::list)                        :138-157  (max - min before every step)
    Same mass read-back / print at t % tdump == 0 as the reference, the steps in between inside one launch."""
    import math

    from . import SwalbeError, field_stats
    from . import inclination as inclination2d

    p = sys_.param
    if isinstance(sys_, SysConstWithBound_1D):
        return _time_loop_bound(sys_, st, verbose)
    θ = dh = incl = measure = None
    if len(extra) == 1 and isinstance(extra[0], list):
        dh = extra[0]
    elif len(extra) == 1:
        θ = extra[0]
    elif len(extra) == 2:  # time_loop(sys, state, f::Function, measure)  :159-179 -- f(measure, state) between F and equilibrium!
        f, measure = extra
        if f not in (inclination, inclination2d):
            raise SwalbeError("time_loop (1-D): only the Swalbe.inclination! callback is fused on the device")
        incl = (measure, 0.5 + 0.5 * math.tanh((1000 - 0) / 1))  # keyword defaults of src/forcing.jl:379
    elif extra:
        raise TypeError("MethodError: no method matching time_loop with these arguments")
    t, tdump = 1, max(1, p.tdump)
    while t <= p.Tmax:
        if t % tdump == 0:
            mass = field_stats(st.height)[2]
            if verbose:
                print(f"Time step {t} mass is {round(mass, 3)}")
        nxt = min(p.Tmax + 1, (t // tdump + 1) * tdump)
        mn, mx = fused_steps(st, sys_, nxt - t, θ=θ, log_minmax=dh is not None, skip_aux=nxt <= p.Tmax, incl=incl)
        if dh is not None:
            dh.extend((mx - mn).cpu().tolist())
        t = nxt
    return st if measure is None else (st, measure)


def _time_loop_bound(sys_: SysConstWithBound_1D, st: "CuStateWithBound_1D", verbose):
    """time_loop(sys::SysConstWithBound_1D, state::StateWithBound_1D)   src/simulate.jl:181-204: operator by operator (the
    bounce-back collision is its own kernel); the mass is that of the fluid nodes, sum(h - interior*h)"""
    from . import equilibrium as eq
    from . import moments as mom
    from . import slippage as slip

    if sys_._border_dev is None:
        obslist(sys_)
    p, b = sys_.param, st.basestate
    for t in range(1, p.Tmax + 1):
        if t % max(1, p.tdump) == 0:
            h = b.height.t
            mass = float((h - sys_._border_dev[2].t * h).sum().item())
            if verbose:
                print(f"Time step {t} mass is {round(mass, 3)}")
        filmpressure_expanded(st, sys_)
        hgradp(st)
        slip(b, sys_)
        update(b)
        eq(b, sys_)
        BGKandStream_bound(st, sys_)
        mom(b)
    return st


def singledroplet_1d(L, radius, θ, center, hcrit=0.05):
    """singledroplet(height::Vector, radius, θ, center; hcrit)   src/initialvalues.jl:226-242 (host construction)"""
    from . import cospi

    i = np.arange(1, L + 1, dtype=np.float64)
    circ = np.sqrt((i - center) ** 2)
    inside = circ <= radius
    h = np.where(inside, (np.cos(np.arcsin(np.where(inside, circ / radius, 0.0))) - cospi(θ)) * radius, hcrit)
    return np.where(h <= hcrit, hcrit, h)


def two_droplets(sys_, r1=230, r2=230, θ1=1 / 9, θ2=1 / 9, center=None):
    """two_droplets(sys::Consts_1D; r₁, r₂, θ₁, θ₂, center)   src/initialvalues.jl:277-323"""
    from . import cospi

    L = sys_.L
    center = (L / 3, 2 * L / 3) if center is None else center
    i = np.arange(1, L + 1, dtype=np.float64)

    def cap(r, θ, c):
        circ = np.sqrt((i - c) ** 2)
        inside = circ <= r
        d = np.where(inside, (np.cos(np.arcsin(np.where(inside, circ / r, 0.0))) - cospi(θ)) * r, 0.05)
        return np.where(d < 0, 0.05, d)

    h = cap(r1, θ1, center[0]) + cap(r2, θ2, center[1]) - 0.05
    return np.where(h < 0, 0.05, h)


def run_dropletforced(sys_: SysConst_1D, radius=20, θ0=1 / 6, center=None, f=0.0, verbos=True):
    """run_dropletforced(sys::SysConst_1D; radius, θ₀, center, θₛ, f)   src/simulate.jl:486-503"""
    print("Simulating a sliding droplet in one dimension")
    st = CuState_1D(sys_.L)
    st.height.set(singledroplet_1d(sys_.L, radius, θ0, sys_.L // 2 if center is None else center))
    equilibrium(st.feq, st.height, st.vel, sys_.param.g)
    print("Starting the lattice Boltzmann time loop")
    time_loop(sys_, st, inclination, f, verbose=verbos)
    return st.height, st.vel


def run_gamma(sys_: SysConst_1D, gamma, r1=115, r2=115, θ0=1 / 9, verbos=True, dump=100, fluid=None):
    """run_gamma(sys::SysConst_1D, gamma::Vector; r₁, r₂, θ₀, verbos, dump, fluid)   src/simulate.jl:518-560: coalescence
    of two droplets under a surface-tension field; the loop body (pressure with the per-site tension, F = -h∇p - slip - ∇γ)
    runs fused between two snapshots; `fluid` rows = height every `dump` steps."""
    print("Simulating droplet coalecense with surface tension gardient")
    p, L = sys_.param, sys_.L
    st = CuState_gamma_1D(L)
    b = st.basestate
    b.height.set(two_droplets(sys_, r1=r1, r2=r2, θ1=θ0, θ2=θ0, center=(L / 3, 2 * L / 3)))
    equilibrium(b.feq, b.height, b.vel, p.g)
    st.γ.set(np.asarray(gamma, dtype=np.float64))
    gradgamma(st)
    fluid = np.zeros((p.Tmax // dump, L)) if fluid is None else fluid
    print("Starting the lattice Boltzmann time loop")
    t = 1
    while t <= p.Tmax:  # chunks end at the steps that print (before the update) or snapshot (after it)
        if verbos and t % max(1, p.tdump) == 0:
            lo, hi = L // 2 - 20, L // 2 + 20  # Julia's height[L÷2-20 : L÷2+20], 1-based inclusive
            print(f"Time step {t} bridge height is {round(float(b.height.t[lo - 1:hi].min().item()), 3)}")
        nxt_print = (t // max(1, p.tdump) + 1) * max(1, p.tdump)
        nxt_snap = ((t - 1) // dump + 1) * dump + 1
        nxt = min(p.Tmax + 1, nxt_snap, nxt_print if verbos else p.Tmax + 1)
        fused_steps(st, sys_, nxt - t, skip_aux=nxt <= p.Tmax, gamma_field=True, marangoni=True)
        if (nxt - 1) % dump == 0 and (nxt - 1) // dump >= 1:
            fluid[(nxt - 1) // dump - 1, :] = b.height.numpy()
        t = nxt
    return fluid


def run_flat(sys_: SysConst_1D, verbos=True):
    """run_flat(sys::SysConst_1D)  src/simulate.jl:247-256"""
    print("Simulating a flat interface without driving forces (nothing should happen) in one dimension")
    st = CuState_1D(sys_.L)
    st.height.set(1.0)
    time_loop(sys_, st, verbose=verbos)
    return st.height


def run_random(sys_: SysConst_1D, h0=1.0, ϵ=0.01, verbos=True, rng=None):
    """run_random(sys::SysConst_1D)  src/simulate.jl:296-304"""
    print("Simulating a random undulated interface in one dimension")
    st = CuState_1D(sys_.L)
    rng = rng if rng is not None else np.random.default_rng()
    st.height.set(h0 * (1.0 + ϵ * rng.standard_normal(sys_.L)))
    equilibrium(st.feq, st.height, st.vel, sys_.param.g)
    time_loop(sys_, st, verbose=verbos)
    return st.height
