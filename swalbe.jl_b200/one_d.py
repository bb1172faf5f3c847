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


class CuState_1D:
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

    def _c_state(self):
        s = _lib.CState1D()
        for name, _ in _lib.CState1D._fields_:
            setattr(s, name, getattr(self, name).ptr)
        return s


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


def hgradp(st: CuState_1D):
    """h∇p!(state::LBM_state_1D)   src/forcing.jl:189-198"""
    gradf(st.hgradp, st.pressure, st.dgrad, st.height)


def update(st: CuState_1D):
    """state.F .= -state.h∇p .- state.slip   src/simulate.jl:110"""
    _lib.call("swalbe_force_sum_1d", st.F.ptr, st.hgradp.ptr, st.slip.ptr, st.L, _stream())


def fused_steps(st: CuState_1D, sys_: SysConst_1D, nsteps: int, θ=None, log_minmax=False, skip_aux=False,
                pressure_variant=_lib.PRESSURE_POWER_BROAD):
    """nsteps iterations of the loop body src/simulate.jl:107-114 through swalbe_time_loop_1d."""
    import torch

    from . import _c_params

    q = _c_params(sys_.param, θ, pressure_variant=pressure_variant)
    cs = st._c_state()
    logs = _lib.CLogs()
    mn = mx = None
    if log_minmax:
        mn = torch.empty(nsteps, dtype=torch.float64, device="cuda")
        mx = torch.empty(nsteps, dtype=torch.float64, device="cuda")
        logs.hmin, logs.hmax = mn.data_ptr(), mx.data_ptr()
    _lib.call("swalbe_time_loop_1d", C.byref(cs), C.byref(q), st.L, int(nsteps), _lib.LOOP_SKIP_AUX if skip_aux else 0,
              C.byref(logs) if log_minmax else None, _stream())
    return mn, mx


def time_loop(sys_: SysConst_1D, st: CuState_1D, *extra, verbose=False):
    """time_loop(sys::SysConst_1D, state::State_1D)         src/simulate.jl:98-116
    time_loop(sys, state, θ)                               :118-136  (θ scalar or length-L Field)
    time_loop(sys, state, Δh::list)                        :138-157  (max - min before every step)
    Same mass read-back / print at t % tdump == 0 as the reference, the steps in between inside one launch."""
    from . import field_stats

    p = sys_.param
    θ = dh = None
    if len(extra) == 1 and isinstance(extra[0], list):
        dh = extra[0]
    elif len(extra) == 1:
        θ = extra[0]
    elif extra:
        raise TypeError("MethodError: no method matching time_loop with these arguments")
    t, tdump = 1, max(1, p.tdump)
    while t <= p.Tmax:
        if t % tdump == 0:
            mass = field_stats(st.height)[2]
            if verbose:
                print(f"Time step {t} mass is {round(mass, 3)}")
        nxt = min(p.Tmax + 1, (t // tdump + 1) * tdump)
        mn, mx = fused_steps(st, sys_, nxt - t, θ=θ, log_minmax=dh is not None, skip_aux=nxt <= p.Tmax)
        if dh is not None:
            dh.extend((mx - mn).cpu().tolist())
        t = nxt
    return st


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
