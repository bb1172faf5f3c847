"""swalbe_b200 -- host-side mirror of Swalbe.jl's 2-D operator API on top of libswalbe_b200.so.

Julia is not available in the build image, so the host side that the reference's scripts would run in
Julia is mirrored here in Python with the same names, argument order and error behaviour
(`SysConst`, `Taumucs`, `Sys(sys, "GPU"; kind)`, `equilibrium!`, `BGKandStream!`, `moments!`,
`filmpressure!`, `h∇p!`, `∇f!`, `∇²f!`, `slippage!`, `slippage2!`, `slippage_ring_riv!`, `thermal!`,
`inclination!`, `update!`, `time_loop`, `run_*`).  `!` and `∇` cannot appear in Python identifiers, so
`equilibrium!` is `equilibrium`, `h∇p!` is `hgradp`, `∇f!` is `gradf`, `∇²f!` is `laplacianf`; the
table ``JULIA_NAMES`` maps the exact Julia spellings.  The Julia glue that binds the same C ABI is in
``julia/SwalbeB200.jl``.

PyTorch is used only as the owner of device memory and streams: every field is a CUDA float64 tensor
whose memory is Julia's column-major layout (x contiguous), handed to the C ABI as a raw pointer
together with the current CUDA stream.  No arithmetic of the path happens in PyTorch or on the CPU.
"""
from __future__ import annotations

import ctypes as C
import math
import time

import numpy as np

from . import _lib, one_d
from ._lib import DomainError, SwalbeError  # noqa: F401
from .one_d import (CuState_1D, CuState_gamma_1D, CuState_thermal_1D, CuStateWithBound_1D, SysConst_1D,  # noqa: F401
                    SysConstWithBound_1D)

__all__ = [
    "Taumucs", "SysConst", "Sys_const", "Sys", "CuState", "CuState_thermal", "Swalbe_state", "Field", "cospi",
    "equilibrium", "BGKandStream", "moments", "filmpressure", "hgradp", "gradf", "laplacianf", "slippage",
    "slippage2", "slippage_ring_riv", "thermal", "inclination", "update", "time_loop", "run_flat", "run_random",
    "run_rayleightaylor", "run_dropletrelax", "run_dropletpatterned", "run_dropletforced", "wetted", "snapshot",
    "field_stats", "DomainError", "SwalbeError", "JULIA_NAMES", "viewdists", "viewneighbors", "power_broad", "power_2",
    "power_3", "fast_93", "fast_32", "fused_steps", "singledroplet", "cospi_field", "torus", "rivulet", "sinewave2d",
    "randinterface", "circshift", "move_substrate", "restart_from_height", "save_heights", "dump_height_slab",
    "load_height_slab", "SnapshotBuffer", "SysConst_1D", "CuState_1D", "one_d",
]


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise SwalbeError("swalbe_b200 needs a CUDA device: the B200 path has no CPU fallback")
    return torch


def _stream() -> C.c_void_p:
    return C.c_void_p(_torch().cuda.current_stream().cuda_stream)


def cospi(x: float) -> float:
    """cos(pi x) with exact range reduction (stand-in for Base.cospi; evaluated once on the host)."""
    x = math.fmod(abs(float(x)), 2.0)
    if x > 1.0:
        x = 2.0 - x
    if x == 0.5:
        return 0.0
    if x <= 0.25:
        return math.cos(math.pi * x)
    if x < 0.75:
        return math.sin(math.pi * (0.5 - x))
    return -math.cos(math.pi * (1.0 - x))


# ------------------------------------------------------------------------------------------------
# constants  (src/initialize.jl:43-81)


class Taumucs:
    """Base.@kwdef struct Taumucs  src/initialize.jl:43-60 -- same fields and defaults (μ = cₛ²(τ-½))."""

    def __init__(self, Tmax=1000, tdump=None, τ=None, cs=None, μ=None, δ=None, kbt=0.0, γ=None, n=9, m=3, hmin=0.1,
                 hcrit=0.05, θ=None, g=0.0, *, tau=None, mu=None, delta=None, gamma=None, theta=None):
        # (Julia's `cₛ` NFKC-normalises to `cs` in Python source, so `Taumucs(cₛ=...)` works as written)
        pick = lambda a, b, d: d if (a is None and b is None) else (a if a is not None else b)  # noqa: E731
        self.Tmax = int(Tmax)
        self.tdump = int(tdump) if tdump is not None else self.Tmax // 10
        self.tau = float(pick(τ, tau, 1.0))
        self.cs = float(cs) if cs is not None else 1 / math.sqrt(3.0)
        self.mu = float(pick(μ, mu, self.cs * self.cs * (self.tau - 0.5)))
        self.delta = float(pick(δ, delta, 1.0))
        self.kbt = float(kbt)
        self.gamma = float(pick(γ, gamma, 0.01))
        self.n, self.m = int(n), int(m)
        self.hmin, self.hcrit = float(hmin), float(hcrit)
        self.theta = float(pick(θ, theta, 1 / 9))
        self.g = float(g)

    # Julia spellings
    τ = property(lambda s: s.tau)
    μ = property(lambda s: s.mu)
    δ = property(lambda s: s.delta)
    γ = property(lambda s: s.gamma)
    θ = property(lambda s: s.theta)


class SysConst:
    """Base.@kwdef struct SysConst  src/initialize.jl:76-81."""

    def __init__(self, Lx=256, Ly=256, param: Taumucs | None = None):
        if param is None:
            raise TypeError("SysConst: keyword argument param not assigned")  # @kwdef field without default
        self.Lx, self.Ly, self.param = int(Lx), int(Ly), param


Sys_const = SysConst  # spelling used by BASELINE.json's north_star


# ------------------------------------------------------------------------------------------------
# device arrays


class Field:
    """A device Float64 array in Julia's column-major layout.

    ``shape`` is Julia's (Lx, Ly) or (Lx, Ly, K).  ``t`` is the owning torch tensor, C-contiguous with the
    reversed shape (K, Ly, Lx); ``jl`` is a permuted view indexed [i, j, k] like Julia (0-based).
    """

    def __init__(self, Lx, Ly=None, K=None, fill=0.0):
        torch = _torch()
        if Ly is None:  # a Julia Vector (the 1-D family, swalbe_b200.one_d)
            self.shape, tshape = (Lx,), (Lx,)
        else:
            self.shape = (Lx, Ly) if K is None else (Lx, Ly, K)
            tshape = (Ly, Lx) if K is None else (K, Ly, Lx)
        self.t = torch.full(tshape, float(fill), dtype=torch.float64, device="cuda")
        self.jl = self.t.permute(*reversed(range(self.t.dim())))
        self._gen = 0  # bumped by touch(): writes through the raw pointer that torch's version counter cannot see

    def touch(self) -> "Field":
        self._gen += 1
        return self

    def _stamp(self):
        return (self.t._version, self._gen)

    @property
    def ptr(self) -> C.c_void_p:
        return C.c_void_p(self.t.data_ptr())

    def numpy(self) -> np.ndarray:
        """Host copy as a Fortran-ordered NumPy array with the Julia shape (== Array(field))."""
        return np.asfortranarray(self.t.cpu().numpy().transpose())

    def set(self, a) -> "Field":
        """field .= a   (a: scalar, NumPy array with the Julia shape, or another Field)."""
        torch = _torch()
        if isinstance(a, Field):
            self.t.copy_(a.t)
        elif np.isscalar(a):
            self.t.fill_(float(a))
        else:
            a = np.asarray(a, dtype=np.float64)
            if a.shape != self.shape:
                raise ValueError(f"DimensionMismatch: {a.shape} vs {self.shape}")
            self.t.copy_(torch.from_numpy(np.ascontiguousarray(a.transpose())))
        return self


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, Field):
        return x.ptr
    raise TypeError(f"expected a swalbe_b200.Field device array, got {type(x).__name__} "
                    "(host arrays are not accepted: there is no CPU path)")


class CuState:
    """CuState  src/initialize.jl:214-233 as built by Sys(sys, "GPU")  :513-535 (height = 1, rest 0)."""

    thermal = False

    def __init__(self, Lx, Ly):
        self.Lx, self.Ly = Lx, Ly
        self.fout, self.ftemp, self.feq = Field(Lx, Ly, 9), Field(Lx, Ly, 9), Field(Lx, Ly, 9)
        self.height = Field(Lx, Ly, fill=1.0)
        self.velx, self.vely, self.vsq, self.pressure = Field(Lx, Ly), Field(Lx, Ly), Field(Lx, Ly), Field(Lx, Ly)
        self.Fx, self.Fy, self.slipx, self.slipy = Field(Lx, Ly), Field(Lx, Ly), Field(Lx, Ly), Field(Lx, Ly)
        self.hgradpx, self.hgradpy = Field(Lx, Ly), Field(Lx, Ly)  # h∇px, h∇py
        self.dgrad = Field(Lx, Ly, 8)
        self._plan = None

    def __getattr__(self, name):  # Julia field spellings
        if name == "h∇px":
            return self.hgradpx
        if name == "h∇py":
            return self.hgradpy
        raise AttributeError(name)

    def _c_state(self) -> _lib.CState:
        s = _lib.CState()
        for name, _ in _lib.CState._fields_:
            f = getattr(self, name, None) if name not in ("kbtx", "kbty") or self.thermal else None
            setattr(s, name, f.ptr if isinstance(f, Field) else None)
        return s

    def plan(self):
        if self._plan is None:
            h = C.c_void_p()
            _lib.call("swalbe_plan_create", C.byref(h), self.Lx, self.Ly)
            self._plan = _Plan(h)
        return self._plan.handle


class CuState_thermal(CuState):
    """CuState_thermal  src/initialize.jl:235-256 (adds kbtx, kbty)."""

    thermal = True

    def __init__(self, Lx, Ly):
        super().__init__(Lx, Ly)
        self.kbtx, self.kbty = Field(Lx, Ly), Field(Lx, Ly)


Swalbe_state = CuState  # spelling used by BASELINE.json's north_star


class _Plan:
    def __init__(self, handle):
        self.handle = handle

    def __del__(self):
        try:
            _lib.load().swalbe_plan_destroy(self.handle)
        except Exception:
            pass


def Sys(sysc: SysConst, device: str = "GPU", *legacy, T=float, kind: str = "simple"):
    """Sys(sysc, device; T, kind)  src/initialize.jl:491-572  -> CuState / CuState_thermal, and the older tuple form
    Sys(sysc, device, exotic::Bool, T)  src/initialize.jl:358-475  -> (fout, ftemp, feq, height, velx, vely, vsq,
    pressure, dgrad, Fx, Fy, slipx, slipy, h∇px, h∇py[, fthermalx, fthermaly]).  Only the "GPU" device string exists."""
    if isinstance(sysc, (SysConst_1D, SysConstWithBound_1D)):  # Sys(sysc::Consts_1D; T, kind)  src/initialize.jl:587-616
        return one_d.Sys_1D(sysc, kind)                         # (no device argument upstream)
    if device != "GPU":
        raise SwalbeError(f'Sys(sys, "{device}"): swalbe_b200 implements the "GPU" path only (no CPU fallback)')
    if legacy:
        if len(legacy) != 2 or not isinstance(legacy[0], bool):
            raise TypeError("MethodError: no method matching Sys(::SysConst, ::String, ...) with these arguments")
        exotic, T = legacy
    if T not in (float, np.float64, "Float64"):
        raise SwalbeError("swalbe_b200 is Float64 only")
    if legacy:
        st = CuState_thermal(sysc.Lx, sysc.Ly) if exotic else CuState(sysc.Lx, sysc.Ly)
        fields = (st.fout, st.ftemp, st.feq, st.height, st.velx, st.vely, st.vsq, st.pressure, st.dgrad, st.Fx, st.Fy,
                  st.slipx, st.slipy, st.hgradpx, st.hgradpy)
        return fields + ((st.kbtx, st.kbty) if exotic else ())
    if kind == "simple":
        return CuState(sysc.Lx, sysc.Ly)
    if kind == "thermal":
        return CuState_thermal(sysc.Lx, sysc.Ly)
    return None  # the reference falls through and returns nothing for unknown kinds


# ------------------------------------------------------------------------------------------------
# operators: array form f(out..., in..., scalars...) and state form f(state, sys; kw...)


def _dims(f: Field):
    return (f.shape[0], f.shape[1]) if len(f.shape) > 1 else (f.shape[0], 1)  # (a Vector is an L x 1 lattice)


def _is_1d(x) -> bool:
    """State_1D / Vector arguments: the operator belongs to the 1-D family (swalbe_b200.one_d)"""
    return isinstance(x, one_d.LBM_state_1D) or (isinstance(x, Field) and len(x.shape) == 1)


def equilibrium(*args):
    """equilibrium!(feq, height, velx, vely, vsq, g) | equilibrium!(state, sys)   src/equilibrium.jl:63-128"""
    if isinstance(args[0], CuState):
        st, sys_ = args
        return equilibrium(st.feq, st.height, st.velx, st.vely, st.vsq, sys_.param.g)
    if isinstance(args[0], one_d.LBM_state_1D):  # equilibrium!(state::State_1D | Expanded_1D, sys)   src/equilibrium.jl:183-190
        st, sys_ = one_d.base(args[0]), args[1]
        return one_d.equilibrium(st.feq, st.height, st.vel, sys_.param.g)
    if len(args) == 4:                   # equilibrium!(feq, height, velocity, gravity)   :169
        return one_d.equilibrium(*args)
    feq, h, ux, uy, vsq, g = args
    _lib.call("swalbe_equilibrium_d2q9", _ptr(feq), _ptr(h), _ptr(ux), _ptr(uy), _ptr(vsq), float(g), *_dims(h), _stream())


def BGKandStream(*args, τ=None, tau=None):
    """BGKandStream!(fout, feq, ftemp, Fx, Fy, τ) | BGKandStream!(state, sys; τ)   src/collide.jl:70-161"""
    if isinstance(args[0], CuState):
        st, sys_ = args
        t = τ if τ is not None else (tau if tau is not None else sys_.param.tau)
        return BGKandStream(st.fout, st.feq, st.ftemp, st.Fx, st.Fy, t)
    if isinstance(args[0], CuStateWithBound_1D) and isinstance(args[1], SysConstWithBound_1D):  # bounce-back  src/collide.jl:214-249
        return one_d.BGKandStream_bound(*args)
    if isinstance(args[0], one_d.LBM_state_1D):  # src/collide.jl:203-211
        st, sys_ = one_d.base(args[0]), args[1]
        return one_d.BGKandStream(st.fout, st.feq, st.ftemp, st.F, sys_.param.tau)
    if len(args) == 5:                   # BGKandStream!(fout, feq, ftemp, F::Vector, τ)   :179
        return one_d.BGKandStream(*args)
    fout, feq, ftemp, Fx, Fy, t = args
    _lib.call("swalbe_bgk_stream_d2q9", _ptr(fout), _ptr(feq), _ptr(ftemp), _ptr(Fx), _ptr(Fy), float(t), *_dims(Fx), _stream())


def moments(*args):
    """moments!(height, velx, vely, fout) | moments!(state)   src/moments.jl:43-73"""
    if isinstance(args[0], CuState):
        st = args[0]
        return moments(st.height, st.velx, st.vely, st.fout)
    if isinstance(args[0], one_d.LBM_state_1D):  # src/moments.jl:66-71
        st = one_d.base(args[0])
        return one_d.moments(st.height, st.vel, st.fout)
    if len(args) == 3:                   # moments!(height::Vector, vel, fout)   :54
        return one_d.moments(*args)
    h, ux, uy, f = args
    _lib.call("swalbe_moments_d2q9", _ptr(h), _ptr(ux), _ptr(uy), _ptr(f), *_dims(h), _stream())


def cospi_field(θ: "Field") -> "Field":
    """cospi.(θ) on the device (swalbe_cospi_field); the result is what filmpressure!/time_loop take as the field."""
    out = Field(*θ.shape[:2])
    _lib.call("swalbe_cospi_field", out.ptr, θ.ptr, int(θ.t.numel()), _stream())
    return out


def _theta_args(θ):
    """θ scalar -> (cospi θ, NULL); θ Field of angles -> (0, device field cospi.(θ)), evaluated on the device once and
    cached on the Field until the Field is modified (move_substrate! etc.)."""
    if isinstance(θ, Field):
        c = getattr(θ, "_cospi", None)
        if c is None or getattr(θ, "_cospi_version", None) != θ._stamp():
            c = cospi_field(θ)
            θ._cospi, θ._cospi_version = c, θ._stamp()
        return 0.0, c.ptr
    return cospi(θ), None


def filmpressure(*args, θ=None, γ=None, n=None, m=None, hmin=None, hcrit=None, theta=None, gamma=None, Gamma=0.0):
    """filmpressure!(output, f, dgrad, γ, θ, n, m, hmin, hcrit)        src/pressure.jl:72-115 (fast_93/fast_32)
    filmpressure!(state, sys; θ, γ, n, m, hmin, hcrit)                src/pressure.jl:119-155 (power_broad)
    filmpressure!(state::CuState_thermal, sys)                         src/pressure.jl:117 (array form, no keywords)"""
    θ = θ if θ is not None else theta
    γ = γ if γ is not None else gamma
    if isinstance(args[0], one_d.Expanded_1D):  # src/pressure.jl:258-282 (Expanded_1D), :284-315 (State_gamma_1D)
        return one_d.filmpressure_expanded(args[0], args[1], θ=θ, n=n, m=m, hmin=hmin, hcrit=hcrit, γ=γ)
    if isinstance(args[0], CuState_1D):  # filmpressure!(state::LBM_state_1D, sys; θ, n, m, hmin, hcrit, γ)   src/pressure.jl:230-256
        st, sys_ = args
        p = sys_.param
        return one_d.filmpressure(st.pressure, st.height, st.dgrad, p.gamma if γ is None else γ, p.theta if θ is None else θ,
                                  p.n if n is None else n, p.m if m is None else m, p.hmin if hmin is None else hmin,
                                  p.hcrit if hcrit is None else hcrit, variant=_lib.PRESSURE_POWER_BROAD)
    if _is_1d(args[0]) and len(args) == 10:  # filmpressure!(output::Vector, f, dgrad, rho, γ, θ, n, m, hmin, hcrit; Gamma)   :318
        return one_d.filmpressure_rho(*args, Gamma=Gamma)
    if _is_1d(args[0]):                  # filmpressure!(output::Vector, f, dgrad, γ, θ, n, m, hmin, hcrit)   :196
        return one_d.filmpressure(*args)
    if isinstance(args[0], CuState):
        st, sys_ = args
        p = sys_.param
        if isinstance(st, CuState_thermal):
            if any(v is not None for v in (θ, γ, n, m, hmin, hcrit)):
                raise TypeError("MethodError: filmpressure!(::CuState_thermal, ::SysConst) accepts no keyword arguments")
            return filmpressure(st.pressure, st.height, st.dgrad, p.gamma, p.theta, p.n, p.m, p.hmin, p.hcrit)
        ct, ctf = _theta_args(p.theta if θ is None else θ)
        _lib.call("swalbe_filmpressure", st.pressure.ptr, st.height.ptr, st.dgrad.ptr,
                  float(p.gamma if γ is None else γ), ct, ctf, int(p.n if n is None else n), int(p.m if m is None else m),
                  float(p.hmin if hmin is None else hmin), float(p.hcrit if hcrit is None else hcrit),
                  _lib.PRESSURE_POWER_BROAD, st.Lx, st.Ly, _stream())
        return None
    out, f, dgrad, γa, θa, na, ma, hmina, hcrita = args
    ct, ctf = _theta_args(θa)
    _lib.call("swalbe_filmpressure", _ptr(out), _ptr(f), _ptr(dgrad), float(γa), ct, ctf, int(na), int(ma), float(hmina),
              float(hcrita), _lib.PRESSURE_FAST, *_dims(f), _stream())


def hgradp(st: CuState):
    """h∇p!(state)   src/forcing.jl:168-187"""
    if isinstance(st, one_d.LBM_state_1D):
        return one_d.hgradp(st)
    _lib.call("swalbe_hgradp", st.hgradpx.ptr, st.hgradpy.ptr, st.pressure.ptr, st.height.ptr, st.Lx, st.Ly, _stream())


def gradf(outx, outy, f, *rest):
    """∇f!(outx, outy, f) | ∇f!(outx, outy, f, a) | ∇f!(outx, outy, f, dgrad, a)   src/differences.jl:153-206"""
    if _is_1d(outx):  # ∇f!(output::Vector, f, dgrad[, a])   src/differences.jl:208-230
        return one_d.gradf(outx, outy, f, *rest)
    a = rest[-1] if rest else None
    if a is not None and not isinstance(a, Field):  # a scalar multiplier broadcasts like the reference's `a .* (...)`
        a = Field(*_dims(f), fill=float(a))
    _lib.call("swalbe_grad9", _ptr(outx), _ptr(outy), _ptr(f), _ptr(a), *_dims(f), _stream())


def laplacianf(out, f, γ):
    """∇²f!(output, f, γ)   src/differences.jl:57-75"""
    if _is_1d(out):  # ∇²f!(output, f::Vector, dgrad)   :77-85
        return one_d.laplacianf(out, f, γ)
    _lib.call("swalbe_lap9", _ptr(out), _ptr(f), float(γ), *_dims(f), _stream())


def _slip(variant, args):
    if isinstance(args[0], one_d.LBM_state_1D):  # slippage!(state::LBM_state_1D | Expanded_1D, sys)   src/forcing.jl:73-83
        st, sys_ = one_d.base(args[0]), args[1]
        return one_d.slippage(st.slip, st.height, st.vel, sys_.param.delta, sys_.param.mu)
    if _is_1d(args[0]):                  # slippage!(slip, height, vel, δ, μ)   src/forcing.jl:68-71
        return one_d.slippage(*args)
    if isinstance(args[0], CuState):
        st, sys_ = args
        p = sys_.param
        a = (st.slipx, st.slipy, st.height, st.velx, st.vely, p.delta, p.mu, p.hcrit)
    else:
        a = tuple(args) + ((0.0,) if len(args) == 7 else ())
    sx, sy, h, ux, uy, δ, μ, hcrit = a
    _lib.call("swalbe_slippage", _ptr(sx), _ptr(sy), _ptr(h), _ptr(ux), _ptr(uy), float(δ), float(μ), float(hcrit),
              variant, *_dims(h), _stream())


def slippage(*args):
    """slippage!(slipx, slipy, height, velx, vely, δ, μ) | slippage!(state, sys)   src/forcing.jl:42-56"""
    _slip(_lib.SLIP_STANDARD, args)


def slippage2(st, sys_):
    """slippage2!(state, sys)   src/forcing.jl:85-99"""
    _slip(_lib.SLIP_HCRIT, (st, sys_))


def slippage_ring_riv(*args):
    """slippage_ring_riv!(slipx, slipy, height, velx, vely, δ, μ, hcrit) | (state, sys)   src/forcing.jl:107-122"""
    _slip(_lib.SLIP_RING_RIV, args)


_noise = {"seed": None, "calls": 0}


def _noise_seed() -> int:
    """the process-wide default Philox key: drawn from the OS once, like Julia's unseeded default RNG"""
    if _noise["seed"] is None:
        import os

        _noise["seed"] = int.from_bytes(os.urandom(8), "little")
    return _noise["seed"]


def thermal(*args, seed=None, step=None):
    """thermal!(kbtx, kbty, height, kbt, μ, δ) | thermal!(state, sys)   src/forcing.jl:297-320
    (counter-based Philox normals keyed on (seed, step, cell) instead of Julia's randn! stream).

    Like `randn!`, every call draws FRESH noise: without `step` a process-wide call counter advances the Philox
    stream, without `seed` the key is drawn from the OS once per process.  Pass both for a reproducible field (the
    decomposition-independent noise of the fused loop uses step = time step)."""
    if step is None:
        step = _noise["calls"]
        _noise["calls"] += 1
    if seed is None:
        seed = _noise_seed()
    if isinstance(args[0], CuState_thermal_1D):  # thermal!(state::State_thermal_1D, sys)   src/forcing.jl:335-336
        st, sys_ = args
        p = sys_.param
        return one_d.thermal(st.kbt, st.basestate.height, p.kbt, p.mu, p.delta, seed, step)
    if len(args) == 5:                           # thermal!(fluc, height, kᵦT, μ, δ)   src/forcing.jl:322-333
        return one_d.thermal(*args, seed, step)
    if isinstance(args[0], CuState):
        st, sys_ = args
        p = sys_.param
        args = (st.kbtx, st.kbty, st.height, p.kbt, p.mu, p.delta)
    kx, ky, h, kbt, μ, δ = args
    _lib.call("swalbe_thermal", _ptr(kx), _ptr(ky), _ptr(h), float(kbt), float(μ), float(δ), int(seed), int(step),
              *_dims(h), _stream())


def inclination(α, st: CuState, t=1000, tstart=0, tsmooth=1):
    """inclination!(α, state; t, tstart, tsmooth)   src/forcing.jl:363-368"""
    if isinstance(st, one_d.LBM_state_1D):  # inclination!(α::Float64, state::State_1D | Expanded_1D)   src/forcing.jl:379-389
        return one_d.inclination(α, st, t=t, tstart=tstart, tsmooth=tsmooth)
    factor = 0.5 + 0.5 * math.tanh((t - tstart) / tsmooth)
    _lib.call("swalbe_inclination", st.Fx.ptr, st.Fy.ptr, st.height.ptr, float(α[0]), float(α[1]), factor, st.Lx, st.Ly,
              _stream())


def update(st: CuState):
    """The inline force sum of the drivers, `state.Fx .= -state.h∇px .- state.slipx` (src/simulate.jl:18-19);
    for thermal states `... .- state.kbtx` (scripts/Rivulet_stability.jl:123-124).  north_star calls it update!."""
    if isinstance(st, one_d.LBM_state_1D):
        return one_d.update(st)
    kx, ky = (st.kbtx.ptr, st.kbty.ptr) if st.thermal else (None, None)
    _lib.call("swalbe_force_sum", st.Fx.ptr, st.Fy.ptr, st.hgradpx.ptr, st.hgradpy.ptr, st.slipx.ptr, st.slipy.ptr, kx, ky,
              st.Lx, st.Ly, _stream())


def field_stats(f: Field, thresh=0.055):
    """(min, max, sum, count(f > thresh)) computed on the device; one 32-byte read-back."""
    torch = _torch()
    out = torch.empty(4, dtype=torch.float64, device="cuda")
    _lib.call("swalbe_field_stats", C.c_void_p(out.data_ptr()), f.ptr, float(thresh), *_dims(f), _stream())
    mn, mx, sm, cnt = out.cpu().tolist()
    return mn, mx, sm, int(cnt)


def wetted(area_size, *args, hthresh=0.055):
    """wetted!(area_size, state; hthresh)                 src/measures.jl:13-17  (push! the wetted-site count)
    wetted!(area_size, maxheight, height, t; hthresh)   src/measures.jl:6-11   (slot t, Julia's 1-based step index)

    Both reductions run on the device (swalbe_field_stats); only the two numbers cross the bus."""
    if len(args) == 1 or (len(args) == 2 and not isinstance(args[1], Field) and isinstance(args[0], CuState)):
        st = args[0]
        if len(args) == 2:  # positional hthresh, as the first version of this mirror took it
            hthresh = args[1]
        area_size.append(field_stats(st.height, hthresh)[3])
        return
    if len(args) != 3:
        raise TypeError("MethodError: no method matching wetted! with these arguments")
    maxheight, height, t = args
    _, mx, _, cnt = field_stats(height, hthresh)
    area_size[t - 1] = cnt
    maxheight[t - 1] = mx


class SnapshotBuffer:
    """The scripts' `snap = zeros(ndumps, Lx*Ly)` matrix in PINNED host memory, filled by asynchronous copies
    (SURVEY.md 8f1): `snapshot!` copies the field into a device staging plane on the compute stream (the field itself is
    overwritten by the very next step) and a side stream moves the staging plane to the host while the loop goes on.
    `array()` waits for the copies in flight and returns the NumPy matrix (rows = dumps, Julia's `vec` order)."""

    def __init__(self, ndumps: int, Lx: int, Ly: int):
        torch = _torch()
        self.host = torch.zeros((ndumps, Lx * Ly), dtype=torch.float64).pin_memory()
        self.stage = torch.empty(Lx * Ly, dtype=torch.float64, device="cuda")
        self.stream = torch.cuda.Stream()
        self.done = None

    def push(self, row: int, field: Field):
        torch = _torch()
        cur = torch.cuda.current_stream()
        if self.done is not None:
            cur.wait_event(self.done)  # the staging plane is free again once the previous dump has left the device
        self.stage.copy_(field.t.reshape(-1))
        ready = torch.cuda.Event()
        ready.record(cur)
        self.stream.wait_event(ready)
        with torch.cuda.stream(self.stream):
            self.host[row].copy_(self.stage, non_blocking=True)
            self.done = torch.cuda.Event()
            self.done.record(self.stream)

    def array(self) -> np.ndarray:
        self.stream.synchronize()
        return self.host.numpy()


def snapshot(snap, field: Field, t: int, dumping=1000):
    """snapshot!(snap, field, t; dumping)   src/measures.jl:99-105.  `snap`: a SnapshotBuffer (asynchronous pinned
    copy, the loop is not stalled) or a plain NumPy matrix (synchronous, like the reference's `Array(field)`)."""
    if t % dumping == 0:
        if isinstance(snap, SnapshotBuffer):
            snap.push(t // dumping - 1, field)
        else:
            snap[t // dumping - 1, :] = field.numpy().reshape(-1, order="F")


# ------------------------------------------------------------------------------------------------
# small helpers of the reference that user code calls directly


def viewdists(f: Field):
    """viewdists(f)   src/collide.jl:270-282 -- the nine population planes as views (torch tensors, [j, i] indexed)."""
    return tuple(f.t[k] for k in range(9))


def viewneighbors(f: Field):
    """viewneighbors(f)   src/differences.jl:251-262 -- the eight planes of a dgrad-like array as views."""
    return tuple(f.t[k] for k in range(8))


def power_broad(arg, n: int):
    """power_broad(arg, n)   src/pressure.jl:363-385 -- temp = 1; temp *= arg, n times (host scalar helper), with the
    reference's three methods: Float64 (:363), Float32 (:371, every product rounded to single precision) and Int (:379)."""
    if isinstance(arg, np.float32):
        temp = np.float32(1.0)
    else:
        temp = 1.0 if isinstance(arg, (float, np.floating)) else 1
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
# drivers  (src/simulate.jl)


def _c_params(p: Taumucs, θ=None, slip_variant=_lib.SLIP_STANDARD, incl=None, thermal_seed=None,
              pressure_variant=_lib.PRESSURE_POWER_BROAD):
    q = _lib.CParams()
    q.tau, q.mu, q.delta, q.kbt, q.gamma, q.hmin, q.hcrit, q.g = p.tau, p.mu, p.delta, p.kbt, p.gamma, p.hmin, p.hcrit, p.g
    q.n, q.m = p.n, p.m
    ct, ctf = _theta_args(p.theta if θ is None else θ)
    q.cospi_theta, q.cospi_theta_field = ct, ctf
    q.pressure_variant, q.slip_variant = pressure_variant, slip_variant
    if incl is not None:
        α, factor = incl
        q.use_inclination, q.incl_ax, q.incl_ay, q.incl_factor = 1, float(α[0]), float(α[1]), float(factor)
    if thermal_seed is not None:
        q.use_thermal, q.seed = 1, int(thermal_seed)
    return q


def fused_steps(st: CuState, sys_: SysConst, nsteps: int, *, θ=None, slip_variant=_lib.SLIP_STANDARD, incl=None,
                thermal_seed=None, step0=0, lazy_populations=False, log_minmax=False, log_wetted=False, hthresh=0.055,
                pressure_variant=None, skip_aux=False, moments_consistent=False, host_in=None, host_out=None,
                mass_log=None):
    """nsteps iterations of the loop body src/simulate.jl:15-22 through swalbe_time_loop (one fused kernel/step).

    ``host_in`` / ``host_out`` (pinned CPU torch tensors or NumPy arrays of Lx*Ly float64 in the memory order of
    ``state.height``, i.e. ``h.T`` C-contiguous): the loop starts from the height in ``host_in`` and / or leaves the final
    height in ``host_out`` (swalbe_time_loop_host: on large lattices the copies travel in row bands behind which / ahead of
    which the first / last steps run).  The download is complete once the current stream has been synchronised.

    ``mass_log = (first, every, out)``: out[m] = sum(height) BEFORE step first + m*every of this call (0-based), written by
    the device as soon as it is known; ``out``: float64 CUDA tensor or pinned CPU tensor (pre-set to NaN to poll it).

    ``moments_consistent`` (τ ≠ 1): the caller vouches that height/velx/vely are the moments of ftemp -- true after any
    earlier fused_steps/time_loop/moments! call on this state, false after writing an initial condition into height --
    so that the first step, too, derives them from the populations (SWALBE_LOOP_MOMENTS_CONSISTENT).

    Returns (hmin[nsteps], hmax[nsteps], wetted[nsteps]) device tensors for the requested logs (else None)."""
    torch = _torch()
    if pressure_variant is None:  # CuState_thermal goes through the array form (src/pressure.jl:117)
        pressure_variant = _lib.PRESSURE_FAST if isinstance(st, CuState_thermal) else _lib.PRESSURE_POWER_BROAD
    q = _c_params(sys_.param, θ, slip_variant, incl, thermal_seed, pressure_variant)
    cs = st._c_state()
    logs = _lib.CLogs()
    mn = mx = wet = None
    if log_minmax:
        mn = torch.empty(nsteps, dtype=torch.float64, device="cuda")
        mx = torch.empty(nsteps, dtype=torch.float64, device="cuda")
        logs.hmin, logs.hmax = mn.data_ptr(), mx.data_ptr()
    if log_wetted:
        wet = torch.empty(nsteps, dtype=torch.int64, device="cuda")
        logs.wetted = wet.data_ptr()
    logs.hthresh = hthresh
    flags = ((_lib.LOOP_LAZY_POPULATIONS if lazy_populations else 0) | (_lib.LOOP_SKIP_AUX if skip_aux else 0) |
             (_lib.LOOP_MOMENTS_CONSISTENT if moments_consistent else 0))
    if mass_log is not None:
        logs.hsum_first, logs.hsum_every, out = int(mass_log[0]), int(mass_log[1]), mass_log[2]
        if out.dtype != torch.float64 or not out.is_contiguous() or (out.device.type == "cpu" and not out.is_pinned()):
            raise ValueError("mass_log: contiguous float64 CUDA tensor or pinned CPU tensor expected")
        need = 0 if nsteps <= logs.hsum_first else (nsteps - 1 - logs.hsum_first) // max(1, logs.hsum_every) + 1
        if out.numel() < need:
            raise ValueError(f"mass_log: {need} slots needed, {out.numel()} given")
        logs.hsum = out.data_ptr()
    lg = C.byref(logs) if (log_minmax or log_wetted or mass_log is not None) else None
    if host_in is None and host_out is None:
        _lib.call("swalbe_time_loop", st.plan(), C.byref(cs), C.byref(q), int(nsteps), int(step0), flags, lg, _stream())
    else:
        _lib.call("swalbe_time_loop_host", st.plan(), C.byref(cs), C.byref(q), int(nsteps), int(step0), flags, lg,
                  _host_ptr(host_in, st), _host_ptr(host_out, st), _stream())
        st.height.touch()
    return mn, mx, wet


def _host_ptr(x, st):
    """address of a host plane of Lx*Ly float64 in the memory order of ``state.height`` (i fastest): a NumPy array of
    shape (Lx, Ly) in Fortran order, of shape (Ly, Lx) in C order, or flat; a torch CPU tensor (ideally pinned) of shape
    (Ly, Lx) or flat.  Anything whose memory order would silently transpose the lattice is refused."""
    if x is None:
        return None
    Lx, Ly = st.Lx, st.Ly
    shape = tuple(x.shape)
    if isinstance(x, np.ndarray):
        ok = x.dtype == np.float64 and ((shape == (Lx * Ly,) and x.flags.c_contiguous) or
                                        (shape == (Ly, Lx) and x.flags.c_contiguous) or (shape == (Lx, Ly) and x.flags.f_contiguous))
        if not ok:
            raise ValueError(f"host plane: float64 array of shape ({Lx}, {Ly}) in Fortran order, ({Ly}, {Lx}) in C order or flat "
                             f"expected, got shape {shape}, C-contiguous={x.flags.c_contiguous}, F-contiguous={x.flags.f_contiguous}")
        return C.c_void_p(x.ctypes.data)
    import torch  # (the layout check itself needs no device)

    ok = (x.device.type == "cpu" and x.dtype == torch.float64 and x.is_contiguous() and shape in ((Lx * Ly,), (Ly, Lx)))
    if not ok:
        raise ValueError(f"host plane: contiguous float64 CPU tensor of shape ({Ly}, {Lx}) or flat expected, got {shape}")
    return C.c_void_p(x.data_ptr())


_ONE_CALL_SITES = 1 << 21  # time_loop: lattices from this size on run as one library call (see time_loop)


def _mass_buffer(n):
    """Pinned host slots for the in-loop mass log, pre-set to NaN.  The buffer is owned by the module and never returned to
    torch's pinned-memory cache: the device writes into it from inside a loop that may still be running when time_loop
    returns, so it is only reused once the loop that used it last has finished."""
    torch = _torch()
    busy = getattr(_mass_buffer, "busy", None)
    if busy is not None:
        busy.synchronize()
        _mass_buffer.busy = None
    buf = getattr(_mass_buffer, "buf", None)
    if buf is None or buf.numel() < max(1, n):
        buf = torch.empty(max(16, 2 * n), dtype=torch.float64).pin_memory()
        _mass_buffer.old = getattr(_mass_buffer, "old", []) + [getattr(_mass_buffer, "buf", None)]  # (kept alive)
        _mass_buffer.buf = buf
    buf.fill_(float("nan"))
    return buf


class _MassDumps:
    """`mass = sum(state.height)` at every dump step (src/simulate.jl:8-14) without stalling the loop: the reduction runs
    on the device in stream order (swalbe_field_stats), its 32 bytes travel to pinned host memory asynchronously, and the
    line is printed as soon as the copy has landed -- same lines, same order, never a host synchronisation in the loop."""

    _ring = None  # pinned staging shared by all loops of the process: 256 dumps in flight

    def __init__(self, verbose):
        self.verbose, self.pending, self.masses = verbose, [], []

    def push(self, t, height: Field):
        torch = _torch()
        cls = _MassDumps
        if cls._ring is None:
            cls._ring = (torch.empty((256, 4), dtype=torch.float64).pin_memory(),
                         torch.empty((256, 4), dtype=torch.float64, device="cuda"), [0])
        host, dev, nxt = cls._ring
        if len(self.pending) >= 128:
            self.flush(block=True)
        i = nxt[0] % 256
        nxt[0] += 1
        _lib.call("swalbe_field_stats", C.c_void_p(dev[i].data_ptr()), height.ptr, 0.055, *_dims(height), _stream())
        host[i].copy_(dev[i], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.pending.append((t, ev, i))
        self.flush(block=False)

    def flush(self, block):
        host = _MassDumps._ring[0] if _MassDumps._ring else None
        while self.pending and (block or self.pending[0][1].query()):
            t, ev, i = self.pending.pop(0)
            ev.synchronize()
            self._emit(t, float(host[i, 2]))

    def _emit(self, t, mass):
        self.masses.append((t, mass))
        if self.verbose:
            print(f"Time step {t} mass is {round(mass, 3)}")


def time_loop(sys_: SysConst, st: CuState, *extra, verbose=False, chunk=None, lazy_populations=True, host_in=None,
              host_out=None):
    """The four 2-D time_loop methods  src/simulate.jl:6-96:

    time_loop(sys, state)                       plain                          :6-25
    time_loop(sys, state, θ)                    θ scalar or Field              :26-45
    time_loop(sys, state, Δh::list)             logs max-min every step        :47-67
    time_loop(sys, state, f, measure::list)     callback slot; f ∈ {wetted, inclination}  :69-96

    The loop body runs as fused kernels in chunks of `tdump` steps; the mass print happens at the same
    steps as in the reference (t % tdump == 0, before that step's update).  On return every field of the state holds
    what the reference's holds; in between, the loop moves as little as the arithmetic allows: at τ = 1 the populations
    are only written by the last step of a chunk (`lazy_populations`: ω = 0, nothing reads them -- 48 instead of 120
    bytes per lattice update, same bits on return), at τ ≠ 1 every chunk after the first vouches for its moments so that
    all of its steps derive h and u from the populations (144 instead of 192 bytes).

    ``host_in`` / ``host_out``: host planes (see fused_steps) the first chunk starts from / the last chunk leaves the final
    height in -- `state.height .= CUDA.adapt(CuArray, h)` before and `Array(state.height)` after the loop, with the
    copies hidden behind the first and last steps."""
    if isinstance(sys_, (SysConst_1D, SysConstWithBound_1D)):
        return one_d.time_loop(sys_, st, *extra, verbose=verbose)
    p = sys_.param
    θ, dh, cb, measure = None, None, None, None
    if len(extra) == 1 and isinstance(extra[0], list):
        dh = extra[0]
    elif len(extra) == 1:
        θ = extra[0]
    elif len(extra) == 2:
        cb, measure = extra
        if cb not in (wetted, inclination):
            raise SwalbeError("time_loop: only Swalbe.wetted! and Swalbe.inclination! callbacks are fused on the device")
    elif extra:
        raise TypeError("MethodError: no method matching time_loop with these arguments")
    incl = (measure, 0.5 + 0.5 * math.tanh((1000 - 0) / 1)) if cb is inclination else None  # defaults of :363
    t = 1
    tdump = max(1, p.tdump)
    if sys_.Lx * sys_.Ly >= _ONE_CALL_SITES and not chunk and p.Tmax >= 1:
        # Large lattices: the whole loop is ONE library call.  The mass of every dump step is summed on the device inside
        # the loop (logs.hsum) and lands in pinned host memory by itself; a verbose loop prints each line as it arrives.
        # (Small lattices keep the chunks between two dumps: repeated chunks are what the CUDA-graph replay needs.)
        torch = _torch()
        ndumps = p.Tmax // tdump
        host = _mass_buffer(ndumps)
        mn, mx, wet = fused_steps(st, sys_, p.Tmax, θ=θ, incl=incl, log_minmax=dh is not None, log_wetted=cb is wetted,
                                  lazy_populations=bool(lazy_populations) and p.tau == 1.0, host_in=host_in, host_out=host_out,
                                  mass_log=(tdump - 1, tdump, host) if ndumps else None)
        done = torch.cuda.Event()
        done.record()
        _mass_buffer.busy = done
        if verbose:
            for m in range(ndumps):
                while math.isnan(host[m].item()):
                    if done.query() and math.isnan(host[m].item()):
                        raise SwalbeError("time_loop: the loop finished without writing the mass of a dump step")
                    time.sleep(5e-5)
                print(f"Time step {(m + 1) * tdump} mass is {round(host[m].item(), 3)}")
        if dh is not None:
            dh.extend((mx - mn).cpu().tolist())
        if cb is wetted:
            measure.extend(wet.cpu().tolist())
        return st if cb is None else (st, measure)
    dumps = _MassDumps(verbose)
    while t <= p.Tmax:
        if t % tdump == 0:
            dumps.push(t, st.height)
        nxt = min(p.Tmax + 1, (t // tdump + 1) * tdump)  # run up to (not including) the next dump step
        if chunk:
            nxt = min(nxt, t + chunk)
        n = nxt - t
        # only the final chunk needs feq/pressure/h∇p/slip/F materialised (the state the reference returns)
        mn, mx, wet = fused_steps(st, sys_, n, θ=θ, incl=incl, log_minmax=dh is not None, log_wetted=cb is wetted,
                                  skip_aux=nxt <= p.Tmax, lazy_populations=bool(lazy_populations) and p.tau == 1.0,
                                  moments_consistent=t > 1, host_in=host_in if t == 1 else None,
                                  host_out=host_out if nxt > p.Tmax else None)
        if dh is not None:
            dh.extend((mx - mn).cpu().tolist())
        if cb is wetted:
            measure.extend(wet.cpu().tolist())
        t = nxt
    dumps.flush(block=True)
    if p.Tmax < 1:  # (no step ran: the copies still happen)
        fused_steps(st, sys_, 0, host_in=host_in, host_out=host_out)
    return st if cb is None else (st, measure)


def run_host(sys_: SysConst, h_host, out_host=None, θ=None, kind="simple", verbos=False, state=None):
    """A whole film job from and to host memory, the shape of every shipped GPU script and of the run_* drivers
    (src/simulate.jl:338-358): Sys(sys, "GPU"); state.height .= h; time_loop(sys, state[, θ]); Array(state.height).
    The reference's `equilibrium!` before the loop is skipped when at least one step runs (every field it writes is
    written again by the last step, and at τ = 1 nothing reads it in between).  ``h_host`` / ``out_host``: see fused_steps;
    returns (state, out_host) with the download complete."""
    torch = _torch()
    st = state if state is not None else Sys(sys_, "GPU", kind=kind)
    if out_host is None:
        out_host = torch.empty(sys_.Lx * sys_.Ly, dtype=torch.float64).pin_memory()
    if sys_.param.Tmax < 1 or sys_.param.tau != 1.0:
        st.height.t.copy_(h_host.reshape(st.height.t.shape) if not isinstance(h_host, np.ndarray)
                          else torch.from_numpy(h_host).reshape(st.height.t.shape), non_blocking=True)
        equilibrium(st, sys_)
        h_host = None
    time_loop(sys_, st, *(() if θ is None else (θ,)), verbose=verbos, host_in=h_host, host_out=out_host)
    torch.cuda.current_stream().synchronize()
    return st, out_host


def run_flat(sys_: SysConst, device: str = "GPU", verbos=True):
    """run_flat  src/simulate.jl:236-245 (2-D), :247-256 (1-D: run_flat(sys::SysConst_1D))"""
    if isinstance(sys_, SysConst_1D):
        return one_d.run_flat(sys_, verbos=verbos)
    print("Simulating a flat interface without driving forces (nothing should happen) in two dimensions")
    st = Sys(sys_, device)
    st.height.set(1.0)
    time_loop(sys_, st, verbose=verbos)
    return st.height


def run_random(sys_: SysConst, device: str = "GPU", h0=1.0, ϵ=0.01, verbos=True, rng=None):
    """run_random  src/simulate.jl:286-294 (randinterface! src/initialvalues.jl:23-33 with a NumPy generator); 1-D :296-304"""
    if isinstance(sys_, SysConst_1D):
        return one_d.run_random(sys_, h0=h0, ϵ=ϵ, verbos=verbos, rng=rng)
    print("Simulating a random undulated interface in two dimensions")
    st = Sys(sys_, device)
    rng = rng if rng is not None else np.random.default_rng()
    st.height.set(h0 * (1.0 + ϵ * rng.standard_normal((sys_.Lx, sys_.Ly))))
    equilibrium(st, sys_)
    time_loop(sys_, st, 1 / 9, verbose=verbos)
    return st.height


def run_rayleightaylor(sys_: SysConst, device: str, kx=15, ky=18, h0=1.0, ϵ=0.001, verbos=True):
    """run_rayleightaylor  src/simulate.jl:338-358 (divides by Lx-1, Ly-1 like the reference)"""
    print("Simulating the Rayleigh Taylor instability in two dimensions")
    st = Sys(sys_, device)
    i = np.arange(1, sys_.Lx + 1, dtype=np.float64)[:, None]
    j = np.arange(1, sys_.Ly + 1, dtype=np.float64)[None, :]
    st.height.set(h0 * (1 + ϵ * np.sin(2 * np.pi * kx * i / (sys_.Lx - 1)) * np.sin(2 * np.pi * ky * j / (sys_.Ly - 1))))
    diff: list = []
    equilibrium(st, sys_)
    time_loop(sys_, st, diff, verbose=verbos)
    return st.height, diff


def _slab(field, j_begin, Ly):
    Lx, Lyl = field.shape[:2]
    return Lx, Lyl, int(j_begin), int(Ly if Ly is not None else j_begin + Lyl)


def singledroplet(*args, precursor=0.05, j_begin=0):
    """singledroplet(height, radius, θ, center)  src/initialvalues.jl:203-224.

    ``singledroplet(field, radius, θ, center)`` fills the device Field in place (swalbe_ic_singledroplet; ``j_begin``
    places a row slab in the global lattice); ``singledroplet(Lx, Ly, radius, θ, center)`` is the reference's
    host-side construction and returns a NumPy array (what the run_* drivers upload, like upstream)."""
    if isinstance(args[0], Field):
        f, radius, θ, center = args
        Lx, Lyl, jb, _ = _slab(f, j_begin, None)
        _lib.call("swalbe_ic_singledroplet", f.ptr, float(radius), cospi(θ), float(center[0]), float(center[1]),
                  float(precursor), Lx, Lyl, jb, _stream())
        return f.touch()
    Lx, Ly, radius, θ, center = args
    i = np.arange(1, Lx + 1, dtype=np.float64)[:, None]
    j = np.arange(1, Ly + 1, dtype=np.float64)[None, :]
    circ = np.sqrt((i - center[0]) ** 2 + (j - center[1]) ** 2)
    inside = circ <= radius
    cap = (np.cos(np.arcsin(np.where(inside, circ / radius, 0.0))) - cospi(θ)) * radius
    h = np.where(inside, cap, precursor)
    return np.where(h < 0, precursor, h)


def torus(field: "Field", r1, R2, θ, center, hmin=0.05, noise=0.0, seed=0, j_begin=0):
    """torus(lx, ly, r₁, R₂, θ, center, hmin; noise)  src/initialvalues.jl:144-168, written into a device Field."""
    Lx, Lyl, jb, _ = _slab(field, j_begin, None)
    _lib.call("swalbe_ic_torus", field.ptr, float(r1), float(R2), cospi(θ), float(center[0]), float(center[1]),
              float(hmin), float(noise), int(seed), Lx, Lyl, jb, _stream())
    return field.touch()


def rivulet(field: "Field", radius, θ, orientation, center, hmin=0.05, noise=0.0, seed=0, j_begin=0):
    """rivulet(Lx, Ly, radius, θ, orientation, center, hmin; noise)  src/initialvalues.jl:69-104 (orientation "y"|"x",
    Julia's :y / :x), written into a device Field."""
    if orientation not in ("y", "x"):
        raise ValueError("orientation must be 'y' or 'x'")
    Lx, Lyl, jb, _ = _slab(field, j_begin, None)
    _lib.call("swalbe_ic_rivulet", field.ptr, float(radius), cospi(θ), 0 if orientation == "y" else 1, float(center),
              float(hmin), float(noise), int(seed), Lx, Lyl, jb, _stream())
    return field.touch()


def sinewave2d(field: "Field", h0=1.0, ϵ=0.001, kx=15, ky=18, j_begin=0, Ly=None):
    """The initial condition loop of run_rayleightaylor (src/simulate.jl:350-353) on the device."""
    Lx, Lyl, jb, Lyg = _slab(field, j_begin, Ly)
    _lib.call("swalbe_ic_sinewave2d", field.ptr, float(h0), float(ϵ), float(kx), float(ky), Lx, Lyg, Lyl, jb, _stream())
    return field.touch()


def randinterface(field: "Field", h0, ϵ, seed=None, j_begin=0):
    """randinterface!(height, h₀, ϵ)  src/initialvalues.jl:23-33 with counter-based normals on the device.  Without
    `seed` every call draws a new interface (like the reference's randn!); slabs of one lattice must share a seed."""
    if seed is None:
        seed = (_noise_seed() + 0x9E3779B97F4A7C15 * (1 + _noise["calls"])) & 0xFFFFFFFFFFFFFFFF
        _noise["calls"] += 1
    Lx, Lyl, jb, _ = _slab(field, j_begin, None)
    _lib.call("swalbe_ic_randinterface", field.ptr, float(h0), float(ϵ), int(seed), Lx, Lyl, jb, _stream())
    return field.touch()


def circshift(dst: "Field", src: "Field", shifts):
    """circshift!(dst, src, (sx, sy)) on device Fields: dst[i, j] = src[i - sx, j - sy] (periodic)."""
    if dst.shape != src.shape or len(dst.shape) != 2:
        raise ValueError(f"DimensionMismatch: {dst.shape} vs {src.shape}")
    _lib.call("swalbe_circshift", dst.ptr, src.ptr, int(shifts[0]), int(shifts[1]), *dst.shape, _stream())
    return dst.touch()


def move_substrate(θ: "Field", input: "Field", t, tmove, direction="diagonal"):
    """move_substrate!(θ, input, t, tmove; direction)  scripts/Moving_wettability_structs.jl:139-152."""
    if (t % tmove == 0) and (t > 0):
        shift = {"diagonal": (1, 1), "x": (1, 0), "y": (0, 1)}.get(direction)
        if shift is not None:
            circshift(θ, input, shift)
        input.set(θ)


def run_dropletrelax(sys_: SysConst, device: str, radius=20, θ0=1 / 6, center=None, verbos=True):
    """run_dropletrelax  src/simulate.jl:384-399"""
    print("Simulating an out of equilibrium droplet in two dimensions")
    center = center or (sys_.Lx // 2, sys_.Ly // 2)
    st = Sys(sys_, device)
    st.height.set(singledroplet(sys_.Lx, sys_.Ly, radius, θ0, center))
    equilibrium(st, sys_)
    area: list = []
    time_loop(sys_, st, wetted, area, verbose=verbos)
    return st.height, area


def run_dropletpatterned(sys_: SysConst, device: str, radius=20, θ0=1 / 6, center=None, θs=None, verbos=True):
    """run_dropletpatterned  src/simulate.jl:422-439 (θₛ: Lx x Ly NumPy array or Field of angles)"""
    print("Simulating a droplet on a patterned substrate in two dimensions")
    center = center or (sys_.Lx // 2, sys_.Ly // 2)
    st = Sys(sys_, device)
    st.height.set(singledroplet(sys_.Lx, sys_.Ly, radius, θ0, center))
    equilibrium(st, sys_)
    if θs is None:
        θs = np.full((sys_.Lx, sys_.Ly), 1 / 9)
    if not isinstance(θs, Field):
        θs = Field(sys_.Lx, sys_.Ly).set(θs)
    time_loop(sys_, st, θs, verbose=verbos)
    return st.height


def run_dropletforced(sys_: SysConst, device: str = "GPU", radius=20, θ0=1 / 6, center=None, fx=0.0, fy=0.0, verbos=True,
                      f=None):
    """run_dropletforced  src/simulate.jl:462-484"""
    if isinstance(sys_, SysConst_1D):  # run_dropletforced(sys::SysConst_1D; radius, θ₀, center, θₛ, f)   src/simulate.jl:486-503
        return one_d.run_dropletforced(sys_, radius=radius, θ0=θ0, center=center, f=fx if f is None else f, verbos=verbos)
    bodyforce = [fx, fy]
    print("Simulating a sliding droplet in two dimensions")
    center = center or (sys_.Lx // 2, sys_.Ly // 2)
    st = Sys(sys_, device)
    st.height.set(singledroplet(sys_.Lx, sys_.Ly, radius, θ0, center))
    equilibrium(st, sys_)
    print("Starting the lattice Boltzmann time loop")
    time_loop(sys_, st, inclination, bodyforce, verbose=verbos)
    return st.height, st.velx, st.vely


gradgamma, update_rho, obslist, run_gamma, two_droplets = (one_d.gradgamma, one_d.update_rho, one_d.obslist, one_d.run_gamma,
                                                           one_d.two_droplets)
from .io import (dump_height_slab, load_height_slab, load_heights_bson, restart_from_height, save_heights,  # noqa: E402
                 save_heights_bson)

JULIA_NAMES = {
    "equilibrium!": equilibrium, "BGKandStream!": BGKandStream, "moments!": moments, "filmpressure!": filmpressure,
    "h∇p!": hgradp, "∇f!": gradf, "∇²f!": laplacianf, "slippage!": slippage, "slippage2!": slippage2,
    "slippage_ring_riv!": slippage_ring_riv, "thermal!": thermal, "inclination!": inclination, "update!": update,
    "wetted!": wetted, "snapshot!": snapshot, "time_loop": time_loop, "run_flat": run_flat, "run_random": run_random,
    "run_rayleightaylor": run_rayleightaylor, "run_dropletrelax": run_dropletrelax,
    "run_dropletpatterned": run_dropletpatterned, "run_dropletforced": run_dropletforced, "Sys": Sys,
    "SysConst": SysConst, "Sys_const": Sys_const, "Taumucs": Taumucs, "CuState": CuState,
    "CuState_thermal": CuState_thermal, "Swalbe_state": Swalbe_state, "viewdists": viewdists,
    "viewneighbors": viewneighbors, "singledroplet": singledroplet, "torus": torus, "rivulet": rivulet,
    "∇γ!": one_d.gradgamma, "update_rho!": one_d.update_rho, "obslist!": one_d.obslist, "run_gamma": one_d.run_gamma,
    "two_droplets": one_d.two_droplets, "randinterface!": randinterface, "circshift!": circshift, "move_substrate!": move_substrate, "restart_from_height": restart_from_height, "power_broad": power_broad, "fast_93": fast_93, "fast_32": fast_32,
}
