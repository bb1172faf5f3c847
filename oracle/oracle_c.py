"""ctypes binding of oracle/libswalbe_oracle.so (the C restatement) -- TEST INFRASTRUCTURE ONLY.

Same call signatures as oracle_np so the golden cases can run against either.  See the header of
oracle_np.py for what the oracle is pinned against.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libswalbe_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "swalbe_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None
_dp = C.POINTER(C.c_double)


class _State(C.Structure):
    _fields_ = [(n, _dp) for n in ("fout", "ftemp", "feq", "height", "velx", "vely", "vsq", "pressure", "Fx", "Fy",
                                   "slipx", "slipy", "hgradpx", "hgradpy", "dgrad", "kbtx", "kbty")]


class _Params(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("tau", "mu", "delta", "kbt", "gamma", "hmin", "hcrit", "g")] + [
        ("n", C.c_int), ("m", C.c_int), ("cospi_theta", C.c_double), ("cospi_theta_field", _dp),
        ("pressure_variant", C.c_int), ("slip_variant", C.c_int), ("use_inclination", C.c_int),
        ("incl_ax", C.c_double), ("incl_ay", C.c_double), ("incl_factor", C.c_double)]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oracle_filmpressure.restype = C.c_int
        _lib.oracle_step.restype = C.c_int
        _lib.oracle_time_loop.restype = C.c_int
        _lib.oracle_max_threads.restype = C.c_int
        _lib.oracle_time_loop_lowmem.restype = C.c_int
    return _lib


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.f_contiguous, "oracle arrays must be float64, Fortran order"
    return a.ctypes.data_as(_dp)


def _dims(a):
    return C.c_int(a.shape[0]), C.c_int(a.shape[1])


_d = C.c_double
THREADS = 1  # module-level knob: OpenMP threads used by every call below


def equilibrium(feq, h, ux, uy, vsq, g):
    lib().oracle_equilibrium(_p(feq), _p(h), _p(ux), _p(uy), _p(vsq), _d(g), *_dims(h), C.c_int(THREADS))


def BGKandStream(fout, feq, ftemp, Fx, Fy, tau):
    lib().oracle_bgk_stream(_p(fout), _p(feq), _p(ftemp), _p(Fx), _p(Fy), _d(tau), *_dims(Fx), C.c_int(THREADS))


def moments(h, ux, uy, fout):
    lib().oracle_moments(_p(h), _p(ux), _p(uy), _p(fout), *_dims(h), C.c_int(THREADS))


def filmpressure(output, f, dgrad, gamma, cospi_theta, n, m, hmin, hcrit, variant="fast"):
    field = cospi_theta if isinstance(cospi_theta, np.ndarray) else None
    rc = lib().oracle_filmpressure(_p(output), _p(f), _p(dgrad), _d(gamma), _d(0.0 if field is not None else cospi_theta),
                                   _p(field), C.c_int(n), C.c_int(m), _d(hmin), _d(hcrit),
                                   C.c_int(1 if variant == "fast" else 0), *_dims(f), C.c_int(THREADS))
    if rc:
        raise ValueError(f"DomainError({(n, m)})")


def lap9(output, f, gamma):
    d = np.zeros(f.shape + (8,), order="F")
    lib().oracle_lap9(_p(output), _p(f), _p(d), _d(gamma), *_dims(f), C.c_int(THREADS))


def grad9(ox, oy, f, a=None, dgrad=None):
    d = dgrad if dgrad is not None else np.zeros(f.shape + (8,), order="F")
    lib().oracle_grad9(_p(ox), _p(oy), _p(f), _p(d), _p(a), *_dims(f), C.c_int(THREADS))


def hgradp(gx, gy, pressure, height, dgrad):
    lib().oracle_grad9(_p(gx), _p(gy), _p(pressure), _p(dgrad), _p(height), *_dims(height), C.c_int(THREADS))


def slippage(sx, sy, h, ux, uy, delta, mu, hcrit=0.0, variant=0):
    lib().oracle_slippage(_p(sx), _p(sy), _p(h), _p(ux), _p(uy), _d(delta), _d(mu), _d(hcrit), C.c_int(variant),
                          *_dims(h), C.c_int(THREADS))


def thermal(kx, ky, h, kbt, mu, delta, nx, ny):
    lib().oracle_thermal(_p(kx), _p(ky), _p(h), _d(kbt), _d(mu), _d(delta), _p(nx), _p(ny), *_dims(h), C.c_int(THREADS))


def force_sum(Fx, Fy, gx, gy, sx, sy, kx=None, ky=None):
    lib().oracle_force_sum(_p(Fx), _p(Fy), _p(gx), _p(gy), _p(sx), _p(sy), _p(kx), _p(ky), *_dims(Fx), C.c_int(THREADS))


def inclination(Fx, Fy, h, alpha, factor):
    lib().oracle_inclination(_p(Fx), _p(Fy), _p(h), _d(alpha[0]), _d(alpha[1]), _d(factor), *_dims(h), C.c_int(THREADS))


def _mk_state(st):
    s = _State()
    for name, _ in _State._fields_:
        setattr(s, name, _p(getattr(st, name, None)))
    return s


def _mk_params(p, cospi_theta, pvariant, slip_variant, incl):
    from . import oracle_np as onp

    q = _Params()
    for name in ("tau", "mu", "delta", "kbt", "gamma", "hmin", "hcrit", "g", "n", "m"):
        setattr(q, name, getattr(p, name))
    ct = onp.cospi(p.theta) if cospi_theta is None else cospi_theta
    keep = None
    if isinstance(ct, np.ndarray):
        keep = ct
        q.cospi_theta, q.cospi_theta_field = 0.0, _p(ct)
    else:
        q.cospi_theta, q.cospi_theta_field = float(ct), None
    q.pressure_variant = 1 if pvariant == "fast" else 0
    q.slip_variant = slip_variant
    if incl is not None:
        q.use_inclination, q.incl_ax, q.incl_ay, q.incl_factor = 1, incl[0][0], incl[0][1], incl[1]
    return q, keep


def step(st, p, cospi_theta=None, pvariant="power_broad", slip_variant=0, incl=None, threads=None):
    q, _keep = _mk_params(p, cospi_theta, pvariant, slip_variant, incl)
    s = _mk_state(st)
    rc = lib().oracle_step(C.byref(s), C.byref(q), C.c_int(st.Lx), C.c_int(st.Ly), C.c_int(threads or THREADS))
    if rc:
        raise ValueError("DomainError")


def time_loop(st, p, nsteps=None, cospi_theta=None, pvariant="power_broad", slip_variant=0, incl=None,
              log_dh=False, log_wetted=False, hthresh=0.055, threads=None):
    """nsteps iterations of src/simulate.jl:15-22; returns (dh, wetted) logs (None when not requested)."""
    n = p.Tmax if nsteps is None else nsteps
    q, _keep = _mk_params(p, cospi_theta, pvariant, slip_variant, incl)
    s = _mk_state(st)
    dh = np.zeros(n) if log_dh else None
    wet = np.zeros(n, dtype=np.int64) if log_wetted else None
    rc = lib().oracle_time_loop(C.byref(s), C.byref(q), C.c_int(st.Lx), C.c_int(st.Ly), C.c_int(n),
                                dh.ctypes.data_as(_dp) if log_dh else None,
                                wet.ctypes.data_as(C.POINTER(C.c_longlong)) if log_wetted else None,
                                _d(hthresh), C.c_int(threads or THREADS))
    if rc:
        raise ValueError("DomainError")
    return dh, wet


class _LowmemState(C.Structure):
    _fields_ = [(n, _dp) for n in ("height", "velx", "vely", "pressure", "f", "g")]


def time_loop_lowmem(height, velx, vely, f, p, nsteps, cospi_theta=None, pvariant="power_broad", slip_variant=0, incl=None,
                     threads=None):
    """The low-memory fused restatement (22 planes; see swalbe_oracle.c): nsteps iterations of src/simulate.jl:15-22 on
    height/velx/vely (Lx, Ly) and the populations f (Lx, Ly, 9; the reference's ftemp == fout), all updated IN PLACE.
    Returns the pressure field of the last step.  Pinned bit for bit to time_loop() by tests/test_oracle_golden.py."""
    q, _keep = _mk_params(p, cospi_theta, pvariant, slip_variant, incl)
    Lx, Ly = height.shape
    pressure = np.zeros((Lx, Ly), order="F")
    g = np.zeros((Lx, Ly, 9), order="F")
    s = _LowmemState(_p(height), _p(velx), _p(vely), _p(pressure), _p(f), _p(g))
    rc = lib().oracle_time_loop_lowmem(C.byref(s), C.byref(q), C.c_int(Lx), C.c_int(Ly), C.c_int(nsteps),
                                       C.c_int(threads or THREADS))
    if rc:
        raise ValueError("DomainError")
    if nsteps % 2:  # the two population buffers swap every step: the newest set is in g after an odd number of steps
        f[...] = g
    return pressure
