"""ctypes binding of libswalbe_b200.so (the C ABI in include/swalbe_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C swalbe.jl_b200/csrc``.  There is no
fallback: if the shared object is missing or a CUDA call fails, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("SWALBE_B200_SO") or os.path.join(_HERE, "libswalbe_b200.so")  # (override: A/B builds)

OK, ERR_DOMAIN, ERR_EXTENT, ERR_CUDA, ERR_NCCL, ERR_ARG = range(6)

PRESSURE_POWER_BROAD, PRESSURE_FAST = 0, 1
SLIP_STANDARD, SLIP_HCRIT, SLIP_RING_RIV = 0, 1, 2
LOOP_DEFAULT, LOOP_LAZY_POPULATIONS, LOOP_SKIP_AUX, LOOP_MOMENTS_CONSISTENT = 0, 1, 2, 4
LOOP_GAMMA_FIELD, LOOP_MARANGONI = 8, 16  # swalbe_time_loop_1d
NCCL_UNIQUE_ID_BYTES = 128


class DomainError(ValueError):
    """Julia's DomainError for unsupported (n, m) in the array-form filmpressure! (src/pressure.jl:101-107)."""


class SwalbeError(RuntimeError):
    pass


_vp = C.c_void_p
_d = C.c_double
_i = C.c_int
_u64 = C.c_ulonglong


class CState(C.Structure):
    """struct swalbe_state"""
    _fields_ = [(n, _vp) for n in ("fout", "ftemp", "feq", "height", "velx", "vely", "vsq", "pressure", "Fx", "Fy",
                                   "slipx", "slipy", "hgradpx", "hgradpy", "dgrad", "kbtx", "kbty")]


class CParams(C.Structure):
    """struct swalbe_params"""
    _fields_ = [(n, _d) for n in ("tau", "mu", "delta", "kbt", "gamma", "hmin", "hcrit", "g")] + [
        ("n", _i), ("m", _i), ("cospi_theta", _d), ("cospi_theta_field", _vp), ("pressure_variant", _i),
        ("slip_variant", _i), ("use_inclination", _i), ("incl_ax", _d), ("incl_ay", _d), ("incl_factor", _d),
        ("use_thermal", _i), ("seed", _u64)]


class CState1D(C.Structure):
    """struct swalbe_state_1d"""
    _fields_ = [(n, _vp) for n in ("fout", "ftemp", "feq", "height", "vel", "pressure", "F", "slip", "hgradp", "dgrad",
                                   "gamma", "dgamma", "kbt", "fbound")]


class CLogs(C.Structure):
    """struct swalbe_loop_logs"""
    _fields_ = [("hmin", _vp), ("hmax", _vp), ("wetted", _vp), ("hthresh", _d), ("hsum", _vp), ("hsum_first", _i),
                ("hsum_every", _i)]


# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/swalbe_b200.h one to one
SIGNATURES = {
    "swalbe_version": [],
    "swalbe_last_error": [],
    "swalbe_launch_count": [],
    "swalbe_equilibrium_d2q9": [_vp, _vp, _vp, _vp, _vp, _d, _i, _i, _vp],
    "swalbe_bgk_stream_d2q9": [_vp, _vp, _vp, _vp, _vp, _d, _i, _i, _vp],
    "swalbe_moments_d2q9": [_vp, _vp, _vp, _vp, _i, _i, _vp],
    "swalbe_filmpressure": [_vp, _vp, _vp, _d, _d, _vp, _i, _i, _d, _d, _i, _i, _i, _vp],
    "swalbe_hgradp": [_vp, _vp, _vp, _vp, _i, _i, _vp],
    "swalbe_grad9": [_vp, _vp, _vp, _vp, _i, _i, _vp],
    "swalbe_lap9": [_vp, _vp, _d, _i, _i, _vp],
    "swalbe_slippage": [_vp, _vp, _vp, _vp, _vp, _d, _d, _d, _i, _i, _i, _vp],
    "swalbe_force_sum": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "swalbe_thermal": [_vp, _vp, _vp, _d, _d, _d, _u64, _u64, _i, _i, _vp],
    "swalbe_inclination": [_vp, _vp, _vp, _d, _d, _d, _i, _i, _vp],
    "swalbe_field_stats": [_vp, _vp, _d, _i, _i, _vp],
    "swalbe_selftest_division": [_u64, _u64, _vp, _vp],
    "swalbe_selftest_philox": [C.POINTER(C.c_uint * 4), C.POINTER(C.c_uint * 2), _vp, _vp],
    "swalbe_cospi_field": [_vp, _vp, C.c_size_t, _vp],
    "swalbe_ic_singledroplet": [_vp, _d, _d, _d, _d, _d, _i, _i, _i, _vp],
    "swalbe_ic_torus": [_vp, _d, _d, _d, _d, _d, _d, _d, _u64, _i, _i, _i, _vp],
    "swalbe_ic_rivulet": [_vp, _d, _d, _i, _d, _d, _d, _u64, _i, _i, _i, _vp],
    "swalbe_ic_sinewave2d": [_vp, _d, _d, _d, _d, _i, _i, _i, _i, _vp],
    "swalbe_ic_randinterface": [_vp, _d, _d, _u64, _i, _i, _i, _vp],
    "swalbe_circshift": [_vp, _vp, _i, _i, _i, _i, _vp],
    "swalbe_plan_create": [C.POINTER(_vp), _i, _i],
    "swalbe_plan_destroy": [_vp],
    "swalbe_time_loop": [_vp, C.POINTER(CState), C.POINTER(CParams), _i, _u64, _i, C.POINTER(CLogs), _vp],
    "swalbe_time_loop_host": [_vp, C.POINTER(CState), C.POINTER(CParams), _i, _u64, _i, C.POINTER(CLogs), _vp, _vp, _vp],
    "swalbe_selftest_host_loop_schedule": [_i, _i, _i, _i, _i, _i, _i, _i, C.POINTER(_i), _i, C.POINTER(_i)],
    "swalbe_equilibrium_d1q3": [_vp, _vp, _vp, _d, _i, _vp],
    "swalbe_bgk_stream_d1q3": [_vp, _vp, _vp, _vp, _d, _i, _vp],
    "swalbe_moments_d1q3": [_vp, _vp, _vp, _i, _vp],
    "swalbe_filmpressure_1d": [_vp, _vp, _vp, _d, _d, _vp, _i, _i, _d, _d, _i, _i, _vp],
    "swalbe_grad_1d": [_vp, _vp, _vp, _i, _vp],
    "swalbe_lap_1d": [_vp, _vp, _i, _vp],
    "swalbe_slippage_1d": [_vp, _vp, _vp, _d, _d, _i, _vp],
    "swalbe_force_sum_1d": [_vp, _vp, _vp, _i, _vp],
    "swalbe_force_sum3_1d": [_vp, _vp, _vp, _vp, _i, _vp],
    "swalbe_thermal_1d": [_vp, _vp, _d, _d, _d, _u64, _u64, _i, _vp],
    "swalbe_inclination_1d": [_vp, _vp, _d, _d, _i, _vp],
    "swalbe_gradgamma_1d": [_vp, _vp, _vp, _d, _i, _vp],
    "swalbe_filmpressure_gamma_1d": [_vp, _vp, _d, _vp, _vp, _d, _d, _vp, _i, _i, _d, _d, _vp, _i, _vp],
    "swalbe_bgk_stream_bound_d1q3": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _d, _i, _vp],
    "swalbe_update_rho_1d": [_vp, _vp, _vp, _vp, _d, _d, _i, _vp],
    "swalbe_time_loop_1d": [C.POINTER(CState1D), C.POINTER(CParams), _i, _i, _i, C.POINTER(CLogs), _vp],
    "swalbe_dist_unique_id": [_vp],
    "swalbe_dist_create": [C.POINTER(_vp), _vp, _i, _i, _i, _i, C.POINTER(CParams)],
    "swalbe_dist_destroy": [_vp],
    "swalbe_dist_local_rows": [_vp, C.POINTER(_i), C.POINTER(_i)],
    "swalbe_dist_set_state": [_vp, _vp, _vp, _vp, _vp, _vp],
    "swalbe_dist_set_theta": [_vp, _vp, _vp],
    "swalbe_dist_shift_theta": [_vp, _i, _i, _vp],
    "swalbe_dist_height_stats": [_vp, _vp, _d, _vp],
    "swalbe_dist_time_loop": [_vp, _i, _u64, _vp],
    "swalbe_dist_time_loop_host": [_vp, _i, _u64, _vp, _vp, _vp, _vp, _vp],
    "swalbe_dist_get_state": [_vp, _vp, _vp, _vp, _vp, _vp],
    "swalbe_dist_uses_peer_memory": [_vp, C.POINTER(_i)],
    "swalbe_dist_last_loop_ms": [_vp, C.POINTER(C.c_float)],
}
_RESTYPES = {"swalbe_last_error": C.c_char_p, "swalbe_launch_count": _u64}

_lib = None


def load() -> C.CDLL:
    """Load libswalbe_b200.so (no fallback: raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise SwalbeError(
                f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback on this path)")
        lib = C.CDLL(SO_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here == ABI symbol missing
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, _i)
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc == OK:
        return
    msg = load().swalbe_last_error().decode("utf-8", "replace")
    if rc == ERR_DOMAIN:
        raise DomainError(msg)
    if rc in (ERR_ARG, ERR_EXTENT):
        raise ValueError(msg)
    raise SwalbeError(f"libswalbe_b200 error {rc}: {msg}")


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args))
