"""The kernels' arithmetic SOURCE (swalbe.jl_b200/csrc/common.cuh: site functions, exact division, Philox, normals)
compiled as plain C++ on the host (tests/host_emulation.cpp) and compared with the oracle bit for bit -- a CPU-side net
under the GPU parity tests: an edit to common.cuh that changes a result bit fails here, without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle_c as oc
from oracle import oracle_np as onp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"
_d, _i, _p = C.c_double, C.c_int, C.c_void_p


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    if not os.path.isfile(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not installed")
    so = str(tmp_path_factory.mktemp("emul") / "libemul.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-DSW_HOST_EMULATION", "-w", "-I", CUDA_INC, "-shared", "-fPIC",
                    os.path.join(ROOT, "tests", "host_emulation.cpp"), "-o", so, "-lm"], check=True)
    lib = C.CDLL(so)
    lib.emul_step.argtypes = [_p] * 6 + [_i, _i] + [_d] * 7 + [_i, _i, _d, _p, _i, _i, _p, _d, C.c_ulonglong, C.c_ulonglong]
    lib.emul_division_mismatches.argtypes = [_p, _p, C.c_long]
    lib.emul_division_mismatches.restype = C.c_long
    lib.emul_philox.argtypes = [_p, _p, _p]
    lib.emul_thermal.argtypes = [_p, _p, _p, C.c_long, C.c_int, _d, _d, _d, C.c_ulonglong, C.c_ulonglong]
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _emul_steps(lib, st, p, nsteps, ct=None, pvariant=0, slip_variant=0):
    Lx, Ly = st.Lx, st.Ly
    scratch = np.zeros(10 * Lx * Ly)
    for _ in range(nsteps):
        rc = lib.emul_step(_ptr(st.height), _ptr(st.velx), _ptr(st.vely), _ptr(st.fout), _ptr(st.ftemp), _ptr(st.pressure),
                           Lx, Ly, p.tau, p.mu, p.delta, p.gamma, p.hmin, p.hcrit, p.g, p.n, p.m, onp.cospi(p.theta),
                           _ptr(ct), pvariant, slip_variant, _ptr(scratch), 0.0, 0, 0)
        assert rc == 0
        st.ftemp[...] = st.fout  # fout == ftemp after every step (src/collide.jl:103)


def _state(Lx, Ly, seed, pops=False):
    rng = np.random.default_rng(seed)
    st = onp.State(Lx, Ly)
    st.height[...] = np.abs(1.0 + 0.2 * rng.standard_normal((Lx, Ly))) + 0.06
    st.velx[...] = 0.01 * rng.standard_normal((Lx, Ly))
    st.vely[...] = 0.01 * rng.standard_normal((Lx, Ly))
    if pops:
        st.ftemp[...] = 0.1 + 0.01 * rng.random((Lx, Ly, 9))
    return st


@pytest.mark.parametrize("Lx,Ly", [(5, 5), (25, 26), (1, 7), (64, 33)])
@pytest.mark.parametrize("kw", [dict(g=-0.001, gamma=0.0005), dict(n=3, m=2, hmin=0.07), dict(n=4, m=2), dict(tau=0.8),
                                dict(tau=1.3, n=3, m=2, hmin=0.07, delta=2.0)])
def test_device_arithmetic_source_equals_oracle(emul, Lx, Ly, kw):
    p = onp.Params(**kw)
    a, b = _state(Lx, Ly, Lx * 100 + Ly, pops=p.tau != 1.0), _state(Lx, Ly, Lx * 100 + Ly, pops=p.tau != 1.0)
    _emul_steps(emul, a, p, 4)
    oc.time_loop(b, p, nsteps=4)
    for name in ("height", "velx", "vely", "pressure", "fout"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name


def test_device_arithmetic_variants(emul):
    """array-form pressure (fast_93 / fast_32), the three slip variants, a contact-angle field"""
    Lx, Ly = 30, 21
    rng = np.random.default_rng(3)
    ct = np.asfortranarray(np.cos(np.pi * (1 / 9 + rng.random((Lx, Ly)) / 36)))
    for kw, pv, sv, field in [(dict(), 1, 0, None), (dict(n=3, m=2, hmin=0.07), 1, 1, None), (dict(), 0, 2, None),
                              (dict(n=3, m=2, hmin=0.07), 0, 0, ct), (dict(), 1, 0, ct)]:
        p = onp.Params(**kw)
        a, b = _state(Lx, Ly, 9), _state(Lx, Ly, 9)
        _emul_steps(emul, a, p, 3, ct=field, pvariant=pv, slip_variant=sv)
        oc.time_loop(b, p, nsteps=3, cospi_theta=field, pvariant="fast" if pv else "power_broad", slip_variant=sv)
        for name in ("height", "velx", "vely", "pressure", "fout"):
            assert np.array_equal(getattr(a, name), getattr(b, name)), (kw, pv, sv, name)


def test_exact_division_helper_on_host(emul):
    """div_exact / div2_exact (shared reciprocal, +-0 fast path, acceptance test) against `/` on every operand class"""
    rng = np.random.default_rng(0)
    n = 400_000
    bits = rng.integers(0, 2 ** 64, size=2 * n, dtype=np.uint64)
    raw = bits.view(np.float64)
    a, b = raw[:n].copy(), raw[n:].copy()
    mid = rng.standard_normal(n) * 10.0 ** rng.integers(-30, 30, n)
    mid2 = rng.standard_normal(n) * 10.0 ** rng.integers(-30, 30, n)
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, 2.2e-308, 1.7e308, 1.0, -1.0, 3.0, 1e-300, 1e300])
    sa, sb = [v.ravel() for v in np.meshgrid(special, special)]
    near = 1.0 + rng.integers(-4, 5, n) * 2.0 ** -52
    with np.errstate(all="ignore"):
        for x, y in ((a, b), (mid, mid2), (sa.copy(), sb.copy()), (np.zeros(n), mid2), (mid * near, mid)):
            x, y = np.ascontiguousarray(x), np.ascontiguousarray(y)
            assert emul.emul_division_mismatches(_ptr(x), _ptr(y), len(x)) == 0


def test_philox_known_answers_on_host(emul):
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        c, k, out = np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32), np.zeros(4, dtype=np.uint32)
        emul.emul_philox(_ptr(c), _ptr(k), _ptr(out))
        assert tuple(int(v) for v in out) == want


def test_thermal_pair_statistics_on_host(emul):
    """thermal!  src/forcing.jl:297-311 through the device source: variance 2 kbt mu 6 h / (2h^2 + 6 h delta + 3 delta^2),
    zero mean, independent components (test/forcing.jl:141-164 asks for 10 %; 1 % here)."""
    n = 1_000_000
    h = np.full(n, 1.3)
    kx, ky = np.zeros(n), np.zeros(n)
    kbt, mu, delta = 1e-6, 1 / 6, 1.0
    emul.emul_thermal(_ptr(kx), _ptr(ky), _ptr(h), n, 500, kbt, mu, delta, 42, 7)
    var = 2 * kbt * mu * 6 * 1.3 / (2 * 1.3 * 1.3 + 6 * 1.3 * delta + 3 * delta * delta)
    for k in (kx, ky):
        assert abs(k.mean()) < 4 * np.sqrt(var / n) and abs(k.var() / var - 1) < 0.01
    assert abs(np.mean(kx * ky)) < 4 * var / np.sqrt(n)
    kx2 = np.zeros(n)
    emul.emul_thermal(_ptr(kx2), _ptr(ky), _ptr(h), n, 500, kbt, mu, delta, 42, 7)
    assert np.array_equal(kx, kx2)  # counter-based: reproducible
