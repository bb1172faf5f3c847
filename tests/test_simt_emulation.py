"""The step KERNELS themselves (csrc/fused.cuh marching kernel in its strict-lean, run-time-option, full and bulk-prefetch
flavours; csrc/tile.cuh) executed on the CPU by a small SIMT emulation (tests/simt_emulation.cpp: one OS thread per
CUDA thread, pthread barrier for __syncthreads, static storage for __shared__) and compared with the oracle bit for bit.
This exercises, without a GPU, the kernels' index logic: strips and halo columns, row cursors with the periodic wrap,
the software pipeline and its ring slots, the fill / steady / drain instantiations, chunk seams, the tile phases."""
import ctypes as C
import ctypes as C_
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle_c as oc
from oracle import oracle_np as onp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"
_d, _i, _p = C.c_double, C.c_int, C.c_void_p
STRICT, OPTS, FULL, BULK, TILE, THERMAL, STRICT224, FM_LEAN, FM_FULL, NS_LDGSTS, NS_BULK, NS_THERMAL = range(12)


class SimtStep(C.Structure):
    _fields_ = ([("flavour", _i)] + [(n, _i) for n in ("Lx", "Ly", "jbeg", "jend", "W", "rows_per_cta", "wrap_y")] +
                [(n, _d) for n in ("tau", "mu", "delta", "gamma", "hmin", "hcrit", "g", "cospi_theta")] +
                [(n, _i) for n in ("n", "m", "pressure_variant", "slip_variant", "use_incl")] +
                [(n, _d) for n in ("incl_ax", "incl_ay", "incl_factor")] +
                [(n, _p) for n in ("h_in", "ux_in", "uy_in", "f_in", "ct_field", "h_out", "ux_out", "uy_out", "f_out", "f_out2",
                                   "pressure", "hgx", "hgy", "slipx", "slipy", "Fx", "Fy", "feq", "vsq")] +
                [("fstride", C.c_size_t), ("kbt", _d), ("seed", C.c_ulonglong), ("step", C.c_ulonglong),
                 ("jglobal0", C.c_longlong), ("Ly_global", C.c_longlong), ("log_min", _p), ("log_max", _p), ("log_wet", _p),
                 ("hthresh", _d), ("fm_prefetch", _i)])


@pytest.fixture(scope="module")
def simt(tmp_path_factory):
    if not os.path.isfile(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not installed")
    so = str(tmp_path_factory.mktemp("simt") / "libsimt.so")
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-DSW_HOST_EMULATION", "-w", "-I", CUDA_INC, "-shared", "-fPIC",
                    "-pthread", os.path.join(ROOT, "tests", "simt_emulation.cpp"), "-o", so, "-lm"], check=True)
    lib = C.CDLL(so)
    lib.simt_step.argtypes = [C.POINTER(SimtStep)]
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _state(Lx, Ly, seed, pops=False):
    rng = np.random.default_rng(seed)
    st = onp.State(Lx, Ly)
    st.height[...] = np.abs(1.0 + 0.2 * rng.standard_normal((Lx, Ly))) + 0.06
    st.velx[...] = 0.01 * rng.standard_normal((Lx, Ly))
    st.vely[...] = 0.01 * rng.standard_normal((Lx, Ly))
    if pops:
        st.ftemp[...] = 0.1 + 0.01 * rng.random((Lx, Ly, 9))
    return st


def _run(simt, st, p, nsteps, flavour, W, rows, ct=None, pvariant=0, slip_variant=0, incl=None, last_full=False):
    """nsteps launches with the moment ping-pong of swalbe_time_loop; the last one optionally through the FULL kernel,
    which also materialises pressure / h∇p / slip / F / feq / vsq and writes ftemp."""
    Lx, Ly = st.Lx, st.Ly
    N = Lx * Ly
    cur = [st.height, st.velx, st.vely]
    alt = [np.zeros_like(a) for a in cur]
    fsrc, fdst = st.ftemp, st.fout
    for s in range(nsteps):
        last = s == nsteps - 1
        q = SimtStep()
        q.flavour = FULL if (last and last_full) else flavour
        q.Lx, q.Ly, q.jbeg, q.jend, q.W, q.rows_per_cta, q.wrap_y = Lx, Ly, 0, Ly, W, rows, 1
        q.tau, q.mu, q.delta, q.gamma, q.hmin, q.hcrit, q.g = p.tau, p.mu, p.delta, p.gamma, p.hmin, p.hcrit, p.g
        q.cospi_theta, q.n, q.m, q.pressure_variant, q.slip_variant = onp.cospi(p.theta), p.n, p.m, pvariant, slip_variant
        if incl is not None:
            q.use_incl, q.incl_ax, q.incl_ay, q.incl_factor = 1, incl[0][0], incl[0][1], incl[1]
        q.h_in, q.ux_in, q.uy_in = (_ptr(a) for a in cur)
        q.h_out, q.ux_out, q.uy_out = (_ptr(a) for a in alt)
        q.ct_field, q.fstride = _ptr(ct), N
        if p.tau == 1.0:
            q.f_in, q.f_out = None, _ptr(st.fout)
            q.f_out2 = _ptr(st.ftemp) if (last and last_full) else None
        else:
            q.f_in, q.f_out = _ptr(fsrc), _ptr(fdst)
        if last and last_full:
            for name, fld in (("pressure", st.pressure), ("hgx", st.hgradpx), ("hgy", st.hgradpy), ("slipx", st.slipx),
                              ("slipy", st.slipy), ("Fx", st.Fx), ("Fy", st.Fy), ("feq", st.feq), ("vsq", st.vsq)):
                setattr(q, name, _ptr(fld))
        assert simt.simt_step(C.byref(q)) == 0
        cur, alt = alt, cur
        fsrc, fdst = fdst, fsrc
    for dst, src in zip((st.height, st.velx, st.vely), cur):
        if dst is not src:
            dst[...] = src
    if p.tau != 1.0 and fsrc is not st.fout:  # newest populations are in the array written last
        st.fout[...] = fsrc


FIELDS = ("height", "velx", "vely", "fout")
AUX = ("pressure", "hgradpx", "hgradpy", "slipx", "slipy", "Fx", "Fy", "feq", "vsq", "ftemp")


def _same(a, b, names):
    for name in names:
        assert np.array_equal(getattr(a, name), getattr(b, name)), name


@pytest.mark.parametrize("Lx,Ly,W,rows", [(25, 26, 120, 26), (25, 26, 9, 5), (150, 40, 120, 13), (130, 9, 64, 64), (5, 5, 4, 2),
                                          (3, 30, 120, 7)])
def test_marching_kernel_strict_lean_on_cpu(simt, Lx, Ly, W, rows):
    for kw in (dict(g=-0.001, gamma=0.0005), dict(n=3, m=2, hmin=0.07)):  # GZ = false / true instantiations
        p = onp.Params(**kw)
        a, b = _state(Lx, Ly, 7), _state(Lx, Ly, 7)
        _run(simt, a, p, 3, STRICT, W, rows)
        oc.time_loop(b, p, nsteps=3)
        _same(a, b, FIELDS)


def test_marching_kernel_full_flavour_materialises_every_field(simt):
    for kw, pops in ((dict(g=-0.001), False), (dict(tau=0.8, n=3, m=2, hmin=0.07), True), (dict(n=4, m=2), False)):
        p = onp.Params(**kw)
        a, b = _state(70, 23, 5, pops), _state(70, 23, 5, pops)
        _run(simt, a, p, 3, FULL if (p.tau != 1.0 or kw.get("n") == 4) else STRICT, 62, 8, last_full=True)
        oc.time_loop(b, p, nsteps=3)
        _same(a, b, FIELDS + (AUX if p.tau == 1.0 else AUX[:-1]))


def test_marching_kernel_runtime_options_on_cpu(simt):
    """OPTS flavour: contact-angle field, slip variants, inclination, array-form pressure"""
    Lx, Ly = 40, 31
    rng = np.random.default_rng(3)
    ct = np.asfortranarray(np.cos(np.pi * (1 / 9 + rng.random((Lx, Ly)) / 36)))
    cases = [(dict(), 0, 0, ct, None), (dict(n=3, m=2, hmin=0.07), 1, 2, None, None),
             (dict(), 1, 1, ct, ([1e-4, -2e-4], 0.75)), (dict(g=-0.002), 0, 0, None, ([3e-4, 0.0], 1.0))]
    for kw, pv, sv, field, incl in cases:
        p = onp.Params(**kw)
        a, b = _state(Lx, Ly, 11), _state(Lx, Ly, 11)
        _run(simt, a, p, 3, OPTS, 36, 11, ct=field, pvariant=pv, slip_variant=sv, incl=incl)
        oc.time_loop(b, p, nsteps=3, cospi_theta=field, pvariant="fast" if pv else "power_broad", slip_variant=sv, incl=incl)
        _same(a, b, FIELDS)


def test_marching_kernel_bulk_prefetch_flavour_on_cpu(simt):
    """cp.async.bulk row segments (thread 0 copies whole rows, split where a strip crosses the periodic x boundary)"""
    p = onp.Params(n=3, m=2, hmin=0.07)
    for Lx, Ly, W, rows in ((128, 12, 120, 12), (200, 10, 100, 4)):
        a, b = _state(Lx, Ly, 2), _state(Lx, Ly, 2)
        _run(simt, a, p, 2, BULK, W, rows)
        oc.time_loop(b, p, nsteps=2)
        _same(a, b, FIELDS)


@pytest.mark.parametrize("Lx,Ly", [(5, 5), (33, 9), (70, 20), (1, 3), (32, 8), (64, 16)])
def test_tile_kernel_on_cpu(simt, Lx, Ly):
    for kw in (dict(g=-0.001), dict(n=3, m=2, hmin=0.07)):
        p = onp.Params(**kw)
        a, b = _state(Lx, Ly, 13), _state(Lx, Ly, 13)
        _run(simt, a, p, 3, TILE, 0, 0)
        oc.time_loop(b, p, nsteps=3)
        _same(a, b, FIELDS)


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    """the device arithmetic source composed on the host (tests/host_emulation.cpp): the reference for the noise path"""
    so = str(tmp_path_factory.mktemp("emul2") / "libemul.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-DSW_HOST_EMULATION", "-w", "-I", CUDA_INC, "-shared", "-fPIC",
                    os.path.join(ROOT, "tests", "host_emulation.cpp"), "-o", so, "-lm"], check=True)
    lib = C.CDLL(so)
    lib.emul_step.argtypes = [_p] * 6 + [_i, _i] + [_d] * 7 + [_i, _i, _d, _p, _i, _i, _p, _d, C.c_ulonglong, C.c_ulonglong]
    return lib


def test_thermal_kernel_noise_plumbing_on_cpu(simt, emul):
    """The marching kernel with in-kernel noise against the same device source composed site by site: seed, step counter
    and the running global-cell cursor (strip offsets, chunk seams, periodic wrap) must select the same Philox block
    for every site."""
    Lx, Ly, kbt, seed = 50, 23, 1e-5, 77
    p = onp.Params(kbt=kbt, n=3, m=2, hmin=0.07)
    a, b = _state(Lx, Ly, 21), _state(Lx, Ly, 21)
    N = Lx * Ly
    cur = [a.height, a.velx, a.vely]
    alt = [np.zeros_like(x) for x in cur]
    scratch = np.zeros(10 * N)
    for s in range(3):
        q = SimtStep()
        q.flavour, q.Lx, q.Ly, q.jbeg, q.jend, q.W, q.rows_per_cta, q.wrap_y = THERMAL, Lx, Ly, 0, Ly, 21, 6, 1
        q.tau, q.mu, q.delta, q.gamma, q.hmin, q.hcrit, q.g = p.tau, p.mu, p.delta, p.gamma, p.hmin, p.hcrit, p.g
        q.cospi_theta, q.n, q.m, q.pressure_variant, q.slip_variant = onp.cospi(p.theta), p.n, p.m, 1, 0
        q.h_in, q.ux_in, q.uy_in = (_ptr(x) for x in cur)
        q.h_out, q.ux_out, q.uy_out = (_ptr(x) for x in alt)
        q.f_out, q.fstride, q.kbt, q.seed, q.step = _ptr(a.fout), N, kbt, seed, 5 + s
        assert simt.simt_step(C.byref(q)) == 0
        cur, alt = alt, cur
        assert emul.emul_step(_ptr(b.height), _ptr(b.velx), _ptr(b.vely), _ptr(b.fout), _ptr(b.ftemp), _ptr(b.pressure), Lx, Ly,
                              p.tau, p.mu, p.delta, p.gamma, p.hmin, p.hcrit, p.g, p.n, p.m, onp.cospi(p.theta), None, 1, 0,
                              _ptr(scratch), kbt, seed, 5 + s) == 0
    assert np.array_equal(cur[0], b.height) and np.array_equal(cur[1], b.velx) and np.array_equal(a.fout, b.fout)
    quiet = _state(Lx, Ly, 21)
    oc.time_loop(quiet, onp.Params(n=3, m=2, hmin=0.07), nsteps=3, pvariant="fast")
    assert not np.array_equal(cur[0], quiet.height)  # (the noise is really there)


def test_slab_launches_with_ghost_rows_on_cpu(simt):
    """What swalbe_dist_time_loop launches on one rank: planes with 3 ghost rows per side handed over at logical row 0
    (wrap_y = 0), two 3-row edge strips and the interior as separate launches; two emulated ranks with the ghost rows
    exchanged in NumPy must reproduce the global oracle bit for bit."""
    Lx, Ly, ranks, GH = 45, 24, 2, 3
    n = Ly // ranks
    p = onp.Params(g=-0.001)
    ref = _state(Lx, Ly, 31)
    h0, ux0, uy0 = ref.height.copy(), ref.velx.copy(), ref.vely.copy()

    def padded(a, r):  # rows [r*n - GH, (r+1)*n + GH) of the periodic global field
        return np.asfortranarray(np.take(a, np.arange(r * n - GH, (r + 1) * n + GH), axis=1, mode="wrap"))

    cur = [[padded(a, r) for a in (h0, ux0, uy0)] for r in range(ranks)]
    fout = [np.zeros((Lx, n, 9), order="F") for _ in range(ranks)]
    for s in range(3):
        nxt = [[np.zeros_like(a) for a in cur[r]] for r in range(ranks)]
        for r in range(ranks):
            off = GH * Lx * 8  # pointers at logical row 0
            for jbeg, jend, rows in ((0, GH, GH), (n - GH, n, GH), (GH, n - GH, 4)):  # edge strips, then the interior
                q = SimtStep()
                q.flavour, q.Lx, q.Ly, q.jbeg, q.jend, q.W, q.rows_per_cta, q.wrap_y = STRICT, Lx, n, jbeg, jend, 40, rows, 0
                q.tau, q.mu, q.delta, q.gamma, q.hmin, q.hcrit, q.g = p.tau, p.mu, p.delta, p.gamma, p.hmin, p.hcrit, p.g
                q.cospi_theta, q.n, q.m = onp.cospi(p.theta), p.n, p.m
                q.h_in, q.ux_in, q.uy_in = (C.c_void_p(a.ctypes.data + off) for a in cur[r])
                q.h_out, q.ux_out, q.uy_out = (C.c_void_p(a.ctypes.data + off) for a in nxt[r])
                q.f_out, q.fstride = _ptr(fout[r]), Lx * n
                assert simt.simt_step(C.byref(q)) == 0
        for r in range(ranks):  # halo exchange: ghost rows from the ring neighbours' owned rows
            lo, hi = (r - 1) % ranks, (r + 1) % ranks
            for k in range(3):
                nxt[r][k][:, :GH] = nxt[lo][k][:, n:n + GH]
                nxt[r][k][:, n + GH:] = nxt[hi][k][:, GH:2 * GH]
        cur = nxt
    oc.time_loop(ref, p, nsteps=3)
    got_h = np.concatenate([cur[r][0][:, GH:GH + n] for r in range(ranks)], axis=1)
    got_f = np.concatenate(fout, axis=1)
    assert np.array_equal(got_h, ref.height) and np.array_equal(got_f, ref.fout)


def test_wider_cta_and_moments_only_steps_on_cpu(simt):
    """CTAs of 224 threads (the width the 8192^2 film step runs with) and steps that do not write the populations"""
    p = onp.Params(n=3, m=2, hmin=0.07)
    a, b = _state(230, 12, 17), _state(230, 12, 17)
    _run(simt, a, p, 2, STRICT224, 216, 5)
    oc.time_loop(b, p, nsteps=2)
    _same(a, b, FIELDS)
    # lazy populations: f_out == NULL on all but the last step
    a, b = _state(40, 17, 19), _state(40, 17, 19)
    N = 40 * 17
    cur, alt = [a.height, a.velx, a.vely], [np.zeros((40, 17), order="F") for _ in range(3)]
    for s in range(4):
        q = SimtStep()
        q.flavour, q.Lx, q.Ly, q.jbeg, q.jend, q.W, q.rows_per_cta, q.wrap_y = STRICT, 40, 17, 0, 17, 33, 6, 1
        q.tau, q.mu, q.delta, q.gamma, q.hmin, q.hcrit, q.g = p.tau, p.mu, p.delta, p.gamma, p.hmin, p.hcrit, p.g
        q.cospi_theta, q.n, q.m = onp.cospi(p.theta), p.n, p.m
        q.h_in, q.ux_in, q.uy_in = (_ptr(x) for x in cur)
        q.h_out, q.ux_out, q.uy_out = (_ptr(x) for x in alt)
        q.f_out, q.fstride = (_ptr(a.fout) if s == 3 else None), N
        assert simt.simt_step(C.byref(q)) == 0
        cur, alt = alt, cur
    oc.time_loop(b, p, nsteps=4)
    assert np.array_equal(cur[0], b.height) and np.array_equal(cur[2], b.vely) and np.array_equal(a.fout, b.fout)


def test_general_tau_slab_with_population_ghost_rows_on_cpu(simt):
    """tau != 1 on a slab: the old populations are read through pointers at logical row 0 of planes with ONE ghost row
    per side (the moments have three), plane stride including the ghosts -- the layout of swalbe_dist_* at tau != 1."""
    Lx, Ly, ranks, GH = 37, 20, 2, 3
    n = Ly // ranks
    p = onp.Params(tau=0.8)
    ref = _state(Lx, Ly, 41, pops=True)
    h0, ux0, uy0, f0 = ref.height.copy(), ref.velx.copy(), ref.vely.copy(), ref.ftemp.copy()
    rows_m = lambda r: np.arange(r * n - GH, (r + 1) * n + GH)  # noqa: E731
    rows_f = lambda r: np.arange(r * n - 1, (r + 1) * n + 1)    # noqa: E731
    mom = [[np.asfortranarray(np.take(a, rows_m(r), axis=1, mode="wrap")) for a in (h0, ux0, uy0)] for r in range(ranks)]
    pop = [np.asfortranarray(np.take(f0, rows_f(r), axis=1, mode="wrap")) for r in range(ranks)]  # (Lx, n+2, 9)
    for s in range(3):
        mom2 = [[np.zeros_like(a) for a in mom[r]] for r in range(ranks)]
        pop2 = [np.zeros_like(a) for a in pop]
        for r in range(ranks):
            mo, fo = GH * Lx * 8, 1 * Lx * 8
            for jbeg, jend, rows in ((0, GH, GH), (n - GH, n, GH), (GH, n - GH, 3)):
                q = SimtStep()
                q.flavour, q.Lx, q.Ly, q.jbeg, q.jend, q.W, q.rows_per_cta, q.wrap_y = FULL, Lx, n, jbeg, jend, 30, rows, 0
                q.tau, q.mu, q.delta, q.gamma, q.hmin, q.hcrit, q.g = p.tau, p.mu, p.delta, p.gamma, p.hmin, p.hcrit, p.g
                q.cospi_theta, q.n, q.m = onp.cospi(p.theta), p.n, p.m
                q.h_in, q.ux_in, q.uy_in = (C.c_void_p(a.ctypes.data + mo) for a in mom[r])
                q.h_out, q.ux_out, q.uy_out = (C.c_void_p(a.ctypes.data + mo) for a in mom2[r])
                q.f_in, q.f_out = C.c_void_p(pop[r].ctypes.data + fo), C.c_void_p(pop2[r].ctypes.data + fo)
                q.fstride = Lx * (n + 2)
                assert simt.simt_step(C.byref(q)) == 0
        for r in range(ranks):
            lo, hi = (r - 1) % ranks, (r + 1) % ranks
            for k in range(3):
                mom2[r][k][:, :GH] = mom2[lo][k][:, n:n + GH]
                mom2[r][k][:, n + GH:] = mom2[hi][k][:, GH:2 * GH]
            pop2[r][:, 0, :] = pop2[lo][:, n, :]
            pop2[r][:, n + 1, :] = pop2[hi][:, 1, :]
        mom, pop = mom2, pop2
    oc.time_loop(ref, p, nsteps=3)
    assert np.array_equal(np.concatenate([mom[r][0][:, GH:GH + n] for r in range(ranks)], axis=1), ref.height)
    assert np.array_equal(np.concatenate([pop[r][:, 1:n + 1, :] for r in range(ranks)], axis=1), ref.fout)


def test_per_step_logs_on_cpu(simt):
    """time_loop(sys, state, Δh) / the wetted! callback: min, max and count(h > thresh) of the height BEFORE each step,
    reduced per thread, per warp (shuffles), per CTA (shared memory) and across CTAs (atomics)"""
    Lx, Ly, nsteps = 70, 19, 4
    p = onp.Params(g=-0.001, gamma=0.0005)
    a, b = _state(Lx, Ly, 23), _state(Lx, Ly, 23)
    mn, mx = np.full(nsteps, np.inf), np.full(nsteps, -np.inf)
    wet = np.zeros(nsteps, dtype=np.uint64)
    cur, alt = [a.height, a.velx, a.vely], [np.zeros((Lx, Ly), order="F") for _ in range(3)]
    for s in range(nsteps):
        q = SimtStep()
        q.flavour, q.Lx, q.Ly, q.jbeg, q.jend, q.W, q.rows_per_cta, q.wrap_y = OPTS, Lx, Ly, 0, Ly, 50, 7, 1
        q.tau, q.mu, q.delta, q.gamma, q.hmin, q.hcrit, q.g = p.tau, p.mu, p.delta, p.gamma, p.hmin, p.hcrit, p.g
        q.cospi_theta, q.n, q.m = onp.cospi(p.theta), p.n, p.m
        q.h_in, q.ux_in, q.uy_in = (_ptr(x) for x in cur)
        q.h_out, q.ux_out, q.uy_out = (_ptr(x) for x in alt)
        q.f_out, q.fstride = _ptr(a.fout), Lx * Ly
        q.log_min, q.log_max = C.c_void_p(mn.ctypes.data + 8 * s), C.c_void_p(mx.ctypes.data + 8 * s)
        q.log_wet, q.hthresh = C.c_void_p(wet.ctypes.data + 8 * s), 1.0
        assert simt.simt_step(C.byref(q)) == 0
        cur, alt = alt, cur
    dh, w = oc.time_loop(b, p, nsteps=nsteps, log_dh=True, log_wetted=True, hthresh=1.0)
    assert np.array_equal(cur[0], b.height)
    assert np.array_equal(mx - mn, np.asarray(dh)) and [int(v) for v in wet] == [int(v) for v in w]


@pytest.mark.parametrize("Lx,Ly", [(33, 9), (70, 20), (5, 3)])
def test_tile_kernel_with_theta_field_on_cpu(simt, Lx, Ly):
    """the opt-in (SWALBE_TILE_THETA=1) tile instantiations compiled for a contact-angle field"""
    rng = np.random.default_rng(5)
    ct = np.asfortranarray(np.cos(np.pi * (1 / 9 + rng.random((Lx, Ly)) / 36)))
    for kw, pv in ((dict(n=3, m=2, hmin=0.07), 0), (dict(g=-0.001), 1)):
        p = onp.Params(**kw)
        a, b = _state(Lx, Ly, 29), _state(Lx, Ly, 29)
        _run(simt, a, p, 3, TILE, 0, 0, ct=ct, pvariant=pv)
        oc.time_loop(b, p, nsteps=3, cospi_theta=ct, pvariant="fast" if pv else "power_broad")
        _same(a, b, FIELDS)



@pytest.mark.parametrize("Lx,Ly,W,rows", [(25, 26, 120, 26), (150, 40, 120, 13), (70, 23, 62, 8), (5, 5, 4, 2), (3, 30, 120, 7)])
def test_general_tau_from_moments_flavour_on_cpu(simt, Lx, Ly, W, rows):
    """tau != 1, FM kernels: the first step reads the caller's h / u planes (which are NOT the moments of ftemp -- an
    initial condition), every later step derives h from the nine old populations it loads for the h ring and u from the
    populations it collides; the moment planes are written by the last step only.  This is the launch sequence of
    swalbe_time_loop at tau != 1 (csrc/fused.cu), compared with the oracle bit for bit, materialised fields included."""
    for kw, last_full, pf in ((dict(tau=0.8, n=3, m=2, hmin=0.07), True, 3), (dict(tau=1.3, g=-0.001, gamma=0.0005), False, 0)):
        p = onp.Params(**kw)
        a, b = _state(Lx, Ly, 47, pops=True), _state(Lx, Ly, 47, pops=True)
        nsteps, N = 4, Lx * Ly
        fsrc, fdst = a.ftemp, a.fout
        for s in range(nsteps):
            last = s == nsteps - 1
            q = SimtStep()
            q.flavour = FULL if s == 0 else (FM_FULL if (last and last_full) else FM_LEAN)
            q.Lx, q.Ly, q.jbeg, q.jend, q.W, q.rows_per_cta, q.wrap_y = Lx, Ly, 0, Ly, W, rows, 1
            q.tau, q.mu, q.delta, q.gamma, q.hmin, q.hcrit, q.g = p.tau, p.mu, p.delta, p.gamma, p.hmin, p.hcrit, p.g
            q.cospi_theta, q.n, q.m, q.fm_prefetch = onp.cospi(p.theta), p.n, p.m, pf
            q.h_in, q.ux_in, q.uy_in = _ptr(a.height), _ptr(a.velx), _ptr(a.vely)  # (read by step 0 only)
            if last:
                q.h_out, q.ux_out, q.uy_out = _ptr(a.height), _ptr(a.velx), _ptr(a.vely)
            q.f_in, q.f_out, q.fstride = _ptr(fsrc), _ptr(fdst), N
            if last and last_full:
                for name, fld in (("pressure", a.pressure), ("hgx", a.hgradpx), ("hgy", a.hgradpy), ("slipx", a.slipx),
                                  ("slipy", a.slipy), ("Fx", a.Fx), ("Fy", a.Fy), ("feq", a.feq), ("vsq", a.vsq)):
                    setattr(q, name, _ptr(fld))
            assert simt.simt_step(C.byref(q)) == 0
            fsrc, fdst = fdst, fsrc
        if fsrc is not a.fout:
            a.fout[...] = fsrc
        oc.time_loop(b, p, nsteps=nsteps)
        _same(a, b, FIELDS + (AUX[:-1] if last_full else ()))


@pytest.mark.parametrize("flavour,Lx,Ly,W,rows", [(NS_LDGSTS, 25, 26, 120, 26), (NS_LDGSTS, 150, 40, 120, 13), (NS_LDGSTS, 130, 9, 64, 64),
                                                  (NS_LDGSTS, 5, 5, 4, 2), (NS_BULK, 128, 12, 120, 12), (NS_BULK, 200, 10, 100, 4),
                                                  (NS_BULK, 64, 14, 120, 5)])
def test_neighbour_sync_flavour_on_cpu(simt, flavour, Lx, Ly, W, rows):
    """NS kernels: no CTA barrier in the row loop, each warp hands over to its two neighbour warps through mbarriers (here:
    real atomics between OS threads that do run ahead of each other), rows prefetched per thread or per warp."""
    for kw in (dict(g=-0.001, gamma=0.0005), dict(n=3, m=2, hmin=0.07)):
        p = onp.Params(**kw)
        a, b = _state(Lx, Ly, 7), _state(Lx, Ly, 7)
        _run(simt, a, p, 3, flavour, W, rows)
        oc.time_loop(b, p, nsteps=3)
        _same(a, b, FIELDS)


@pytest.mark.parametrize("Lx,Ly,C,nsteps", [(25, 26, 4, 5), (100, 100, 8, 3), (33, 12, 4, 4), (40, 17, 2, 6), (9, 7, 1, 3), (64, 48, 16, 2),
                                            (30, 30, 4, 1)])
def test_persistent_cluster_kernel_on_cpu(simt, Lx, Ly, C, nsteps):
    """k_cluster_steps (csrc/cluster.cuh): all steps of a call inside one launch on a thread-block cluster whose CTAs own
    row slabs in shared memory and read each other's halo rows (distributed shared memory) behind one cluster barrier per
    step.  Emulated with all CTAs of the cluster running concurrently; uneven slabs (Ly % C != 0), in place on the caller's
    planes, per-step logs, populations every step or on the last one only -- against the oracle, bit for bit."""
    simt.simt_cluster_steps.argtypes = [C_.POINTER(SimtStep), C_.c_int, C_.c_int, C_.c_int]
    for kw, lazy in ((dict(g=-0.001, gamma=0.0005), 0), (dict(n=3, m=2, hmin=0.07), 1)):
        p = onp.Params(**kw)
        a, b = _state(Lx, Ly, 61), _state(Lx, Ly, 61)
        mn, mx = np.full(nsteps, np.inf), np.full(nsteps, -np.inf)
        wet = np.zeros(nsteps, dtype=np.uint64)
        q = SimtStep()
        q.Lx, q.Ly, q.jbeg, q.jend, q.wrap_y = Lx, Ly, 0, Ly, 1
        q.tau, q.mu, q.delta, q.gamma, q.hmin, q.hcrit, q.g = p.tau, p.mu, p.delta, p.gamma, p.hmin, p.hcrit, p.g
        q.cospi_theta, q.n, q.m = onp.cospi(p.theta), p.n, p.m
        q.h_in, q.ux_in, q.uy_in = _ptr(a.height), _ptr(a.velx), _ptr(a.vely)
        q.h_out, q.ux_out, q.uy_out = _ptr(a.height), _ptr(a.velx), _ptr(a.vely)  # in place
        q.f_out, q.f_out2, q.fstride = _ptr(a.fout), _ptr(a.ftemp), Lx * Ly
        q.log_min, q.log_max, q.log_wet, q.hthresh = _ptr(mn), _ptr(mx), _ptr(wet), 1.0
        assert simt.simt_cluster_steps(C_.byref(q), nsteps, C, lazy) == 0
        dh, w = oc.time_loop(b, p, nsteps=nsteps, log_dh=True, log_wetted=True, hthresh=1.0)
        _same(a, b, FIELDS + ("ftemp",))
        assert np.array_equal(mx - mn, np.asarray(dh)) and [int(v) for v in wet] == [int(v) for v in w]


# ---- swalbe_time_loop_host: the skewed sweeps behind the upload front / ahead of the download (csrc/sweep.h) ------------

def _host_loop_ops(Lx, Ly, nsteps, has_in, has_out, band_rows, kmax):
    from swalbe_b200 import _lib

    lib = _lib.load()
    n = C.c_int(0)
    _lib.call("swalbe_selftest_host_loop_schedule", Lx, Ly, nsteps, int(has_in), int(has_out), band_rows, kmax, 1, None, 0, C.byref(n))
    buf = (C.c_int * (7 * n.value))()
    _lib.call("swalbe_selftest_host_loop_schedule", Lx, Ly, nsteps, int(has_in), int(has_out), band_rows, kmax, 1, buf, n.value,
              C.byref(n))
    assert lib is not None
    return [tuple(buf[7 * q:7 * q + 7]) for q in range(n.value)]


def _columns_order(ops):
    """The other extreme of what the two compute streams of a sweep allow: inside every sweep, step k of ALL band stages
    before step k+1 of any (the schedule lists stage after stage).  Legal because launch (stage b, k) is only ordered
    after (b, k-1) and (b-1, k-1); uploads first, downloads last."""
    out, run = [], []

    def flush():
        if run:
            kmin = min(o[1] for o in run)
            out.extend(sorted(run, key=lambda o: (o[1] - kmin, o[6])))
            run.clear()

    for o in ops:
        if o[0] != 1:
            continue
        if o[6] >= 0:
            if run and o[6] == 0 and run[-1][6] != 0:  # a new sweep starts
                flush()
            run.append(o)
        else:
            flush()
            out.append(o)
    flush()
    return [o for o in ops if o[0] == 0] + out + [o for o in ops if o[0] == 2]


def _replay_host_loop(simt, ops, st, p, nsteps, h_host, out_host, early_copies, lazy, W=40, rows=16):
    """What enqueue_steps_host (csrc/fused.cu) issues, on NumPy planes: A = the state's planes, B = the plan's scratch;
    `early_copies`: every upload happens before the first launch and every download after the last one (the other legal
    extreme of the stream order); else they happen exactly where the schedule lists them."""
    Lx, Ly = st.Lx, st.Ly
    A = [st.height, st.velx, st.vely]
    B = [np.full_like(st.height, np.nan), np.full_like(st.height, np.nan), np.full_like(st.height, np.nan)]
    src0 = A if nsteps % 2 == 0 else B
    if h_host is not None:
        src0[0][...] = np.nan  # rows that have not arrived yet must never be used
    elif src0 is B:
        B[0][...] = A[0]
    if src0 is B:
        B[1][...], B[2][...] = A[1], A[2]
    ups = [o for o in ops if o[0] == 0]
    downs = [o for o in ops if o[0] == 2]
    if early_copies:
        for _, _, j0, j1, _, _, _ in ups:
            src0[0][:, j0:j1] = h_host[:, j0:j1]
    arrived = set()
    for kind, s, j0, j1, band, seam, _stage in ops:
        if kind == 0:
            if not early_copies:
                src0[0][:, j0:j1] = h_host[:, j0:j1]
            arrived.add(band)
        elif kind == 2:
            if not early_copies:
                out_host[:, j0:j1] = A[0][:, j0:j1]
        else:
            assert band < 0 or band in arrived, "a launch waits for a band that is queued after it"
            if j1 <= j0:
                continue
            last = s == nsteps - 1
            reads_A = (nsteps - s) % 2 == 0
            src, dst = (A, B) if reads_A else (B, A)
            q = SimtStep()
            q.flavour = FULL if last else STRICT
            q.Lx, q.Ly, q.jbeg, q.jend, q.W, q.rows_per_cta, q.wrap_y = Lx, Ly, j0, j1, W, rows, 1
            q.tau, q.mu, q.delta, q.gamma, q.hmin, q.hcrit, q.g = p.tau, p.mu, p.delta, p.gamma, p.hmin, p.hcrit, p.g
            q.cospi_theta, q.n, q.m, q.pressure_variant, q.slip_variant = onp.cospi(p.theta), p.n, p.m, 0, 0
            q.h_in, q.ux_in, q.uy_in = (_ptr(a) for a in src)
            q.h_out, q.ux_out, q.uy_out = (_ptr(a) for a in dst)
            q.fstride = Lx * Ly
            q.f_out = _ptr(st.fout) if (last or not lazy) else None
            if last:
                q.f_out2 = _ptr(st.ftemp)
                for name, fld in (("pressure", st.pressure), ("hgx", st.hgradpx), ("hgy", st.hgradpy), ("slipx", st.slipx),
                                  ("slipy", st.slipy), ("Fx", st.Fx), ("Fy", st.Fy), ("feq", st.feq), ("vsq", st.vsq)):
                    setattr(q, name, _ptr(fld))
            assert simt.simt_step(C.byref(q)) == 0
    if early_copies:
        for _, _, j0, j1, _, _, _ in downs:
            out_host[:, j0:j1] = A[0][:, j0:j1]


@pytest.mark.parametrize("nsteps,has_in,has_out", [(3, True, True), (11, True, True), (5, False, True), (9, True, False)])
def test_host_loop_sweeps_on_cpu(simt, nsteps, has_in, has_out):
    """Every launch of the schedule through the emulated kernels (partial row ranges with the periodic wrap), two moment
    buffers, rows that have not been uploaded poisoned with NaN: the state and the downloaded plane equal the oracle's
    after the same number of steps, with the copies replayed at both extremes of what the streams allow."""
    Lx, Ly, band, kmax = 40, 126, 31, 4
    ops = _host_loop_ops(Lx, Ly, nsteps, has_in, has_out, band, kmax)
    steps = [o for o in ops if o[0] == 1]
    assert any(o[5] for o in steps) and len([o for o in ops if o[0] == 0]) == (4 if has_in else 0)
    # every step covers every row exactly once
    for s in range(nsteps):
        cover = np.zeros(Ly, dtype=int)
        for _, ss, j0, j1, _, _, _ in steps:
            if ss == s and j1 > j0:
                cover[j0:j1] += 1
        assert (cover == 1).all(), (s, cover)
    if has_out:
        cover = np.zeros(Ly, dtype=int)
        for o in ops:
            if o[0] == 2:
                cover[o[2]:o[3]] += 1
        assert (cover == 1).all()
    p = onp.Params(g=-0.001, gamma=0.0005)
    ref = _state(Lx, Ly, 11)
    h0 = ref.height.copy()
    oc.time_loop(ref, p, nsteps=nsteps)
    cols = _columns_order(ops)
    assert sorted(cols) == sorted(ops) and cols != ops
    for early, lazy, order in ((False, False, ops), (True, True, cols)):
        st = _state(Lx, Ly, 11)
        if has_in:
            st.height[...] = 7.0  # the device plane holds something else: the job starts from the host plane
        out = np.full((Lx, Ly), np.nan)
        _replay_host_loop(simt, order, st, p, nsteps, h0 if has_in else None, out if has_out else None, early, lazy)
        _same(st, ref, FIELDS + AUX)
        if has_out:
            assert np.array_equal(out, ref.height)


@pytest.mark.parametrize("nsteps,has_in,has_out", [(3, True, True), (9, True, True), (5, False, True)])
def test_slab_host_loop_sweeps_on_cpu(simt, nsteps, has_in, has_out):
    """swalbe_dist_time_loop_host (csrc/dist.cu) on two emulated ranks: the same schedule on the rows of one slab, planes
    with ghost rows (wrap_y = 0); band stages never touch a ghost row, the strips next to the slab boundaries are stepped
    after the last band with one halo exchange per strip pair; whole-slab steps in between exchange after every step.
    Rows that have not been uploaded and ghost rows that have not been exchanged are NaN.  Global oracle, bit for bit."""
    Lx, n, ranks, GH, band, kmax = 40, 66, 2, 3, 22, 3
    Ly = n * ranks
    ops = _host_loop_ops(Lx, n, nsteps, has_in, has_out, band, kmax)
    assert any(o[5] for o in ops) and any(o[6] >= 0 for o in ops)
    p = onp.Params(g=-0.001, gamma=0.0005)
    ref = _state(Lx, Ly, 23)
    h0, ux0, uy0 = ref.height.copy(), ref.velx.copy(), ref.vely.copy()
    oc.time_loop(ref, p, nsteps=nsteps)

    def padded(a, r, ghosts=True):
        out = np.asfortranarray(np.take(a, np.arange(r * n - GH, (r + 1) * n + GH), axis=1, mode="wrap"))
        if not ghosts:
            out[:, :GH] = np.nan; out[:, n + GH:] = np.nan
        return out

    # two moment sets per rank; state 0 lives in set 0.  With an upload only the velocities are there (their ghost rows
    # too: zero / caller-provided slabs are exchanged with the height at the first exchange -- here they start un-exchanged)
    sets = [[[padded(a, r, ghosts=not has_in) for a in (h0, ux0, uy0)], [np.full((Lx, n + 2 * GH), np.nan, order="F") for _ in range(3)]]
            for r in range(ranks)]
    if has_in:
        for r in range(ranks):
            sets[r][0][0][...] = np.nan
    fout = [np.zeros((Lx, n, 9), order="F") for _ in range(ranks)]
    out = [np.full((Lx, n), np.nan) for _ in range(ranks)]

    def exchange(si):
        for r in range(ranks):
            lo, hi = (r - 1) % ranks, (r + 1) % ranks
            for k in range(3):
                sets[r][si][k][:, :GH] = sets[lo][si][k][:, n:n + GH]
                sets[r][si][k][:, n + GH:] = sets[hi][si][k][:, GH:2 * GH]

    def launch(r, s, j0, j1, rows):
        src, dst = sets[r][s & 1], sets[r][(s & 1) ^ 1]
        off = GH * Lx * 8
        q = SimtStep()
        q.flavour, q.Lx, q.Ly, q.jbeg, q.jend, q.W, q.rows_per_cta, q.wrap_y = STRICT, Lx, n, j0, j1, 40, rows, 0
        q.tau, q.mu, q.delta, q.gamma, q.hmin, q.hcrit, q.g = p.tau, p.mu, p.delta, p.gamma, p.hmin, p.hcrit, p.g
        q.cospi_theta, q.n, q.m = onp.cospi(p.theta), p.n, p.m
        q.h_in, q.ux_in, q.uy_in = (C.c_void_p(a.ctypes.data + off) for a in src)
        q.h_out, q.ux_out, q.uy_out = (C.c_void_p(a.ctypes.data + off) for a in dst)
        q.f_out, q.fstride = _ptr(fout[r]), Lx * n
        assert simt.simt_step(C.byref(q)) == 0

    ghosts_current, seam_launches = not has_in, 0
    for kind, s, j0, j1, b, seam, stage in ops:
        if kind == 0:
            for r in range(ranks):
                sets[r][0][0][:, GH + j0:GH + j1] = h0[:, r * n + j0:r * n + j1]
        elif kind == 2:
            for r in range(ranks):
                out[r][:, j0:j1] = sets[r][nsteps & 1][0][:, GH + j0:GH + j1]
        elif stage < 0 and not seam:  # whole-slab step of the ordinary loop: edge strips, interior, exchange
            if not ghosts_current:
                exchange(s & 1); ghosts_current = True
            for r in range(ranks):
                for a, c, rows in ((0, GH, GH), (n - GH, n, GH), (GH, n - GH, 16)):
                    launch(r, s, a, c, rows)
            exchange((s & 1) ^ 1)
        elif seam:
            if not ghosts_current:
                exchange(s & 1); ghosts_current = True
            for r in range(ranks):
                launch(r, s, j0, j1, 4)
            seam_launches += 1
            if seam_launches % 2 == 0:
                exchange((s & 1) ^ 1)
        elif j1 > j0:
            for r in range(ranks):
                launch(r, s, j0, j1, 16)
    fin = nsteps & 1
    got_h = np.concatenate([sets[r][fin][0][:, GH:GH + n] for r in range(ranks)], axis=1)
    assert np.array_equal(got_h, ref.height)
    assert np.array_equal(np.concatenate(fout, axis=1), ref.fout)
    for r in range(ranks):  # the runtime's invariant: the ghost rows of the current set are current on return
        lo, hi = (r - 1) % ranks, (r + 1) % ranks
        assert np.array_equal(sets[r][fin][0][:, :GH], sets[lo][fin][0][:, n:n + GH])
        assert np.array_equal(sets[r][fin][0][:, n + GH:], sets[hi][fin][0][:, GH:2 * GH])
    if has_out:
        assert np.array_equal(np.concatenate(out, axis=1), ref.height)
