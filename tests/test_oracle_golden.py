"""Pins the oracle (NumPy and C restatements) against the reference's own known-answer tests.

Transcribed from /root/reference/test/*.jl and the jldoctests in /root/reference/src/*.jl (see
tests/golden_cases.py for per-case citations).  CPU only.
"""
import numpy as np
import pytest

from oracle import oracle_c as oc
from oracle import oracle_np as onp
from tests import golden_cases as gc

BACKENDS = {"numpy": onp, "c": oc}


@pytest.mark.parametrize("backend", sorted(BACKENDS))
@pytest.mark.parametrize("case", gc.ALL_OPERATOR_CASES, ids=lambda c: c.__name__)
def test_operator_known_answers(case, backend):
    case(BACKENDS[backend])


def test_power_broad():
    gc.case_power_broad()


def test_cospi_exact_points():
    assert onp.cospi(0.0) == 1.0 and onp.cospi(0.5) == 0.0 and onp.cospi(1.0) == -1.0
    assert onp.cospi(1 / 3) == pytest.approx(0.5, abs=1e-16)
    assert onp.cospi(1 / 9) == pytest.approx(0.9396926207859084, abs=2e-16)


@pytest.mark.parametrize("backend", sorted(BACKENDS))
def test_thermal_statistics(backend):
    """test/forcing.jl:141-164 with NumPy normals standing in for Julia's randn! (statistical only)."""
    B = BACKENDS[backend]
    rng = np.random.default_rng(1234)

    def draw(kb):
        kx, ky = onp.zeros(50, 50), onp.zeros(50, 50)
        nx = np.asfortranarray(rng.standard_normal((50, 50)))
        ny = np.asfortranarray(rng.standard_normal((50, 50)))
        B.thermal(kx, ky, onp.ones(50, 50), kb, 1 / 6, 1.0, nx, ny)
        return kx, ky

    gc.case_thermal_statistics(B, draw)


# ------------------------------------------------------------------------------------------------
# NumPy and C restatements must agree bit for bit (they were written independently from Appendix A)


def _random_state(Lx, Ly, seed, thermal=False):
    rng = np.random.default_rng(seed)
    st = onp.State(Lx, Ly, thermal=thermal)
    st.height[...] = 1.0 + 0.3 * rng.standard_normal((Lx, Ly))
    st.height[...] = np.abs(st.height) + 0.06
    st.velx[...] = 0.05 * rng.standard_normal((Lx, Ly))
    st.vely[...] = 0.05 * rng.standard_normal((Lx, Ly))
    st.ftemp[...] = 0.1 * rng.random((Lx, Ly, 9))
    return st


def _copy_state(st):
    c = onp.State(st.Lx, st.Ly, thermal=hasattr(st, "kbtx"))
    for k, v in vars(st).items():
        if isinstance(v, np.ndarray):
            getattr(c, k)[...] = v
    return c


FIELDS = ("fout", "ftemp", "feq", "height", "velx", "vely", "vsq", "pressure", "Fx", "Fy", "slipx", "slipy",
          "hgradpx", "hgradpy", "dgrad")


@pytest.mark.parametrize("Lx,Ly", [(5, 5), (25, 26), (33, 7)])
@pytest.mark.parametrize("tau", [1.0, 0.75])
@pytest.mark.parametrize("nm,pvariant", [((9, 3), "power_broad"), ((3, 2), "fast"), ((9, 3), "fast"),
                                          ((4, 2), "power_broad")])
def test_c_matches_numpy_bitwise(Lx, Ly, tau, nm, pvariant):
    p = onp.Params(tau=tau, n=nm[0], m=nm[1], g=-0.001, gamma=0.01, delta=1.5, hmin=0.07)
    a = _random_state(Lx, Ly, seed=Lx * 100 + Ly)
    b = _copy_state(a)
    rng = np.random.default_rng(7)
    ct_field = np.asfortranarray(np.cos(np.pi * (1 / 9 + 1 / 36 * rng.random((Lx, Ly)))))
    for it, (ct, sv, incl) in enumerate([(None, 0, None), (ct_field, 1, None), (0.5, 2, ([1e-4, -2e-4], 0.75))]):
        for _ in range(3):
            onp.step(a, p, cospi_theta=ct, pvariant=pvariant, slip_variant=sv, incl=incl)
            oc.step(b, p, cospi_theta=ct, pvariant=pvariant, slip_variant=sv, incl=incl)
        for name in FIELDS:
            x, y = getattr(a, name), getattr(b, name)
            assert np.array_equal(x, y, equal_nan=True), (it, name, np.abs(x - y).max())


def test_c_threads_do_not_change_bits():
    p = onp.Params(g=0.002, hmin=0.07, n=3, m=2)
    a = _random_state(40, 37, seed=3)
    b = _copy_state(a)
    oc.time_loop(a, p, nsteps=20, threads=1)
    oc.time_loop(b, p, nsteps=20, threads=4)
    for name in FIELDS:
        assert np.array_equal(getattr(a, name), getattr(b, name)), name


# ------------------------------------------------------------------------------------------------
# test/simulate.jl -- whole-loop known answers


def test_flat_film_stays_exactly_flat():  # test/simulate.jl:4-7
    p = onp.Params(Tmax=200, tdump=100)
    for B in (onp, oc):
        st = onp.State(25, 25)
        st.height[...] = 1.0
        B.time_loop(st, p)
        assert np.all(st.height == 1.0)
        assert st.height.sum() == 25 * 25


def test_random_interface_flattens():  # test/simulate.jl:17-21 (run_random: theta = 1/9 via time_loop(...,θ))
    p = onp.Params(Tmax=10000, tdump=5000)
    st = onp.State(25, 25)
    st.height[...] = onp.randinterface(25, 25, 1.0, 0.1, np.random.default_rng(42))
    onp.equilibrium(st.feq, st.height, st.velx, st.vely, st.vsq, p.g)
    oc.time_loop(st, p, cospi_theta=onp.cospi(1 / 9))
    assert st.height.max() - st.height.min() < 0.1


def test_rayleigh_taylor_grows():  # test/simulate.jl:33-35
    p = onp.Params(Tmax=1000, tdump=500, g=-0.002)
    st = onp.State(100, 100)
    st.height[...] = onp.rayleightaylor_ic(100, 100, kx=4, ky=5, eps=0.01)
    onp.equilibrium(st.feq, st.height, st.velx, st.vely, st.vsq, p.g)
    dh, _ = oc.time_loop(st, p, log_dh=True)
    assert dh[0] < dh[-1]
    # the NumPy restatement walks the same trajectory bit for bit (first 50 steps)
    st2 = onp.State(100, 100)
    st2.height[...] = onp.rayleightaylor_ic(100, 100, kx=4, ky=5, eps=0.01)
    dh2 = onp.time_loop(st2, p, nsteps=50)
    assert np.array_equal(np.asarray(dh2), dh[:50])


def _droplet_checks(h, rads=35):
    cospi = onp.cospi
    vol = np.pi / 3 * rads ** 3 * (2 + cospi(1 / 6)) * (1 - cospi(1 / 6)) ** 2
    R1 = np.cbrt((rads ** 3 * (2 + cospi(1 / 6)) * (1 - cospi(1 / 6)) ** 2) / ((2 + cospi(1 / 9)) * (1 - cospi(1 / 9)) ** 2))
    r1 = np.sin(np.pi / 9) * R1
    drop1d = np.nonzero(h[74, :] > 0.055)[0]
    droprad = len(drop1d) / 2
    droph = h.max()
    vnum = 1 / 6 * np.pi * droph * (3 * droprad ** 2 + droph ** 2)
    assert abs(vol - vnum) < vol / 100 * 10
    assert abs(r1 - droprad) < r1 / 100 * 10


def test_relaxing_droplet():  # test/simulate.jl:44-60
    p = onp.Params(Tmax=10000, delta=3.0)
    st = onp.State(150, 150)
    st.height[...] = onp.singledroplet(150, 150, 35, 1 / 6, (75, 75))
    onp.equilibrium(st.feq, st.height, st.velx, st.vely, st.vsq, p.g)
    mass0 = st.height.sum()
    _, wet = oc.time_loop(st, p, log_wetted=True, threads=oc.max_threads())
    _droplet_checks(st.height)
    assert wet[0] < wet[-1]
    assert abs(st.height.sum() - mass0) / mass0 < 1e-12  # mass conserved to round-off


def test_patterned_droplet():  # test/simulate.jl:96-112 (run_dropletpatterned: θₛ = fill(1/9) as a FIELD)
    p = onp.Params(Tmax=10000, delta=3.0)
    st = onp.State(150, 150)
    st.height[...] = onp.singledroplet(150, 150, 35, 1 / 6, (75, 75))
    onp.equilibrium(st.feq, st.height, st.velx, st.vely, st.vsq, p.g)
    ct = np.asfortranarray(np.full((150, 150), onp.cospi(1 / 9)))
    oc.time_loop(st, p, cospi_theta=ct, threads=oc.max_threads())
    _droplet_checks(st.height)
    # a constant field must give exactly what the scalar gives (same arithmetic, site by site)
    st2 = onp.State(150, 150)
    st2.height[...] = onp.singledroplet(150, 150, 35, 1 / 6, (75, 75))
    oc.time_loop(st2, onp.Params(Tmax=300, delta=3.0), threads=oc.max_threads())
    st3 = onp.State(150, 150)
    st3.height[...] = onp.singledroplet(150, 150, 35, 1 / 6, (75, 75))
    oc.time_loop(st3, onp.Params(Tmax=300, delta=3.0), cospi_theta=ct, threads=oc.max_threads())
    assert np.array_equal(st2.height, st3.height)


def test_sliding_droplet():  # test/simulate.jl:147-155 (inclination! in the callback slot, factor(t=1000)=1)
    import math

    p = onp.Params(Tmax=5000, delta=2.0)
    st = onp.State(150, 150)
    st.height[...] = onp.singledroplet(150, 150, 35, 1 / 6, (75, 75))
    onp.equilibrium(st.feq, st.height, st.velx, st.vely, st.vsq, p.g)
    factor = 0.5 + 0.5 * math.tanh((1000 - 0) / 1)
    oc.time_loop(st, p, incl=([1e-4, 0.0], factor), threads=oc.max_threads())
    i, j = np.unravel_index(np.argmax(st.height), st.height.shape)
    assert i + 1 != 75 and j + 1 == 75
    assert np.all(st.velx < 0.1) and np.all(st.vely < 0.1)


def _findmax_index(a):  # Julia's findmax: first maximum in column-major order, as a 0-based (i, j)
    return tuple(int(v) for v in np.unravel_index(a.ravel(order="F").argmax(), a.shape, order="F"))


def test_initial_conditions_known_answers():  # test/initialvalues.jl:16-76
    h = onp.singledroplet(100, 100, 50, 1 / 3, (50, 50))
    assert h.max() == 50 * (1 - onp.cospi(1 / 3)) and _findmax_index(h) == (49, 49)
    rad, th, lx, ly, c = 45, 1 / 4, 150, 200, 80
    top = rad * (1 - onp.cospi(th))
    r = onp.rivulet(lx, ly, rad, th, "y", c, 0.05)
    assert r.shape == (lx, ly) and abs(r.max() - top) < 1e-4 and abs(r[c - 1, :].sum() - top * ly) < 1e-4
    r = onp.rivulet(lx, ly, rad, th, "x", c, 0.05)
    assert abs(r[:, c - 1].sum() - top * lx) < 1e-4
    t = onp.torus(lx, ly, 10, 45, 1 / 9, (80, 80), 0.05)
    assert t.min() == 0.05 and np.isclose(t.max(), (1 - onp.cospi(1 / 9)) * 10)
    assert _findmax_index(t) == (79, 34)  # CartesianIndex(80, 35)
    t = onp.torus(256, 256, 45, 80, 1 / 9, (128, 128))  # doctest src/initialvalues.jl:126-137
    assert np.isclose(t.max(), 45 * (1 - onp.cospi(1 / 9))) and _findmax_index(t) == (127, 47)



@pytest.mark.parametrize("case", [
    dict(prm=dict(g=-0.001, gamma=0.0005)),
    dict(prm=dict(n=3, m=2, hmin=0.07, gamma=0.01)),
    dict(prm=dict(tau=0.8, n=4, m=2), pops=True),
    dict(prm=dict(n=3, m=2, hmin=0.07, mu=1 / 12), theta_field=True, slip_variant=1, incl=([1e-4, -2e-5], 0.75)),
    dict(prm=dict(tau=1.3), pops=True, pvariant="fast", slip_variant=2),
], ids=["rt", "c3", "tau0.8", "theta-slip2-incl", "tau1.3-fast-ring"])
def test_lowmem_fused_restatement_equals_the_pass_structured_oracle(case):
    """oracle_time_loop_lowmem (22 planes, three passes: what the 4096^2 / 8192^2 GPU parity tests run) against
    oracle_time_loop (the reference's pass structure, 46 planes), bit for bit, on every variant; 1 vs 4 threads."""
    Lx, Ly, nsteps = 37, 29, 5
    rng = np.random.default_rng(99)
    p = onp.Params(**case["prm"])
    ref = onp.State(Lx, Ly)
    ref.height[...] = np.abs(1.0 + 0.2 * rng.standard_normal((Lx, Ly))) + 0.06
    ref.velx[...] = 0.01 * rng.standard_normal((Lx, Ly))
    ref.vely[...] = 0.01 * rng.standard_normal((Lx, Ly))
    if case.get("pops"):
        ref.ftemp[...] = 0.1 + 0.01 * rng.random((Lx, Ly, 9))
    ct = np.asfortranarray(np.cos(np.pi * (1 / 9 + rng.random((Lx, Ly)) / 36))) if case.get("theta_field") else None
    kw = dict(cospi_theta=ct, pvariant=case.get("pvariant", "power_broad"), slip_variant=case.get("slip_variant", 0),
              incl=case.get("incl"))
    for threads, steps in ((1, nsteps), (4, nsteps - 1)):
        h, ux, uy, f = (np.asfortranarray(a.copy()) for a in (ref.height, ref.velx, ref.vely, ref.ftemp))
        full = onp.State(Lx, Ly)
        for name in ("height", "velx", "vely", "ftemp"):
            getattr(full, name)[...] = getattr(ref, name)
        pr = oc.time_loop_lowmem(h, ux, uy, f, p, steps, threads=threads, **kw)
        oc.time_loop(full, p, nsteps=steps, threads=threads, **kw)
        assert np.array_equal(h, full.height) and np.array_equal(ux, full.velx) and np.array_equal(uy, full.vely)
        assert np.array_equal(f, full.fout) and np.array_equal(pr, full.pressure)


def test_appendix_a_constants_bit_patterns():
    """SURVEY.md Appendix A lists the IEEE-754 bit patterns of the literals the reference's broadcasts fold (2/3, 1/6, 10/3,
    -1/3, 1/12, 1/3, 1/24, 1/9, 1/36, 5/6), of c_s = 1/sqrt(3.0) and of the default viscosity mu = (c_s*c_s)*(tau - 0.5), which
    is 2 ulp above 1/6.  The oracle's Python expressions and the host mirror's Taumucs must produce exactly these doubles."""
    import math
    import struct

    import swalbe_b200 as sw

    def bits(x):
        return struct.pack(">d", float(x)).hex().upper()

    want = {2 / 3: "3FE5555555555555", 1 / 6: "3FC5555555555555", 10 / 3: "400AAAAAAAAAAAAB", -1 / 3: "BFD5555555555555",
            1 / 12: "3FB5555555555555", 1 / 3: "3FD5555555555555", 1 / 24: "3FA5555555555555", 1 / 9: "3FBC71C71C71C71C",
            1 / 36: "3F9C71C71C71C71C", 5 / 6: "3FEAAAAAAAAAAAAB"}
    for value, pattern in want.items():
        assert bits(value) == pattern, (value, bits(value))
    cs = 1 / math.sqrt(3.0)
    assert bits(cs) == "3FE279A74590331D" and bits(cs * cs) == "3FD5555555555557"
    for prm in (onp.Params(), sw.Taumucs()):
        assert bits(prm.cs) == "3FE279A74590331D"
        assert bits(prm.mu) == "3FC5555555555557" and prm.mu != 1 / 6  # (cs*cs)*(1 - 0.5): 2 ulp above 1/6
