"""The 1-D (D1Q3) oracle (oracle/oracle_1d.py) against the reference's own 1-D known answers: test/collide.jl:141-226,
test/equilibrium.jl:87-153, test/moments.jl:14-134 (1-D entries), test/pressure.jl:56-104, test/differences.jl:79-138,
test/forcing.jl:71-102 and :125-138, test/simulate.jl:8-30.  `B` is the backend under test: the oracle here, the C ABI
in tests/test_gpu_1d.py."""
import numpy as np
import pytest

from oracle import oracle_1d as o1
from oracle import oracle_np as onp


def case_collide_1d(B):  # test/collide.jl:141-226
    onebytau, omega = 1.0 / 0.75, 1.0 - 1.0 / 0.75
    for tau, force in ((1.0, 0.0), (0.75, 0.0), (1.0, 0.1), (0.75, 0.1)):
        feq, ftemp, fout = np.ones((30, 3), order="F"), np.ones((30, 3), order="F"), np.ones((30, 3), order="F")
        feq[0, :] = 2.0
        e1 = feq[:, 0].copy()
        B.BGKandStream(fout, feq, ftemp, np.full(30, force), tau)
        base = e1 if tau == 1.0 else omega * 1.0 + onebytau * e1
        half = 1 / 20 if force else 0.0
        assert np.array_equal(fout[:, 0], base)
        assert np.array_equal(fout[:, 1], np.roll(base + half, 1))
        assert np.array_equal(fout[:, 2], np.roll(base - half, -1))
        assert np.array_equal(fout, ftemp)  # fout .= ftemp (src/collide.jl:199)


def case_equilibrium_1d(B):  # test/equilibrium.jl:87-153
    feq = np.zeros((30, 3), order="F")
    B.equilibrium(feq, np.zeros(30), np.zeros(30), 0.0)
    assert np.all(feq == 0.0)
    B.equilibrium(feq, np.ones(30), np.zeros(30), 0.0)
    assert np.all(feq[:, 0] == 1.0) and np.all(feq[:, 1:] == 0.0)
    B.equilibrium(feq, np.ones(30), np.zeros(30), 0.1)
    assert np.all(feq[:, 0] == 1.0 - 0.05) and np.all(feq[:, 1:] == 0.025)
    B.equilibrium(feq, np.ones(30), np.full(30, 0.1), 0.0)
    assert np.all(feq[:, 0] == 1.0 - 0.01)
    assert np.allclose(feq[:, 1], 0.5 * 0.1 + 0.5 * 0.01, rtol=1e-14) and np.allclose(feq[:, 2], -0.5 * 0.1 + 0.5 * 0.01, rtol=1e-14)
    B.equilibrium(feq, np.ones(30), np.full(30, 0.1), 0.1)
    assert np.all(feq[:, 0] == 1.0 - 0.5 * 0.1 - 0.01)
    assert np.allclose(feq[:, 1], 0.025 + 0.5 * 0.1 + 0.5 * 0.01, rtol=1e-14)
    feq10 = np.zeros((10, 3), order="F")  # doctest src/equilibrium.jl:141-163
    B.equilibrium(feq10, np.ones(10), np.full(10, 0.1), 0.1)
    assert np.allclose(feq10[:, 0], 0.94, rtol=1e-15)


def case_moments_1d(B):  # test/moments.jl (1-D entries): h = sum of the three columns, v = (f1 - f2) / h
    f = np.zeros((30, 3), order="F")
    f[:, 0] = 1.0; f[:, 1] = 0.1
    h, v = np.zeros(30), np.zeros(30)
    B.moments(h, v, f)
    assert np.all(h == 1.1) and np.all(v == 0.1 / 1.1)
    f[:, 2] = 0.2
    B.moments(h, v, f)
    assert np.all(h == (1.0 + 0.1) + 0.2) and np.all(v == (0.1 - 0.2) / ((1.0 + 0.1) + 0.2))


def case_pressure_1d(B):  # test/pressure.jl:56-104
    f = np.arange(1.0, 31.0)
    sol = np.zeros(30); sol[0] = 30; sol[-1] = -30
    res, dummy = np.zeros(30), np.zeros((30, 2), order="F")
    B.filmpressure(res, f, dummy, 1.0, onp.cospi(0.0), 3, 2, 0.1, 0.1)
    assert np.all(res == -sol)
    B.filmpressure(res, f, dummy, 1.0, onp.cospi(1 / 2), 3, 2, 0.1, 0.0)
    assert np.allclose(res, -1 * (sol + 20 * ((0.1 / f) ** 3 - (0.1 / f) ** 2)), atol=1e-10, rtol=0)
    B.filmpressure(res, np.ones(30), dummy, 1.0, onp.cospi(1 / 2), 3, 2, 0.1, 0.0)
    assert np.allclose(res, -2 * (0.1 ** 2 - 0.1), atol=1e-10, rtol=0)
    with pytest.raises(ValueError):  # DomainError((4, 2), ...)  src/pressure.jl:218-225
        B.filmpressure(res, f, dummy, 1.0, 0.0, 4, 2, 0.1, 0.0)


def case_stencils_1d(B):  # test/differences.jl:79-138
    f1 = np.arange(1.0, 26.0)
    out = np.zeros(25)
    sol = np.ones(25); sol[0] = sol[-1] = -11.5
    B.grad(out, f1, np.ones(25))
    assert np.array_equal(out, sol)
    B.grad(out, f1)
    assert np.array_equal(out, sol)
    B.lap(out, f1)
    sol = np.zeros(25); sol[0] = 25; sol[-1] = -25
    assert np.array_equal(out, sol)


def case_slippage_1d(B):  # test/forcing.jl:71-102
    s = np.zeros(30)
    B.slippage(s, np.ones(30), np.zeros(30), 1.0, 1 / 6)
    assert np.all(s == 0.0)
    B.slippage(s, np.ones(30), np.full(30, 0.1), 1.0, 1 / 6)
    assert np.allclose(s, 0.1 / 11, atol=1e-10, rtol=0)
    B.slippage(s, np.ones(30), np.full(30, -0.1), 0.0, 1 / 6)
    assert np.allclose(s, -0.1 / 2, atol=1e-10, rtol=0)


def case_hgradp_1d(B):  # test/forcing.jl:125-138: pressure = height = 1..30 -> h∇p = 1 except -14 at both ends
    p = np.arange(1.0, 31.0)
    out = np.zeros(30)
    B.grad(out, p, np.ones(30))
    sol = np.ones(30); sol[0] = sol[-1] = -14
    assert np.array_equal(out, sol)


def case_bounce_back_1d(B):  # test/collide.jl:133-150: walls at both ends, tau = 1, no forces, feq = ftemp = fout = 1 except
    # feq[1, :] = 2 ... the reference's state there is the freshly allocated one (all zero), for which the known answer is
    # fout[:, k] == circshift(feq[:, 1], c_k); the dummy-distribution case below adds the reflected populations
    obst = np.zeros(30); obst[0] = obst[-1] = 1
    interior, border = o1.obslist1D(obst)   # src/obstacle.jl:6-35
    assert interior.sum() == 0 and np.array_equal(border[0], np.eye(30)[29]) and np.array_equal(border[1], np.eye(30)[0])
    z = lambda *sh: np.zeros(sh, order="F")  # noqa: E731
    fout, feq, ftemp, fb = z(30, 3), z(30, 3), z(30, 3), z(30, 3)
    B.BGKandStream_bound(fout, feq, ftemp, fb, np.zeros(30), border, 1.0)
    assert np.array_equal(fout[:, 0], feq[:, 0]) and np.array_equal(fout[:, 1], np.roll(feq[:, 0], 1))
    # a uniform gas between two walls: what would leave through a wall link comes back in the opposite direction, the
    # total is conserved (the property bounce-back exists for)
    feq[...] = 1.0; ftemp[...] = 1.0
    feq[3, :] = 2.0
    B.BGKandStream_bound(fout, feq, ftemp, fb, np.full(30, 0.1), border, 0.75)
    want_o, want_t, want_b = z(30, 3), ftemp.copy(), z(30, 3)
    want_t[...] = 1.0
    o1.BGKandStream_bound(want_o, feq, want_t, want_b, np.full(30, 0.1), border, 0.75)
    assert np.array_equal(fout, want_o) and np.array_equal(ftemp, want_t) and np.array_equal(fb, want_b)
    omega, it = 1 - 1 / 0.75, 1 / 0.75
    total = (omega * 1.0 + it * feq).sum()
    assert abs(fout.sum() - total) < 1e-12
    assert fb[29, 1] == (omega + it) + 0.05 and fb[0, 2] == (omega + it) - 0.05 and fb[:, 0].sum() == 0


def case_inclination_1d(B):  # test/forcing.jl:165-183
    sols = {0: 0.05, 1: 0.1 * (0.5 + 0.5 * np.tanh(1.0))}
    for t in (0, 1):
        F = np.zeros(30)
        B.inclination(F, np.ones(30), 0.1, t=t, tstart=0, tsmooth=1)
        assert np.all(F == sols[t])


def case_gradgamma_1d(B):  # test/forcing.jl:185-194
    out = np.zeros(30)
    B.gradgamma(out, np.arange(1.0, 31.0))
    sol = np.full(30, 3 / 2); sol[0] = sol[-1] = -21.0
    assert np.array_equal(out, sol)
    # ∇γ!(state, sys) (src/forcing.jl:434-447) has no known answer upstream: flat film h = 1, δ = 1 -> (2+6+3)/6 * 1/2 * dγ
    B.gradgamma(out, np.arange(1.0, 31.0), np.ones(30), 1.0)
    sol = np.full(30, 11 / 6 * 1.0 / 2 * -1.0); sol[0] = sol[-1] = 11 / 6 * 1.0 / 2 * 14.0
    assert np.array_equal(out, sol)


def case_rho_update_1d(B):  # test/forcing.jl:196-204: constant fields stay constant
    rho, out = np.ones(25), np.zeros(25)
    B.update_rho(rho, out, np.ones(25), np.zeros((25, 4), order="F"))
    assert np.all(rho == 1) and np.all(out == 0)


def case_pressure_gamma_1d(B):  # test/pressure.jl:56-131
    f = np.arange(1.0, 31.0)
    sol = np.zeros(30); sol[0] = 30; sol[-1] = -30
    res = np.zeros(30)
    ft = np.zeros((30, 3), order="F")
    B.filmpressure_gamma(res, f, 1.0, onp.cospi(0.0), 3, 2, 0.1, 0.0, ftemp=ft)       # state3, sys3 (θ = 0)   :73-79
    assert np.all(res == -sol) and np.all(ft[:, 1] == 0.0) and np.array_equal(ft[:, 2], -sol)
    B.filmpressure_gamma(res, f, np.full(30, 1.0), onp.cospi(1 / 2), 3, 2, 0.1, 0.0)   # γ as a per-site field   :81-88
    assert np.allclose(res, -1 * (sol + 20 * ((0.1 / f) ** 3 - (0.1 / f) ** 2)), atol=1e-10, rtol=0)
    B.filmpressure_gamma(res, np.ones(30), 1.0, onp.cospi(1 / 2), 3, 2, 0.1, 0.0)      # :90-100
    assert np.allclose(res, -2 * (0.1 ** 2 - 0.1), atol=1e-10, rtol=0)
    rho = np.full(30, 0.1)                                                            # active matter   :104-131
    B.filmpressure_gamma(res, f, 1.0, onp.cospi(0.0), 3, 2, 0.1, 0.1, rho=rho)
    assert np.all(res == -sol)
    B.filmpressure_gamma(res, np.ones(30), 1.0, onp.cospi(1 / 2), 3, 2, 0.1, 0.0, rho=np.zeros(30), Gamma=0.0)
    assert np.allclose(res, -2 * (0.1 ** 2 - 0.1), atol=1e-10, rtol=0)
    B.filmpressure_gamma(res, f, 1.0, onp.cospi(1 / 2), 3, 2, 0.1, 0.0, rho=rho, Gamma=0.1)
    assert np.allclose(res, -(1 + 0.01) * (sol + 20 * ((0.1 / f) ** 3 - (0.1 / f) ** 2)), atol=1e-10, rtol=0)


ALL_1D_CASES = [case_collide_1d, case_equilibrium_1d, case_moments_1d, case_pressure_1d, case_stencils_1d, case_slippage_1d,
                case_hgradp_1d, case_bounce_back_1d, case_inclination_1d, case_gradgamma_1d, case_rho_update_1d,
                case_pressure_gamma_1d]


@pytest.mark.parametrize("case", ALL_1D_CASES, ids=lambda c: c.__name__)
def test_oracle_1d_reproduces_the_reference_known_answers(case):
    case(o1)


def test_oracle_1d_whole_loop_known_answers():
    """test/simulate.jl:8-14: a flat 1-D film stays exactly 1.0 for 200 steps (sum == 25); :25-30: a randomly perturbed
    interface (ϵ = 0.1) flattens below 0.02 within 10 000 steps."""
    st = o1.State1D(25)
    o1.time_loop(st, onp.Params(), nsteps=200)
    assert np.all(st.height == 1.0) and st.height.sum() == 25
    st = o1.State1D(25)
    st.height[...] = 1.0 * (1.0 + 0.1 * np.random.default_rng(42).standard_normal(25))
    o1.equilibrium(st.feq, st.height, st.vel, 0.0)
    dh = o1.time_loop(st, onp.Params(), nsteps=10000)
    assert st.height.max() - st.height.min() < 0.02 and dh[0] > 0.1
    assert abs(st.height.sum() - 25 * st.height.mean()) < 1e-12
