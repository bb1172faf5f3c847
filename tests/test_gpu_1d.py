"""The 1-D (D1Q3) family on the GPU through the C ABI: the reference's 1-D known answers
(tests/test_oracle_1d.py transcribes test/*.jl) with every operator on the device, then the fused loop
(swalbe_time_loop_1d: persistent single-CTA kernel / step-by-step kernel) against the oracle, bit for bit."""
import numpy as np
import pytest

from oracle import oracle_1d as o1
from oracle import oracle_np as onp
from tests import test_oracle_1d as cases

pytestmark = pytest.mark.gpu


class _Backend:
    """oracle_1d's call signatures on device Fields: NumPy in -> device -> kernel -> NumPy out"""

    def __init__(self):
        import swalbe_b200 as sw

        self.sw = sw

    def _up(self, a):
        sw = self.sw
        f = sw.Field(a.shape[0]) if a.ndim == 1 else sw.Field(a.shape[0], a.shape[1])
        return f.set(a)

    def equilibrium(self, feq, h, v, g):
        d = [self._up(x) for x in (feq, h, v)]
        self.sw.equilibrium(*d, g)
        feq[...] = d[0].numpy()

    def BGKandStream(self, fout, feq, ftemp, F, tau):
        d = [self._up(x) for x in (fout, feq, ftemp, F)]
        self.sw.BGKandStream(*d, tau)
        fout[...] = d[0].numpy(); ftemp[...] = d[2].numpy()

    def moments(self, h, v, f):
        d = [self._up(x) for x in (h, v, f)]
        self.sw.moments(*d)
        h[...] = d[0].numpy(); v[...] = d[1].numpy()

    def filmpressure(self, out, f, dgrad, gamma, cospi_theta, n, m, hmin, hcrit):
        import ctypes as C

        from swalbe_b200 import _lib

        o, fd = self._up(out), self._up(f)
        _lib.call("swalbe_filmpressure_1d", o.ptr, fd.ptr, None, float(gamma), float(cospi_theta), None, n, m, float(hmin),
                  float(hcrit), _lib.PRESSURE_FAST, f.shape[0], self.sw._stream())
        out[...] = o.numpy()

    def grad(self, out, f, a=None):
        o, fd = self._up(out), self._up(f)
        self.sw.gradf(o, fd, None, self._up(a)) if a is not None else self.sw.gradf(o, fd, None)
        out[...] = o.numpy()

    def lap(self, out, f):
        o, fd = self._up(out), self._up(f)
        self.sw.laplacianf(o, fd, None)
        out[...] = o.numpy()

    def slippage(self, s, h, v, delta, mu):
        d = [self._up(x) for x in (s, h, v)]
        self.sw.slippage(*d, delta, mu)
        s[...] = d[0].numpy()

    # ---- the expanded kinds (State_thermal_1D, State_gamma_1D, StateWithBound_1D) ----
    def inclination(self, F, h, alpha, t=1000, tstart=0, tsmooth=1):
        sw = self.sw
        st = sw.Sys(sw.SysConst_1D(L=len(F), param=sw.Taumucs()), kind="thermal")  # (Expanded_1D method, src/forcing.jl:385)
        st.basestate.F.set(F); st.basestate.height.set(h)
        sw.inclination(alpha, st, t=t, tstart=tstart, tsmooth=tsmooth)
        F[...] = st.basestate.F.numpy()

    def gradgamma(self, out, gamma, height=None, delta=None):
        sw = self.sw
        sysc = sw.SysConst_1D(L=len(out), param=sw.Taumucs(δ=1.0 if delta is None else delta))
        st = sw.Sys(sysc, kind="gamma")
        st.γ.set(gamma)
        if height is None:
            sw.gradgamma(st)
        else:
            st.basestate.height.set(height)
            sw.gradgamma(st, sysc)
        out[...] = getattr(st, "∇γ").numpy()

    def filmpressure_gamma(self, out, f, gamma, cospi_theta, n, m, hmin, hcrit, rho=None, Gamma=0.0, ftemp=None):
        from swalbe_b200 import _lib

        o, fd = self._up(out), self._up(f)
        gf = self._up(gamma) if isinstance(gamma, np.ndarray) else None
        rf = self._up(rho) if rho is not None else None
        ft = self._up(ftemp) if ftemp is not None else None
        _lib.call("swalbe_filmpressure_gamma_1d", o.ptr, fd.ptr, 0.0 if gf is not None else float(gamma),
                  gf.ptr if gf is not None else None, rf.ptr if rf is not None else None, float(Gamma), float(cospi_theta), None,
                  n, m, float(hmin), float(hcrit), ft.ptr if ft is not None else None, len(f), self.sw._stream())
        out[...] = o.numpy()
        if ftemp is not None:
            ftemp[...] = ft.numpy()

    def BGKandStream_bound(self, fout, feq, ftemp, fbound, F, border, tau):
        from swalbe_b200 import _lib

        d = [self._up(x) for x in (fout, feq, ftemp, fbound, F, border[0], border[1])]
        _lib.call("swalbe_bgk_stream_bound_d1q3", *(x.ptr for x in d), float(tau), len(F), self.sw._stream())
        fout[...] = d[0].numpy(); ftemp[...] = d[2].numpy(); fbound[...] = d[3].numpy()

    def update_rho(self, rho, rho_int, height, differentials, D=1.0, M=0.0):
        d = [self._up(x) for x in (rho, rho_int, height, differentials)]
        self.sw.update_rho(d[0], d[1], d[2], None, d[3], D=D, M=M)
        rho[...] = d[0].numpy(); rho_int[...] = d[1].numpy(); differentials[...] = d[3].numpy()


@pytest.mark.parametrize("case", cases.ALL_1D_CASES, ids=lambda c: c.__name__)
def test_reference_1d_known_answers_on_gpu(case):
    case(_Backend())


def _mk(L, seed, pops, **kw):
    import swalbe_b200 as sw

    rng = np.random.default_rng(seed)
    ref = o1.State1D(L)
    ref.height[...] = np.abs(1.0 + 0.2 * rng.standard_normal(L)) + 0.06
    ref.vel[...] = 0.01 * rng.standard_normal(L)
    if pops:
        ref.ftemp[...] = 0.3 + 0.01 * rng.random((L, 3))
    sysc = sw.SysConst_1D(L=L, param=sw.Taumucs(**kw))
    st = sw.Sys(sysc)
    st.height.set(ref.height); st.vel.set(ref.vel); st.ftemp.set(ref.ftemp)
    okw = {{"γ": "gamma", "δ": "delta", "τ": "tau", "μ": "mu", "θ": "theta"}.get(k, k): v for k, v in kw.items()}
    return sw, st, sysc, ref, onp.Params(**okw)


FIELDS = ("height", "vel", "fout", "ftemp", "feq", "pressure", "hgradp", "slip", "F")


@pytest.mark.parametrize("L", [1, 2, 3, 30, 257, 1024, 5000, 20000])
@pytest.mark.parametrize("kw,pops", [(dict(g=-0.001), False), (dict(n=3, m=2, hmin=0.07, γ=0.02), False), (dict(τ=0.8), True)],
                         ids=["g", "32", "tau0.8"])
def test_fused_1d_loop_bitwise(L, kw, pops):
    """persistent kernel (lattice in one CTA's shared memory) for the first nsteps - 1 steps, the materialising step on
    global memory; 20000 sites do not fit and go step by step.  Every field of State_1D against the oracle."""
    sw, st, sysc, ref, p = _mk(L, L + 1, pops, **kw)
    for n in (1, 4, 7):
        sw.one_d.fused_steps(st, sysc, n)
        o1.time_loop(ref, p, nsteps=n)
        for name in FIELDS:
            assert np.array_equal(getattr(st, name).numpy(), getattr(ref, name), equal_nan=True), (name, n)


def test_fused_1d_theta_vector_logs_and_drivers():
    """time_loop(sys, state, θ) with a contact-angle VECTOR, the Δh variant, run_flat / run_random (test/simulate.jl:8-30)"""
    L = 300
    sw, st, sysc, ref, p = _mk(L, 5, False, n=3, m=2, hmin=0.07, Tmax=25, tdump=10)
    theta = 1 / 9 + 1 / 36 * np.random.default_rng(1).random(L)
    thf = sw.Field(L).set(theta)
    ct = sw.cospi_field(thf).numpy()
    sw.time_loop(sysc, st, thf)
    o1.time_loop(ref, p, nsteps=25, cospi_theta=ct)
    for name in FIELDS:
        assert np.array_equal(getattr(st, name).numpy(), getattr(ref, name)), name
    dh = []
    sw.time_loop(sysc, st, dh)
    want = o1.time_loop(ref, p, nsteps=25)
    assert dh == want and np.array_equal(st.height.numpy(), ref.height)
    h = sw.run_flat(sw.SysConst_1D(L=25, param=sw.Taumucs(Tmax=200, tdump=100)), verbos=False).numpy()
    assert np.all(h == 1.0) and h.sum() == 25
    h = sw.run_random(sw.SysConst_1D(L=25, param=sw.Taumucs(Tmax=10000, tdump=5000)), ϵ=0.1, verbos=False,
                      rng=np.random.default_rng(42)).numpy()
    assert h.max() - h.min() < 0.02
    with pytest.raises(sw.DomainError):
        sw.filmpressure(sw.Field(30), sw.Field(30, fill=1.0), sw.Field(30, 2), 1.0, 0.0, 4, 2, 0.1, 0.0)


def test_operator_by_operator_1d_equals_fused():
    """the seven-call loop body of src/simulate.jl:107-114 on the per-operator 1-D kernels == the fused loop"""
    sw, st, sysc, _, _ = _mk(500, 9, True, τ=0.9, g=0.001)
    _, st2, _, _, _ = _mk(500, 9, True, τ=0.9, g=0.001)
    sw.one_d.fused_steps(st, sysc, 5)
    for _ in range(5):
        sw.filmpressure(st2, sysc)
        sw.hgradp(st2)
        sw.slippage(st2, sysc)
        sw.update(st2)
        sw.equilibrium(st2, sysc)
        sw.BGKandStream(st2, sysc)
        sw.moments(st2)
    for name in FIELDS:
        assert np.array_equal(getattr(st, name).numpy(), getattr(st2, name).numpy()), name


def test_expanded_1d_operators_random_inputs_against_oracle():
    """the operators of the expanded 1-D kinds on random fields, bit for bit against the NumPy restatement"""
    B, rng, L = _Backend(), np.random.default_rng(3), 777
    h = np.abs(1.0 + 0.3 * rng.standard_normal(L)) + 0.06
    gam = 0.01 * (1.0 + 0.2 * rng.random(L))
    rho = 0.1 * rng.random(L)
    for fn, args, kw in (
            ("gradgamma", (gam,), {}), ("gradgamma", (gam, h, 0.7), {}),
            ("filmpressure_gamma", (h, gam, onp.cospi(1 / 9), 9, 3, 0.1, 0.05), {}),
            ("filmpressure_gamma", (h, 0.013, onp.cospi(1 / 7), 3, 2, 0.07, 0.05), {"rho": rho, "Gamma": 0.4})):
        a, b = np.zeros(L), np.zeros(L)
        getattr(B, fn)(a, *args, **kw)
        getattr(o1, fn)(b, *args, **kw)
        assert np.array_equal(a, b), fn
    fa, fb = np.zeros((L, 3), order="F"), np.zeros((L, 3), order="F")
    a, b = np.zeros(L), np.zeros(L)
    B.filmpressure_gamma(a, h, gam, onp.cospi(1 / 9), 9, 3, 0.1, 0.05, ftemp=fa)
    o1.filmpressure_gamma(b, h, gam, onp.cospi(1 / 9), 9, 3, 0.1, 0.05, ftemp=fb)
    assert np.array_equal(a, b) and np.array_equal(fa, fb)
    obs = np.zeros(L); obs[:4] = 1; obs[300:305] = 1
    _, border = o1.obslist1D(obs)
    import swalbe_b200 as sw
    i2, b2 = sw.one_d.obslist1D(obs)
    assert np.array_equal(border[0], b2[0]) and np.array_equal(border[1], b2[1]) and np.array_equal(i2, o1.obslist1D(obs)[0])
    feq, ft = 0.3 + 0.01 * rng.random((L, 3)), 0.3 + 0.01 * rng.random((L, 3))
    F = 1e-3 * rng.standard_normal(L)
    outs = []
    for Bk in (B, o1):
        fo, fe, ftc, fbd = np.zeros((L, 3), order="F"), np.asfortranarray(feq), np.asfortranarray(ft.copy()), np.zeros((L, 3), order="F")
        Bk.BGKandStream_bound(fo, fe, ftc, fbd, F, border, 0.8)
        outs.append((fo, ftc, fbd))
    for x, y in zip(*outs):
        assert np.array_equal(x, y)
    ra, rb = rho.copy(), rho.copy()
    ia, ib, da, db = np.zeros(L), np.zeros(L), np.zeros((L, 4), order="F"), np.zeros((L, 4), order="F")
    B.update_rho(ra, ia, h, da, D=0.8, M=0.3)
    o1.update_rho(rb, ib, h, db, D=0.8, M=0.3)
    assert np.array_equal(ra, rb) and np.array_equal(ia, ib) and np.array_equal(da, db)


def test_thermal_1d_statistics():
    """thermal!(state::State_thermal_1D, sys): unit normals times the reference's amplitude (src/forcing.jl:322-333);
    test/forcing.jl:141-164 asks for mean ~ 0 and variance ~ 2 kbt mu 6 h / (2h² + 6hδ + 3δ²) within 10 %"""
    import swalbe_b200 as sw

    L = 1 << 18
    sysc = sw.SysConst_1D(L=L, param=sw.Taumucs(kbt=0.2))
    st = sw.Sys(sysc, kind="thermal")
    sw.thermal(st, sysc, seed=11, step=0)
    k = st.kbt.numpy()
    var = o1.thermal_amplitude(1.0, 0.2, sysc.param.mu, sysc.param.delta) ** 2
    assert abs(k.mean()) < 4 * np.sqrt(var / L) and abs(k.var() / var - 1) < 0.02
    sw.thermal(st, sysc, seed=11, step=1)
    k2 = st.kbt.numpy()
    assert abs(np.corrcoef(k, k2)[0, 1]) < 0.01  # fresh noise every step
    sw.thermal(st.kbt, st.basestate.height, 0.2, sysc.param.mu, sysc.param.delta, seed=11, step=0)  # array form
    assert np.array_equal(st.kbt.numpy(), k)


def test_fused_1d_gamma_and_inclination_loops():
    """the loop body of run_gamma (per-site tension in the pressure, F = -h∇p - slip - ∇γ; src/simulate.jl:541-547) and
    the inclination! callback slot (:159-179) inside the fused 1-D loop, persistent and step-by-step, against the oracle;
    then the drivers run_gamma / run_dropletforced(::SysConst_1D) against the same loops written out"""
    import math

    import swalbe_b200 as sw

    for L in (300, 20000):
        rng = np.random.default_rng(L)
        kw = dict(n=9, m=3, hmin=0.1, γ=0.01)
        sysc = sw.SysConst_1D(L=L, param=sw.Taumucs(**kw))
        p = onp.Params(n=9, m=3, hmin=0.1, gamma=0.01)
        h0 = np.abs(1.0 + 0.2 * rng.standard_normal(L)) + 0.06
        gam = 0.01 * (1.0 + 0.3 * np.sin(2 * np.pi * np.arange(L) / L))
        st = sw.Sys(sysc, kind="gamma")
        st.basestate.height.set(h0)
        st.γ.set(gam)
        sw.gradgamma(st)
        ref = o1.State1D(L)
        ref.height[...] = h0
        dg = np.zeros(L)
        o1.gradgamma(dg, gam)
        for n in (1, 6):
            sw.one_d.fused_steps(st, sysc, n, gamma_field=True, marangoni=True)
            for _ in range(n):
                o1.step_gamma(ref, p, gam, dg)
            for name in FIELDS:
                assert np.array_equal(getattr(st.basestate, name).numpy(), getattr(ref, name)), (L, name, n)
        st2 = sw.Sys(sysc)
        st2.height.set(h0)
        ref2 = o1.State1D(L)
        ref2.height[...] = h0
        fac = 0.5 + 0.5 * math.tanh(1000.0)
        sw.one_d.fused_steps(st2, sysc, 5, incl=(1e-4, fac))
        for _ in range(5):
            o1.step_gamma(ref2, p, None, None, alpha=1e-4, incl_factor=fac)
        for name in FIELDS:
            assert np.array_equal(getattr(st2, name).numpy(), getattr(ref2, name)), (L, name)
    # drivers
    L = 512
    sysg = sw.SysConst_1D(L=L, param=sw.Taumucs(Tmax=230, tdump=100, γ=0.01, n=9, m=3))
    gam = 0.01 * (1.0 - 0.2 * np.arange(L) / L)
    fluid = sw.run_gamma(sysg, gam, r1=60, r2=60, verbos=False, dump=100)
    ref = o1.State1D(L)
    ref.height[...] = sw.two_droplets(sysg, r1=60, r2=60)
    dg = np.zeros(L)
    o1.gradgamma(dg, gam)
    want = np.zeros((2, L))
    p = onp.Params(gamma=0.01)
    for t in range(1, 231):
        o1.step_gamma(ref, p, gam, dg)
        if t % 100 == 0:
            want[t // 100 - 1] = ref.height
    assert fluid.shape == (2, L) and np.array_equal(fluid, want)
    sysf = sw.SysConst_1D(L=256, param=sw.Taumucs(Tmax=60, tdump=25))
    hgt, vel = sw.run_dropletforced(sysf, radius=40, f=1e-4, verbos=False)
    ref = o1.State1D(256)
    ref.height[...] = sw.one_d.singledroplet_1d(256, 40, 1 / 6, 128)
    for _ in range(60):
        o1.step_gamma(ref, onp.Params(), None, None, alpha=1e-4, incl_factor=0.5 + 0.5 * math.tanh(1000.0))
    assert np.array_equal(hgt.numpy(), ref.height) and np.array_equal(vel.numpy(), ref.vel)


def test_bounce_back_1d_time_loop():
    """time_loop(sys::SysConstWithBound_1D, state::StateWithBound_1D)  src/simulate.jl:181-204 against the oracle loop; the
    fluid mass between the walls is conserved"""
    import swalbe_b200 as sw

    L = 200
    obs = np.zeros(L); obs[:4] = 1; obs[-4:] = 1
    sysb = sw.SysConstWithBound_1D(L=L, param=sw.Taumucs(Tmax=40, tdump=10, τ=0.9), obs=obs)
    sw.obslist(sysb)
    st = sw.Sys(sysb, kind="gamma_bound")
    h0 = 1.0 + 0.1 * np.sin(2 * np.pi * np.arange(L) / L)
    st.basestate.height.set(h0)
    sw.equilibrium(st, sysb)
    ref = o1.State1D(L)
    ref.height[...] = h0
    o1.equilibrium(ref.feq, ref.height, ref.vel, 0.0)
    fb = np.zeros((L, 3), order="F")
    sw.time_loop(sysb, st)
    _, border = o1.obslist1D(obs)
    p = onp.Params(tau=0.9)
    for _ in range(40):
        o1.step_bound(ref, fb, p, border)
    for name in FIELDS:
        assert np.array_equal(getattr(st.basestate, name).numpy(), getattr(ref, name), equal_nan=True), name
    assert np.array_equal(st.fbound.numpy(), fb)
