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
