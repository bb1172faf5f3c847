"""Parity of every per-operator CUDA kernel (through the C ABI) with (a) the reference's own known-answer
vectors and (b) the oracle, bit for bit, on random fields at awkward sizes."""
import numpy as np
import pytest

from oracle import oracle_c as oc
from oracle import oracle_np as onp
from tests import golden_cases as gc

pytestmark = pytest.mark.gpu

SIZES = [(5, 5), (25, 26), (150, 200), (257, 3), (1, 7), (1030, 33)]


@pytest.fixture(scope="module")
def G():
    from tests import gpu_backend

    return gpu_backend


@pytest.mark.parametrize("case", gc.ALL_OPERATOR_CASES, ids=lambda c: c.__name__)
def test_reference_known_answers_on_gpu(case, G):
    case(G)


def _rand(shape, rng, scale=1.0, offset=0.0):
    return np.asfortranarray(offset + scale * rng.standard_normal(shape))


def _eq(a, b, name=""):
    assert np.array_equal(a, b, equal_nan=True), f"{name}: max abs diff {np.nanmax(np.abs(a - b)):.3e}"


@pytest.mark.parametrize("Lx,Ly", SIZES)
def test_equilibrium_bitwise(G, Lx, Ly):
    rng = np.random.default_rng(Lx + Ly)
    h, ux, uy = np.abs(_rand((Lx, Ly), rng, 0.3, 1.0)), _rand((Lx, Ly), rng, 0.05), _rand((Lx, Ly), rng, 0.05)
    for g in (0.0, -0.001, 0.1):
        a, b = onp.zeros(Lx, Ly, 9), onp.zeros(Lx, Ly, 9)
        va, vb = onp.zeros(Lx, Ly), onp.zeros(Lx, Ly)
        oc.equilibrium(a, h, ux, uy, va, g)
        G.equilibrium(b, h, ux, uy, vb, g)
        _eq(a, b, "feq"); _eq(va, vb, "vsq")


@pytest.mark.parametrize("Lx,Ly", SIZES)
@pytest.mark.parametrize("tau", [1.0, 0.75, 1.3])
def test_bgk_stream_bitwise(G, Lx, Ly, tau):
    rng = np.random.default_rng(Lx * 7 + Ly)
    feq, ft = _rand((Lx, Ly, 9), rng, 0.1, 0.1), _rand((Lx, Ly, 9), rng, 0.1, 0.1)
    Fx, Fy = _rand((Lx, Ly), rng, 1e-3), _rand((Lx, Ly), rng, 1e-3)
    fo_a, ft_a = onp.zeros(Lx, Ly, 9), ft.copy(order="F")
    fo_b, ft_b = onp.zeros(Lx, Ly, 9), ft.copy(order="F")
    oc.BGKandStream(fo_a, feq, ft_a, Fx, Fy, tau)
    G.BGKandStream(fo_b, feq, ft_b, Fx, Fy, tau)
    _eq(fo_a, fo_b, "fout"); _eq(ft_a, ft_b, "ftemp"); _eq(fo_b, ft_b, "fout==ftemp")


@pytest.mark.parametrize("Lx,Ly", SIZES)
def test_moments_bitwise(G, Lx, Ly):
    rng = np.random.default_rng(Lx * 3 + Ly)
    f = _rand((Lx, Ly, 9), rng, 0.05, 0.11)
    a = [onp.zeros(Lx, Ly) for _ in range(3)]
    b = [onp.zeros(Lx, Ly) for _ in range(3)]
    oc.moments(*a, f)
    G.moments(*b, f)
    for x, y, n in zip(a, b, "h ux uy".split()):
        _eq(x, y, n)


@pytest.mark.parametrize("Lx,Ly", SIZES)
@pytest.mark.parametrize("n,m,variant", [(9, 3, "fast"), (3, 2, "fast"), (9, 3, "power_broad"), (3, 2, "power_broad"),
                                         (4, 2, "power_broad"), (6, 3, "power_broad")])
def test_filmpressure_bitwise(G, Lx, Ly, n, m, variant):
    rng = np.random.default_rng(Lx + 11 * Ly)
    h = np.abs(_rand((Lx, Ly), rng, 0.5, 1.0)) + 0.04
    ct_field = np.asfortranarray(np.cos(np.pi * (1 / 9 + 1 / 36 * rng.random((Lx, Ly)))))
    for ct in (onp.cospi(1 / 9), 0.0, ct_field):
        a, b = onp.zeros(Lx, Ly), onp.zeros(Lx, Ly)
        oc.filmpressure(a, h, onp.zeros(Lx, Ly, 8), 0.01, ct, n, m, 0.07, 0.05, variant=variant)
        G.filmpressure(b, h, onp.zeros(Lx, Ly, 8), 0.01, ct, n, m, 0.07, 0.05, variant=variant)
        _eq(a, b, "pressure")


@pytest.mark.parametrize("Lx,Ly", SIZES)
def test_stencils_bitwise(G, Lx, Ly):
    rng = np.random.default_rng(Lx + 13 * Ly)
    f, a = _rand((Lx, Ly), rng), _rand((Lx, Ly), rng)
    for mult in (None, a):
        xa, ya, xb, yb = (onp.zeros(Lx, Ly) for _ in range(4))
        oc.grad9(xa, ya, f, a=mult)
        G.grad9(xb, yb, f, a=mult)
        _eq(xa, xb, "gradx"); _eq(ya, yb, "grady")
    la, lb = onp.zeros(Lx, Ly), onp.zeros(Lx, Ly)
    oc.lap9(la, f, 0.37)
    G.lap9(lb, f, 0.37)
    _eq(la, lb, "lap")
    xa, ya, xb, yb = (onp.zeros(Lx, Ly) for _ in range(4))
    oc.hgradp(xa, ya, f, a, onp.zeros(Lx, Ly, 8))
    G.hgradp(xb, yb, f, a, None)
    _eq(xa, xb, "h∇px"); _eq(ya, yb, "h∇py")


@pytest.mark.parametrize("Lx,Ly", SIZES)
@pytest.mark.parametrize("variant", [0, 1, 2])
def test_slippage_force_inclination_bitwise(G, Lx, Ly, variant):
    rng = np.random.default_rng(Lx + 17 * Ly + variant)
    h, ux, uy = np.abs(_rand((Lx, Ly), rng, 0.3, 1.0)), _rand((Lx, Ly), rng, 0.05), _rand((Lx, Ly), rng, 0.05)
    mu = onp.Params().mu
    sa = [onp.zeros(Lx, Ly) for _ in range(2)]
    sb = [onp.zeros(Lx, Ly) for _ in range(2)]
    oc.slippage(*sa, h, ux, uy, 1.5, mu, 0.05, variant)
    G.slippage(*sb, h, ux, uy, 1.5, mu, 0.05, variant)
    _eq(sa[0], sb[0], "slipx"); _eq(sa[1], sb[1], "slipy")
    gx, gy, kx, ky = (_rand((Lx, Ly), rng, 1e-3) for _ in range(4))
    for k in ((None, None), (kx, ky)):
        Fa = [onp.zeros(Lx, Ly) for _ in range(2)]
        Fb = [onp.zeros(Lx, Ly) for _ in range(2)]
        oc.force_sum(*Fa, gx, gy, *sa, *k)
        G.force_sum(*Fb, gx, gy, *sb, *k)
        _eq(Fa[0], Fb[0], "Fx"); _eq(Fa[1], Fb[1], "Fy")
        oc.inclination(*Fa, h, [1e-4, -3e-5], 0.8807970779778823)
        G.inclination(*Fb, h, [1e-4, -3e-5], 0.8807970779778823)
        _eq(Fa[0], Fb[0], "Fx+incl"); _eq(Fa[1], Fb[1], "Fy+incl")


def test_thermal_statistics_and_amplitude(G):
    """test/forcing.jl:141-164: mean ~ 0, var ~ 2kbt/11 (+-10 %) at h = 1; plus: the deterministic amplitude
    is bit-identical to the oracle's (k / normal is the same field for both components' generators)."""
    import swalbe_b200 as sw

    def draw(kb):
        kx, ky, h = sw.Field(50, 50), sw.Field(50, 50), sw.Field(50, 50, fill=1.0)
        sw.thermal(kx, ky, h, kb, 1 / 6, 1.0, seed=1234, step=int(kb * 1000))
        return kx.numpy(), ky.numpy()

    gc.case_thermal_statistics(G, draw)
    # seeded: same (seed, step) -> same field; different step -> different field
    a1, _ = draw(0.01)
    a2, _ = draw(0.01)
    assert np.array_equal(a1, a2)
    # large-sample moments of the normals themselves (h such that the amplitude is exactly 1 is awkward; use ratio)
    Lx = Ly = 1024
    h = sw.Field(Lx, Ly).set(np.asfortranarray(1.0 + 0.5 * np.random.default_rng(1).random((Lx, Ly))))
    kx, ky = sw.Field(Lx, Ly), sw.Field(Lx, Ly)
    sw.thermal(kx, ky, h, 1e-3, 1 / 6, 1.0, seed=7, step=3)
    amp = onp.thermal_amplitude(h.numpy(), 1e-3, 1 / 6, 1.0)
    zx, zy = kx.numpy() / amp, ky.numpy() / amp
    for z in (zx, zy):
        assert abs(z.mean()) < 5e-3 and abs(z.var() - 1.0) < 5e-3
        assert abs(np.mean(z ** 3)) < 2e-2 and abs(np.mean(z ** 4) - 3.0) < 5e-2
    assert abs(np.mean(zx * zy)) < 5e-3  # x and y draws are independent


def test_field_stats(G):
    import swalbe_b200 as sw

    rng = np.random.default_rng(5)
    a = _rand((301, 77), rng, 0.2, 0.1)
    f = sw.Field(301, 77).set(a)
    mn, mx, sm, cnt = sw.field_stats(f, 0.055)
    assert mn == a.min() and mx == a.max() and cnt == int((a > 0.055).sum())
    assert abs(sm - a.sum()) < 1e-9 * abs(a).sum()


def test_thermal_moments_at_scale(G):
    """SURVEY 8d parity check for thermal configs: sample mean / variance of the in-kernel noise against
    2*kbt*mu*6h/(2h^2+6h*delta+3*delta^2) within 1 % on a 4096^2 field (16.8 M samples per component), and the fused
    loop's materialised kbtx/kbty equal to the stand-alone thermal! operator for the same (seed, step) bit for bit."""
    import torch

    import swalbe_b200 as sw

    L = 4096
    kbt, mu, delta = 1e-7, 1 / 12, 1.0
    h = sw.Field(L, L)
    i = torch.arange(L, device="cuda", dtype=torch.float64)
    h.t.copy_(1.0 + 0.1 * torch.sin(2 * np.pi * i / L)[None, :] * torch.sin(2 * np.pi * i / L)[:, None])
    kx, ky = sw.Field(L, L), sw.Field(L, L)
    sw.thermal(kx, ky, h, kbt, mu, delta, seed=1234, step=17)
    hh = h.t
    var_expected = 2 * kbt * mu * 6 * hh / (2 * hh * hh + 6 * hh * delta + 3 * delta * delta)
    for k in (kx.t, ky.t):
        z = k / torch.sqrt(var_expected)
        assert abs(z.mean().item()) < 1e-3
        assert abs(z.var().item() - 1.0) < 1e-2
        assert abs((k * k).mean().item() / var_expected.mean().item() - 1.0) < 1e-2
    # the fused loop draws the same normals for the same (seed, step, cell)
    sysc = sw.SysConst(Lx=256, Ly=192, param=sw.Taumucs(kbt=kbt, μ=mu, δ=delta))
    st = sw.Sys(sysc, "GPU", kind="thermal")
    st.height.set(np.asfortranarray(1.0 + 0.1 * np.random.default_rng(3).random((256, 192))))
    h0 = sw.Field(256, 192).set(st.height)
    sw.fused_steps(st, sysc, 1, thermal_seed=99, step0=5)
    k2x, k2y = sw.Field(256, 192), sw.Field(256, 192)
    sw.thermal(k2x, k2y, h0, kbt, mu, delta, seed=99, step=5)
    assert np.array_equal(st.kbtx.numpy(), k2x.numpy()) and np.array_equal(st.kbty.numpy(), k2y.numpy())


def test_exact_division_helper_matches_ieee_division(G):
    """div_exact/div2_exact (shared reciprocal, +-0 numerators on the fast path) must be bit-identical to `/` for every
    operand class; 2^28 random triples x 3 quotients each."""
    import ctypes as C

    import torch

    import swalbe_b200 as sw
    from swalbe_b200 import _lib

    out = torch.zeros(1, dtype=torch.int64, device="cuda")
    for seed in (1, 20261017):
        _lib.call("swalbe_selftest_division", 1 << 27, seed, C.c_void_p(out.data_ptr()), sw._stream())
        assert out.item() == 0, f"{out.item()} mismatching quotients (seed {seed})"


def test_legacy_tuple_allocator_and_array_forms(G):
    """Sys(sysc, "GPU", exotic, T) (src/initialize.jl:358-475) hands out bare arrays; a docs-style loop over the array
    forms (docs/src/tutorials.md) must equal the oracle."""
    import swalbe_b200 as sw

    sysc = sw.SysConst(Lx=40, Ly=30, param=sw.Taumucs(n=3, m=2, hmin=0.07))
    p = sysc.param
    arrs = sw.Sys(sysc, "GPU", False, float)
    assert len(arrs) == 15 and len(sw.Sys(sysc, "GPU", True, float)) == 17
    fout, ftemp, feq, height, velx, vely, vsq, pressure, dgrad, Fx, Fy, slipx, slipy, hpx, hpy = arrs
    assert np.all(height.numpy() == 1.0) and np.all(fout.numpy() == 0.0)
    rng = np.random.default_rng(2)
    h0 = np.asfortranarray(np.abs(1.0 + 0.1 * rng.standard_normal((40, 30))) + 0.06)
    height.set(h0)
    ref = onp.State(40, 30)
    ref.height[...] = h0
    op = onp.Params(n=3, m=2, hmin=0.07)
    for _ in range(4):
        sw.filmpressure(pressure, height, dgrad, p.gamma, p.theta, p.n, p.m, p.hmin, p.hcrit)
        sw.gradf(hpx, hpy, pressure, dgrad, height)
        sw.slippage(slipx, slipy, height, velx, vely, p.delta, p.mu)
        st = sw.CuState.__new__(sw.CuState)
        st.Lx, st.Ly, st.Fx, st.Fy, st.hgradpx, st.hgradpy, st.slipx, st.slipy = 40, 30, Fx, Fy, hpx, hpy, slipx, slipy
        sw.update(st)
        sw.equilibrium(feq, height, velx, vely, vsq, p.g)
        sw.BGKandStream(fout, feq, ftemp, Fx, Fy, p.tau)
        sw.moments(height, velx, vely, fout)
        oc.step(ref, op, pvariant="fast")
    for got, want in ((height, ref.height), (velx, ref.velx), (fout, ref.fout), (pressure, ref.pressure)):
        assert np.array_equal(got.numpy(), want)


def test_cospi_field_on_device(G):
    """swalbe_cospi_field (libdevice cospi, what CUDA.jl's cospi.(θ) broadcast calls) against the host cospi: exact at the
    exact points, within 1 ulp elsewhere."""
    import swalbe_b200 as sw

    rng = np.random.default_rng(0)
    th = np.asfortranarray(rng.random((64, 33)) * 2 - 0.5)
    th[0, :8] = [0.0, 0.5, 1.0, 1.5, 2.0, 1 / 9, 1 / 6, 1 / 3]
    got = sw.cospi_field(sw.Field(64, 33).set(th)).numpy()
    want = np.vectorize(sw.cospi)(th)
    assert list(got[0, :5]) == [1.0, 0.0, -1.0, 0.0, 1.0]
    assert np.max(np.abs(got - want)) <= 1.2e-16


def test_philox_known_answers(G):
    """The device's Philox4x32-10 against the known-answer vectors shipped with Random123 (kat_vectors)."""
    import ctypes as C

    import torch

    from swalbe_b200 import _lib

    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    out = torch.zeros(4, dtype=torch.int32, device="cuda")
    for ctr, key, want in kat:
        _lib.call("swalbe_selftest_philox", C.byref((C.c_uint * 4)(*ctr)), C.byref((C.c_uint * 2)(*key)),
                  C.c_void_p(out.data_ptr()), None)
        got = tuple(int(v) & 0xffffffff for v in out.cpu().tolist())
        assert got == want, (ctr, key, [hex(g) for g in got])


def test_thermal_normals_distribution(G):
    """Distribution of the in-kernel normals through swalbe_thermal on a flat film (k = N * const): moments to 6th order,
    tail fractions, Kolmogorov-Smirnov distance, independence of the two components and of neighbouring cells/steps."""
    import swalbe_b200 as sw
    from scipy import stats

    L = 2048
    sysc = sw.SysConst(Lx=L, Ly=L, param=sw.Taumucs(kbt=1e-6))
    st = sw.Sys(sysc, "GPU", kind="thermal")
    st.height.set(1.0)
    p = sysc.param
    amp = np.sqrt(2 * p.kbt * p.μ * 6 * 1.0 / (2 * 1.0 + 6 * 1.0 * p.δ + 3 * p.δ * p.δ))
    sw.thermal(st, sysc, seed=99, step=5)
    zx, zy = st.kbtx.numpy() / amp, st.kbty.numpy() / amp
    n = zx.size
    for z in (zx, zy):
        assert abs(z.mean()) < 4 / np.sqrt(n) and abs(z.var() - 1) < 4 * np.sqrt(2 / n)
        assert abs((z ** 3).mean()) < 4 * np.sqrt(15 / n) and abs((z ** 4).mean() - 3) < 4 * np.sqrt(96 / n)
        assert abs((z ** 6).mean() - 15) < 4 * np.sqrt(10170 / n)
        for t, pt in ((3.0, 2.6998e-3), (4.0, 6.334e-5)):
            frac = (np.abs(z) > t).mean()
            assert abs(frac - pt) < 5 * np.sqrt(pt / n), (t, frac)
        assert stats.kstest(z.ravel()[:2_000_000], "norm").statistic < 1.63 / np.sqrt(2_000_000)  # 1 % level
    c = lambda a, b: abs(np.mean(a * b))  # noqa: E731
    lim = 4 / np.sqrt(n)
    assert c(zx, zy) < lim and c(zx[1:, :], zx[:-1, :]) < lim and c(zx[:, 1:], zx[:, :-1]) < lim
    assert c(zx * zx - 1, zy * zy - 1) < 4 * 2 / np.sqrt(n)
    sw.thermal(st, sysc, seed=99, step=6)
    assert c(zx, st.kbtx.numpy() / amp) < lim
    sw.thermal(st, sysc, seed=99, step=5)
    assert np.array_equal(zx, st.kbtx.numpy() / amp)  # counter-based: reproducible


def test_thermal_draws_fresh_noise_per_call_like_randn(G):
    """thermal!(state, sys) in a user loop (scripts/Rivulet_stability.jl:120-124) must not repeat its field: without
    seed/step the call counter advances the Philox stream; with both, the field is reproducible."""
    import swalbe_b200 as sw

    sysc = sw.SysConst(Lx=64, Ly=48, param=sw.Taumucs(kbt=1e-6))
    st = sw.Sys(sysc, "GPU", kind="thermal")
    sw.thermal(st, sysc)
    a = st.kbtx.numpy()
    sw.thermal(st, sysc)
    b = st.kbtx.numpy()
    assert not np.array_equal(a, b) and abs(np.corrcoef(a.ravel(), b.ravel())[0, 1]) < 0.1
    sw.thermal(st, sysc, seed=5, step=9)
    c = st.kbtx.numpy()
    sw.thermal(st, sysc, seed=5, step=9)
    assert np.array_equal(c, st.kbtx.numpy())
    f1 = sw.randinterface(sw.Field(64, 48), 1.0, 0.01).numpy()
    f2 = sw.randinterface(sw.Field(64, 48), 1.0, 0.01).numpy()
    assert not np.array_equal(f1, f2)


def test_gradf_scalar_multiplier_and_async_snapshots(G):
    """∇f!(outx, outy, f, a) with a scalar a (src/differences.jl:171-187 broadcasts it); snapshot! into pinned host memory
    through a staging plane while the loop keeps running (src/measures.jl:99-105)."""
    import swalbe_b200 as sw

    Lx, Ly = 40, 33
    rng = np.random.default_rng(8)
    f = np.asfortranarray(rng.random((Lx, Ly)))
    ox, oy, fd = sw.Field(Lx, Ly), sw.Field(Lx, Ly), sw.Field(Lx, Ly).set(f)
    sw.gradf(ox, oy, fd, 0.37)
    wx, wy = np.zeros((Lx, Ly), order="F"), np.zeros((Lx, Ly), order="F")
    oc.grad9(wx, wy, f, a=np.full((Lx, Ly), 0.37, order="F"))
    assert np.array_equal(ox.numpy(), wx) and np.array_equal(oy.numpy(), wy)
    sysc = sw.SysConst(Lx=Lx, Ly=Ly, param=sw.Taumucs(Tmax=12, tdump=100, g=-0.001))
    st = sw.Sys(sysc, "GPU")
    st.height.set(np.asfortranarray(1.0 + 0.1 * rng.random((Lx, Ly))))
    snaps, plain = sw.SnapshotBuffer(3, Lx, Ly), np.zeros((3, Lx * Ly))
    for t in range(1, 13):
        sw.fused_steps(st, sysc, 1, skip_aux=True)
        sw.snapshot(snaps, st.height, t, dumping=4)
        sw.snapshot(plain, st.height, t, dumping=4)
    assert np.array_equal(snaps.array(), plain) and plain[2].any() and not np.array_equal(plain[0], plain[2])
