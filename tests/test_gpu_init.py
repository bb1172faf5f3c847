"""On-device initial conditions and substrate motion through the C ABI, against the oracle's
restatement of src/initialvalues.jl and the reference's own known answers (test/initialvalues.jl).

sin/cos/asin on the device are CUDA's, the oracle's are glibc's (Julia's are openlibm's): all within an ulp or two of the
true value, so these comparisons are bounded at 4 ulp of the field's scale instead of bitwise.  Index logic (slabs,
periodic shifts, branch structure) is exact and is tested exactly."""
import numpy as np
import pytest

from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sw():
    import swalbe_b200 as sw_

    return sw_


def _findmax_index(a):  # Julia's findmax: first maximum in column-major order, 1-based
    i, j = np.unravel_index(a.ravel(order="F").argmax(), a.shape, order="F")
    return int(i) + 1, int(j) + 1


def _close(got, want, scale):
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) <= 4 * np.finfo(np.float64).eps * scale


def test_singledroplet_device(sw):  # test/initialvalues.jl:16-25
    f = sw.singledroplet(sw.Field(100, 100), 50, 1 / 3, (50, 50))
    h = f.numpy()
    assert h.max() == 50 * (1 - sw.cospi(1 / 3)) and _findmax_index(h) == (50, 50)
    _close(h, onp.singledroplet(100, 100, 50, 1 / 3, (50, 50)), 50)
    for (Lx, Ly, r, th, c) in [(150, 150, 35, 1 / 6, (75, 75)), (257, 130, 60.5, 1 / 9, (100, 31)), (64, 64, 200, 1 / 2, (1, 64))]:
        got = sw.singledroplet(sw.Field(Lx, Ly), r, th, c).numpy()
        want = onp.singledroplet(Lx, Ly, r, th, c)
        _close(got, want, r)
        assert np.array_equal(got == 0.05, want == 0.05)  # same precursor region


def test_rivulet_device(sw):  # test/initialvalues.jl:46-63
    rad, th, lx, ly, c = 45, 1 / 4, 150, 200, 80
    top = rad * (1 - sw.cospi(th))
    h = sw.rivulet(sw.Field(lx, ly), rad, th, "y", c, 0.05).numpy()
    assert abs(h.max() - top) < 1e-4 and abs(h[c - 1, :].sum() - top * ly) < 1e-4
    _close(h, onp.rivulet(lx, ly, rad, th, "y", c, 0.05), rad)
    h = sw.rivulet(sw.Field(lx, ly), rad, th, "x", c, 0.05).numpy()
    assert abs(h[:, c - 1].sum() - top * lx) < 1e-4
    _close(h, onp.rivulet(lx, ly, rad, th, "x", c, 0.05), rad)
    h = sw.rivulet(sw.Field(200, 200), 45, 1 / 9, "y", 100, 0.05).numpy()  # doctest src/initialvalues.jl:50-62
    assert h.max() == 45 * (1 - sw.cospi(1 / 9)) and _findmax_index(h) == (100, 1)
    with pytest.raises(ValueError):
        sw.rivulet(sw.Field(8, 8), 3, 1 / 9, "z", 4)


def test_torus_device(sw):  # test/initialvalues.jl:65-76 and the doctest src/initialvalues.jl:126-137
    h = sw.torus(sw.Field(150, 200), 10, 45, 1 / 9, (80, 80), 0.05).numpy()
    assert h.min() == 0.05 and np.isclose(h.max(), (1 - sw.cospi(1 / 9)) * 10) and _findmax_index(h) == (80, 35)
    _close(h, onp.torus(150, 200, 10, 45, 1 / 9, (80, 80), 0.05), 10)
    h = sw.torus(sw.Field(256, 256), 45, 80, 1 / 9, (128, 128)).numpy()
    assert np.isclose(h.max(), 45 * (1 - sw.cospi(1 / 9))) and _findmax_index(h) == (128, 48)
    # noise: only the wetted part is perturbed, with the requested standard deviation
    hn = sw.torus(sw.Field(256, 256), 45, 80, 1 / 9, (128, 128), noise=0.01, seed=5).numpy()
    wet = h > 0.05 + 0.1
    d = (hn - h)[wet]
    assert np.array_equal(hn[h == 0.05], h[h == 0.05]) and abs(d.std() - 0.01) < 0.001 and abs(d.mean()) < 0.001


def test_sinewave2d_device_matches_rayleightaylor_ic(sw):  # src/simulate.jl:350-353
    for (Lx, Ly, kx, ky, eps) in [(100, 100, 15, 18, 0.01), (130, 77, 4, 5, 0.001)]:
        got = sw.sinewave2d(sw.Field(Lx, Ly), 1.0, eps, kx, ky).numpy()
        want = onp.rayleightaylor_ic(Lx, Ly, kx, ky, 1.0, eps)
        _close(got, want, 1.0)
    # README configuration: a trajectory started from the device-built field stays within 1e-12 of the oracle's
    sysc = sw.SysConst(Lx=100, Ly=100, param=sw.Taumucs(g=-0.001, γ=0.0005, Tmax=200))
    st = sw.Sys(sysc, "GPU")
    sw.sinewave2d(st.height, 1.0, 0.01)
    sw.equilibrium(st, sysc)
    sw.time_loop(sysc, st, verbose=False)
    ref = onp.State(100, 100)
    ref.height[...] = onp.rayleightaylor_ic(100, 100, eps=0.01)
    p = onp.Params(g=-0.001, gamma=0.0005)
    onp.equilibrium(ref.feq, ref.height, ref.velx, ref.vely, ref.vsq, p.g)
    onp.time_loop(ref, p, 200)
    assert np.max(np.abs(st.height.numpy() - ref.height)) <= 1e-12 * np.max(np.abs(ref.height))


def test_randinterface_device(sw):  # test/initialvalues.jl:9-14 + statistics
    h = sw.randinterface(sw.Field(10, 10), 2.0, 0.1, seed=1).numpy()
    assert h.max() <= 3.0 and h.min() >= 1.0
    a = sw.randinterface(sw.Field(512, 512), 1.0, 0.01, seed=7).numpy()
    z = (a - 1.0) / 0.01
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01 and abs(((z ** 4).mean()) - 3.0) < 0.1
    b = sw.randinterface(sw.Field(512, 512), 1.0, 0.01, seed=7).numpy()
    c = sw.randinterface(sw.Field(512, 512), 1.0, 0.01, seed=8).numpy()
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert abs(np.corrcoef(z[:-1, :].ravel(), z[1:, :].ravel())[0, 1]) < 0.01


def test_slab_construction_equals_global(sw):
    """Every rank building its own rows gives the same field as one device building all of it (bitwise)."""
    Lx, Ly, parts = 96, 120, 4
    rows = Ly // parts
    builders = {
        "droplet": lambda f, jb: sw.singledroplet(f, 40, 1 / 6, (48, 60), j_begin=jb),
        "torus": lambda f, jb: sw.torus(f, 10, 30, 1 / 9, (48, 60), noise=0.01, seed=3, j_begin=jb),
        "rivulet": lambda f, jb: sw.rivulet(f, 20, 1 / 9, "x", 60, noise=0.01, seed=4, j_begin=jb),
        "sine": lambda f, jb: sw.sinewave2d(f, 1.0, 0.01, 3, 4, j_begin=jb, Ly=Ly),
        "rand": lambda f, jb: sw.randinterface(f, 1.0, 0.01, seed=9, j_begin=jb),
    }
    for name, build in builders.items():
        whole = build(sw.Field(Lx, Ly), 0).numpy()
        for r in range(parts):
            slab = build(sw.Field(Lx, rows), r * rows).numpy()
            assert np.array_equal(slab, whole[:, r * rows:(r + 1) * rows]), (name, r)


def test_circshift_and_move_substrate(sw):  # scripts/Moving_wettability_structs.jl:139-152
    rng = np.random.default_rng(0)
    for (Lx, Ly) in [(5, 5), (33, 18), (130, 257)]:
        a = np.asfortranarray(rng.random((Lx, Ly)))
        src = sw.Field(Lx, Ly).set(a)
        for sh in [(1, 1), (1, 0), (0, 1), (-1, 2), (Lx + 3, -Ly - 2), (0, 0)]:
            got = sw.circshift(sw.Field(Lx, Ly), src, sh).numpy()
            assert np.array_equal(got, onp.circshift(a, sh)), (Lx, Ly, sh)
    with pytest.raises(ValueError):
        sw.circshift(src, src, (1, 1))
    Lx, Ly = 24, 20
    a = np.asfortranarray(rng.random((Lx, Ly)) * 0.1 + 0.1)
    th, inp = sw.Field(Lx, Ly).set(a), sw.Field(Lx, Ly).set(a)
    want = a
    for t in range(0, 31):
        sw.move_substrate(th, inp, t, 10, direction="diagonal")
        if t % 10 == 0 and t > 0:
            want = onp.circshift(want, (1, 1))
        assert np.array_equal(th.numpy(), want) and np.array_equal(inp.numpy(), want)
    sw.move_substrate(th, inp, 40, 10, direction="x")
    assert np.array_equal(th.numpy(), onp.circshift(want, (1, 0)))


def test_moving_substrate_loop_matches_oracle(sw):
    """The loop of scripts/Moving_wettability_structs.jl (pressure with a θ field that moves every `tmove` steps):
    device circshift + device cospi + fused steps against the oracle stepping with the same cospi values."""
    Lx, Ly, tmove, T = 64, 48, 7, 30
    rng = np.random.default_rng(2)
    i = np.arange(Lx)[:, None]; j = np.arange(Ly)[None, :]
    theta = np.asfortranarray(1 / 9 + (1 / 36) * np.sin(2 * np.pi * 2 * i / Lx) * np.sin(2 * np.pi * 2 * j / Ly))
    h0 = np.asfortranarray(1 + 0.1 * rng.standard_normal((Lx, Ly)) * 0.1)
    sysc = sw.SysConst(Lx=Lx, Ly=Ly, param=sw.Taumucs(n=3, m=2, hmin=0.07, γ=0.01))
    st = sw.Sys(sysc, "GPU")
    st.height.set(h0)
    th, inp = sw.Field(Lx, Ly).set(theta), sw.Field(Lx, Ly).set(theta)
    ref = onp.State(Lx, Ly)
    ref.height[...] = h0
    ct = sw.cospi_field(th).numpy()
    for t in range(T):
        sw.move_substrate(th, inp, t, tmove)
        if t % tmove == 0 and t > 0:
            ct = onp.circshift(ct, (1, 1))
        sw.fused_steps(st, sysc, 1, θ=th)
        onp.step(ref, onp.Params(n=3, m=2, hmin=0.07, gamma=0.01), cospi_theta=ct)
    for name in ("height", "velx", "vely", "pressure"):
        assert np.array_equal(getattr(st, name).numpy(), getattr(ref, name)), name
