"""Parity of the fused time loop (swalbe_time_loop, one fused kernel per step) with the oracle: every field of
the state, bit for bit (numerically equal doubles), after N steps -- plus the reference's whole-loop known answers
(test/simulate.jl) and size-independent properties at BASELINE sizes."""
import math

import numpy as np
import pytest

from oracle import oracle_c as oc
from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu

STATE_FIELDS = ("height", "velx", "vely", "fout", "ftemp", "feq", "vsq", "pressure", "hgradpx", "hgradpy", "slipx",
                "slipy", "Fx", "Fy")


@pytest.fixture(scope="module")
def sw():
    import swalbe_b200

    return swalbe_b200


def _mk(sw, Lx, Ly, seed, prm_kw, tau_pops=False):
    rng = np.random.default_rng(seed)
    h0 = np.asfortranarray(np.abs(1.0 + 0.2 * rng.standard_normal((Lx, Ly))) + 0.06)
    ux0 = np.asfortranarray(0.01 * rng.standard_normal((Lx, Ly)))
    uy0 = np.asfortranarray(0.01 * rng.standard_normal((Lx, Ly)))
    ft0 = np.asfortranarray(0.1 + 0.01 * rng.random((Lx, Ly, 9)))
    ref = onp.State(Lx, Ly)
    ref.height[...] = h0; ref.velx[...] = ux0; ref.vely[...] = uy0
    sysc = sw.SysConst(Lx=Lx, Ly=Ly, param=sw.Taumucs(**prm_kw))
    st = sw.Sys(sysc, "GPU")
    st.height.set(h0); st.velx.set(ux0); st.vely.set(uy0)
    if tau_pops:
        ref.ftemp[...] = ft0
        st.ftemp.set(ft0)
    okw = {{"γ": "gamma", "δ": "delta", "τ": "tau", "μ": "mu", "θ": "theta"}.get(k, k): v for k, v in prm_kw.items()}
    return st, sysc, ref, onp.Params(**okw)


@pytest.fixture(params=["cluster", "tile", "march"])
def small_lattice_flavour(request, monkeypatch):
    """Lean steps of lattices that fit one thread-block cluster run inside the persistent cluster kernel, lattices
    <= 512^2 through the tile kernel; "tile" switches the cluster kernel off, "march" both, so that all three flavours
    see every small parity case."""
    if request.param != "cluster":
        monkeypatch.setenv("SWALBE_CLUSTER", "0")
    if request.param == "march":
        monkeypatch.setenv("SWALBE_TILE_MAX", "0")
    return request.param


def _compare(st, ref, fields=STATE_FIELDS, what=""):
    for name in fields:
        got, want = getattr(st, name).numpy(), getattr(ref, name)
        assert np.array_equal(got, want, equal_nan=True), \
            f"{what}{name}: {np.count_nonzero(got != want)} sites differ, max abs {np.nanmax(np.abs(got - want)):.3e}"


@pytest.mark.parametrize("Lx,Ly", [(5, 5), (25, 26), (150, 200), (100, 96), (257, 19), (600, 40), (7, 300)])
@pytest.mark.parametrize("nsteps", [1, 2, 7])
def test_fused_loop_bitwise_default_params(sw, Lx, Ly, nsteps, small_lattice_flavour):
    st, sysc, ref, p = _mk(sw, Lx, Ly, seed=Lx * 1000 + Ly, prm_kw=dict(g=-0.001, γ=0.0005))
    sw.fused_steps(st, sysc, nsteps)
    oc.time_loop(ref, p, nsteps=nsteps)
    _compare(st, ref)


@pytest.mark.parametrize("Lx,Ly", [(1, 1), (2, 3), (3, 1), (1, 40), (9, 2)])
def test_fused_loop_tiny_extents(sw, Lx, Ly, small_lattice_flavour):
    """Degenerate periodic lattices: every neighbour is a wrapped copy of the few sites there are."""
    for kw, pops in ((dict(g=-0.001), False), (dict(τ=0.9), True)):
        st, sysc, ref, p = _mk(sw, Lx, Ly, seed=Lx * 10 + Ly, prm_kw=kw, tau_pops=pops)
        sw.fused_steps(st, sysc, 3)
        oc.time_loop(ref, p, nsteps=3)
        _compare(st, ref, what=f"{Lx}x{Ly}:")


@pytest.mark.parametrize("prm_kw", [
    dict(n=3, m=2, hmin=0.07, γ=0.01),
    dict(n=4, m=2, δ=2.0),
    dict(τ=0.75, g=0.002),
    dict(τ=1.3, n=3, m=2),
    dict(μ=1 / 12, δ=0.25, θ=1 / 6),
], ids=lambda d: ",".join(f"{k}={v:.3g}" if isinstance(v, float) else f"{k}={v}" for k, v in d.items()))
@pytest.mark.parametrize("nsteps", [1, 4, 5])
def test_fused_loop_bitwise_param_sets(sw, prm_kw, nsteps, small_lattice_flavour):
    st, sysc, ref, p = _mk(sw, 70, 45, seed=11, prm_kw=prm_kw, tau_pops="τ" in prm_kw)
    sw.fused_steps(st, sysc, nsteps)
    oc.time_loop(ref, p, nsteps=nsteps)
    _compare(st, ref)


def test_fused_loop_variants_theta_field_slip_inclination(sw):
    Lx, Ly = 64, 50
    rng = np.random.default_rng(3)
    theta = np.asfortranarray(1 / 9 + 1 / 36 * rng.random((Lx, Ly)))
    ct = sw.cospi_field(sw.Field(Lx, Ly).set(theta)).numpy()  # the device's cospi.(θ) is handed to the oracle as data
    for sv in (0, 1, 2):
        st, sysc, ref, p = _mk(sw, Lx, Ly, seed=sv, prm_kw=dict(n=3, m=2, hmin=0.07))
        thf = sw.Field(Lx, Ly).set(theta)
        factor = 0.5 + 0.5 * math.tanh(1.0)
        sw.fused_steps(st, sysc, 5, θ=thf, slip_variant=sv, incl=([1e-4, -2e-5], factor))
        oc.time_loop(ref, p, nsteps=5, cospi_theta=ct, slip_variant=sv, incl=([1e-4, -2e-5], factor))
        _compare(st, ref, what=f"slip{sv}:")
    # array-form (fast_93) pressure through the fused loop, as CuState_thermal uses it (src/pressure.jl:117)
    st, sysc, ref, p = _mk(sw, Lx, Ly, seed=9, prm_kw=dict())
    from swalbe_b200 import _lib

    sw.fused_steps(st, sysc, 3, pressure_variant=_lib.PRESSURE_FAST)
    oc.time_loop(ref, p, nsteps=3, pvariant="fast")
    _compare(st, ref, what="fast93:")


@pytest.mark.parametrize("prm_kw", [dict(τ=0.9), dict(τ=0.75, n=3, m=2, hmin=0.07, g=0.001), dict(τ=1.4, n=4, m=2)],
                         ids=["tau0.9", "tau0.75-32-g", "tau1.4-generic"])
def test_general_tau_from_moments_kernels(sw, prm_kw, monkeypatch, small_lattice_flavour):
    """tau != 1: after the first step the loop derives h and u from the populations (FM kernels, 144 B/LU).  Chunked
    calls that vouch for their moments (SWALBE_LOOP_MOMENTS_CONSISTENT), one long call, the plane-reading kernels
    (SWALBE_FM=0) and the oracle must all agree bit for bit, on every field of the state."""
    Lx, Ly, chunks = 150, 61, (1, 3, 2, 1)
    st, sysc, ref, p = _mk(sw, Lx, Ly, seed=77, prm_kw=prm_kw, tau_pops=True)
    done = 0
    for n in chunks:  # the first chunk starts from an initial condition: height is NOT the moment of ftemp
        sw.fused_steps(st, sysc, n, moments_consistent=done > 0, skip_aux=done + n < sum(chunks))
        done += n
    oc.time_loop(ref, p, nsteps=sum(chunks))
    _compare(st, ref, what="chunked:")
    st1, _, _, _ = _mk(sw, Lx, Ly, seed=77, prm_kw=prm_kw, tau_pops=True)
    sw.fused_steps(st1, sysc, sum(chunks))
    _compare(st1, ref, what="one call:")
    monkeypatch.setenv("SWALBE_FM", "0")
    st0, _, _, _ = _mk(sw, Lx, Ly, seed=77, prm_kw=prm_kw, tau_pops=True)
    sw.fused_steps(st0, sysc, sum(chunks))
    _compare(st0, ref, what="FM off:")
    monkeypatch.delenv("SWALBE_FM")
    # options that send the FM steps through the run-time-option kernel: theta field, slip variant, per-step logs
    rng = np.random.default_rng(5)
    theta = np.asfortranarray(1 / 9 + 1 / 36 * rng.random((Lx, Ly)))
    thf = sw.Field(Lx, Ly).set(theta)
    ct = sw.cospi_field(thf).numpy()
    st2, _, ref2, _ = _mk(sw, Lx, Ly, seed=78, prm_kw=prm_kw, tau_pops=True)
    mn, mx, wet = sw.fused_steps(st2, sysc, 5, θ=thf, slip_variant=1, log_minmax=True, log_wetted=True, hthresh=1.0)
    dh, w = oc.time_loop(ref2, p, nsteps=5, cospi_theta=ct, slip_variant=1, log_dh=True, log_wetted=True, hthresh=1.0)
    _compare(st2, ref2, what="options:")
    assert np.array_equal((mx - mn).cpu().numpy(), np.asarray(dh)) and wet.cpu().tolist() == [int(v) for v in w]


@pytest.mark.parametrize("tile_theta", ["1", "0"], ids=["tile", "marching"])
@pytest.mark.parametrize("Lx,Ly", [(33, 9), (70, 20), (130, 64), (256, 96)])
def test_contact_angle_field_small_lattices(sw, monkeypatch, tile_theta, Lx, Ly):
    """θ(x, y) and nothing else -- the moving-wettability scripts (scripts/Moving_wettability_structs.jl:28-71) at their
    lattice sizes: the tile kernel instantiated for a contact-angle field (default below SWALBE_TILE_MAX sites) and the
    run-time-option marching kernel (SWALBE_TILE_THETA=0) against the oracle, with a substrate move in the middle."""
    monkeypatch.setenv("SWALBE_TILE_THETA", tile_theta)
    rng = np.random.default_rng(Lx)
    theta = np.asfortranarray(1 / 9 + 1 / 36 * rng.random((Lx, Ly)))
    for kw, pv in ((dict(n=3, m=2, hmin=0.07), None), (dict(g=-0.001), "fast")):
        st, sysc, ref, p = _mk(sw, Lx, Ly, seed=Lx + Ly, prm_kw=kw)
        th, inp = sw.Field(Lx, Ly).set(theta), sw.Field(Lx, Ly).set(theta)
        from swalbe_b200 import _lib

        opt = dict(pressure_variant=_lib.PRESSURE_FAST) if pv else {}
        ct0 = sw.cospi_field(th).numpy()
        sw.fused_steps(st, sysc, 4, θ=th, skip_aux=True, **opt)
        sw.move_substrate(th, inp, 98, 98)
        sw.fused_steps(st, sysc, 3, θ=th, **opt)
        okw = dict(pvariant="fast") if pv else {}
        oc.time_loop(ref, p, nsteps=4, cospi_theta=ct0, **okw)
        oc.time_loop(ref, p, nsteps=3, cospi_theta=onp.circshift(ct0, (1, 1)), **okw)
        _compare(st, ref, what=f"theta {Lx}x{Ly}:")


def test_fused_equals_operator_by_operator_on_gpu(sw):
    """The fused kernel against the seven per-operator kernels run in the reference's order (both on the GPU)."""
    st, sysc, _, _ = _mk(sw, 130, 61, seed=21, prm_kw=dict(g=0.001))
    st2, _, _, _ = _mk(sw, 130, 61, seed=21, prm_kw=dict(g=0.001))
    sw.fused_steps(st, sysc, 6)
    for _ in range(6):  # src/simulate.jl:15-22
        sw.filmpressure(st2, sysc)
        sw.hgradp(st2)
        sw.slippage(st2, sysc)
        sw.update(st2)
        sw.equilibrium(st2, sysc)
        sw.BGKandStream(st2, sysc)
        sw.moments(st2)
    for name in STATE_FIELDS:
        assert np.array_equal(getattr(st, name).numpy(), getattr(st2, name).numpy()), name


def test_lazy_populations_same_result(sw, small_lattice_flavour):
    st, sysc, ref, p = _mk(sw, 90, 70, seed=5, prm_kw=dict())
    sw.fused_steps(st, sysc, 9, lazy_populations=True)
    oc.time_loop(ref, p, nsteps=9)
    _compare(st, ref)
    st3, sysc3, _, _ = _mk(sw, 20, 20, seed=5, prm_kw=dict(τ=0.8))
    with pytest.raises(ValueError):
        sw.fused_steps(st3, sysc3, 2, lazy_populations=True)


def test_logs_minmax_wetted(sw):
    st, sysc, ref, p = _mk(sw, 150, 33, seed=8, prm_kw=dict(g=-0.002))
    mn, mx, wet = sw.fused_steps(st, sysc, 12, log_minmax=True, log_wetted=True, hthresh=1.0)
    dh, w = oc.time_loop(ref, p, nsteps=12, log_dh=True, log_wetted=True, hthresh=1.0)
    assert np.array_equal((mx - mn).cpu().numpy(), dh)
    assert np.array_equal(wet.cpu().numpy(), w)


def test_geometry_overrides_do_not_change_bits(sw, monkeypatch):
    """Every CTA width / rows-per-CTA choice must give identical fields (tiling is not allowed to matter)."""
    base = None
    for nt, rows in [(128, 16), (192, 7), (256, 64), (128, 1000)]:
        monkeypatch.setenv("SWALBE_NT", str(nt))
        monkeypatch.setenv("SWALBE_ROWS", str(rows))
        st, sysc, _, _ = _mk(sw, 300, 90, seed=2, prm_kw=dict(g=-0.001))
        sw.fused_steps(st, sysc, 3)
        cur = {n: getattr(st, n).numpy() for n in STATE_FIELDS}
        if base is None:
            base = cur
        for n in STATE_FIELDS:
            assert np.array_equal(base[n], cur[n]), (nt, rows, n)


@pytest.mark.parametrize("bulk", ["0", "2"], ids=["ldgsts", "per-warp-bulk"])
def test_neighbour_sync_kernels_same_bits(sw, monkeypatch, bulk):
    """SWALBE_NSYNC=1: the strict lean kernels that replace the per-row CTA barrier by warp-to-warp mbarrier hand-shakes
    (per-thread LDGSTS rows and per-warp TMA rows) against the oracle and against the thermal kernel with the barrier."""
    monkeypatch.setenv("SWALBE_NSYNC", "1")
    monkeypatch.setenv("SWALBE_TILE_MAX", "0")
    monkeypatch.setenv("SWALBE_BULK", bulk)
    for (Lx, Ly), nt in (((600, 130), 0), ((512, 96), 224), ((300, 90), 128), ((258, 40), 160)):
        if nt:
            monkeypatch.setenv("SWALBE_NT", str(nt))
        for kw in (dict(g=-0.001, γ=0.0005), dict(n=3, m=2, hmin=0.07)):
            st, sysc, ref, p = _mk(sw, Lx, Ly, seed=Lx + nt, prm_kw=kw)
            sw.fused_steps(st, sysc, 6)
            oc.time_loop(ref, p, nsteps=6)
            _compare(st, ref, what=f"NS {Lx}x{Ly} nt={nt}:")
    monkeypatch.delenv("SWALBE_NT", raising=False)
    # thermal: same noise, so the two synchronisation flavours must agree bit for bit
    sysc = sw.SysConst(Lx=520, Ly=70, param=sw.Taumucs(kbt=1e-6))
    out = []
    for ns in ("1", "0"):
        monkeypatch.setenv("SWALBE_NSYNC", ns)
        st = sw.Sys(sysc, "GPU", kind="thermal")
        st.height.set(np.asfortranarray(1.0 + 0.1 * np.random.default_rng(3).random((520, 70))))
        sw.fused_steps(st, sysc, 5, thermal_seed=11, step0=3)
        out.append({n: getattr(st, n).numpy() for n in ("height", "velx", "fout", "kbtx")})
    for n in out[0]:
        assert np.array_equal(out[0][n], out[1][n]), n


@pytest.mark.parametrize("csize", [0, 1, 2, 4, 8, 16])
def test_persistent_cluster_kernel_sizes_logs_and_chunks(sw, monkeypatch, csize):
    """The persistent cluster kernel on every cluster size (0 = the library's choice): chunked calls (the drivers'
    tdump chunks), lazy populations, per-step logs, in place on the state's planes; uneven slabs (100 rows on 8 / 16 CTAs)."""
    if csize:
        monkeypatch.setenv("SWALBE_CLUSTER_SIZE", str(csize))
    for (Lx, Ly), kw in (((100, 100), dict(g=-0.001, γ=0.0005)), ((64, 50), dict(n=3, m=2, hmin=0.07)), ((37, 128), dict())):
        if csize and Ly // csize < 3:
            continue
        st, sysc, ref, p = _mk(sw, Lx, Ly, seed=Lx, prm_kw=kw)
        logs, done = [], 0
        for n, lazy in ((3, False), (1, True), (6, True), (2, False)):
            mn, mx, wet = sw.fused_steps(st, sysc, n, log_minmax=True, log_wetted=True, hthresh=1.0, lazy_populations=lazy,
                                         skip_aux=done + n < 12)
            logs.append(((mx - mn).cpu().numpy(), wet.cpu().numpy()))
            done += n
        dh, w = oc.time_loop(ref, p, nsteps=12, log_dh=True, log_wetted=True, hthresh=1.0)
        _compare(st, ref, what=f"cluster {csize} {Lx}x{Ly}:")
        assert np.array_equal(np.concatenate([a for a, _ in logs]), np.asarray(dh))
        assert np.array_equal(np.concatenate([b for _, b in logs]), np.asarray(w))


# ---- reference whole-loop known answers (test/simulate.jl) through the drop-in drivers ---------------------


def test_run_flat_stays_exactly_flat(sw):  # test/simulate.jl:4-7
    sysc = sw.SysConst(Lx=25, Ly=25, param=sw.Taumucs(Tmax=200, tdump=100))
    h = sw.run_flat(sysc, "GPU", verbos=False).numpy()
    assert np.all(h == 1.0) and h.sum() == 25 * 25


def test_run_random_flattens(sw):  # test/simulate.jl:17-21
    sysc = sw.SysConst(Lx=25, Ly=25, param=sw.Taumucs(Tmax=10000, tdump=5000))
    h = sw.run_random(sysc, "GPU", ϵ=0.1, verbos=False, rng=np.random.default_rng(42)).numpy()
    assert h.max() - h.min() < 0.1


def test_run_rayleightaylor_grows_and_matches_oracle(sw):  # test/simulate.jl:33-35 + BASELINE config 1 (README)
    sysc = sw.SysConst(Lx=100, Ly=100, param=sw.Taumucs(Tmax=1000, tdump=500, g=-0.002))
    h, diff = sw.run_rayleightaylor(sysc, "GPU", kx=4, ky=5, ϵ=0.01, verbos=False)
    assert len(diff) == 1000 and diff[0] < diff[-1]
    p = onp.Params(Tmax=1000, tdump=500, g=-0.002)
    ref = onp.State(100, 100)
    ref.height[...] = onp.rayleightaylor_ic(100, 100, kx=4, ky=5, eps=0.01)
    oc.equilibrium(ref.feq, ref.height, ref.velx, ref.vely, ref.vsq, p.g)
    dh, _ = oc.time_loop(ref, p, log_dh=True)
    assert np.array_equal(np.asarray(diff), dh)
    assert np.array_equal(h.numpy(), ref.height)  # 0 ulp after 1000 steps


def test_readme_rayleightaylor_config(sw):  # BASELINE.json configs[0]: Lx=Ly=100, g=-0.001, γ=0.0005, Tmax=1000
    sysc = sw.SysConst(Lx=100, Ly=100, param=sw.Taumucs(Tmax=1000, g=-0.001, γ=0.0005))
    h, diff = sw.run_rayleightaylor(sysc, "GPU", h0=1.0, ϵ=0.01, verbos=False)
    p = onp.Params(Tmax=1000, g=-0.001, gamma=0.0005)
    ref = onp.State(100, 100)
    ref.height[...] = onp.rayleightaylor_ic(100, 100, kx=15, ky=18, eps=0.01)
    oc.time_loop(ref, p)
    hg = h.numpy()
    rel = np.abs(hg - ref.height).max() / np.abs(ref.height).max()
    assert rel <= 1e-12, rel            # north_star tolerance
    assert np.array_equal(hg, ref.height)  # and in fact 0 ulp
    m0 = onp.rayleightaylor_ic(100, 100, kx=15, ky=18, eps=0.01).sum()  # (the IC divides by Lx-1: not exactly periodic)
    assert abs(hg.sum() - m0) / m0 < 1e-12


def test_run_dropletrelax(sw):  # test/simulate.jl:44-60 (shortened to 2000 steps for volume/area checks vs oracle)
    sysc = sw.SysConst(Lx=150, Ly=150, param=sw.Taumucs(Tmax=2000, δ=3.0))
    h, area = sw.run_dropletrelax(sysc, "GPU", radius=35, verbos=False)
    p = onp.Params(Tmax=2000, delta=3.0)
    ref = onp.State(150, 150)
    ref.height[...] = onp.singledroplet(150, 150, 35, 1 / 6, (75, 75))
    _, wet = oc.time_loop(ref, p, log_wetted=True, threads=oc.max_threads())
    assert area == wet.tolist() and area[0] < area[-1]
    assert np.array_equal(h.numpy(), ref.height)
    m0 = onp.singledroplet(150, 150, 35, 1 / 6, (75, 75)).sum()
    assert abs(h.numpy().sum() - m0) / m0 < 1e-12


def test_run_dropletforced_moves_in_x_only(sw):  # test/simulate.jl:147-155
    sysc = sw.SysConst(Lx=150, Ly=150, param=sw.Taumucs(Tmax=5000, δ=2.0))
    h, ux, uy = sw.run_dropletforced(sysc, "GPU", radius=35, fx=1e-4, verbos=False)
    hn = h.numpy()
    i, j = np.unravel_index(np.argmax(hn), hn.shape)
    assert i + 1 != 75 and j + 1 == 75
    assert np.all(ux.numpy() < 0.1) and np.all(uy.numpy() < 0.1)


def test_run_dropletpatterned(sw):  # test/simulate.jl:97-112 (radius/volume within 10 % after 10^4 steps)
    sysc = sw.SysConst(Lx=150, Ly=150, param=sw.Taumucs(Tmax=10000, δ=3.0))
    h = sw.run_dropletpatterned(sysc, "GPU", radius=35, θs=np.full((150, 150), 1 / 9), verbos=False).numpy()
    c = sw.cospi
    vol = np.pi / 3 * 35 ** 3 * (2 + c(1 / 6)) * (1 - c(1 / 6)) ** 2
    R1 = np.cbrt((35 ** 3 * (2 + c(1 / 6)) * (1 - c(1 / 6)) ** 2) / ((2 + c(1 / 9)) * (1 - c(1 / 9)) ** 2))
    r1 = np.sin(np.pi / 9) * R1
    droprad = np.count_nonzero(h[74, :] > 0.055) / 2
    droph = h.max()
    vnum = 1 / 6 * np.pi * droph * (3 * droprad ** 2 + droph ** 2)
    assert abs(vol - vnum) < vol / 10 and abs(r1 - droprad) < r1 / 10


# ---- BASELINE sizes: oracle at 1024^2 for a few steps, size-independent properties at 4096^2 / 8192^2 --------


def test_1024_droplet_bitwise_20_steps(sw):  # BASELINE configs[1] shape
    Lx = Ly = 1024
    sysc = sw.SysConst(Lx=Lx, Ly=Ly, param=sw.Taumucs(n=3, m=2, hmin=0.07, δ=1.0))
    st = sw.Sys(sysc, "GPU")
    h0 = onp.singledroplet(Lx, Ly, 256, 1 / 6, (512, 512))
    st.height.set(h0)
    sw.equilibrium(st, sysc)
    sw.fused_steps(st, sysc, 20)
    ref = onp.State(Lx, Ly)
    ref.height[...] = h0
    oc.time_loop(ref, onp.Params(n=3, m=2, hmin=0.07, delta=1.0), nsteps=20, threads=oc.max_threads())
    _compare(st, ref, fields=("height", "velx", "vely", "pressure", "fout"))


@pytest.mark.parametrize("L", [4096, 8192])
def test_large_grid_properties(sw, L):
    """Mass conservation to round-off, translation equivariance (periodic shift of the input == shift of the output,
    bit for bit, which exercises every CTA seam) and flat-film invariance at BASELINE sizes."""
    import torch

    sysc = sw.SysConst(Lx=L, Ly=L, param=sw.Taumucs())
    st = sw.Sys(sysc, "GPU")
    i = torch.arange(L, device="cuda", dtype=torch.float64)
    h0 = 1.0 + 1e-3 * torch.sin(2 * math.pi * i / L)[None, :] * torch.sin(2 * math.pi * i / L)[:, None]
    h0 += 1e-4 * torch.rand((L, L), device="cuda", dtype=torch.float64, generator=torch.Generator("cuda").manual_seed(1))
    st.height.t.copy_(h0)
    m0 = st.height.t.sum().item()
    sw.fused_steps(st, sysc, 10)
    h1 = st.height.t.clone()
    assert abs(h1.sum().item() - m0) / m0 < 1e-13
    # shifted run
    sx, sy = 1237, 411
    st.height.t.copy_(torch.roll(h0, (sy, sx), (0, 1)))
    st.velx.t.zero_(); st.vely.t.zero_()
    sw.fused_steps(st, sysc, 10)
    assert torch.equal(st.height.t, torch.roll(h1, (sy, sx), (0, 1)))
    # flat film
    st.height.t.fill_(1.0); st.velx.t.zero_(); st.vely.t.zero_()
    sw.fused_steps(st, sysc, 5)
    assert torch.all(st.height.t == 1.0)
    del st
    torch.cuda.empty_cache()


def test_wide_slab_32768_tiling_equivariance(sw):
    """The per-GPU slab of the 32768^2 scaling config (32768 x 4096, 1 GiB per plane, > 2^32 bytes into the population
    block): an initial condition made of four identical 8192-column tiles must stay four identical tiles, bit for bit,
    and equal the 8192 x 4096 periodic lattice started from one tile (strip seams fall differently in each copy);
    mass is conserved to round-off."""
    import torch

    Lt, Ly, ntile = 8192, 4096, 4
    g = torch.Generator("cuda").manual_seed(7)
    i = torch.arange(Lt, device="cuda", dtype=torch.float64)
    j = torch.arange(Ly, device="cuda", dtype=torch.float64)
    tile = 1.0 + 1e-2 * torch.sin(2 * math.pi * 3 * j / Ly)[:, None] * torch.sin(2 * math.pi * 5 * i / Lt)[None, :]
    tile += 1e-3 * torch.rand((Ly, Lt), device="cuda", dtype=torch.float64, generator=g)
    small = sw.SysConst(Lx=Lt, Ly=Ly, param=sw.Taumucs())
    st1 = sw.Sys(small, "GPU")
    st1.height.t.copy_(tile)
    sw.fused_steps(st1, small, 6)
    want_h, want_f = st1.height.t.clone(), st1.fout.t[8].clone()
    del st1
    torch.cuda.empty_cache()
    wide = sw.SysConst(Lx=Lt * ntile, Ly=Ly, param=sw.Taumucs())
    st = sw.Sys(wide, "GPU")
    st.height.t.copy_(tile.repeat(1, ntile))
    m0 = st.height.t.sum().item()
    sw.fused_steps(st, wide, 6)
    assert abs(st.height.t.sum().item() - m0) / m0 < 1e-13
    for q in range(ntile):
        assert torch.equal(st.height.t[:, q * Lt:(q + 1) * Lt], want_h), q
        assert torch.equal(st.fout.t[8][:, q * Lt:(q + 1) * Lt], want_f), q  # last population plane: offsets > 2^32 B
    del st
    torch.cuda.empty_cache()


@pytest.mark.parametrize("prm_kw,tau_pops,nsteps", [(dict(g=-0.001), False, 10), (dict(), False, 9),
                                                     (dict(τ=0.8), True, 8), (dict(τ=0.8, n=3, m=2, hmin=0.07), True, 11)])
def test_repeated_loops_replay_a_cuda_graph_bitwise(sw, prm_kw, tau_pops, nsteps, small_lattice_flavour):
    """The chunks of a driver repeat the same call; from the second repetition on the library replays a captured CUDA
    graph of the loop.  Four identical calls (plain launches, capture + launch, two replays) against the oracle, bit for
    bit, with the launch counter advancing by nsteps every time; a changed parameter must not hit the stale graph."""
    from swalbe_b200 import _lib

    lib = _lib.load()
    st, sysc, ref, p = _mk(sw, 70, 52, seed=11, prm_kw=prm_kw, tau_pops=tau_pops)
    for rep in range(4):
        c0 = lib.swalbe_launch_count()
        sw.fused_steps(st, sysc, nsteps)
        # (one kernel per step -- or, for a lattice that fits a thread-block cluster at tau == 1, the persistent kernel
        #  for the first nsteps - 1 steps plus the materialising step)
        assert lib.swalbe_launch_count() - c0 == (2 if (small_lattice_flavour == "cluster" and not tau_pops) else nsteps), rep
        oc.time_loop(ref, p, nsteps=nsteps)
        _compare(st, ref, what=f"repetition {rep}: ")
    # same shapes, different surface tension: a new key -> plain launches again, no stale replay
    kw2 = dict(prm_kw, γ=0.02)
    sysc2 = sw.SysConst(Lx=70, Ly=52, param=sw.Taumucs(**kw2))
    p2 = onp.Params(**{{"τ": "tau", "γ": "gamma"}.get(k, k): v for k, v in kw2.items()})
    for rep in range(3):
        sw.fused_steps(st, sysc2, nsteps)
        oc.time_loop(ref, p2, nsteps=nsteps)
        _compare(st, ref, what=f"second parameter set, repetition {rep}: ")
    # skip_aux / lazy chunks (what time_loop issues between mass prints) replay as well
    for rep in range(3):
        sw.fused_steps(st, sysc2, nsteps, skip_aux=True)
        oc.time_loop(ref, p2, nsteps=nsteps)
        _compare(st, ref, fields=("height", "velx", "vely", "fout", "ftemp"), what=f"skip_aux repetition {rep}: ")


def test_graph_replay_can_be_switched_off(sw, monkeypatch):
    monkeypatch.setenv("SWALBE_GRAPH", "0")
    st, sysc, ref, p = _mk(sw, 64, 40, seed=3, prm_kw=dict(g=-0.001))
    for _ in range(3):
        sw.fused_steps(st, sysc, 8)
        oc.time_loop(ref, p, nsteps=8)
    _compare(st, ref)


def test_skip_aux_keeps_moments_and_populations_current(sw):
    """SWALBE_LOOP_SKIP_AUX: intermediate chunks of a driver skip the materialisation of feq/pressure/h∇p/slip/F; the
    moments and fout == ftemp are still exactly the reference's, and a later default call materialises everything."""
    for kw, pops in ((dict(g=-0.001), False), (dict(τ=0.8), True)):
        st, sysc, ref, p = _mk(sw, 96, 70, seed=31, prm_kw=kw, tau_pops=pops)
        sw.fused_steps(st, sysc, 5, skip_aux=True)
        oc.time_loop(ref, p, nsteps=5)
        _compare(st, ref, fields=("height", "velx", "vely", "fout", "ftemp"))
        assert np.all(st.pressure.numpy() == 0.0) and np.all(st.feq.numpy() == 0.0)  # untouched since Sys()
        sw.fused_steps(st, sysc, 4)
        oc.time_loop(ref, p, nsteps=4)
        _compare(st, ref)


def test_time_loop_driver_matches_oracle_with_chunking(sw):
    """time_loop(sys, state) with tdump chunking (mass read-back between fused chunks) == oracle, all fields."""
    sysc = sw.SysConst(Lx=60, Ly=44, param=sw.Taumucs(Tmax=23, tdump=5, g=-0.001))
    st = sw.Sys(sysc, "GPU")
    rng = np.random.default_rng(9)
    h0 = np.asfortranarray(np.abs(1.0 + 0.2 * rng.standard_normal((60, 44))) + 0.06)
    st.height.set(h0)
    sw.time_loop(sysc, st)
    ref = onp.State(60, 44)
    ref.height[...] = h0
    oc.time_loop(ref, onp.Params(Tmax=23, tdump=5, g=-0.001))
    _compare(st, ref)


def test_special_values_propagate_like_the_reference(sw):
    """Zeros, negatives, denormals, huge values, Inf and NaN in the input fields: every IEEE corner (division by zero,
    0*Inf, overflow in the power chain, the division slow paths) must come out exactly as in the oracle (NaN == NaN)."""
    Lx, Ly = 48, 20
    rng = np.random.default_rng(77)
    h0 = np.asfortranarray(np.abs(1.0 + 0.2 * rng.standard_normal((Lx, Ly))) + 0.06)
    specials = [0.0, -0.0, -0.05, -1.0, 1e-310, 5e-324, 1e-160, 1e160, 1e300, np.inf, -np.inf, np.nan, 0.05, -0.0500000001]
    for k, v in enumerate(specials):
        h0[3 * k % Lx, (5 * k + 2) % Ly] = v
    u0 = np.asfortranarray(0.01 * rng.standard_normal((Lx, Ly)))
    u0[7, 3], u0[9, 9], u0[11, 1] = np.nan, np.inf, 1e308
    for kw, pops in ((dict(g=0.01), False), (dict(τ=0.75, n=3, m=2), True)):
        okw = {{"τ": "tau"}.get(k, k): v for k, v in kw.items()}
        sysc = sw.SysConst(Lx=Lx, Ly=Ly, param=sw.Taumucs(**kw))
        st = sw.Sys(sysc, "GPU")
        ref = onp.State(Lx, Ly)
        st.height.set(h0); st.velx.set(u0)
        ref.height[...] = h0; ref.velx[...] = u0
        if pops:
            f0 = np.asfortranarray(0.1 + 0.01 * rng.random((Lx, Ly, 9)))
            st.ftemp.set(f0); ref.ftemp[...] = f0
        with np.errstate(all="ignore"):
            sw.fused_steps(st, sysc, 2)
            oc.time_loop(ref, onp.Params(**okw), nsteps=2)
        if pops:  # tau != 1: the kernel evaluates omega*ftemp + feq/tau literally -> identical, NaN for NaN
            _compare(st, ref, what=f"{kw}:")
            continue
        # tau == 1: the kernel uses feq for 0*ftemp + 1*feq.  That is exact for finite ftemp; where ftemp is already
        # Inf/NaN the reference's 0*Inf yields NaN while feq may be +-Inf.  So: identical finite/non-finite pattern,
        # identical finite values, and the moments (what the next step reads) identical including NaN.
        _compare(st, ref, fields=("height", "velx", "vely", "pressure", "hgradpx", "hgradpy", "slipx", "slipy", "Fx", "Fy",
                                  "feq", "vsq"), what=f"{kw}:")
        for name in ("fout", "ftemp"):
            got, want = getattr(st, name).numpy(), getattr(ref, name)
            fin = np.isfinite(want)
            assert np.array_equal(np.isfinite(got), fin), name
            assert np.array_equal(got[fin], want[fin]), name
            assert np.isfinite(want).sum() > 0.3 * want.size  # a good part of the lattice is still healthy


def test_spinodal_dewetting_long_run_bitwise(sw, small_lattice_flavour):
    """A physically UNSTABLE configuration (thin random film, disjoining pressure n=3, m=2: spinodal dewetting, C3-style)
    amplifies any rounding difference exponentially: the perturbation first decays, then grows 6x within 3000 steps.
    The whole max-min history and the final fields must still equal the oracle bit for bit."""
    Lx, Ly, nsteps = 192, 160, 3000
    rng = np.random.default_rng(20261017)
    h0 = np.asfortranarray(0.2 * (1.0 + 0.01 * rng.standard_normal((Lx, Ly))))
    sysc = sw.SysConst(Lx=Lx, Ly=Ly, param=sw.Taumucs(n=3, m=2, hmin=0.07, γ=0.02))
    st = sw.Sys(sysc, "GPU")
    st.height.set(h0)
    mn, mx, _ = sw.fused_steps(st, sysc, nsteps, θ=1 / 9, log_minmax=True)
    ref = onp.State(Lx, Ly)
    ref.height[...] = h0
    p = onp.Params(n=3, m=2, hmin=0.07, gamma=0.02)
    dh, _ = oc.time_loop(ref, p, nsteps=nsteps, cospi_theta=onp.cospi(1 / 9), threads=oc.max_threads(), log_dh=True)
    assert dh[-1] > 3 * dh.min() and dh.argmin() > 100  # decayed first, then the instability took over
    assert np.array_equal((mx - mn).cpu().numpy(), dh)
    _compare(st, ref, fields=("height", "velx", "vely", "pressure", "fout"))


@pytest.mark.parametrize("Lx,Ly", [(256, 40), (600, 50), (1030, 33), (2048, 16)])
def test_bulk_copy_prefetch_variant_bitwise(sw, monkeypatch, Lx, Ly):
    """The cp.async.bulk (TMA unit) row-prefetch flavour of the lean kernel, forced on for small lattices: strips that
    cross the periodic x boundary are fed by two bulk copies; results must not move by a bit."""
    monkeypatch.setenv("SWALBE_BULK", "2")
    for nt in (0, 192, 256):
        if nt:
            monkeypatch.setenv("SWALBE_NT", str(nt))
        st, sysc, ref, p = _mk(sw, Lx, Ly, seed=Lx + Ly, prm_kw=dict(g=-0.001, n=3, m=2, hmin=0.07))
        sw.fused_steps(st, sysc, 6)
        oc.time_loop(ref, p, nsteps=6)
        _compare(st, ref, what=f"NT={nt}:")


def test_large_odd_lattice_fused_equals_operator_kernels(sw):
    """Odd extents at multi-strip / multi-wave scale (2049 x 1025: unaligned rows, LDGSTS prefetch, a last strip that
    wraps): the fused loop against the seven per-operator kernels, both on the GPU, every field bit for bit."""
    import torch

    Lx, Ly = 2049, 1025
    sysc = sw.SysConst(Lx=Lx, Ly=Ly, param=sw.Taumucs(g=-0.0005, n=3, m=2, hmin=0.07))
    st, st2 = sw.Sys(sysc, "GPU"), sw.Sys(sysc, "GPU")
    gen = torch.Generator("cuda").manual_seed(11)
    h0 = 1.0 + 0.05 * torch.randn((Ly, Lx), device="cuda", dtype=torch.float64, generator=gen)
    for s in (st, st2):
        s.height.t.copy_(h0)
    sw.fused_steps(st, sysc, 5)
    for _ in range(5):
        sw.filmpressure(st2, sysc); sw.hgradp(st2); sw.slippage(st2, sysc); sw.update(st2)
        sw.equilibrium(st2, sysc); sw.BGKandStream(st2, sysc); sw.moments(st2)
    for name in STATE_FIELDS:
        assert torch.equal(getattr(st, name).t, getattr(st2, name).t), name


def test_non_default_stream_is_honoured(sw):
    """Every entry point runs on the caller's stream: work queued on a side stream must be ordered with that stream's
    other work (a fill that precedes it, a copy that follows it) without any device-wide synchronisation."""
    import torch

    st, sysc, ref, p = _mk(sw, 200, 64, seed=4, prm_kw=dict(g=-0.001))
    side = torch.cuda.Stream()
    h0 = st.height.numpy()
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        st.height.t.mul_(1.0)            # same-stream producer
        sw.fused_steps(st, sysc, 4)      # uses torch.cuda.current_stream() == side
        out = st.height.t.clone()        # same-stream consumer
    side.synchronize()
    oc.time_loop(ref, p, nsteps=4)
    assert np.array_equal(np.asfortranarray(out.cpu().numpy().T), ref.height)
    assert not np.array_equal(h0, ref.height)
