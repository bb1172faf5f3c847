"""swalbe_time_loop_host: the time loop of a job whose initial height comes from, and whose final height goes back to,
host memory -- row bands, skewed sweeps behind the upload front and ahead of the download (csrc/sweep.h).  The contract:
state and downloaded plane equal, bit for bit, what copy + swalbe_time_loop + copy leave; that in turn equals the oracle
(tests/test_gpu_step.py), which the small cases here check directly as well."""
import numpy as np
import pytest

from oracle import oracle_c as oc
from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu

FIELDS = ("height", "velx", "vely", "fout", "ftemp", "feq", "vsq", "pressure", "hgradpx", "hgradpy", "slipx", "slipy",
          "Fx", "Fy")


@pytest.fixture(scope="module")
def sw():
    import swalbe_b200

    return swalbe_b200


def _pinned(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a.transpose())).pin_memory()


def _inputs(Lx, Ly, seed):
    rng = np.random.default_rng(seed)
    h0 = np.asfortranarray(np.abs(1.0 + 0.2 * rng.standard_normal((Lx, Ly))) + 0.06)
    ux0 = np.asfortranarray(0.01 * rng.standard_normal((Lx, Ly)))
    uy0 = np.asfortranarray(0.01 * rng.standard_normal((Lx, Ly)))
    return h0, ux0, uy0


def _bands(monkeypatch, rows, kmax=0):
    monkeypatch.setenv("SWALBE_HOST_MIN_SITES", "1")
    monkeypatch.setenv("SWALBE_BAND_ROWS", str(rows))
    if kmax:
        monkeypatch.setenv("SWALBE_HOST_KMAX", str(kmax))


@pytest.mark.parametrize("Lx,Ly,band", [(150, 200, 50), (257, 301, 64), (64, 1000, 100), (600, 97, 40)])
@pytest.mark.parametrize("nsteps", [1, 2, 5, 9, 14, 31])
def test_host_loop_equals_oracle(sw, monkeypatch, Lx, Ly, band, nsteps):
    import torch

    _bands(monkeypatch, band)
    kw = dict(g=-0.001, γ=0.0005)
    h0, ux0, uy0 = _inputs(Lx, Ly, Lx + Ly + nsteps)
    sysc = sw.SysConst(Lx=Lx, Ly=Ly, param=sw.Taumucs(**kw))
    st = sw.Sys(sysc, "GPU")
    st.velx.set(ux0); st.vely.set(uy0)
    st.height.set(3.0)  # the job must start from the host plane, not from this
    hin, hout = _pinned(h0), torch.full((Ly, Lx), float("nan"), dtype=torch.float64).pin_memory()
    for lazy in (False, True):
        st.velx.set(ux0); st.vely.set(uy0); st.height.set(3.0)
        hout.fill_(float("nan"))
        sw.fused_steps(st, sysc, nsteps, host_in=hin, host_out=hout, lazy_populations=lazy)
        torch.cuda.synchronize()
        ref = onp.State(Lx, Ly)
        ref.height[...] = h0; ref.velx[...] = ux0; ref.vely[...] = uy0
        oc.time_loop(ref, onp.Params(g=-0.001, gamma=0.0005), nsteps=nsteps)
        for name in FIELDS:
            got, want = getattr(st, name).numpy(), getattr(ref, name)
            assert np.array_equal(got, want), f"{name} (lazy={lazy}): {np.count_nonzero(got != want)} sites differ"
        assert np.array_equal(hout.numpy().transpose(), ref.height)


def _plain_vs_host(sw, monkeypatch, sysc, h0, ux0, uy0, nsteps, kind="simple", only=("in", "out", "both"), **kw):
    """copy + swalbe_time_loop + copy against swalbe_time_loop_host on the same inputs"""
    import torch

    def fresh():
        st = sw.Sys(sysc, "GPU", kind=kind)
        st.velx.set(ux0); st.vely.set(uy0)
        return st

    a = fresh()
    a.height.set(h0)
    logs_a = sw.fused_steps(a, sysc, nsteps, **kw)
    want_out = a.height.numpy()
    for mode in only:
        b = fresh()
        hin = _pinned(h0) if mode in ("in", "both") else None
        hout = torch.full((sysc.Ly, sysc.Lx), float("nan"), dtype=torch.float64).pin_memory() if mode in ("out", "both") else None
        b.height.set(h0 if hin is None else 9.0)
        logs_b = sw.fused_steps(b, sysc, nsteps, host_in=hin, host_out=hout, **kw)
        torch.cuda.synchronize()
        fields = FIELDS + (("kbtx", "kbty") if kind == "thermal" else ())
        for name in fields:
            got, want = getattr(b, name).numpy(), getattr(a, name).numpy()
            assert np.array_equal(got, want), f"{mode}: {name}: {np.count_nonzero(got != want)} sites differ"
        if hout is not None:
            assert np.array_equal(hout.numpy().transpose(), want_out), mode
        for la, lb in zip(logs_a, logs_b):
            if la is not None:
                assert torch.equal(la, lb), mode


@pytest.mark.parametrize("nsteps", [3, 10, 27])
def test_host_loop_options_equal_plain_loop(sw, monkeypatch, nsteps):
    """contact-angle field, slip variant, inclination, per-step logs, gravity-free lean kernels, thermal noise"""
    _bands(monkeypatch, 60)
    Lx, Ly = 192, 330
    h0, ux0, uy0 = _inputs(Lx, Ly, nsteps)
    rng = np.random.default_rng(5)
    θ = sw.Field(Lx, Ly).set(1 / 9 + 0.02 * rng.standard_normal((Lx, Ly)))
    sysc = sw.SysConst(Lx=Lx, Ly=Ly, param=sw.Taumucs(n=3, m=2, hmin=0.07, γ=0.01))
    _plain_vs_host(sw, monkeypatch, sysc, h0, ux0, uy0, nsteps)
    _plain_vs_host(sw, monkeypatch, sysc, h0, ux0, uy0, nsteps, θ=θ, only=("both",))
    _plain_vs_host(sw, monkeypatch, sysc, h0, ux0, uy0, nsteps, slip_variant=2, incl=((1e-4, -2e-4), 0.7), only=("both",))
    _plain_vs_host(sw, monkeypatch, sysc, h0, ux0, uy0, nsteps, log_minmax=True, log_wetted=True, hthresh=1.0, only=("both",))
    _plain_vs_host(sw, monkeypatch, sysc, h0, ux0, uy0, nsteps, skip_aux=True, lazy_populations=True, only=("both",))
    sysk = sw.SysConst(Lx=Lx, Ly=Ly, param=sw.Taumucs(n=3, m=2, hmin=0.07, γ=0.01, kbt=1e-6))
    _plain_vs_host(sw, monkeypatch, sysk, h0, ux0, uy0, nsteps, kind="thermal", thermal_seed=77, step0=5, only=("both",))


def test_host_loop_general_tau_and_small_lattices_take_the_plain_path(sw, monkeypatch):
    """tau != 1 and lattices below the size bound: the copies bracket the ordinary loop -- same contract"""
    Lx, Ly = 100, 96
    h0, ux0, uy0 = _inputs(Lx, Ly, 3)
    _plain_vs_host(sw, monkeypatch, sw.SysConst(Lx=Lx, Ly=Ly, param=sw.Taumucs()), h0, ux0, uy0, 7)
    _bands(monkeypatch, 30)
    _plain_vs_host(sw, monkeypatch, sw.SysConst(Lx=Lx, Ly=Ly, param=sw.Taumucs(τ=0.9)), h0, ux0, uy0, 7)


def test_host_loop_default_bands_large_lattice(sw):
    """default configuration (bands of 8 Mi sites, 12-step sweeps) on a lattice big enough to take it, TMA-prefetch
    kernels included: 4096 x 4096, 20 and 33 steps, against the plain loop"""
    Lx = Ly = 4096
    rng = np.random.default_rng(1)
    h0 = np.asfortranarray(1.0 + 0.01 * rng.standard_normal((Lx, Ly)))
    z = np.zeros((Lx, Ly), order="F")
    sysc = sw.SysConst(Lx=Lx, Ly=Ly, param=sw.Taumucs())
    for nsteps in (20, 33):
        _plain_vs_host(sw, None, sysc, h0, z, z, nsteps, only=("both",))
        _plain_vs_host(sw, None, sysc, h0, z, z, nsteps, only=("both",), lazy_populations=True)


@pytest.mark.parametrize("nsteps,first,every", [(23, 2, 4), (9, 0, 1), (30, 9, 10), (5, 7, 3)])
def test_mass_log_inside_the_loop(sw, monkeypatch, nsteps, first, every):
    """logs.hsum: sum(height) BEFORE the selected steps, produced inside the loop -- by the plain loop (device slots), by the
    host loop's sweeps (pinned host slots, rows summed band by band behind the launches that produce them): the same
    bits either way (fixed row-wise order), and the oracle's masses to round-off."""
    import torch

    _bands(monkeypatch, 50)
    Lx, Ly = 150, 200
    h0, ux0, uy0 = _inputs(Lx, Ly, nsteps)
    sysc = sw.SysConst(Lx=Lx, Ly=Ly, param=sw.Taumucs(g=-0.001, γ=0.0005))
    ref = onp.State(Lx, Ly)
    ref.height[...] = h0; ref.velx[...] = ux0; ref.vely[...] = uy0
    want = []
    for s in range(nsteps):
        if s >= first and (s - first) % every == 0:
            want.append(ref.height.sum())
        oc.time_loop(ref, onp.Params(g=-0.001, gamma=0.0005), nsteps=1)
    nslots = max(1, len(want))
    a = sw.Sys(sysc, "GPU")
    a.height.set(h0); a.velx.set(ux0); a.vely.set(uy0)
    dev = torch.full((nslots,), float("nan"), dtype=torch.float64, device="cuda")
    sw.fused_steps(a, sysc, nsteps, mass_log=(first, every, dev))
    b = sw.Sys(sysc, "GPU")
    b.velx.set(ux0); b.vely.set(uy0)
    host = torch.full((nslots,), float("nan"), dtype=torch.float64).pin_memory()
    hout = torch.empty(Lx * Ly, dtype=torch.float64).pin_memory()
    sw.fused_steps(b, sysc, nsteps, mass_log=(first, every, host), host_in=_pinned(h0), host_out=hout, lazy_populations=True)
    torch.cuda.synchronize()
    got_a, got_b = dev.cpu().numpy()[:len(want)], host.numpy()[:len(want)]
    assert np.array_equal(got_a, got_b)
    assert np.allclose(got_a, np.array(want), rtol=1e-13, atol=0)
    for name in FIELDS:
        assert np.array_equal(getattr(a, name).numpy(), getattr(ref, name)), name
        assert np.array_equal(getattr(b, name).numpy(), getattr(ref, name)), name
    if len(want) < nslots:
        assert np.isnan(dev.cpu().numpy()[len(want):]).all()


@pytest.mark.parametrize("one_call", [False, True])
def test_mass_prints_are_asynchronous_but_complete(sw, capsys, monkeypatch, one_call):
    """time_loop(verbose=True): one line per dump step (t % tdump == 0), in order, with the mass of the state BEFORE that
    step (src/simulate.jl:8-14); the read-back does not synchronise the loop, the lines are all there on return."""
    Lx, Ly = 96, 80
    h0, _, _ = _inputs(Lx, Ly, 4)
    sysc = sw.SysConst(Lx=Lx, Ly=Ly, param=sw.Taumucs(Tmax=35, tdump=5))
    st = sw.Sys(sysc, "GPU")
    st.height.set(h0)
    if one_call:  # the path of large lattices: the whole loop is one library call with the in-loop mass log
        monkeypatch.setattr(sw, "_ONE_CALL_SITES", 1)
    sw.time_loop(sysc, st, verbose=True)
    other = sw.Sys(sysc, "GPU")
    other.height.set(h0)
    monkeypatch.setattr(sw, "_ONE_CALL_SITES", 1 if not one_call else 1 << 40)
    sw.time_loop(sysc, other)
    for name in FIELDS:
        assert np.array_equal(getattr(st, name).numpy(), getattr(other, name).numpy()), name
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("Time step")]
    assert [int(ln.split()[2]) for ln in lines] == [5, 10, 15, 20, 25, 30, 35]
    for ln in lines:
        assert abs(float(ln.split()[-1]) - h0.sum()) < 2e-3, ln


def test_time_loop_and_run_host_drivers(sw, monkeypatch):
    """time_loop(host_in=, host_out=) over several tdump chunks and run_host == the plain driver"""
    import torch

    _bands(monkeypatch, 40)
    monkeypatch.setattr(sw, "_ONE_CALL_SITES", 1)
    Lx, Ly = 128, 250
    h0, _, _ = _inputs(Lx, Ly, 8)
    sysc = sw.SysConst(Lx=Lx, Ly=Ly, param=sw.Taumucs(Tmax=50, tdump=20))
    a = sw.Sys(sysc, "GPU")
    a.height.set(h0)
    sw.equilibrium(a, sysc)
    sw.time_loop(sysc, a)
    b, out = sw.run_host(sysc, _pinned(h0))
    for name in FIELDS:
        assert np.array_equal(getattr(a, name).numpy(), getattr(b, name).numpy()), name
    assert np.array_equal(out.numpy().reshape(Ly, Lx).transpose(), a.height.numpy())
    c = sw.Sys(sysc, "GPU")
    hout = torch.empty(Lx * Ly, dtype=torch.float64).pin_memory()
    sw.time_loop(sysc, c, 1 / 9, host_in=_pinned(h0), host_out=hout)
    torch.cuda.synchronize()
    d = sw.Sys(sysc, "GPU")
    d.height.set(h0)
    sw.time_loop(sysc, d, 1 / 9)
    assert np.array_equal(hout.numpy().reshape(Ly, Lx).transpose(), d.height.numpy())
    assert np.array_equal(c.pressure.numpy(), d.pressure.numpy())
