"""CPU tests of the host mirror's driver logic (no GPU, no kernels): the fused time loop must be chunked so that the
mass print of the reference (`t % tdump == 0`, src/simulate.jl:8-14) happens at the same steps, the chunks add up to
Tmax, and the four time_loop call shapes dispatch like the reference's methods (src/simulate.jl:6-96)."""
import math

import numpy as np
import pytest

import swalbe_b200 as sw


class _FakeTensor:
    def __init__(self, a):
        self.a = np.asarray(a, dtype=np.float64)

    def __sub__(self, o):
        return _FakeTensor(self.a - o.a)

    def cpu(self):
        return self

    def tolist(self):
        return self.a.tolist()


@pytest.fixture()
def fake(monkeypatch):
    calls = {"steps": [], "stats_at": [], "kw": []}
    state = {"t": 1}

    def fused_steps(st, sysc, n, **kw):
        calls["steps"].append(n)
        calls["kw"].append(kw)
        state["t"] += n
        mn = _FakeTensor(np.zeros(n)) if kw.get("log_minmax") else None
        mx = _FakeTensor(np.arange(n) + 1.0) if kw.get("log_minmax") else None
        wet = _FakeTensor(np.full(n, 7)) if kw.get("log_wetted") else None
        return mn, mx, wet

    def field_stats(f, thresh=0.055):
        calls["stats_at"].append(state["t"])
        return 0.0, 1.0, 625.0, 3

    class Dumps(sw._MassDumps):  # the device-side reduction and its asynchronous read-back, done on the spot
        def push(self, t, height):
            self._emit(t, field_stats(height)[2])

        def flush(self, block):
            pass

    monkeypatch.setattr(sw, "fused_steps", fused_steps)
    monkeypatch.setattr(sw, "field_stats", field_stats)
    monkeypatch.setattr(sw, "_MassDumps", Dumps)
    return calls


class _State(sw.CuState):
    def __init__(self):  # no device allocation
        self.height = None


@pytest.mark.parametrize("Tmax,tdump", [(1000, 100), (200, 100), (10, 3), (7, 10), (5, 1), (1000, None)])
def test_time_loop_chunks_and_mass_prints(fake, capsys, Tmax, tdump):
    sysc = sw.SysConst(Lx=5, Ly=5, param=sw.Taumucs(Tmax=Tmax, tdump=tdump))
    td = sysc.param.tdump
    sw.time_loop(sysc, _State(), verbose=True)
    assert sum(fake["steps"]) == Tmax
    expected = [t for t in range(1, Tmax + 1) if t % td == 0]  # the reference prints BEFORE the update of step t
    assert fake["stats_at"] == expected
    out = capsys.readouterr().out.strip().splitlines()
    assert out == [f"Time step {t} mass is 625.0" for t in expected]


def test_time_loop_method_shapes(fake):
    sysc = sw.SysConst(Lx=5, Ly=5, param=sw.Taumucs(Tmax=12, tdump=5))
    st = _State()
    assert sw.time_loop(sysc, st) is st                                   # src/simulate.jl:6-25
    assert all(k.get("θ") is None for k in fake["kw"])
    fake["kw"].clear()
    sw.time_loop(sysc, st, 1 / 9)                                          # :26-45 (θ scalar)
    assert all(k["θ"] == 1 / 9 for k in fake["kw"])
    dh = []
    sw.time_loop(sysc, st, dh)                                             # :47-67 (Δh log, one entry per step)
    assert len(dh) == 12
    area = []
    ret = sw.time_loop(sysc, st, sw.wetted, area)                          # :69-96 (callback slot)
    assert ret == (st, area) and area == [7] * 12
    force = [1e-4, 0.0]
    fake["kw"].clear()
    sw.time_loop(sysc, st, sw.inclination, force)
    alpha, factor = fake["kw"][0]["incl"]
    assert alpha is force and factor == 0.5 + 0.5 * math.tanh(1000.0)     # defaults t=1000, tstart=0, tsmooth=1 (forcing.jl:363)
    with pytest.raises(sw.SwalbeError):
        sw.time_loop(sysc, st, lambda m, s: None, [])
    with pytest.raises(TypeError):
        sw.time_loop(sysc, st, 1, 2, 3)


def test_julia_name_table_and_aliases():
    for name in ("equilibrium!", "BGKandStream!", "moments!", "filmpressure!", "h∇p!", "slippage!", "update!", "time_loop",
                 "run_flat", "run_random", "run_rayleightaylor", "run_dropletrelax", "run_dropletpatterned",
                 "run_dropletforced", "∇f!", "∇²f!", "thermal!", "inclination!", "slippage2!", "slippage_ring_riv!"):
        assert callable(sw.JULIA_NAMES[name]), name
    assert sw.Sys_const is sw.SysConst and sw.Swalbe_state is sw.CuState
    assert issubclass(sw.CuState_thermal, sw.CuState)


def test_cospi_matches_reference_points():
    assert sw.cospi(0) == 1.0 and sw.cospi(0.5) == 0.0 and sw.cospi(1) == -1.0 and sw.cospi(1.5) == 0.0
    assert sw.cospi(-1 / 9) == sw.cospi(1 / 9) and sw.cospi(2.25) == sw.cospi(0.25) == sw.cospi(-1.75)
    assert abs(sw.cospi(1 / 6) - math.sqrt(3) / 2) < 2e-16


def test_wetted_call_shapes(fake):  # src/measures.jl:6-17; known answers of test/measures.jl:2-24 need the device
    area = []
    sw.wetted(area, _State())
    sw.wetted(area, _State(), hthresh=0.1)
    assert area == [3, 3]
    area_size, maxheight = [1.0] * 5, [0.0] * 5
    sw.wetted(area_size, maxheight, sw.Field.__new__(sw.Field), 2)
    assert area_size == [1.0, 3, 1.0, 1.0, 1.0] and maxheight == [0.0, 1.0, 0.0, 0.0, 0.0]
    with pytest.raises(TypeError):
        sw.wetted(area_size, maxheight, 2)



def test_cluster_kernel_division_free_row_index():
    """csrc/cluster.cuh computes the row of a slab-local site index as int((idx + 0.5f) * (1.0f / Lx)) inside the step
    loop; csrc/cluster.cu admits Lx <= 1024 and idx < 65536.  Exact over that whole range."""
    idx = np.arange(0, 65536, dtype=np.int64)
    for Lx in range(1, 1025):
        inv = np.float32(1.0) / np.float32(Lx)
        got = ((idx.astype(np.float32) + np.float32(0.5)) * inv).astype(np.int64)
        assert np.array_equal(got, idx // Lx), Lx


def test_power_broad_methods():
    """power_broad(arg::Float64 | Float32 | Int, n)   src/pressure.jl:363-385 and its doctest (:344-356)"""
    import numpy as np

    assert sw.power_broad(3, 3) == 27 and isinstance(sw.power_broad(3, 3), int)
    assert [sw.power_broad(x, 2) for x in (2.0, 5.0, 6.0)] == [4.0, 25.0, 36.0]
    x32 = np.float32(0.1)
    got = sw.power_broad(x32, 9)
    want = np.float32(1.0)
    for _ in range(9):
        want = np.float32(want * x32)  # every product rounded to single precision, as `temp = 1.0f0; temp *= arg` does
    assert isinstance(got, np.float32) and got == want and float(got) != sw.power_broad(0.1, 9)
    assert sw.fast_93(0.5) == (0.5 ** 3) ** 3 - 0.5 ** 3 and sw.fast_32(0.5) == 0.5 ** 3 - 0.5 ** 2


def test_host_plane_layout_is_checked():
    """host_in / host_out of the host loop: only memory orders that match state.height (i fastest) are accepted"""
    import numpy as np
    import torch

    class S:
        Lx, Ly = 6, 4

    f = np.zeros((6, 4), order="F")
    assert sw._host_ptr(f, S).value == f.ctypes.data
    c = np.zeros((4, 6))
    assert sw._host_ptr(c, S).value == c.ctypes.data
    assert sw._host_ptr(np.zeros(24), S) is not None and sw._host_ptr(None, S) is None
    t = torch.zeros(4, 6, dtype=torch.float64)
    assert sw._host_ptr(t, S).value == t.data_ptr() and sw._host_ptr(torch.zeros(24, dtype=torch.float64), S) is not None
    for bad in (np.zeros((6, 4)), np.zeros((4, 6), order="F"), np.zeros((6, 4), dtype=np.float32, order="F"), np.zeros(23),
                torch.zeros(6, 4, dtype=torch.float64), torch.zeros(4, 6), torch.zeros(4, 6, dtype=torch.float64).t()):
        with pytest.raises(ValueError):
            sw._host_ptr(bad, S)


def _schedule(Lx, Ly, nsteps, has_in, has_out, band, kmax):
    import ctypes as C

    from swalbe_b200 import _lib

    n = C.c_int(0)
    _lib.call("swalbe_selftest_host_loop_schedule", Lx, Ly, nsteps, int(has_in), int(has_out), band, kmax, 1, None, 0, C.byref(n))
    buf = (C.c_int * (7 * max(1, n.value)))()
    _lib.call("swalbe_selftest_host_loop_schedule", Lx, Ly, nsteps, int(has_in), int(has_out), band, kmax, 1, buf, n.value, C.byref(n))
    return [tuple(buf[7 * q:7 * q + 7]) for q in range(n.value)]


def _symbolic_run(ops, Ly, nsteps, has_in, has_out, order, early_uploads):
    """Execute a host-loop schedule on row LABELS instead of numbers: each of the two moment buffers holds, per row, the
    index of the state it contains (-1: garbage).  A launch of step s over [j0, j1) must find state s in rows
    [j0 - 3, j1 + 3) (periodic) of its source buffer and leaves state s + 1 in [j0, j1) of the other one; a download must
    find the final state.  Any read of a row that has not been produced yet, or that a later step already overwrote,
    fails -- for whichever order of the operations the streams allow."""
    import numpy as np

    buf = [np.full(Ly, -1), np.full(Ly, -1)]  # [A (the state's planes), B (scratch)]
    src0 = 0 if nsteps % 2 == 0 else 1
    if not has_in:
        buf[src0][:] = 0
    ups = [o for o in ops if o[0] == 0]
    if early_uploads:
        for _, _, j0, j1, *_ in ups:
            buf[src0][j0:j1] = 0
    got = np.full(Ly, -1)
    for kind, s, j0, j1, band, seam, stage in order:
        if kind == 0:
            if not early_uploads:
                buf[src0][j0:j1] = 0
        elif kind == 2:
            assert (buf[0][j0:j1] == nsteps).all(), ("download of rows that are not final", j0, j1)
            got[j0:j1] = nsteps
        elif kind == 3:  # mass log: the rows of state s, summed behind the launch (or upload wait) that made them final
            holder = buf[0] if (nsteps - s) % 2 == 0 else buf[1]  # state s lives in the buffer step s reads
            assert (holder[j0:j1] == s).all(), ("row sums of state", s, "rows", j0, j1, "find", sorted(set(holder[j0:j1].tolist())))
        elif j1 > j0:
            reads_A = (nsteps - s) % 2 == 0
            src, dst = (buf[0], buf[1]) if reads_A else (buf[1], buf[0])
            rows = np.arange(j0 - 3, j1 + 3) % Ly
            assert (src[rows] == s).all(), ("step", s, "rows", j0, j1, "stage", stage, "finds", sorted(set(src[rows].tolist())))
            dst[j0:j1] = s + 1
    assert (buf[0] == nsteps).all()
    if has_out:
        assert (got == nsteps).all()


def test_host_loop_schedule_symbolic():
    """csrc/sweep.h over a few hundred lattice / band / sweep-length / step-count combinations, executed symbolically in the
    listed order (uploads as late as possible) and in the step-major order the two compute streams allow (uploads first)"""
    import random

    rnd = random.Random(7)
    nswept = 0
    for trial in range(400):
        Ly = rnd.choice([24, 61, 97, 128, 200, 301, 1000, 4096])
        band = rnd.choice([0, 8, 16, 31, 50, 64, 100, 512, 1024])
        kmax = rnd.choice([0, 1, 2, 3, 5, 12, 32])
        nsteps = rnd.choice([1, 2, 3, 4, 7, 12, 13, 24, 25, 40])
        has_in, has_out = rnd.choice([(True, True), (True, False), (False, True)])
        ops = _schedule(64, Ly, nsteps, has_in, has_out, band, kmax)
        steps = [o for o in ops if o[0] == 1]
        assert sorted({o[1] for o in steps}) == list(range(nsteps))
        nswept += any(o[6] >= 0 for o in steps)
        _symbolic_run(ops, Ly, nsteps, has_in, has_out, ops, early_uploads=False)
        # step-major inside every sweep (what two streams make possible), uploads first, downloads last
        out, run = [], []

        def flush():
            if run:
                k0 = min(o[1] for o in run)
                out.extend(sorted(run, key=lambda o: (o[1] - k0, o[6])))
                run.clear()

        for o in steps:
            if o[6] >= 0:
                if run and o[6] == 0 and run[-1][6] != 0:
                    flush()
                run.append(o)
            else:
                flush()
                out.append(o)
        flush()
        order = [o for o in ops if o[0] == 0] + out + [o for o in ops if o[0] == 2]
        _symbolic_run(ops, Ly, nsteps, has_in, has_out, order, early_uploads=True)
        # random interleavings of the two lanes (even / odd band stages, each in issue order) that respect the one
        # cross-lane dependency the library enforces with an event: (stage b, k) after (stage b-1, k-1)
        band_rows = {o[4]: (o[2], o[3]) for o in ops if o[0] == 0}
        for _ in range(3):
            out, run = [], []

            def interleave():
                if not run:
                    return
                k0 = min(o[1] for o in run)
                lanes = [[], []]
                for o in run:  # every launch is followed, on its own lane, by the row sums of the rows it produced; the first
                    ln = lanes[o[6] % 2]  # launch of a band stage that waits for an upload is preceded by those of the band
                    if o[4] >= 0:
                        ln.append((3, 0, band_rows[o[4]][0], band_rows[o[4]][1], -1, 0, o[6]))
                    ln.append(o)
                    if o[1] + 1 < nsteps:
                        ln.append((3, o[1] + 1, o[2], o[3], -1, 0, o[6]))
                done = set()
                while lanes[0] or lanes[1]:
                    ready = [ln for ln in lanes if ln and (ln[0][0] == 3 or ln[0][6] == 0 or ln[0][1] == k0 or
                                                          (ln[0][6] - 1, ln[0][1] - 1) in done)]
                    assert ready, "the two lanes deadlock"
                    ln = rnd.choice(ready)
                    o = ln.pop(0)
                    if o[0] == 1:
                        done.add((o[6], o[1]))
                    out.append(o)
                run.clear()

            for o in steps:
                if o[6] >= 0:
                    if run and o[6] == 0 and run[-1][6] != 0:
                        interleave()
                    run.append(o)
                else:
                    interleave()
                    out.append(o)
            interleave()
            order = [o for o in ops if o[0] == 0] + out + [o for o in ops if o[0] == 2]
            assert sorted(o for o in order if o[0] != 3) == sorted(ops)
            _symbolic_run(ops, Ly, nsteps, has_in, has_out, order, early_uploads=True)
    assert nswept > 150  # (most of the combinations really are banded sweeps, not the plain fallback)
