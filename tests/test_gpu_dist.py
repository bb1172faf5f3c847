"""The slab runtime (swalbe_dist_*): ghost-row kernel path + halo exchange.

nranks = 1 runs everywhere (self-exchange through device copies) and must equal the oracle bit for bit; the 2-rank
NCCL test runs only where two GPUs are visible (gpurun --gpus 2) and launches one process per GPU with torchrun."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle_c as oc
from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("Lx,Ly", [(64, 12), (300, 50), (25, 6)])
@pytest.mark.parametrize("tau", [1.0, 0.8])
def test_single_rank_slab_matches_oracle(Lx, Ly, tau):
    import swalbe_b200 as sw
    from swalbe_b200.dist import DistSim

    rng = np.random.default_rng(Lx + Ly)
    h0 = np.asfortranarray(np.abs(1.0 + 0.2 * rng.standard_normal((Lx, Ly))) + 0.06)
    u0 = np.asfortranarray(0.01 * rng.standard_normal((Lx, Ly)))
    f0 = np.asfortranarray(0.1 + 0.01 * rng.random((Lx, Ly, 9)))
    sysc = sw.SysConst(Lx=Lx, Ly=Ly, param=sw.Taumucs(τ=tau, g=-0.001))
    sim = DistSim(sysc, 0, 1, None)
    h, ux, uy, f = sw.Field(Lx, Ly).set(h0), sw.Field(Lx, Ly).set(u0), sw.Field(Lx, Ly), sw.Field(Lx, Ly, 9).set(f0)
    sim.set_state(h, ux, uy, f if tau != 1.0 else None)
    sim.time_loop(4)
    sim.time_loop(3, step0=4)
    sim.get_state(h, ux, uy, f)
    assert sim.last_loop_ms() > 0
    ref = onp.State(Lx, Ly)
    ref.height[...] = h0; ref.velx[...] = u0; ref.ftemp[...] = f0
    oc.time_loop(ref, onp.Params(tau=tau, g=-0.001), nsteps=7)
    assert np.array_equal(h.numpy(), ref.height)
    assert np.array_equal(ux.numpy(), ref.velx) and np.array_equal(uy.numpy(), ref.vely)
    assert np.array_equal(f.numpy(), ref.fout)
    sim.close()


def test_single_rank_slab_theta_field_and_stats():
    import swalbe_b200 as sw
    from swalbe_b200.dist import DistSim

    Lx, Ly = 96, 30
    rng = np.random.default_rng(4)
    h0 = np.asfortranarray(np.abs(1.0 + 0.2 * rng.standard_normal((Lx, Ly))) + 0.06)
    theta = np.asfortranarray(1 / 9 + 1 / 36 * rng.random((Lx, Ly)))
    ct = np.asfortranarray(np.vectorize(sw.cospi)(theta))
    sysc = sw.SysConst(Lx=Lx, Ly=Ly, param=sw.Taumucs(n=3, m=2, hmin=0.07))
    sim = DistSim(sysc, 0, 1, None)
    h, z1, z2 = sw.Field(Lx, Ly).set(h0), sw.Field(Lx, Ly), sw.Field(Lx, Ly)
    sim.set_state(h, z1, z2)
    sim.set_theta(sw.Field(Lx, Ly).set(ct))
    sim.time_loop(5)
    mn, mx, sm, cnt = sim.height_stats(1.0)
    sim.get_state(h)
    ref = onp.State(Lx, Ly)
    ref.height[...] = h0
    p = onp.Params(n=3, m=2, hmin=0.07)
    oc.time_loop(ref, p, nsteps=5, cospi_theta=ct)
    assert np.array_equal(h.numpy(), ref.height)
    assert mn == ref.height.min() and mx == ref.height.max() and cnt == int((ref.height > 1.0).sum())
    assert abs(sm - ref.height.sum()) < 1e-10 * ref.height.sum()
    # move_substrate! on the slab runtime: periodic shifts of the field, rows crossing the slab edge via the ghost rows
    for shift in [(1, 1), (0, -3), (-5, 2), (Lx + 2, 0)]:
        sim.shift_theta(*shift)
        ct = onp.circshift(ct, shift)
        sim.time_loop(2)
        oc.time_loop(ref, p, nsteps=2, cospi_theta=ct)
        sim.get_state(h)
        assert np.array_equal(h.numpy(), ref.height), shift
    with pytest.raises(ValueError):
        sim.shift_theta(0, 4)  # deeper than the ghost rows
    sim.set_theta(None)  # back to the scalar theta of the params
    sim.time_loop(3)
    sim.get_state(h)
    oc.time_loop(ref, p, nsteps=3)
    assert np.array_equal(h.numpy(), ref.height)
    with pytest.raises(ValueError):
        sim.shift_theta(1, 1)  # no field to move
    sim.close()


def test_single_rank_slab_bulk_copy_variant(monkeypatch):
    """Ghost-row layout + cp.async.bulk row prefetch (what the interior kernel of a large slab uses), forced on."""
    import swalbe_b200 as sw
    from swalbe_b200.dist import DistSim

    monkeypatch.setenv("SWALBE_BULK", "2")
    Lx, Ly = 512, 40
    rng = np.random.default_rng(12)
    h0 = np.asfortranarray(np.abs(1.0 + 0.2 * rng.standard_normal((Lx, Ly))) + 0.06)
    sysc = sw.SysConst(Lx=Lx, Ly=Ly, param=sw.Taumucs(g=-0.001))
    sim = DistSim(sysc, 0, 1, None)
    h, z1, z2 = sw.Field(Lx, Ly).set(h0), sw.Field(Lx, Ly), sw.Field(Lx, Ly)
    sim.set_state(h, z1, z2)
    sim.time_loop(6)
    sim.get_state(h)
    ref = onp.State(Lx, Ly)
    ref.height[...] = h0
    oc.time_loop(ref, onp.Params(g=-0.001), nsteps=6)
    assert np.array_equal(h.numpy(), ref.height)
    sim.close()


@pytest.mark.parametrize("nsteps,rows", [(9, 16), (3, 24), (14, 12), (6, 0)])
def test_single_rank_slab_host_loop_matches_oracle(monkeypatch, nsteps, rows):
    """swalbe_dist_time_loop_host on one rank (the ring neighbour is the slab itself): height from pinned host memory in
    row bands, the strips at the slab boundary stepped last with an exchange per step, height back to the host;
    rows = 0: slabs too small for bands take the plain path (copy, exchange, loop, copy)."""
    import torch

    import swalbe_b200 as sw
    from swalbe_b200.dist import DistSim

    monkeypatch.setenv("SWALBE_HOST_MIN_SITES", "1" if rows else str(1 << 30))
    monkeypatch.setenv("SWALBE_BAND_ROWS", str(rows or 16))
    Lx, Ly = 130, 72
    rng = np.random.default_rng(nsteps)
    h0 = np.asfortranarray(np.abs(1.0 + 0.2 * rng.standard_normal((Lx, Ly))) + 0.06)
    ux0 = np.asfortranarray(0.01 * rng.standard_normal((Lx, Ly)))
    sysc = sw.SysConst(Lx=Lx, Ly=Ly, param=sw.Taumucs(g=-0.001))
    sim = DistSim(sysc, 0, 1, None)
    junk = sw.Field(Lx, Ly).set(5.0)
    sim.set_state(junk, junk, junk)
    hin = torch.from_numpy(np.ascontiguousarray(h0.transpose())).pin_memory()
    hout = torch.full_like(hin, float("nan")).pin_memory()
    sim.time_loop_host(nsteps, host_in=hin, host_out=hout, velx=sw.Field(Lx, Ly).set(ux0))
    torch.cuda.synchronize()
    ref = onp.State(Lx, Ly)
    ref.height[...] = h0; ref.velx[...] = ux0
    oc.time_loop(ref, onp.Params(g=-0.001), nsteps=nsteps)
    assert np.array_equal(hout.numpy().transpose(), ref.height)
    h, ux, uy = sw.Field(Lx, Ly), sw.Field(Lx, Ly), sw.Field(Lx, Ly)
    sim.time_loop(2)  # the runtime is left consistent (current set, ghost rows): two more ordinary steps
    sim.get_state(h, ux, uy)
    oc.time_loop(ref, onp.Params(g=-0.001), nsteps=2)
    assert np.array_equal(h.numpy(), ref.height) and np.array_equal(ux.numpy(), ref.velx) and np.array_equal(uy.numpy(), ref.vely)
    sim.close()


def test_slab_too_thin_is_rejected():
    import swalbe_b200 as sw
    from swalbe_b200.dist import DistSim

    with pytest.raises(ValueError):
        DistSim(sw.SysConst(Lx=16, Ly=4, param=sw.Taumucs()), 0, 1, None)


def _ngpus():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("mode", ["plain", "thermal", "theta_field", "moving_theta_thermal", "host_loop", "host_loop_two_calls"])
def test_two_rank_nccl_matches_single_gpu(tmp_path, mode):
    """2 ranks over NCCL == 1 GPU, bit for bit -- including the thermal noise (counter-based on the global cell) and a
    contact-angle field whose ghost rows travel through the same exchange."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    out = tmp_path / "res.npy"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "dist_worker.py"), str(out), mode]
    subprocess.run(cmd, check=True, cwd=ROOT, timeout=600)
    got = np.load(out)
    import swalbe_b200 as sw

    Lx, Ly = 520, 96
    thermal = mode in ("thermal", "moving_theta_thermal")
    sysc = sw.SysConst(Lx=Lx, Ly=Ly, param=sw.Taumucs(kbt=1e-6 if thermal else 0.0, g=-0.001))
    st = sw.Sys(sysc, "GPU", kind="thermal" if thermal else "simple")
    rng = np.random.default_rng(5)
    st.height.set(np.asfortranarray(np.abs(1.0 + 0.2 * rng.standard_normal((Lx, Ly))) + 0.06))
    theta = np.asfortranarray(1 / 9 + 1 / 36 * rng.random((Lx, Ly)))
    from swalbe_b200 import _lib

    kw = dict(thermal_seed=77 if thermal else None, pressure_variant=_lib.PRESSURE_POWER_BROAD)
    if mode.startswith("host_loop"):
        sw.fused_steps(st, sysc, 9, **kw)
    elif mode == "moving_theta_thermal":  # C4: noise + a contact-angle pattern that moves by (1,1) between the two loops
        th, inp = sw.Field(Lx, Ly).set(theta), sw.Field(Lx, Ly).set(theta)
        sw.fused_steps(st, sysc, 5, θ=th, **kw)
        sw.move_substrate(th, inp, 98, 98)
        sw.fused_steps(st, sysc, 4, θ=th, step0=5, **kw)
    else:
        sw.fused_steps(st, sysc, 9, θ=sw.Field(Lx, Ly).set(theta) if mode == "theta_field" else None, **kw)
    assert np.array_equal(got, st.height.numpy())
