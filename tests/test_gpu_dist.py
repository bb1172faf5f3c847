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


def test_slab_too_thin_is_rejected():
    import swalbe_b200 as sw
    from swalbe_b200.dist import DistSim

    with pytest.raises(ValueError):
        DistSim(sw.SysConst(Lx=16, Ly=4, param=sw.Taumucs()), 0, 1, None)


def _ngpus():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("thermal", [False, True])
def test_two_rank_nccl_matches_single_gpu(tmp_path, thermal):
    """2 ranks over NCCL == 1 GPU, bit for bit -- including the thermal noise (counter-based on the global cell)."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    out = tmp_path / "res.npy"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "dist_worker.py"), str(out), "1" if thermal else "0"]
    subprocess.run(cmd, check=True, cwd=ROOT, timeout=600)
    got = np.load(out)
    import swalbe_b200 as sw

    Lx, Ly = 520, 96
    sysc = sw.SysConst(Lx=Lx, Ly=Ly, param=sw.Taumucs(kbt=1e-6 if thermal else 0.0, g=-0.001))
    st = sw.Sys(sysc, "GPU", kind="thermal" if thermal else "simple")
    rng = np.random.default_rng(5)
    st.height.set(np.asfortranarray(np.abs(1.0 + 0.2 * rng.standard_normal((Lx, Ly))) + 0.06))
    from swalbe_b200 import _lib

    sw.fused_steps(st, sysc, 9, thermal_seed=77 if thermal else None, pressure_variant=_lib.PRESSURE_POWER_BROAD)
    assert np.array_equal(got, st.height.numpy())
