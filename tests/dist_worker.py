"""torchrun worker for tests/test_gpu_dist.py: 2 ranks, one GPU each, NCCL halos; rank 0 saves the gathered height."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import swalbe_b200 as sw
    from swalbe_b200.dist import DistSim, broadcast_unique_id_torch, slab_of

    out, mode = sys.argv[1], sys.argv[2]
    thermal = mode in ("thermal", "moving_theta_thermal")
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    Lx, Ly = 520, 96
    sysc = sw.SysConst(Lx=Lx, Ly=Ly, param=sw.Taumucs(kbt=1e-6 if thermal else 0.0, g=-0.001))
    sim = DistSim(sysc, rank, world, broadcast_unique_id_torch(), thermal_seed=77 if thermal else None)
    rng = np.random.default_rng(5)
    hg = np.asfortranarray(np.abs(1.0 + 0.2 * rng.standard_normal((Lx, Ly))) + 0.06)
    n = sim.j_count
    h, z1, z2 = sw.Field(Lx, n).set(slab_of(hg, sim.decomp, rank)), sw.Field(Lx, n), sw.Field(Lx, n)
    sim.set_state(h, z1, z2)
    theta = np.asfortranarray(1 / 9 + 1 / 36 * rng.random((Lx, Ly)))  # same draw order as the single-GPU reference run
    if mode in ("theta_field", "moving_theta_thermal"):
        ct = sw.cospi_field(sw.Field(Lx, Ly).set(theta)).numpy()  # same device cospi.(θ) as the single-GPU run
        sim.set_theta(sw.Field(Lx, n).set(slab_of(ct, sim.decomp, rank)))
    if mode.startswith("host_loop"):  # the slab's loop from / to pinned host memory: banded sweeps, strips at the slab edges
        os.environ["SWALBE_HOST_MIN_SITES"], os.environ["SWALBE_BAND_ROWS"] = "1", "16"
        hin = torch.from_numpy(np.ascontiguousarray(slab_of(hg, sim.decomp, rank).transpose())).pin_memory()
        hout = torch.full_like(hin, float("nan")).pin_memory()
        junk = sw.Field(Lx, n).set(7.0)
        sim.set_state(junk, junk, junk)            # whatever the runtime held before is replaced
        if mode == "host_loop_two_calls":           # upload sweep + whole-slab steps, then whole-slab steps + download sweep
            sim.time_loop_host(5, host_in=hin)
            sim.time_loop_host(4, host_out=hout)
        else:
            sim.time_loop_host(9, host_in=hin, host_out=hout)
        torch.cuda.synchronize()
        sim.get_state(h)
        assert torch.equal(hout.cuda(), h.t), "downloaded rows differ from the runtime's height"
    else:
        sim.time_loop(5)
        if mode == "moving_theta_thermal":
            sim.shift_theta(1, 1)
        sim.time_loop(4, step0=5)
        sim.get_state(h)
    parts = [torch.empty_like(h.t) for _ in range(world)]
    dist.all_gather(parts, h.t)
    if rank == 0:
        full = torch.cat(parts, dim=0)  # torch layout is (rows, Lx)
        np.save(out, np.asfortranarray(full.cpu().numpy().transpose()))
    dist.barrier()
    sim.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
