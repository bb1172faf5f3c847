"""Host-side logic of the slab decomposition, exercised with 2 processes over gloo on the CPU.

Each rank owns Ly/2 rows plus HALO_DEPTH ghost rows per side, exchanges exactly the messages
``SlabDecomposition.messages`` prescribes (the same four per plane that swalbe_dist_* posts through NCCL), and
advances its padded slab with the oracle; the gathered result must equal the oracle run on the whole lattice,
bit for bit.  This pins the halo depth (3), the neighbour ring with periodic wrap and the row bookkeeping."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, Lx, Ly, nsteps, tau, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from oracle import oracle_c as oc
    from oracle import oracle_np as onp
    from swalbe_b200.dist import HALO_DEPTH, SlabDecomposition, slab_of

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dec = SlabDecomposition(Ly, world)
    j0, n = dec.rows(rank)
    d = HALO_DEPTH
    rng = np.random.default_rng(123)  # every rank builds the same global field and cuts its slab out
    hg = np.asfortranarray(np.abs(1.0 + 0.2 * rng.standard_normal((Lx, Ly))) + 0.06)
    fg = np.asfortranarray(0.1 + 0.01 * rng.random((Lx, Ly, 9)))
    p = onp.Params(tau=tau, g=-0.001)
    st = onp.State(Lx, n + 2 * d)  # padded slab: local row r lives at padded row r + d
    st.height[:, d:d + n] = slab_of(hg, dec, rank)
    st.ftemp[:, d:d + n, :] = slab_of(fg, dec, rank)

    def exchange(arr):
        """the four messages of SlabDecomposition.messages, rows given in local coordinates [−d, n+d)"""
        reqs, recvs = [], []
        for kind, peer, (a, b) in dec.messages(rank):
            view = arr[:, a + d:b + d]
            if kind == "send":
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(view)), peer))
            else:
                buf = torch.empty(view.shape, dtype=torch.float64)
                reqs.append(dist.irecv(buf, peer))
                recvs.append((view, buf))
        for r in reqs:
            r.wait()
        for view, buf in recvs:
            view[...] = buf.numpy()

    fields = [st.height, st.velx, st.vely] + ([st.ftemp[:, :, k] for k in range(9)] if tau != 1.0 else [])
    for _ in range(nsteps):
        for f in fields:
            exchange(f)
        oc.step(st, p)  # periodic on the padded slab: only the 3 outermost rows per side get contaminated
    np.save(os.path.join(out_dir, f"h{rank}.npy"), st.height[:, d:d + n])
    np.save(os.path.join(out_dir, f"f{rank}.npy"), st.fout[:, d:d + n, :])
    dist.barrier()
    dist.destroy_process_group()


def _shift_worker(rank, world, port, Lx, Ly, shifts, out_dir):
    """move_substrate! across slabs (swalbe_dist_shift_theta): shift with the ghost rows, exchange, repeat"""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from swalbe_b200.dist import HALO_DEPTH, SlabDecomposition, shift_padded_slab, slab_of

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dec = SlabDecomposition(Ly, world)
    j0, n = dec.rows(rank)
    d = HALO_DEPTH
    theta = np.asfortranarray(np.random.default_rng(7).random((Lx, Ly)))
    pad = np.zeros((Lx, n + 2 * d))
    pad[:, d:d + n] = slab_of(theta, dec, rank)

    def exchange(arr):
        reqs, recvs = [], []
        for kind, peer, (a, b) in dec.messages(rank):
            view = arr[:, a + d:b + d]
            if kind == "send":
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(view)), peer))
            else:
                buf = torch.empty(view.shape, dtype=torch.float64)
                reqs.append(dist.irecv(buf, peer))
                recvs.append((view, buf))
        for r in reqs:
            r.wait()
        for view, buf in recvs:
            view[...] = buf.numpy()

    exchange(pad)
    for q, (sx, sy) in enumerate(shifts):
        pad = shift_padded_slab(pad, sx, sy)
        exchange(pad)
        np.save(os.path.join(out_dir, f"t{q}_{rank}.npy"), pad[:, d:d + n])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_theta_shift_matches_global_circshift(tmp_path):
    import torch.multiprocessing as mp

    from oracle import oracle_np as onp
    from swalbe_b200.dist import shift_padded_slab

    Lx, Ly, world = 17, 16, 2
    shifts = [(1, 1), (0, -3), (-5, 2), (Lx + 2, 0), (3, 3)]
    mp.spawn(_shift_worker, args=(world, _free_port(), Lx, Ly, shifts, str(tmp_path)), nprocs=world, join=True)
    want = np.asfortranarray(np.random.default_rng(7).random((Lx, Ly)))
    for q, sh in enumerate(shifts):
        want = onp.circshift(want, sh)
        got = np.concatenate([np.load(tmp_path / f"t{q}_{r}.npy") for r in range(world)], axis=1)
        assert np.array_equal(got, want), (q, sh)
    with pytest.raises(ValueError):
        shift_padded_slab(np.zeros((4, 10)), 0, 4)


@pytest.mark.parametrize("tau", [1.0, 0.8])
def test_two_rank_slabs_match_global_oracle(tmp_path, tau):
    import torch.multiprocessing as mp

    from oracle import oracle_c as oc
    from oracle import oracle_np as onp

    Lx, Ly, nsteps, world = 33, 24, 5, 2
    mp.spawn(_worker, args=(world, _free_port(), Lx, Ly, nsteps, tau, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(123)
    hg = np.asfortranarray(np.abs(1.0 + 0.2 * rng.standard_normal((Lx, Ly))) + 0.06)
    fg = np.asfortranarray(0.1 + 0.01 * rng.random((Lx, Ly, 9)))
    ref = onp.State(Lx, Ly)
    ref.height[...] = hg
    ref.ftemp[...] = fg
    oc.time_loop(ref, onp.Params(tau=tau, g=-0.001), nsteps=nsteps)
    h = np.concatenate([np.load(tmp_path / f"h{r}.npy") for r in range(world)], axis=1)
    f = np.concatenate([np.load(tmp_path / f"f{r}.npy") for r in range(world)], axis=1)
    assert np.array_equal(h, ref.height)
    assert np.array_equal(f, ref.fout)


def test_slab_decomposition_bookkeeping():
    from swalbe_b200.dist import SlabDecomposition

    dec = SlabDecomposition(32, 4)
    assert [dec.rows(r) for r in range(4)] == [(0, 8), (8, 8), (16, 8), (24, 8)]
    assert dec.neighbours(0) == (3, 1) and dec.neighbours(3) == (2, 0)
    assert dec.ghost_rows(0) == ([29, 30, 31], [8, 9, 10])
    assert dec.ghost_rows(3) == ([21, 22, 23], [0, 1, 2])
    assert dec.owner(-1) == 3 and dec.owner(8) == 1
    # every ghost row is owned by the neighbour the message table says it comes from
    for r in range(4):
        lo, hi = dec.ghost_rows(r)
        down, up = dec.neighbours(r)
        assert all(dec.owner(j) == down for j in lo) and all(dec.owner(j) == up for j in hi)
        kinds = [(k, peer) for k, peer, _ in dec.messages(r)]
        assert kinds == [("send", up), ("send", down), ("recv", down), ("recv", up)]
    with pytest.raises(ValueError):
        SlabDecomposition(30, 4)
    with pytest.raises(ValueError):
        SlabDecomposition(20, 4)  # 5-row slabs are thinner than 2 x 3 halo rows
    one = SlabDecomposition(16, 1)
    assert one.neighbours(0) == (0, 0) and one.ghost_rows(0) == ([13, 14, 15], [0, 1, 2])
