"""The peer-memory halo protocol of the slab runtime (csrc/dist.cu), executed SYMBOLICALLY on the CPU.

What the library claims: per step a rank stores its edge rows into the neighbours' ghost rows (k_halo_push) and publishes a
sequence number; the next step's edge kernels wait for both of their own flags (k_halo_wait); there is NO "ready to
receive" handshake, because a neighbour can only compute -- and push -- the edge strips of step s+1 after it has seen this
rank's push of step s, which is issued after the kernels that read the ghost rows the neighbour is about to overwrite.

The model: every rank runs the operation sequence dist.cu issues (ordinary loop: wait, two edge strips, push, interior;
host loop: band stages, then per step a strip pair and a push, first exchange of a call as a rendezvous), ranks advance in a
RANDOM order (a rank may run arbitrarily far ahead of another as far as the flags let it), rows carry the index of the
state they hold.  Every read must find the state it expects; a push that came too early would overwrite a ghost row that
is still to be read with a later state and fail the reader.  Pure host logic: no GPU, no kernels.
"""
import ctypes as C
import random

import numpy as np
import pytest

GH = 3


def _schedule(Lx, Ly, nsteps, has_in, has_out, band, kmax):
    from swalbe_b200 import _lib

    n = C.c_int(0)
    _lib.call("swalbe_selftest_host_loop_schedule", Lx, Ly, nsteps, int(has_in), int(has_out), band, kmax, 1, None, 0, C.byref(n))
    buf = (C.c_int * (7 * max(1, n.value)))()
    _lib.call("swalbe_selftest_host_loop_schedule", Lx, Ly, nsteps, int(has_in), int(has_out), band, kmax, 1, buf, n.value, C.byref(n))
    return [tuple(buf[7 * q:7 * q + 7]) for q in range(n.value)]


class Rank:
    """one rank's slab: two moment sets of n + 2 GH row labels (index 0 == logical row -GH), flags, push counter"""

    def __init__(self, n, state0):
        self.n = n
        self.sets = [np.full(n + 2 * GH, -1), np.full(n + 2 * GH, -1)]
        self.sets[0][:] = state0  # the runtime keeps the ghost rows of its current set current between calls
        self.flags = [0, 0]       # pushes received from below / from above
        self.seq = 0              # pushes issued
        self.cur = 0
        self.prog = []            # the operations still to run, in issue order
        self.ghost_via_p2p = False

    def rows(self, s, j0, j1):
        return self.sets[s][GH + j0:GH + j1]


def _loop_program(rank, nsteps, state_of_step0, rng):
    """dist_steps(): per step wait (flags), edge strips, push, interior -- the push and the interior in either order"""
    prog, cur, n = [], rank.cur, rank.n
    via = rank.ghost_via_p2p
    seq = rank.seq
    for s in range(nsteps):
        st = state_of_step0 + s
        src, dst = cur, cur ^ 1
        if via:
            prog.append(("wait", seq))
        prog.append(("read", src, -GH, 2 * GH, st)); prog.append(("write", dst, 0, GH, st + 1))           # lower edge strip
        prog.append(("read", src, n - 2 * GH, n + GH, st)); prog.append(("write", dst, n - GH, n, st + 1))  # upper edge strip
        seq += 1
        tail = [[("push", dst, seq, st + 1)], [("read", src, 0, n, st), ("write", dst, GH, n - GH, st + 1)]]
        if rng.random() < 0.5:
            tail.reverse()
        prog += tail[0] + tail[1]
        via = True
        cur = dst
    return prog, cur, seq, via


def _host_program(rank, ops, nsteps, has_in, rng):
    """swalbe_dist_time_loop_host(): band stages on the slab's own rows, strips at the slab boundary stepped after the last
    band with one exchange per strip pair, whole-slab steps through the ordinary loop"""
    prog, n = [], rank.n
    src0 = 0 if has_in else rank.cur
    via, seq = (False if has_in else rank.ghost_via_p2p), rank.seq
    ghosts_current = not has_in
    seam = 0
    if has_in:
        prog.append(("reset",))  # the state is replaced: nothing of the old one may be read any more
    i = 0
    while i < len(ops):
        kind, s, j0, j1, band, is_seam, stage = ops[i]
        if kind == 0:
            prog.append(("upload", src0, j0, j1))
        elif kind == 2:
            prog.append(("final", src0 ^ (nsteps & 1), j0, j1, nsteps))
        elif stage < 0 and not is_seam:
            cnt = 1
            while i + cnt < len(ops) and ops[i + cnt][0] == 1 and ops[i + cnt][6] < 0 and not ops[i + cnt][5]:
                cnt += 1
            if not ghosts_current:
                prog.append(("rendezvous", src0 ^ (s & 1), s)); ghosts_current, via = True, False
            sub = Rank(n, 0)
            sub.cur, sub.seq, sub.ghost_via_p2p = src0 ^ (s & 1), seq, via
            p, _, seq, via = _loop_program(sub, cnt, s, rng)
            prog += p
            i += cnt
            continue
        else:
            src, dst = src0 ^ (s & 1), src0 ^ (s & 1) ^ 1
            if is_seam:
                if not ghosts_current:
                    prog.append(("rendezvous", src, s)); ghosts_current, via = True, False
                if seam % 2 == 0 and via:
                    prog.append(("wait", seq))
                prog.append(("read", src, j0 - GH, j1 + GH, s)); prog.append(("write", dst, j0, j1, s + 1))
                seam += 1
                if seam % 2 == 0:
                    seq += 1
                    prog.append(("push", dst, seq, s + 1)); via = True
            elif j1 > j0:
                prog.append(("read", src, j0 - GH, j1 + GH, s)); prog.append(("write", dst, j0, j1, s + 1))
        i += 1
    if not ghosts_current:
        prog.append(("rendezvous", src0 ^ (nsteps & 1), nsteps)); via = False
    return prog, src0 ^ (nsteps & 1), seq, via


def _run(ranks, rng):
    """advance randomly chosen ranks one operation at a time; rendezvous operations need every rank to have arrived"""
    R = len(ranks)
    while any(r.prog for r in ranks):
        ready = []
        for q, r in enumerate(ranks):
            if not r.prog:
                continue
            op = r.prog[0]
            if op[0] == "wait" and not (r.flags[0] >= op[1] and r.flags[1] >= op[1]):
                continue
            if op[0] == "rendezvous" and not all(x.prog and x.prog[0][0] == "rendezvous" for x in ranks):
                continue
            ready.append(q)
        assert ready, "deadlock: " + str([r.prog[0] if r.prog else None for r in ranks])
        q = rng.choice(ready)
        r = ranks[q]
        op = r.prog.pop(0)
        up, down = ranks[(q + 1) % R], ranks[(q - 1) % R]
        if op[0] == "read":
            _, s, j0, j1, want = op
            got = r.rows(s, j0, j1)
            assert (got == want).all(), (f"rank {q}", op, sorted(set(got.tolist())))
        elif op[0] == "write":
            _, s, j0, j1, st = op
            r.rows(s, j0, j1)[:] = st
        elif op[0] == "upload":
            _, s, j0, j1 = op
            r.rows(s, j0, j1)[:] = 0
        elif op[0] == "final":
            _, s, j0, j1, st = op
            assert (r.rows(s, j0, j1) == st).all(), (f"rank {q}", op)
        elif op[0] == "reset":
            r.sets[0][:] = -1; r.sets[1][:] = -1
        elif op[0] == "push":
            _, s, seq, st = op
            assert (r.rows(s, r.n - GH, r.n) == st).all() and (r.rows(s, 0, GH) == st).all(), (f"rank {q}", op)
            up.rows(s, -GH, 0)[:] = r.rows(s, r.n - GH, r.n)
            down.rows(s, down.n, down.n + GH)[:] = r.rows(s, 0, GH)
            up.flags[0], down.flags[1] = seq, seq
        elif op[0] == "rendezvous":  # NCCL exchange: everybody is here; all of them exchange the same set at once
            for x in ranks:
                assert x.prog[0][0] == "rendezvous" or x is r
            group = [r] + [x for x in ranks if x is not r]
            for x in group:
                if x is not r:
                    x.prog.pop(0)
            s = op[1]
            snap = [x.sets[s].copy() for x in ranks]
            for k, x in enumerate(ranks):
                x.sets[s][:GH] = snap[(k - 1) % R][x.n:x.n + GH]
                x.sets[s][x.n + GH:] = snap[(k + 1) % R][GH:2 * GH]


@pytest.mark.parametrize("R", [1, 2, 3, 8])
def test_ordinary_loop_needs_no_receive_handshake(R):
    rng = random.Random(R)
    for trial in range(30):
        n = rng.choice([6, 7, 12, 40])
        ranks = [Rank(n, 0) for _ in range(R)]
        done = 0
        for call in range(3):  # consecutive time_loop calls continue the same flag sequence
            steps = rng.choice([1, 2, 5, 8])
            for r in ranks:
                r.prog, cur, seq, via = _loop_program(r, steps, done, rng)
                r._after = (cur, seq, via)
            _run(ranks, rng)
            for r in ranks:
                r.cur, r.seq, r.ghost_via_p2p = r._after
                assert (r.rows(r.cur, 0, n) == done + steps).all()
            done += steps


@pytest.mark.parametrize("R", [1, 2, 3, 4])
def test_slab_host_loop_exchanges(R):
    """the slab host loop: upload sweeps, boundary strips with a push per step, whole-slab steps, download sweeps; then an
    ordinary loop on top (the runtime is left consistent: current set, ghost rows, flag sequence)"""
    rng = random.Random(100 + R)
    for trial in range(40):
        n = rng.choice([48, 66, 97, 200])
        band, kmax = rng.choice([12, 16, 22, 50]), rng.choice([1, 2, 3, 5])
        nsteps = rng.choice([1, 2, 5, 9, 14])
        has_in, has_out = rng.choice([(True, True), (True, False), (False, True)])
        ops = _schedule(32, n, nsteps, has_in, has_out, band, kmax)
        ranks = [Rank(n, 0) for _ in range(R)]
        if rng.random() < 0.5:  # the runtime has a history: some ordinary steps first (then has_in replaces the state)
            for r in ranks:
                r.prog, cur, seq, via = _loop_program(r, 3, 0, rng)
                r._after = (cur, seq, via)
            _run(ranks, rng)
            for r in ranks:
                r.cur, r.seq, r.ghost_via_p2p = r._after
            if not has_in:  # relabel: the state the host loop starts from is "state 0" of its own numbering
                for r in ranks:
                    for s in r.sets:
                        s[s == 3] = 0
                        s[(s != 0)] = -1
        for r in ranks:
            r.prog, cur, seq, via = _host_program(r, ops, nsteps, has_in, rng)
            r._after = (cur, seq, via)
        _run(ranks, rng)
        for r in ranks:
            r.cur, r.seq, r.ghost_via_p2p = r._after
            assert (r.sets[r.cur] == nsteps).all(), "owned rows and ghost rows of the current set hold the final state"
        for r in ranks:
            r.prog, cur, seq, via = _loop_program(r, 2, nsteps, rng)
            r._after = (cur, seq, via)
        _run(ranks, rng)
        for r in ranks:
            assert (r.rows(r._after[0], 0, n) == nsteps + 2).all()
