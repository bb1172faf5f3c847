"""Host side of the multi-GPU slab runtime (one process per GPU; row slabs along y; NCCL halo rows).

The reference has no multi-GPU path; this mirrors what a Julia driver would do around the
``swalbe_dist_*`` entry points of the C ABI: decide the slab of each rank, ship rank 0's NCCL unique id to the
other ranks (here through ``torch.distributed``; any transport works), scatter the initial fields, step, gather.

``SlabDecomposition`` is pure index logic and is what the CPU (gloo, world_size 2) tests exercise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

HALO_DEPTH = 3  # dependency radius of one fused step in h (p <- h, ∇p <- p, streaming <- f*)


class SlabDecomposition:
    """Row-slab decomposition of a periodic Lx x Ly lattice along y (the slow index) over nranks ranks."""

    def __init__(self, Ly: int, nranks: int, depth: int = HALO_DEPTH):
        if nranks < 1:
            raise ValueError("nranks must be >= 1")
        if Ly % nranks:
            raise ValueError(f"Ly={Ly} is not divisible by nranks={nranks}")
        if Ly // nranks < 2 * depth:
            raise ValueError(f"slab of {Ly // nranks} rows is thinner than 2x the halo depth {depth}")
        self.Ly, self.nranks, self.depth, self.rows_per_rank = Ly, nranks, depth, Ly // nranks

    def rows(self, rank: int):
        """(first global row, number of rows) owned by `rank`."""
        return rank * self.rows_per_rank, self.rows_per_rank

    def neighbours(self, rank: int):
        """(down, up): the ranks owning the rows just below / above this slab (periodic ring)."""
        return (rank - 1) % self.nranks, (rank + 1) % self.nranks

    def ghost_rows(self, rank: int):
        """Global row indices that fill the `depth` ghost rows below and above the slab of `rank`."""
        j0, n = self.rows(rank)
        lo = [(j0 - self.depth + k) % self.Ly for k in range(self.depth)]
        hi = [(j0 + n + k) % self.Ly for k in range(self.depth)]
        return lo, hi

    def owner(self, j: int) -> int:
        return (j % self.Ly) // self.rows_per_rank

    def messages(self, rank: int):
        """The four halo messages of one exchange: (kind, peer, local row range) with local rows in [0, n)."""
        down, up = self.neighbours(rank)
        n, d = self.rows_per_rank, self.depth
        return [("send", up, (n - d, n)), ("send", down, (0, d)), ("recv", down, (-d, 0)), ("recv", up, (n, n + d))]


def shift_padded_slab(padded: np.ndarray, sx: int, sy: int, depth: int = HALO_DEPTH) -> np.ndarray:
    """Index logic of swalbe_dist_shift_theta on the host: `padded` is a slab (Lx, n + 2*depth) whose ghost rows are
    current; returns the slab after the GLOBAL periodic shift field[i, j] <- field[i - sx, j - sy] with the owned rows
    filled (rows that cross the slab edge come out of the ghost rows, so |sy| <= depth) and stale ghost rows, which the
    caller exchanges again."""
    if abs(sy) > depth:
        raise ValueError(f"|sy| = {abs(sy)} exceeds the ghost depth {depth}")
    Lx, rows = padded.shape
    n = rows - 2 * depth
    out = padded.copy()
    out[:, depth:depth + n] = np.roll(padded, sx, axis=0)[:, depth - sy:depth - sy + n]
    return out


def nccl_unique_id() -> bytes:
    raw = (C.c_ubyte * _lib.NCCL_UNIQUE_ID_BYTES)()
    _lib.call("swalbe_dist_unique_id", raw)
    return bytes(raw)


def broadcast_unique_id_torch() -> bytes:
    """Rank 0 creates the NCCL unique id; torch.distributed (any backend) ships it to the other ranks."""
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    buf = torch.zeros(_lib.NCCL_UNIQUE_ID_BYTES, dtype=torch.uint8, device=dev)
    if dist.get_rank() == 0:
        buf = torch.tensor(list(nccl_unique_id()), dtype=torch.uint8, device=dev)
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().tolist())


class DistSim:
    """One rank's slab of a multi-GPU thin-film simulation (wraps a swalbe_dist handle)."""

    def __init__(self, sysc, rank: int, nranks: int, unique_id: bytes | None, thermal_seed=None, **param_kw):
        from . import _c_params  # late import: package __init__ imports this module lazily

        self.sysc, self.rank, self.nranks = sysc, rank, nranks
        self.decomp = SlabDecomposition(sysc.Ly, nranks)
        self.j_begin, self.j_count = self.decomp.rows(rank)
        q = _c_params(sysc.param, thermal_seed=thermal_seed, **param_kw)
        raw = (C.c_ubyte * _lib.NCCL_UNIQUE_ID_BYTES)(*unique_id) if unique_id is not None else None
        h = C.c_void_p()
        _lib.call("swalbe_dist_create", C.byref(h), raw, rank, nranks, sysc.Lx, sysc.Ly, C.byref(q))
        self.handle = h
        jb, jc = C.c_int(), C.c_int()
        _lib.call("swalbe_dist_local_rows", h, C.byref(jb), C.byref(jc))
        assert (jb.value, jc.value) == (self.j_begin, self.j_count)

    def _stream(self):
        from . import _stream

        return _stream()

    def set_state(self, height, velx, vely, ftemp=None):
        """height/velx/vely: Fields of shape (Lx, j_count) holding this rank's rows; ftemp: (Lx, j_count, 9) or None."""
        for f in (height, velx, vely):
            assert f.shape == (self.sysc.Lx, self.j_count), f.shape
        _lib.call("swalbe_dist_set_state", self.handle, height.ptr, velx.ptr, vely.ptr,
                  ftemp.ptr if ftemp is not None else None, self._stream())

    def set_theta(self, cospi_theta_slab):
        """cospi.(θ) of this rank's rows as a Field of shape (Lx, j_count), or None for the scalar θ of the params."""
        _lib.call("swalbe_dist_set_theta", self.handle, cospi_theta_slab.ptr if cospi_theta_slab is not None else None,
                  self._stream())

    def shift_theta(self, sx: int, sy: int):
        """move_substrate! across the slabs: θ[i, j] <- θ[i - sx, j - sy] on the global lattice (|sy| <= 3); collective."""
        _lib.call("swalbe_dist_shift_theta", self.handle, int(sx), int(sy), self._stream())

    def height_stats(self, thresh=0.055):
        """(min, max, sum, count(h > thresh)) of this rank's rows; combine across ranks with min/max/+/+."""
        import torch

        out = torch.empty(4, dtype=torch.float64, device="cuda")
        _lib.call("swalbe_dist_height_stats", self.handle, C.c_void_p(out.data_ptr()), float(thresh), self._stream())
        mn, mx, sm, cnt = out.cpu().tolist()
        return mn, mx, sm, int(cnt)

    def time_loop(self, nsteps: int, step0: int = 0):
        _lib.call("swalbe_dist_time_loop", self.handle, int(nsteps), int(step0), self._stream())

    def time_loop_host(self, nsteps: int, host_in=None, host_out=None, velx=None, vely=None, step0: int = 0):
        """time_loop with this rank's rows of the height coming from / going to pinned host memory (CPU torch tensors of
        Lx * j_count float64): the state is replaced by (host_in, velx, vely) -- velocities device Fields or None for zero
        -- and the final height lands in host_out once the current stream has been synchronised (swalbe_dist_time_loop_host)."""
        def hp(x):
            if x is None:
                return None
            import torch

            Lx, n = self.sysc.Lx, self.j_count
            if (x.device.type != "cpu" or x.dtype != torch.float64 or not x.is_contiguous()
                    or tuple(x.shape) not in ((n, Lx), (Lx * n,))):
                raise ValueError(f"host plane: contiguous float64 CPU tensor of shape ({n}, {Lx}) or flat expected (the memory "
                                 f"order of the slab: i fastest), got {tuple(x.shape)}")
            return C.c_void_p(x.data_ptr())

        _lib.call("swalbe_dist_time_loop_host", self.handle, int(nsteps), int(step0), hp(host_in),
                  velx.ptr if velx is not None else None, vely.ptr if vely is not None else None, hp(host_out), self._stream())

    def get_state(self, height=None, velx=None, vely=None, fout=None):
        p = lambda f: f.ptr if f is not None else None  # noqa: E731
        _lib.call("swalbe_dist_get_state", self.handle, p(height), p(velx), p(vely), p(fout), self._stream())

    def uses_peer_memory(self) -> bool:
        """True when the halo rows travel as stores into the neighbours' memory (NVLink), False when through NCCL."""
        yes = C.c_int()
        _lib.call("swalbe_dist_uses_peer_memory", self.handle, C.byref(yes))
        return bool(yes.value)

    def last_loop_ms(self) -> float:
        ms = C.c_float()
        _lib.call("swalbe_dist_last_loop_ms", self.handle, C.byref(ms))
        return ms.value

    def close(self):
        if self.handle is not None:
            _lib.call("swalbe_dist_destroy", self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def slab_of(a: np.ndarray, decomp: SlabDecomposition, rank: int) -> np.ndarray:
    """The rows of a global (Lx, Ly[, K]) array owned by `rank`, as a Fortran-ordered copy."""
    j0, n = decomp.rows(rank)
    return np.asfortranarray(a[:, j0:j0 + n])
