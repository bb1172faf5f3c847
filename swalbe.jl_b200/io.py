"""Height-only checkpoint / restart around the device state.

The reference restarts from `h_<t>` column vectors stored with JLD2 or BSON (`restart_from_height`,
src/initialvalues.jl:358-382; written by the scripts from `snapshot!` matrices, src/measures.jl:99-105).  Those two
container formats need Julia packages; the layout they carry -- one Float64 column vector per dumped time step, the
matrix flattened column-major -- is kept here with NumPy's `.npz` as the container, plus a raw slab-parallel dump in
which every rank of the multi-GPU runtime writes its own rows of one shared file (a row slab is one contiguous byte
range of a column-major matrix, so no gather is needed).

Host-side only: nothing here touches the GPU; `Field.numpy()` / `Field.set()` are the device boundary.
"""
from __future__ import annotations

import os
from collections.abc import Mapping

import numpy as np


def _as_host(a) -> np.ndarray:
    return a.numpy() if hasattr(a, "numpy") and not isinstance(a, np.ndarray) else np.asarray(a, dtype=np.float64)


def _npz_path(path) -> str:
    """np.savez appends ".npz" to names that lack it; use the same name on the way in and on the way out"""
    path = os.fspath(path)
    return path if path.endswith(".npz") else path + ".npz"


def save_heights(path: str, columns: Mapping) -> str:
    """Store {"h_<t>": matrix or vector} the way the scripts do: each entry flattened column-major (Julia's vec).
    Returns the file name written (".npz" appended when missing; restart_from_height accepts either spelling)."""
    out = {k: np.ravel(_as_host(v), order="F").astype(np.float64) for k, v in columns.items()}
    path = _npz_path(path)
    with open(path, "wb") as f:
        np.savez(f, **out)
    return path


def _last_column(cols) -> str:
    """the reference's `df[:, end]`: the column of the highest time step (h_<t> by integer t; other names keep their
    insertion order and come first)"""
    def key(item):
        idx, name = item
        tail = name[2:] if name.startswith("h_") else ""
        return (1, int(tail), idx) if tail.isdigit() else (0, 0, idx)
    return max(enumerate(cols), key=key)[1]


def restart_from_height(data, kind: str = "npz", timestep: int = 0, size=(512, 512)) -> np.ndarray:
    """restart_from_height(data; kind, timestep, size)  src/initialvalues.jl:358-382.

    `data`: a mapping {"h_<t>": column vector} or the path of an `.npz` written by `save_heights`.  timestep == 0 takes
    the last stored column (the reference's `df[:, end]`), otherwise the column `h_<timestep>`.  Returns the Lx x Ly
    matrix (column-major reshape, like Julia's `reshape(v, size[1], size[2])`)."""
    if kind not in ("npz", "dict"):
        raise ValueError(f"kind={kind!r}: only the NumPy container is available here (JLD2/BSON need Julia packages)")
    if isinstance(data, (str, os.PathLike)):
        data = os.fspath(data)
        with np.load(data if os.path.exists(data) else _npz_path(data)) as z:
            cols = {k: z[k] for k in z.files}
    else:
        cols = dict(data)
    if not cols:
        raise ValueError("no stored heights")
    key = _last_column(cols) if timestep == 0 else f"h_{timestep}"
    if key not in cols:
        raise KeyError(key)
    v = np.asarray(cols[key], dtype=np.float64).ravel()
    if v.size != size[0] * size[1]:
        raise ValueError(f"DimensionMismatch: column {key} has {v.size} entries, size={tuple(size)}")
    return np.asfortranarray(v.reshape(size[0], size[1], order="F"))


def dump_height_slab(path: str, slab, Lx: int, Ly: int, j_begin: int = 0) -> None:
    """Write rows [j_begin, j_begin + rows) of a global Lx x Ly height into the shared raw file `path`
    (little-endian Float64, column-major).  Every rank calls this with its own slab; rank order does not matter."""
    a = _as_host(slab)
    if a.ndim != 2 or a.shape[0] != Lx or j_begin < 0 or j_begin + a.shape[1] > Ly:
        raise ValueError(f"slab {a.shape} at row {j_begin} does not fit a {Lx} x {Ly} lattice")
    total = Lx * Ly * 8
    # create / size the file without truncating what other ranks already wrote.  O_CREAT without O_TRUNC and an
    # unconditional ftruncate to the SAME length are idempotent, so concurrent ranks cannot undo each other; a stale
    # file of another size is resized by whoever comes first (ranks that want a clean file remove it before the dump,
    # behind a barrier).
    fd = os.open(path, os.O_RDWR | os.O_CREAT, 0o644)
    try:
        if os.fstat(fd).st_size != total:
            os.ftruncate(fd, total)
    finally:
        os.close(fd)
    mm = np.memmap(path, dtype="<f8", mode="r+", shape=(Ly, Lx))  # C order (Ly, Lx) == column-major (Lx, Ly)
    mm[j_begin:j_begin + a.shape[1], :] = a.T
    mm.flush()
    del mm


def load_height_slab(path: str, Lx: int, Ly: int, j_begin: int = 0, rows: int | None = None) -> np.ndarray:
    """Read rows [j_begin, j_begin + rows) of a raw dump written by `dump_height_slab` (default: all rows)."""
    rows = Ly - j_begin if rows is None else rows
    if os.path.getsize(path) != Lx * Ly * 8:
        raise ValueError(f"{path}: size does not match a {Lx} x {Ly} Float64 lattice")
    mm = np.memmap(path, dtype="<f8", mode="r", shape=(Ly, Lx))
    out = np.asfortranarray(np.array(mm[j_begin:j_begin + rows, :]).T)
    del mm
    return out
