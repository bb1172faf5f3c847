"""Height-only checkpoint / restart around the device state.

The reference restarts from `h_<t>` column vectors stored with JLD2 or BSON (`restart_from_height`,
src/initialvalues.jl:358-382; written by the scripts from `snapshot!` matrices, src/measures.jl:99-105).  The layout they
carry -- one Float64 column vector per dumped time step, the matrix flattened column-major -- is kept here with NumPy's
`.npz` as the container and, for the reference's `kind = "bson"`, a BSON reader / writer in BSON.jl's array layout
(written from the specifications, not checked against a Julia-written file: see the note at the BSON section); JLD2 is an
HDF5 dialect and needs a library this image does not have.  Plus a raw slab-parallel dump in
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


# ---- BSON container --------------------------------------------------------------------------------------------------
# The reference's second restart format (`kind = "bson"`, src/initialvalues.jl:369-376: `DataFrame(BSON.load(data))`).
# The container layer follows the BSON 1.1 specification (bsonspec.org; pinned by the specification's own example in
# tests/test_io.py); the way a Julia `Vector{Float64}` sits inside it follows BSON.jl's documented lowering of arrays of
# bits types -- a document {tag: "array", type: {tag: "datatype", params: [], name: ["Core", "Float64"]}, size: [n],
# data: <raw bytes>} -- and has NOT been checked against a file written by Julia (none exists in this image, and BSON.jl
# is not vendored by the reference): PARITY UNPINNED.  The reader is tolerant (plain BSON arrays of doubles, back
# references, any array key convention); the writer emits exactly the structure above.

import struct


def _bson_cstring(buf, pos):
    end = buf.index(b"\x00", pos)
    return buf[pos:end].decode("utf-8"), end + 1


def _bson_value(buf, pos, t):
    if t == 0x01:
        return struct.unpack_from("<d", buf, pos)[0], pos + 8
    if t == 0x02:
        n = struct.unpack_from("<i", buf, pos)[0]
        return buf[pos + 4:pos + 4 + n - 1].decode("utf-8"), pos + 4 + n
    if t in (0x03, 0x04):
        doc, end = _bson_document(buf, pos)
        if t == 0x04:  # (array: the keys are positions; BSON.jl does not rely on how they are spelled)
            doc = list(doc.values())
        return doc, end
    if t == 0x05:
        n = struct.unpack_from("<i", buf, pos)[0]
        return bytes(buf[pos + 5:pos + 5 + n]), pos + 5 + n
    if t == 0x08:
        return buf[pos] != 0, pos + 1
    if t in (0x0A, 0x06):
        return None, pos
    if t == 0x10:
        return struct.unpack_from("<i", buf, pos)[0], pos + 4
    if t in (0x12, 0x09, 0x11):
        return struct.unpack_from("<q", buf, pos)[0], pos + 8
    raise ValueError(f"BSON element type 0x{t:02x} is not supported")


def _bson_document(buf, pos=0):
    size = struct.unpack_from("<i", buf, pos)[0]
    end, pos, out = pos + size, pos + 4, {}
    while buf[pos] != 0:
        t = buf[pos]
        name, pos = _bson_cstring(buf, pos + 1)
        out[name], pos = _bson_value(buf, pos, t)
    if pos + 1 != end:
        raise ValueError("malformed BSON document (length field and terminator disagree)")
    return out, end


def _bson_encode(v) -> tuple[int, bytes]:
    if isinstance(v, bool):
        return 0x08, bytes([int(v)])
    if isinstance(v, (int, np.integer)):
        return 0x12, struct.pack("<q", int(v))
    if isinstance(v, (float, np.floating)):
        return 0x01, struct.pack("<d", float(v))
    if isinstance(v, str):
        b = v.encode("utf-8") + b"\x00"
        return 0x02, struct.pack("<i", len(b)) + b
    if isinstance(v, (bytes, bytearray)):
        return 0x05, struct.pack("<i", len(v)) + b"\x00" + bytes(v)
    if isinstance(v, Mapping):
        return 0x03, _bson_encode_document(v)
    if isinstance(v, (list, tuple)):
        return 0x04, _bson_encode_document({str(i + 1): x for i, x in enumerate(v)})  # (1-based, like Julia's indices)
    raise TypeError(f"cannot encode {type(v).__name__} as BSON")


def _bson_encode_document(d) -> bytes:
    body = b""
    for k, v in d.items():
        t, payload = _bson_encode(v)
        body += bytes([t]) + str(k).encode("utf-8") + b"\x00" + payload
    return struct.pack("<i", len(body) + 5) + body + b"\x00"


def _bson_lower_vector(v: np.ndarray) -> dict:
    """a Julia Vector{Float64} as BSON.jl lowers arrays of bits types"""
    return {"tag": "array", "type": {"tag": "datatype", "params": [], "name": ["Core", "Float64"]}, "size": [int(v.size)],
            "data": np.ascontiguousarray(v, dtype="<f8").tobytes()}


def _bson_raise_vector(v, backrefs) -> np.ndarray:
    if isinstance(v, Mapping) and v.get("tag") in ("backref", "ref"):
        v = backrefs[int(v["ref"]) - 1]
    if isinstance(v, Mapping) and v.get("tag") == "array":
        name = v["type"]["name"][-1] if isinstance(v.get("type"), Mapping) else v.get("type")
        if name != "Float64":
            raise ValueError(f"BSON array of element type {name!r}: the restart files hold Float64 heights")
        if isinstance(v["data"], (bytes, bytearray)):
            return np.frombuffer(v["data"], dtype="<f8").astype(np.float64)
        return np.asarray(v["data"], dtype=np.float64)
    if isinstance(v, list):  # (a plain BSON array of doubles)
        return np.asarray(v, dtype=np.float64)
    raise ValueError("not a stored height column")


def save_heights_bson(path: str, columns: Mapping) -> str:
    """The `h_<t>` columns as a BSON file in BSON.jl's array layout (see the note above: parity unpinned)."""
    doc = {k: _bson_lower_vector(np.ravel(_as_host(v), order="F")) for k, v in columns.items()}
    path = os.fspath(path)
    with open(path, "wb") as f:
        f.write(_bson_encode_document(doc))
    return path


def load_heights_bson(path: str) -> dict:
    with open(os.fspath(path), "rb") as f:
        doc, _ = _bson_document(f.read())
    backrefs = doc.pop("_backrefs", [])
    return {k: _bson_raise_vector(v, backrefs) for k, v in doc.items()}


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
    if kind not in ("npz", "dict", "bson"):
        raise ValueError(f"kind={kind!r}: the NumPy and BSON containers are available here (JLD2 is an HDF5 dialect that needs "
                         "Julia's JLD2.jl or an HDF5 library; neither is in this image)")
    if kind == "bson":
        cols = load_heights_bson(data)
    elif isinstance(data, (str, os.PathLike)):
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
