"""Height-only checkpoint/restart (swalbe_b200.io): the reference's own restart test (test/initialvalues.jl:84-100,
doctest src/initialvalues.jl:338-352) with the NumPy container, and the slab-parallel raw dump."""
import numpy as np
import pytest

import swalbe_b200  # noqa: F401  (import shim for the package directory swalbe.jl_b200/)
from swalbe_b200 import io


def test_restart_from_height_like_the_reference(tmp_path):
    rng = np.random.default_rng(0)
    h1, h2 = rng.random((10, 10)), rng.random((10, 10))
    f = {"h_1": h1.ravel(order="F"), "h_2": h2.ravel(order="F")}
    path = str(tmp_path / "file.npz")
    io.save_heights(path, f)
    assert np.array_equal(io.restart_from_height(path, timestep=1, size=(10, 10)), h1)
    assert np.array_equal(io.restart_from_height(path, timestep=2, size=(10, 10)), h2)
    assert np.array_equal(io.restart_from_height(path, timestep=0, size=(10, 10)), h2)  # last column
    assert np.array_equal(io.restart_from_height(f, kind="dict", timestep=1, size=(10, 10)), h1)
    with pytest.raises(KeyError):
        io.restart_from_height(path, timestep=3, size=(10, 10))
    with pytest.raises(ValueError):
        io.restart_from_height(path, timestep=1, size=(10, 11))
    with pytest.raises(ValueError):
        io.restart_from_height(path, kind="jld2", timestep=1, size=(10, 10))  # (an HDF5 dialect: no library for it here)
    # a name without the suffix reads back under the same name; "last column" = highest time step, not insertion order
    bare = str(tmp_path / "run7")
    assert io.save_heights(bare, {"h_10": h2.ravel(order="F"), "h_9": h1.ravel(order="F"), "h_2": h1.ravel(order="F")}) == bare + ".npz"
    assert np.array_equal(io.restart_from_height(bare, timestep=9, size=(10, 10)), h1)
    assert np.array_equal(io.restart_from_height(bare, timestep=0, size=(10, 10)), h2)
    assert np.array_equal(io.restart_from_height({"h_10": h2.ravel(order="F"), "h_9": h1.ravel(order="F")}, kind="dict",
                                                 size=(10, 10)), h2)
    # matrices are flattened column-major on the way in, non-square sizes keep their orientation
    m = np.arange(12, dtype=np.float64).reshape(3, 4)
    io.save_heights(path, {"h_5": m})
    assert np.array_equal(io.restart_from_height(path, timestep=5, size=(3, 4)), m)


def test_slab_parallel_dump_roundtrip(tmp_path):
    Lx, Ly, ranks = 12, 20, 4
    rng = np.random.default_rng(1)
    h = np.asfortranarray(rng.random((Lx, Ly)))
    path = str(tmp_path / "h.raw")
    rows = Ly // ranks
    for r in (2, 0, 3, 1):  # any order
        io.dump_height_slab(path, h[:, r * rows:(r + 1) * rows], Lx, Ly, j_begin=r * rows)
    assert np.array_equal(io.load_height_slab(path, Lx, Ly), h)
    assert np.array_equal(io.load_height_slab(path, Lx, Ly, j_begin=5, rows=7), h[:, 5:12])
    assert np.array_equal(np.fromfile(path, dtype="<f8"), h.ravel(order="F"))  # == Julia's write(io, h)
    stale = str(tmp_path / "stale.raw")  # a left-over file of another size is resized, not appended to
    open(stale, "wb").write(b"x" * 17)
    for r in range(ranks):
        io.dump_height_slab(stale, h[:, r * rows:(r + 1) * rows], Lx, Ly, j_begin=r * rows)
    assert np.array_equal(io.load_height_slab(stale, Lx, Ly), h)
    with pytest.raises(ValueError):
        io.dump_height_slab(path, h[:, :3], Lx, Ly, j_begin=18)
    with pytest.raises(ValueError):
        io.load_height_slab(path, Lx, Ly + 1)


def test_bson_container_and_restart(tmp_path):
    """kind = "bson" (src/initialvalues.jl:369-376).  The container layer against the BSON specification's own example
    (bsonspec.org: {"hello": "world"}), the height columns through BSON.jl's documented array layout (round trip; parity with
    a Julia-written file is unpinned: no BSON.jl here), tolerant reading of plain arrays and back references."""
    spec = b"\x16\x00\x00\x00\x02hello\x00\x06\x00\x00\x00world\x00\x00"
    assert io._bson_document(spec)[0] == {"hello": "world"} and io._bson_encode_document({"hello": "world"}) == spec
    spec2 = (b"\x31\x00\x00\x00\x04BSON\x00\x26\x00\x00\x00\x020\x00\x08\x00\x00\x00awesome\x00\x011\x00\x33\x33\x33\x33\x33\x33"
             b"\x14\x40\x102\x00\xc2\x07\x00\x00\x00\x00")  # {"BSON": ["awesome", 5.05, 1986]}
    assert io._bson_document(spec2)[0] == {"BSON": ["awesome", 5.05, 1986]}
    rng = np.random.default_rng(3)
    h1, h2 = rng.random((10, 7)), rng.random((10, 7))
    path = str(tmp_path / "file.bson")
    io.save_heights_bson(path, {"h_1": h1, "h_20": h2})
    assert np.array_equal(io.restart_from_height(path, kind="bson", timestep=1, size=(10, 7)), h1)
    assert np.array_equal(io.restart_from_height(path, kind="bson", timestep=0, size=(10, 7)), h2)
    doc = io._bson_document(open(path, "rb").read())[0]
    assert doc["h_1"]["tag"] == "array" and doc["h_1"]["type"]["name"] == ["Core", "Float64"] and doc["h_1"]["size"] == [70]
    assert doc["h_1"]["data"] == h1.ravel(order="F").astype("<f8").tobytes()
    # what other writers may produce: a plain array of doubles, a back reference to a shared array
    raw = io._bson_encode_document({"h_3": [1.0, 2.0, 3.0, 4.0], "h_4": {"tag": "backref", "ref": 1},
                                    "_backrefs": [io._bson_lower_vector(np.arange(4.0))]})
    p2 = tmp_path / "other.bson"
    p2.write_bytes(raw)
    assert np.array_equal(io.restart_from_height(str(p2), kind="bson", timestep=3, size=(2, 2)), np.array([[1.0, 3.0], [2.0, 4.0]]))
    assert np.array_equal(io.restart_from_height(str(p2), kind="bson", timestep=4, size=(2, 2)), np.array([[0.0, 2.0], [1.0, 3.0]]))
    with pytest.raises(ValueError):
        io._bson_document(spec[:-1] + b"\x01")
