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
        io.restart_from_height(path, kind="jld2", timestep=1, size=(10, 10))
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
