"""Every `file:line` citation into the reference tree (header, docs, kernels, oracle, host mirror, tests) must point at
an existing file and an existing line range.  Runs only where the reference tree is mounted (this container)."""
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
CITE = re.compile(r"(?<![\w/])((?:src|test|scripts|docs/src)/[\w./]+?\.(?:jl|md)):(\d+)(?:-(\d+))?")
BARE = re.compile(r"(?<![\w/.])(\w+\.jl):(\d+)(?:-(\d+))?")  # "pressure.jl:119-155": a file of src/ named without its directory


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_reference_citations_resolve():
    files = [os.path.join(ROOT, f) for f in ("DESIGN.md", "INTEGRATION.md", "README.md", "BASELINE.md", "bench.py")]
    for pat in ("include/*.h", "swalbe.jl_b200/*.py", "swalbe.jl_b200/csrc/*.cu*", "swalbe.jl_b200/csrc/*.h",
                "swalbe.jl_b200/julia/*.jl", "oracle/*.py", "oracle/*.c", "tests/*.py"):
        files += glob.glob(os.path.join(ROOT, pat))
    lengths, bad, n = {}, [], 0
    for f in files:
        for m in CITE.finditer(open(f, encoding="utf-8").read()):
            rel, a, b = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
            path = os.path.join(REF, rel)
            if rel not in lengths:
                lengths[rel] = sum(1 for _ in open(path, encoding="utf-8")) if os.path.isfile(path) else -1
            n += 1
            if lengths[rel] < 0 or not (1 <= a <= b <= lengths[rel]):
                bad.append(f"{os.path.relpath(f, ROOT)}: {m.group(0)} (file has {lengths[rel]} lines)")
        for m in BARE.finditer(open(f, encoding="utf-8").read()):
            rel, a, b = "src/" + m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
            path = os.path.join(REF, rel)
            if not os.path.isfile(path):
                continue  # (a file of this repository, e.g. SwalbeB200.jl)
            if rel not in lengths:
                lengths[rel] = sum(1 for _ in open(path, encoding="utf-8"))
            n += 1
            if not (1 <= a <= b <= lengths[rel]):
                bad.append(f"{os.path.relpath(f, ROOT)}: {m.group(0)} (file has {lengths[rel]} lines)")
    assert n > 100, n  # the citations are there at all
    assert not bad, "\n".join(bad)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_julia_glue_extends_functions_the_reference_defines():
    """Every `Swalbe.f` the Julia glue adds methods to (or the INTEGRATION examples call) is a function of the reference."""
    src = open(os.path.join(ROOT, "swalbe.jl_b200", "julia", "SwalbeB200.jl"), encoding="utf-8").read()
    names = set(re.findall(r"(?:function\s+|^)Swalbe\.([\w!∇²]+)\s*\(", src, flags=re.M))
    assert len(names) >= 12, names
    ref_src = "\n".join(open(f, encoding="utf-8").read() for f in glob.glob(os.path.join(REF, "src", "*.jl")))
    missing = [n for n in sorted(names) if not re.search(r"(?:function\s+|^)%s\s*\(" % re.escape(n), ref_src, flags=re.M)]
    assert not missing, missing
    # the state types and fields the glue touches exist upstream
    for field in ("fout", "ftemp", "feq", "height", "velx", "vely", "vsq", "pressure", "dgrad", "Fx", "Fy", "slipx", "slipy",
                  "h∇px", "h∇py", "kbtx", "kbty"):
        assert re.search(r"\b%s\s*::" % re.escape(field), ref_src), field
    for ty in ("CuState", "CuState_thermal", "SysConst"):
        assert re.search(r"struct\s+%s\b" % ty, ref_src), ty
