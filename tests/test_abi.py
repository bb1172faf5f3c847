"""CPU-side checks of the drop-in boundary: the shared library loads and exports exactly the symbols
include/swalbe_b200.h declares, error codes work without a GPU, and the host mirror refuses a CPU path."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge

    ge.build()
    from swalbe_b200 import _lib

    return _lib.load()


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "swalbe_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(swalbe_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    from swalbe_b200 import _lib

    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/swalbe_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes table and header disagree"


def test_version_and_error_string(lib):
    assert lib.swalbe_version() == 100
    assert isinstance(lib.swalbe_last_error(), bytes)


def test_argument_errors_need_no_gpu(lib):
    from swalbe_b200 import _lib

    # bad extents / NULL pointers are rejected before any CUDA call
    assert lib.swalbe_moments_d2q9(None, None, None, None, 0, 5, None) == _lib.ERR_EXTENT
    assert lib.swalbe_moments_d2q9(None, None, None, None, 5, 5, None) == _lib.ERR_ARG
    assert b"NULL" in lib.swalbe_last_error()
    # DomainError for exponents the array-form pressure does not know (src/pressure.jl:101-107)
    one = C.c_void_p(8)
    rc = lib.swalbe_filmpressure(one, C.c_void_p(16), None, 1.0, 1.0, None, 4, 2, 0.1, 0.1, _lib.PRESSURE_FAST, 5, 5, None)
    assert rc == _lib.ERR_DOMAIN and b"DomainError((4, 2))" in lib.swalbe_last_error()
    with pytest.raises(_lib.DomainError):
        _lib.check(rc)


def test_struct_layouts_match_header():
    from swalbe_b200 import _lib

    assert C.sizeof(_lib.CState) == 17 * 8
    assert C.sizeof(_lib.CParams) == 8 * 8 + 2 * 4 + 8 + 8 + 2 * 4 + 4 + 4 + 3 * 8 + 4 + 4 + 8
    assert C.sizeof(_lib.CLogs) == 48


def test_no_cpu_fallback():
    import torch

    import swalbe_b200 as sw

    sysc = sw.SysConst(Lx=5, Ly=5, param=sw.Taumucs())
    with pytest.raises(sw.SwalbeError):
        sw.Sys(sysc, "CPU")
    if not torch.cuda.is_available():
        with pytest.raises(sw.SwalbeError):
            sw.Sys(sysc, "GPU")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "swalbe.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "oracle" not in text.lower() or f == "README.md", f"{f} mentions the oracle"


def test_taumucs_defaults():
    import swalbe_b200 as sw

    p = sw.Taumucs()
    assert p.Tmax == 1000 and p.tdump == 100 and p.tau == 1.0 and p.n == 9 and p.m == 3
    assert p.mu.hex() == "0x1.5555555555557p-3"  # cs^2*(tau-0.5) is 2 ulp above 1/6 (src/initialize.jl:49)
    assert p.theta == 1 / 9 and p.hmin == 0.1 and p.hcrit == 0.05 and p.gamma == 0.01 and p.delta == 1.0
    with pytest.raises(TypeError):
        sw.SysConst(Lx=5, Ly=5)


def test_header_is_plain_c99(tmp_path):
    """The drop-in boundary must be consumable from C (and therefore from any FFI): no C++-isms outside the guards."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = tmp_path / "use_header.c"
    src.write_text('#include "swalbe_b200.h"\n'
                   "int main(void) { swalbe_state s; swalbe_params p; swalbe_loop_logs l; (void)s; (void)p; (void)l;\n"
                   "  return (int)sizeof(swalbe_state) - 17 * (int)sizeof(void *) + SWALBE_LOOP_SKIP_AUX - 2; }\n")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only",
                           "-I", os.path.join(ROOT, "include"), str(src)])


def _classify_c(param: str) -> str:
    p = param.strip()
    if "*" in p or "[" in p:
        return "ptr"
    if "double" in p:
        return "f64"
    if "unsigned long long" in p:
        return "u64"
    if "size_t" in p:
        return "u64"  # (LP64: size_t and unsigned long long are both 8-byte unsigned; ctypes aliases them)
    if re.search(r"\bint\b", p):
        return "i32"
    raise AssertionError(f"unclassified C parameter: {param!r}")


def _header_prototypes():
    src = open(os.path.join(ROOT, "include", "swalbe_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for name, params in re.findall(r"\b(swalbe_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        params = params.strip()
        protos[name] = [] if params in ("", "void") else [_classify_c(p) for p in params.split(",")]
    return protos


def test_ctypes_signatures_match_header_prototypes():
    """Argument by argument: every ctypes argtypes list agrees with the C prototype it binds."""
    from swalbe_b200 import _lib

    def classify(t):
        if t is C.c_double:
            return "f64"
        if t is C.c_int:
            return "i32"
        if t is C.c_ulonglong:
            return "u64"
        if t is C.c_void_p or t is C.c_char_p or hasattr(t, "contents"):
            return "ptr"
        raise AssertionError(f"unclassified ctypes type {t}")

    protos = _header_prototypes()
    assert sorted(protos) == sorted(_lib.SIGNATURES)
    for name, argtypes in _lib.SIGNATURES.items():
        assert [classify(t) for t in argtypes] == protos[name], name


def test_julia_glue_binds_existing_symbols_with_matching_types():
    """The Julia glue cannot run here (no Julia in the image), but every ccall in it must name a symbol the header
    declares, with an argument-type tuple that matches the C prototype position by position, and its blocks must
    balance."""
    src = open(os.path.join(ROOT, "swalbe.jl_b200", "julia", "SwalbeB200.jl"), encoding="utf-8").read()
    protos = _header_prototypes()
    jl = {"Cdouble": "f64", "Cint": "i32", "Culonglong": "u64", "Csize_t": "u64"}

    def classify(t):
        t = t.strip()
        if t.startswith(("CuPtr{", "Ptr{")) or t == "Cstring":
            return "ptr"
        return jl[t]

    calls = re.findall(r"ccall\(\(:(swalbe_\w+), lib\),\s*(\w+),\s*\(([^)]*)\)", src, flags=re.S)
    assert len(calls) >= 20
    for name, ret, types in calls:
        assert name in protos, f"ccall to undeclared symbol {name}"
        got = [classify(t) for t in types.split(",") if t.strip()]
        assert got == protos[name], f"{name}: Julia {got} vs C {protos[name]}"
        assert ret in ("Cint", "Cstring", "Culonglong"), (name, ret)
    bound = {c[0] for c in calls}
    # every symbol of the header except the two self-tests must be reachable from Julia, the host language north_star
    # names -- the multi-GPU slab runtime (swalbe_dist_*) included
    missing = sorted(n for n in protos if n not in bound and not n.startswith("swalbe_selftest_"))
    assert not missing, f"header symbols without a ccall in SwalbeB200.jl: {missing}"
    # no type piracy: the glue adds methods to Swalbe's functions and its own, never to Base / CUDA functions
    assert not re.search(r"^\s*function\s+(?:Base|CUDA|GPUArrays)\.[\w!]+\(", src, flags=re.M)
    assert not re.search(r"^(?:Base|CUDA|GPUArrays)\.[\w!]+\([^\n]*\)\s*=", src, flags=re.M)
    # a finalizer may only be attached to a mutable object; CuState / CuState_thermal are immutable (src/initialize.jl:214)
    for m in re.finditer(r"finalizer\(([^,]+),\s*([^)]+)\)", src):
        assert "state" not in m.group(2), m.group(0)
    # the flag values of the header
    hdr = open(os.path.join(ROOT, "include", "swalbe_b200.h")).read()
    for name in ("LOOP_LAZY_POPULATIONS", "LOOP_SKIP_AUX", "LOOP_MOMENTS_CONSISTENT", "PRESSURE_POWER_BROAD", "PRESSURE_FAST"):
        c = int(re.search(r"#define SWALBE_%s (\d+)" % name, hdr).group(1))
        j = int(re.search(r"const %s = Cint\((\d+)\)" % name, src).group(1))
        assert c == j, name
    # block balance (function/struct/if/while/do/module ... end), comments and strings stripped
    body = "\n".join(re.sub(r"#.*$", "", re.sub(r'"(?:\\.|[^"\\])*"', '""', ln)) for ln in
                     re.sub(r'"""(?:.|\n)*?"""', '""', src).splitlines())
    openers = re.findall(r"(?<![\w!.])(?:function|struct|if|for|while|do|module|begin|let|try|macro|quote)(?![\w!])", body)
    ends = re.findall(r"(?<![\w!.:\[])end(?![\w!])", body)
    assert len(openers) == len(ends), (len(openers), len(ends))


def _header_struct_fields(name):
    """[(field, kind)] of `typedef struct <name> {...}` in declaration order (kind: ptr / f64 / i32 / u64)."""
    src = open(os.path.join(ROOT, "include", "swalbe_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), src, flags=re.S).group(1)
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        m = re.match(r"((?:const\s+)?(?:unsigned long long|double|int))\s+(.*)", decl)
        base, names = m.group(1), m.group(2)
        for nm in names.split(","):
            nm = nm.strip()
            kind = "ptr" if nm.startswith("*") else {"double": "f64", "int": "i32", "unsigned long long": "u64"}[base.replace("const ", "")]
            out.append((nm.lstrip("*").strip(), kind))
    return out


def test_struct_fields_match_in_ctypes_and_julia():
    """Field by field, in order: the C structs, their ctypes mirrors and the Julia structs of the glue."""
    from swalbe_b200 import _lib

    def ckind(t):
        return {C.c_double: "f64", C.c_int: "i32", C.c_ulonglong: "u64", C.c_void_p: "ptr"}[t]

    jsrc = open(os.path.join(ROOT, "swalbe.jl_b200", "julia", "SwalbeB200.jl"), encoding="utf-8").read()

    def julia_fields(name):
        body = re.search(r"struct %s\b.*?\n(.*?)\nend" % name, jsrc, flags=re.S).group(1)
        body = re.sub(r"#.*", "", body)
        out = []
        for f, t in re.findall(r"(\w+)::([\w{}]+)", body):
            out.append((f, "ptr" if t.startswith(("CuPtr{", "Ptr{")) else {"Cdouble": "f64", "Cint": "i32", "Culonglong": "u64"}[t]))
        return out

    for cname, cty, jname in (("swalbe_state", _lib.CState, "CState"), ("swalbe_params", _lib.CParams, "CParams"),
                              ("swalbe_loop_logs", _lib.CLogs, "CLogs")):
        want = _header_struct_fields(cname)
        assert [(n, ckind(t)) for n, t in cty._fields_] == want, cname
        assert julia_fields(jname) == want, jname
