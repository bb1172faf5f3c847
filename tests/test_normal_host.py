"""The table-driven Box-Muller of swalbe.jl_b200/csrc/normal.cuh compiled as plain C++ and checked on the host: every
output against long double, plus moments and tail fractions (tools/normal_host_test.cpp)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_normal_generator_numerics_on_host(tmp_path):
    exe = str(tmp_path / "normal_host_test")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "swalbe.jl_b200", "csrc"),
                    os.path.join(ROOT, "tools", "normal_host_test.cpp"), "-o", exe, "-lm"], check=True)
    out = subprocess.run([exe, "4000000"], check=True, capture_output=True, text=True).stdout
    num = r"([-+0-9.eE]+)"
    m = re.search(rf"max abs err -2lnu {num} \(rel {num}\)\s+cos {num}\s+sin {num}", out)
    e_ln, _, e_c, e_s = (float(v) for v in m.groups())
    assert e_ln < 1e-14 and e_c < 1e-15 and e_s < 1e-15, out          # (-2 ln u reaches 62: 1e-14 is ~1 ulp there)
    m = re.search(rf"mean {num} var {num} skew {num} kurt {num} m6 {num} corr {num}", out)
    mean, var, skew, kurt, m6, corr = (float(v) for v in m.groups())
    n = 8e6
    assert abs(mean) < 4 / n ** 0.5 and abs(var - 1) < 4 * (2 / n) ** 0.5 and abs(skew) < 4 * (15 / n) ** 0.5
    assert abs(kurt - 3) < 4 * (96 / n) ** 0.5 and abs(m6 - 15) < 4 * (10170 / n) ** 0.5 and abs(corr) < 4 / (n / 2) ** 0.5
    m = re.search(rf">3 {num} .*>4 {num} .*>5 {num}", out)
    t3, t4, _ = (float(v) for v in m.groups())
    assert abs(t3 - 2.6998e-3) < 5 * (2.6998e-3 / n) ** 0.5 and abs(t4 - 6.334e-5) < 5 * (6.334e-5 / n) ** 0.5
