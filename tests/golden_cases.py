"""Known-answer vectors transcribed from the reference's own tests and doctests.

Every case takes a backend ``B`` exposing the array-form operators on NumPy Fortran-order arrays with
Julia shapes (``oracle.oracle_np`` itself, or ``tests/gpu_backend.py`` which routes each call through
the C ABI of libswalbe_b200.so).  The asserts are the reference's asserts, line for line; file:line
citations are into /root/reference.
"""
import math

import numpy as np

from oracle import oracle_np as onp

zeros, ones = onp.zeros, onp.ones


def fill(v, *shape):
    return np.full(shape, v, dtype=np.float64, order="F")


def jl_matrix(rows):
    """A Julia matrix literal [a b; c d] -> F-order array indexed [i,j] like Julia (0-based)."""
    return np.asfortranarray(np.array(rows, dtype=np.float64))


def ramp(Lx=5, Ly=5):
    """reshape(collect(1.0:Lx*Ly), Lx, Ly)"""
    return np.arange(1.0, Lx * Ly + 1.0).reshape((Lx, Ly), order="F")


SHIFTS = [(0, 0), (1, 0), (0, 1), (-1, 0), (0, -1), (1, 1), (-1, 1), (-1, -1), (1, -1)]

# ------------------------------------------------------------------------------------------------
# test/collide.jl:1-123


def _set_dist(feq, ftemp, fout, ft=1.0):
    feq[...] = 1.0
    ftemp[...] = ft
    fout[...] = 1.0
    feq[0, 0, :] = 2.0


def case_collide_tau1_noforce(B):  # test/collide.jl:27-43
    feq, ftemp, fout = ones(5, 5, 9), ones(5, 5, 9), ones(5, 5, 9)
    feq[0, 0, :] = 2.0
    B.BGKandStream(fout, feq, ftemp, zeros(5, 5), zeros(5, 5), 1.0)
    for k in range(9):
        assert np.all(fout[:, :, k] == onp.circshift(feq[:, :, 0], SHIFTS[k]))
    assert np.all(ftemp == fout)  # src/collide.jl:103


def case_collide_tau1_force(B):  # test/collide.jl:44-71
    for ft in (1.0, 1.2):  # array form uses ftemp=1, the state form 1.2 -- irrelevant at tau=1
        feq, ftemp, fout = ones(5, 5, 9), ones(5, 5, 9), ones(5, 5, 9)
        _set_dist(feq, ftemp, fout, ft)
        feq[...] = 1.0
        B.BGKandStream(fout, feq, ftemp, fill(0.1, 5, 5), fill(-0.1, 5, 5), 1.0)
        e = feq[:, :, 0]
        expect = [e, e + 1 / 30, e - 1 / 30, e - 1 / 30, e + 1 / 30, e, e - 1 / 24 * 0.2, e, e + 1 / 24 * 0.2]
        for k in range(9):
            assert np.all(fout[:, :, k] == onp.circshift(expect[k], SHIFTS[k])), k


def case_collide_tau075_noforce(B):  # test/collide.jl:72-97
    onebytau = 1.0 / 0.75
    omega = 1.0 - 1.0 / 0.75
    feq, ftemp, fout = ones(5, 5, 9), ones(5, 5, 9), ones(5, 5, 9)
    _set_dist(feq, ftemp, fout)
    B.BGKandStream(fout, feq, ftemp, zeros(5, 5), zeros(5, 5), 0.75)
    for k in range(9):
        assert np.all(fout[:, :, k] == onp.circshift(omega * 1.0 + onebytau * feq[:, :, 0], SHIFTS[k])), k


def case_collide_tau075_force(B):  # test/collide.jl:99-123
    onebytau = 1.0 / 0.75
    omega = 1.0 - 1.0 / 0.75
    feq, ftemp, fout = ones(5, 5, 9), ones(5, 5, 9), ones(5, 5, 9)
    _set_dist(feq, ftemp, fout)
    B.BGKandStream(fout, feq, ftemp, fill(0.1, 5, 5), fill(-0.1, 5, 5), 0.75)
    c = omega * 1.0 + onebytau * feq[:, :, 0]
    expect = [c, c + 1 / 30, c - 1 / 30, c - 1 / 30, c + 1 / 30, c, c - 1 / 24 * 0.2, c, c + 1 / 24 * 0.2]
    for k in range(9):
        assert np.all(fout[:, :, k] == onp.circshift(expect[k], SHIFTS[k])), k


def case_collide_doctest(B):  # src/collide.jl:33-60
    feq, ftemp, fout = ones(5, 5, 9), zeros(5, 5, 9), zeros(5, 5, 9)
    feq[0, 0, :] = 2.0
    B.BGKandStream(fout, feq, ftemp, zeros(5, 5), zeros(5, 5), 1.0)
    want = ones(5, 5)
    want[1, 1] = 2.0
    assert np.all(fout[:, :, 5] == want)


# ------------------------------------------------------------------------------------------------
# test/equilibrium.jl:1-85


def case_equilibrium_nothing(B):  # :6-16
    feq = ones(5, 5, 9)
    B.equilibrium(feq, zeros(5, 5), zeros(5, 5), zeros(5, 5), zeros(5, 5), 0.0)
    assert np.all(feq == 0.0)


def case_equilibrium_density(B):  # :18-28
    feq = zeros(5, 5, 9)
    B.equilibrium(feq, ones(5, 5), zeros(5, 5), zeros(5, 5), zeros(5, 5), 0.0)
    assert np.all(feq[:, :, 0] == 1.0)
    assert np.all(feq[:, :, 1:9] == 0.0)


def case_equilibrium_gravity(B):  # :30-42
    feq = zeros(5, 5, 9)
    B.equilibrium(feq, ones(5, 5), zeros(5, 5), zeros(5, 5), zeros(5, 5), 0.1)
    assert np.all(feq[:, :, 0] == 1.0 - 1 / 12)
    assert np.allclose(feq[:, :, 1:5], 1 / 9 * 1.5 * 0.1, rtol=1e-8, atol=0)  # Julia ≈ (rtol sqrt(eps))
    assert np.allclose(feq[:, :, 5:9], 1 / 36 * 1.5 * 0.1, rtol=1e-8, atol=0)


def _equilibrium_velocity_expect(g):
    g0 = 1.5 * g
    return [
        None,
        1 / 9 * (g0 + 3 * 0.1 + 4.5 * 0.01 - 3 / 2 * 0.02),
        1 / 9 * (g0 + 3 * -0.1 + 4.5 * 0.01 - 3 / 2 * 0.02),
        1 / 9 * (g0 + 3 * -0.1 + 4.5 * 0.01 - 3 / 2 * 0.02),
        1 / 9 * (g0 + 3 * 0.1 + 4.5 * 0.01 - 3 / 2 * 0.02),
        1 / 36 * (g0 + -3 / 2 * 0.02),
        1 / 36 * (g0 + 3 * -0.2 + 4.5 * 0.2 ** 2 - 3 / 2 * 0.02),
        1 / 36 * (g0 + -3 / 2 * 0.02),
        1 / 36 * (g0 + 3 * 0.2 + 4.5 * 0.2 ** 2 - 3 / 2 * 0.02),
    ]


def case_equilibrium_velocity(B):  # :44-66
    feq = zeros(5, 5, 9)
    vsq = zeros(5, 5)
    B.equilibrium(feq, ones(5, 5), fill(0.1, 5, 5), fill(-0.1, 5, 5), vsq, 0.0)
    assert np.all(feq[:, :, 0] == 1.0 - 2 / 3 * 0.02)
    exp = _equilibrium_velocity_expect(0.0)
    for k in range(1, 9):
        assert np.allclose(feq[:, :, k], exp[k], rtol=1.5e-8, atol=0), k
    assert np.all(vsq == 0.1 * 0.1 + (-0.1) * (-0.1))  # src/equilibrium.jl:71


def case_equilibrium_gravity_velocity(B):  # :68-83
    feq = zeros(5, 5, 9)
    B.equilibrium(feq, ones(5, 5), fill(0.1, 5, 5), fill(-0.1, 5, 5), zeros(5, 5), 0.1)
    assert np.all(feq[:, :, 0] == 1.0 - 1 / 12 - 2 / 3 * 0.02)
    exp = _equilibrium_velocity_expect(0.1)
    for k in range(1, 9):
        assert np.allclose(feq[:, :, k], exp[k], rtol=1.5e-8, atol=0), k


def case_equilibrium_doctest(B):  # src/equilibrium.jl:35-54
    feq = zeros(5, 5, 9)
    B.equilibrium(feq, ones(5, 5), fill(0.1, 5, 5), zeros(5, 5), zeros(5, 5), 0.1)
    assert np.allclose(feq[:, :, 0], 0.91, rtol=0, atol=5e-17 + 1e-15)  # printed as 0.91
    B.equilibrium(feq, ones(5, 5), fill(0.1, 5, 5), zeros(5, 5), zeros(5, 5), 0.0)
    assert np.allclose(feq[:, :, 0], 1 - 2 / 3 * 0.01, rtol=1.5e-8, atol=0)


# ------------------------------------------------------------------------------------------------
# test/moments.jl:14-134 (the 2-D assertions; planes are set cumulatively exactly as upstream)


def case_moments_sequence(B):
    f = zeros(5, 5, 9)
    h, ux, uy = zeros(5, 5), zeros(5, 5), zeros(5, 5)
    f[:, :, 0] = 1.0
    B.moments(h, ux, uy, f)
    assert np.all(h == 1.0)  # :14-33
    f[:, :, 1] = 0.1
    B.moments(h, ux, uy, f)
    assert np.all(h == 1.1) and np.all(ux == 0.1 / 1.1) and np.all(uy == 0)  # :35-62
    f[:, :, 2] = 0.2
    B.moments(h, ux, uy, f)
    assert np.all(h == 1.3) and np.all(ux == 0.1 / 1.3) and np.all(uy == 0.2 / 1.3)  # :66-85
    f[:, :, 3] = -0.2
    B.moments(h, ux, uy, f)
    assert np.all(h == 1.1)  # :90-103
    assert np.allclose(ux, 0.3 / 1.1, rtol=0, atol=1e-10) and np.allclose(uy, 0.2 / 1.1, rtol=0, atol=1e-10)
    f[:, :, 4] = -0.1
    B.moments(h, ux, uy, f)
    assert np.all(h == 1.0)  # :104-117
    assert np.allclose(ux, 0.3, rtol=0, atol=1e-10) and np.allclose(uy, 0.3, rtol=0, atol=1e-10)
    f[:, :, 5] = 0.1
    B.moments(h, ux, uy, f)
    assert np.all(h == 1.1)  # :118-131
    assert np.allclose(ux, 0.4 / 1.1, rtol=0, atol=1e-10) and np.allclose(uy, 0.4 / 1.1, rtol=0, atol=1e-10)


# ------------------------------------------------------------------------------------------------
# test/pressure.jl:1-54,139-161 ; doctest src/pressure.jl:39-63

SOL_LAP = jl_matrix([
    [-30.0, -5.0, -5.0, -5.0, 20],
    [-25.0, 0.0, 0.0, 0.0, 25.0],
    [-25.0, 0.0, 0.0, 0.0, 25.0],
    [-25.0, 0.0, 0.0, 0.0, 25.0],
    [-20.0, 5.0, 5.0, 5.0, 30.0],
])


def case_pressure_no_contact_angle(B):  # :19-35  (theta=0 -> cospi = 1)
    f = ramp()
    for variant in ("fast", "power_broad"):
        res = zeros(5, 5)
        B.filmpressure(res, f, zeros(5, 5, 8), 1.0, onp.cospi(0.0), 3, 2, 0.1, 0.1, variant=variant)
        assert np.allclose(res, SOL_LAP, rtol=0, atol=1e-10), variant


def case_pressure_gradient_and_contact_angle(B):  # :36-44  (theta=1/2 -> cospi = 0)
    f = ramp()
    want = -1 * (-SOL_LAP + 20 * ((0.1 / f) ** 3 - (0.1 / f) ** 2))
    for variant in ("fast", "power_broad"):
        res = zeros(5, 5)
        B.filmpressure(res, f, zeros(5, 5, 8), 1.0, onp.cospi(1 / 2), 3, 2, 0.1, 0.0, variant=variant)
        assert np.allclose(res, want, rtol=0, atol=1e-10), variant


def case_pressure_no_height_gradient(B):  # :45-53
    for variant in ("fast", "power_broad"):
        res = zeros(5, 5)
        B.filmpressure(res, ones(5, 5), zeros(5, 5, 8), 1.0, onp.cospi(1 / 2), 3, 2, 0.1, 0.0, variant=variant)
        assert np.allclose(res, -2 * (0.1 ** 2 - 0.1), rtol=0, atol=1e-10), variant


def case_pressure_doctest(B):  # src/pressure.jl:39-63
    h = ramp()
    p = zeros(5, 5)
    B.filmpressure(p, h, zeros(5, 5, 8), 0.01, onp.cospi(0.0), 3, 2, 0.1, 0.05, variant="fast")
    result = jl_matrix([
        [30.0, 5.0, 5.0, 5.0, -20],
        [25.0, 0.0, 0.0, 0.0, -25.0],
        [25.0, 0.0, 0.0, 0.0, -25.0],
        [25.0, 0.0, 0.0, 0.0, -25.0],
        [20.0, -5.0, -5.0, -5.0, -30.0],
    ])
    assert np.allclose(result, -100 * p, rtol=0, atol=1e-12)


def case_pressure_domain_error(B):  # src/pressure.jl:101-107
    import pytest

    with pytest.raises(ValueError):
        B.filmpressure(zeros(5, 5), ramp(), zeros(5, 5, 8), 1.0, 1.0, 4, 2, 0.1, 0.1, variant="fast")


def case_power_broad():  # test/pressure.jl:139-161 (host helper; oracle only)
    assert onp.power_broad(2, 3) == 8 and onp.power_broad(2, 6) == 64
    assert onp.power_broad(5.0, 3) == 125.0 and onp.power_broad(5.0, 6) == 15625.0
    assert np.all(onp.power_broad(np.array([2.0, 3.0, 4.0]), 3) == [8.0, 27.0, 64.0])
    assert np.all(onp.power_broad(np.array([2.0, 3.0, 4.0]), 6) == [64.0, 729.0, 4096.0])
    assert onp.power_broad(3, 3) == 27
    assert np.all(onp.power_broad(np.array([2.0, 5.0, 6.0]), 2) == [4.0, 25.0, 36.0])


# ------------------------------------------------------------------------------------------------
# test/differences.jl:1-76,104-124 ; test/forcing.jl:104-135

SOLX = jl_matrix([
    [-1.5] * 5,
    [1.0] * 5,
    [1.0] * 5,
    [1.0] * 5,
    [-1.5] * 5,
])
SOLY = jl_matrix([[-7.5, 5.0, 5.0, 5.0, -7.5]] * 5)


def case_grad_simple(B):  # test/differences.jl:7-31
    ox, oy = zeros(5, 5), zeros(5, 5)
    B.grad9(ox, oy, ramp())
    assert np.all(ox == SOLX) and np.all(oy == SOLY)


def case_grad_four_five_args(B):  # test/differences.jl:32-75
    a = fill(0.1, 5, 5)
    for dgrad in (None, zeros(5, 5, 8)):
        ox, oy = zeros(5, 5), zeros(5, 5)
        B.grad9(ox, oy, ramp(), a=a, dgrad=dgrad)
        assert np.all(ox == 0.1 * SOLX) and np.all(oy == 0.1 * SOLY)


def case_laplacian(B):  # test/differences.jl:104-124
    out = zeros(5, 5)
    B.lap9(out, ramp(), -1.0)
    assert np.allclose(out, SOL_LAP, rtol=0, atol=1e-10)


def case_hgradp(B):  # test/forcing.jl:104-135 (height = 1 from Sys)
    gx, gy = zeros(5, 5), zeros(5, 5)
    B.hgradp(gx, gy, ramp(), ones(5, 5), zeros(5, 5, 8))
    assert np.all(gx == SOLX) and np.all(gy == SOLY)


# ------------------------------------------------------------------------------------------------
# test/forcing.jl:13-69, 141-184


def case_slippage(B):
    fx, fy = ones(5, 5), ones(5, 5)
    B.slippage(fx, fy, ones(5, 5), zeros(5, 5), zeros(5, 5), 1.0, 1 / 6)  # :15-21
    assert np.all(fx == 0.0) and np.all(fy == 0.0)
    B.slippage(fx, fy, ones(5, 5), fill(0.1, 5, 5), zeros(5, 5), 1.0, 1 / 6)  # :24-32
    assert np.allclose(fx, 0.1 / 11, rtol=0, atol=1e-10) and np.all(fy == 0.0)
    B.slippage(fx, fy, ones(5, 5), zeros(5, 5), fill(0.1, 5, 5), 1.0, 1 / 6)  # :35-45
    assert np.allclose(fy, 0.1 / 11, rtol=0, atol=1e-10) and np.all(fx == 0.0)
    B.slippage(fx, fy, ones(5, 5), fill(-0.1, 5, 5), fill(0.1, 5, 5), 1.0, 1 / 6)  # :48-58
    assert np.allclose(fx, -0.1 / 11, rtol=0, atol=1e-10) and np.allclose(fy, 0.1 / 11, rtol=0, atol=1e-10)
    B.slippage(fx, fy, ones(5, 5), fill(-0.1, 5, 5), fill(0.1, 5, 5), 0.0, 1 / 6)  # :61-68 (no slip)
    assert np.allclose(fx, -0.1 / 2, rtol=0, atol=1e-10) and np.allclose(fy, 0.1 / 2, rtol=0, atol=1e-10)


def case_inclination(B):  # test/forcing.jl:166-184
    sols = {0: 0.05, 1: 0.1 * (0.5 + 0.5 * math.tanh(1.0))}
    for t in (0, 1):
        Fx, Fy = zeros(5, 5), zeros(5, 5)
        factor = 0.5 + 0.5 * math.tanh((t - 0) / 1)
        B.inclination(Fx, Fy, ones(5, 5), [0.1, 0.1], factor)
        assert np.all(Fx == sols[t]) and np.all(Fy == sols[t])


def case_thermal_statistics(B, draw):  # test/forcing.jl:141-164
    """``draw(shape)`` -> (kx, ky) filled by the backend's thermal! for h=1 on a 50x50 grid."""
    for kb in (0.01, 0.1):
        vartest = 2 * kb / 11
        kx, ky = draw(kb)
        for a in (kx, ky):
            assert abs(a.mean()) < 1e-2
            assert abs(a.var(ddof=1) - vartest) < vartest / 10


ALL_OPERATOR_CASES = [
    case_collide_tau1_noforce,
    case_collide_tau1_force,
    case_collide_tau075_noforce,
    case_collide_tau075_force,
    case_collide_doctest,
    case_equilibrium_nothing,
    case_equilibrium_density,
    case_equilibrium_gravity,
    case_equilibrium_velocity,
    case_equilibrium_gravity_velocity,
    case_equilibrium_doctest,
    case_moments_sequence,
    case_pressure_no_contact_angle,
    case_pressure_gradient_and_contact_angle,
    case_pressure_no_height_gradient,
    case_pressure_doctest,
    case_pressure_domain_error,
    case_grad_simple,
    case_grad_four_five_args,
    case_laplacian,
    case_hgradp,
    case_slippage,
    case_inclination,
]
