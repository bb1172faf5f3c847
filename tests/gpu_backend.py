"""Backend for tests/golden_cases.py that routes every array-form operator through the C ABI of
libswalbe_b200.so (via the swalbe_b200 host mirror): NumPy in -> device -> kernel -> NumPy out."""
import numpy as np

import swalbe_b200 as sw


def _up(a):
    if a is None:
        return None
    shape = a.shape
    return sw.Field(shape[0], shape[1], shape[2] if a.ndim == 3 else None).set(a)


def _down(dst, f):
    dst[...] = f.numpy()


def equilibrium(feq, h, ux, uy, vsq, g):
    d = [_up(x) for x in (feq, h, ux, uy, vsq)]
    sw.equilibrium(*d, g)
    _down(feq, d[0]); _down(vsq, d[4])


def BGKandStream(fout, feq, ftemp, Fx, Fy, tau):
    d = [_up(x) for x in (fout, feq, ftemp, Fx, Fy)]
    sw.BGKandStream(*d, tau)
    _down(fout, d[0]); _down(ftemp, d[2])


def moments(h, ux, uy, f):
    d = [_up(x) for x in (h, ux, uy, f)]
    sw.moments(*d)
    _down(h, d[0]); _down(ux, d[1]); _down(uy, d[2])


class _CospiField(sw.Field):
    """a Field that already holds cospi.(θ) (the oracle's convention) -> bypass the host-side cospi"""


def filmpressure(output, f, dgrad, gamma, cospi_theta, n, m, hmin, hcrit, variant="fast"):
    import ctypes as C

    from swalbe_b200 import _lib

    out, fd, dg = _up(output), _up(f), _up(dgrad)
    if isinstance(cospi_theta, np.ndarray):
        ctf = _up(cospi_theta)
        ct, ctp = 0.0, ctf.ptr
    else:
        ct, ctp = float(cospi_theta), None
    _lib.call("swalbe_filmpressure", out.ptr, fd.ptr, dg.ptr, float(gamma), ct, ctp, int(n), int(m), float(hmin),
              float(hcrit), _lib.PRESSURE_FAST if variant == "fast" else _lib.PRESSURE_POWER_BROAD, f.shape[0], f.shape[1],
              sw._stream())
    _down(output, out)


def lap9(output, f, gamma):
    o, fd = _up(output), _up(f)
    sw.laplacianf(o, fd, gamma)
    _down(output, o)


def grad9(ox, oy, f, a=None, dgrad=None):
    dx, dy, fd, ad = _up(ox), _up(oy), _up(f), _up(a)
    if a is None:
        sw.gradf(dx, dy, fd)
    elif dgrad is None:
        sw.gradf(dx, dy, fd, ad)
    else:
        sw.gradf(dx, dy, fd, _up(dgrad), ad)
    _down(ox, dx); _down(oy, dy)


def hgradp(gx, gy, pressure, height, dgrad):
    st = sw.CuState(*pressure.shape)
    st.pressure.set(pressure); st.height.set(height)
    sw.hgradp(st)
    _down(gx, st.hgradpx); _down(gy, st.hgradpy)


def slippage(sx, sy, h, ux, uy, delta, mu, hcrit=0.0, variant=0):
    from swalbe_b200 import _lib

    d = [_up(x) for x in (sx, sy, h, ux, uy)]
    _lib.call("swalbe_slippage", *[x.ptr for x in d], float(delta), float(mu), float(hcrit), int(variant), h.shape[0],
              h.shape[1], sw._stream())
    _down(sx, d[0]); _down(sy, d[1])


def force_sum(Fx, Fy, gx, gy, sx, sy, kx=None, ky=None):
    from swalbe_b200 import _lib

    d = [_up(x) for x in (Fx, Fy, gx, gy, sx, sy, kx, ky)]
    _lib.call("swalbe_force_sum", *[x.ptr if x is not None else None for x in d], Fx.shape[0], Fx.shape[1], sw._stream())
    _down(Fx, d[0]); _down(Fy, d[1])


def inclination(Fx, Fy, h, alpha, factor):
    from swalbe_b200 import _lib

    d = [_up(x) for x in (Fx, Fy, h)]
    _lib.call("swalbe_inclination", *[x.ptr for x in d], float(alpha[0]), float(alpha[1]), float(factor), h.shape[0],
              h.shape[1], sw._stream())
    _down(Fx, d[0]); _down(Fy, d[1])
