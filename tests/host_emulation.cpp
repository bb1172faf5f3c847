// CPU build of the DEVICE arithmetic (swalbe.jl_b200/csrc/common.cuh) for tests/test_host_emulation.py:
//   g++ -O2 -ffp-contract=off -DSW_HOST_EMULATION -I/usr/local/cuda/include -shared -fPIC tests/host_emulation.cpp
// The site functions the kernels are made of -- film pressure, gradients, slip, equilibrium, collision, moments, the
// exact-division helper (with a host reciprocal in place of MUFU.RCP64H), the Philox round function and the
// table-driven normals -- are compiled as plain C++ and composed into one time step exactly as the tile kernel
// (csrc/tile.cu) composes them, so the CPU suite can compare the kernels' arithmetic SOURCE with the oracle bit for bit
// without a GPU.  Test infrastructure only; nothing in the product links this.
#include <math.h>
#include <stdint.h>
#include <string.h>

// host stand-ins for the CUDA intrinsics common.cuh uses (declared before the header is parsed)
static inline int __double2hiint(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(u >> 32); }
static inline double __hiloint2double(int hi, int lo) {
  const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double d; memcpy(&d, &u, 8); return d;
}
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline void sincospi(double a, double *s, double *c) { *s = sin(a * 3.141592653589793); *c = cos(a * 3.141592653589793); }

#define __noinline__ __attribute__((noinline))
#include "../swalbe.jl_b200/csrc/common.cuh"

using namespace swalbe;

namespace swalbe {  // the two host functions of the library that common.cuh only declares
int set_error(int code, const char *, ...) { return code; }
void count_launch(unsigned) {}
}  // namespace swalbe

static inline size_t at(int i, int j, int Lx, int Ly) { return (size_t)wrapi(j, Ly) * Lx + wrapi(i, Lx); }

extern "C" {

// one time step (tau == 1 or general tau) over the whole periodic lattice, composed like csrc/tile.cu
int emul_step(double *h, double *ux, double *uy, double *fout, const double *ftemp, double *pressure, int Lx, int Ly,
              double tau, double mu, double delta, double gamma, double hmin, double hcrit, double g, int n, int m,
              double cospi_theta, const double *ct_field, int pressure_variant, int slip_variant, double *scratch /* 10*N */,
              double kbt, unsigned long long seed, unsigned long long step) {
  const size_t N = (size_t)Lx * Ly;
  PressureConsts pc;
  if (int e = resolve_pmode(pressure_variant, n, m, &pc.pmode)) return e;
  pc.gamma = gamma; pc.kappa = host_kappa(cospi_theta, n, m, hmin);
  pc.nm1 = (double)(n - 1); pc.mm1 = (double)(m - 1); pc.kden = (double)(n - m) * hmin;
  pc.hmin = hmin; pc.hcrit = hcrit; pc.n = n; pc.m = m;
  const SlipConsts sc = make_slip(delta, mu, hcrit, slip_variant);
  const EqConsts ec = make_eq(g);
  volatile double it = 1.0 / tau;
  volatile double om = 1.0 - it;
  double *p = scratch, *fs = scratch + N;  // pressure, 9 post-collision planes
  const ThermalConsts tc = make_thermal(kbt, mu, delta);
  const PhiloxKey K = make_philox_key(seed);
  for (int j = 0; j < Ly; ++j)
    for (int i = 0; i < Lx; ++i) {
      const double hc = h[at(i, j, Lx, Ly)];
      const double lap = lap9_bracket(hc, h[at(i - 1, j, Lx, Ly)], h[at(i, j - 1, Lx, Ly)], h[at(i + 1, j, Lx, Ly)],
                                      h[at(i, j + 1, Lx, Ly)], h[at(i - 1, j - 1, Lx, Ly)], h[at(i + 1, j - 1, Lx, Ly)],
                                      h[at(i + 1, j + 1, Lx, Ly)], h[at(i - 1, j + 1, Lx, Ly)]);
      const double kappa = ct_field ? kappa_from_field(ct_field[at(i, j, Lx, Ly)], pc) : pc.kappa;
      p[at(i, j, Lx, Ly)] = film_pressure(hc, lap, kappa, pc);
    }
  for (int j = 0; j < Ly; ++j)
    for (int i = 0; i < Lx; ++i) {
      const size_t c = at(i, j, Lx, Ly);
      const double hc = h[c];
      const double pipjp = p[at(i - 1, j - 1, Lx, Ly)], pimjp = p[at(i + 1, j - 1, Lx, Ly)],
                   pimjm = p[at(i + 1, j + 1, Lx, Ly)], pipjm = p[at(i - 1, j + 1, Lx, Ly)];
      const double gx = grad9_x(p[at(i - 1, j, Lx, Ly)], p[at(i + 1, j, Lx, Ly)], pipjp, pimjp, pimjm, pipjm);
      const double gy = grad9_y(p[at(i, j - 1, Lx, Ly)], p[at(i, j + 1, Lx, Ly)], pipjp, pimjp, pimjm, pipjm);
      const double hgx = hc * gx, hgy = hc * gy;
      double sx, sy;
      slip_terms(hc, ux[c], uy[c], sc, slip_variant, sx, sy);
      double Fx = (-hgx) - sx, Fy = (-hgy) - sy;
      if (kbt > 0.0) {  // thermal!: in-kernel noise keyed on (seed, step, global cell)
        double kx, ky;
        thermal_pair(hc, tc, K, step, Lx, (long long)j, i, kx, ky);
        Fx = Fx - kx;
        Fy = Fy - ky;
      }
      double fe[9], vsq, f[9];
      equilibrium_site<false>(hc, ux[c], uy[c], ec, fe, vsq);
      if (tau == 1.0) collide_site_tau1(fe, Fx, Fy, f);
      else {
        double ft[9];
        for (int k = 0; k < 9; ++k) ft[k] = ftemp[c + k * N];
        collide_site(ft, fe, Fx, Fy, om, it, f);
      }
      for (int k = 0; k < 9; ++k) fs[c + k * N] = f[k];
    }
  memcpy(pressure, p, N * sizeof(double));
  static const int cx[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1}, cy[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
  for (int j = 0; j < Ly; ++j)
    for (int i = 0; i < Lx; ++i) {
      double fn[9];
      for (int k = 0; k < 9; ++k) fn[k] = fs[at(i - cx[k], j - cy[k], Lx, Ly) + k * N];
      const size_t c = at(i, j, Lx, Ly);
      double hn, uxn, uyn;
      moments_site(fn, hn, uxn, uyn);
      h[c] = hn; ux[c] = uxn; uy[c] = uyn;
      for (int k = 0; k < 9; ++k) fout[c + k * N] = fn[k];
    }
  return 0;
}

// the exact-division helper against the compiler's `/` on n operand pairs; returns the number of bitwise mismatches
long emul_division_mismatches(const double *a, const double *b, long n) {
  long bad = 0;
  for (long i = 0; i < n; ++i) {
    const double q = div_exact(a[i], b[i]), want = a[i] / b[i];
    double q1, q2;
    div2_exact(a[i], b[i] * 0.5, b[i], q1, q2);
    const double w2 = (b[i] * 0.5) / b[i];
    bad += memcmp(&q, &want, 8) != 0 && !(q != q && want != want);
    bad += memcmp(&q1, &want, 8) != 0 && !(q1 != q1 && want != want);
    bad += memcmp(&q2, &w2, 8) != 0 && !(q2 != q2 && w2 != w2);
  }
  return bad;
}

// Philox4x32-10 through the keyed form the kernels use
void emul_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  const PhiloxKey K = make_philox_key((unsigned long long)key[0] | ((unsigned long long)key[1] << 32));
  philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], K, out);
}

// thermal_pair (single-precision Box-Muller on shared Philox blocks x amplitude) for the cells of an Lx-wide lattice:
// cell c = i + Lx * j for c in [0, n)
void emul_thermal(double *kx, double *ky, const double *h, long n, int Lx, double kbt, double mu, double delta,
                  unsigned long long seed, unsigned long long step) {
  const ThermalConsts tc = make_thermal(kbt, mu, delta);
  const PhiloxKey K = make_philox_key(seed);
  for (long c = 0; c < n; ++c) thermal_pair(h[c], tc, K, step, Lx, (long long)(c / Lx), (int)(c % Lx), kx[c], ky[c]);
}

}  // extern "C"
