/* C restatement of Swalbe.jl's 2-D (D2Q9) thin-film LBM step -- TEST INFRASTRUCTURE ONLY.
 *
 * Oracle / CPU baseline for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline and
 * `--impl reference` legs.  It is never linked into, loaded by, or called from libswalbe_b200.so or
 * the host package.  Pinning status: identical to oracle/oracle_np.py (see its header) -- pinned
 * against the reference's own known-answer tests, NOT against a live Julia run (Julia is absent).
 * tests/test_oracle_golden.py additionally checks this file bit-for-bit against oracle_np.py.
 *
 * Build (oracle/Makefile): gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC.
 * -ffp-contract=off is essential: Julia does not contract a*b+c into an FMA on the CPU.
 *
 * Layout: Julia column-major.  A[i,j,k] (0-based) lives at i + Lx*(j + Ly*k); x (=i) is contiguous.
 * The pass structure follows the reference exactly (one sweep per broadcast line, eight circshift!
 * temporaries in dgrad, fout .= ftemp copy) so that single-thread timings model the Julia CPU path;
 * `threads` > 1 parallelises each sweep over j with OpenMP (the arithmetic per cell is unchanged).
 */
#include <math.h>
#include <stddef.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef long long i64;

#define IDX(i, j) ((size_t)(i) + (size_t)Lx * (size_t)(j))
#define PARFOR _Pragma("omp parallel for schedule(static) num_threads(threads)")

static inline int wrap(int a, int n) { a %= n; return a < 0 ? a + n : a; }

/* Base.circshift!(dest, src, (sx,sy)): dest[i,j] = src[i-sx, j-sy], periodic. */
static void circshift2(double *dest, const double *src, int sx, int sy, int Lx, int Ly, int threads) {
  PARFOR
  for (int j = 0; j < Ly; ++j) {
    int js = wrap(j - sy, Ly);
    for (int i = 0; i < Lx; ++i) dest[IDX(i, j)] = src[IDX(wrap(i - sx, Lx), js)];
  }
}

/* the eight shifts of src/pressure.jl:131-139 == src/forcing.jl:171-179 == src/differences.jl:191-199 */
static void shift8(double *dgrad, const double *f, int Lx, int Ly, int threads) {
  size_t N = (size_t)Lx * Ly;
  static const int s[8][2] = {{1, 0}, {0, 1}, {-1, 0}, {0, -1}, {1, 1}, {-1, 1}, {-1, -1}, {1, -1}};
  for (int k = 0; k < 8; ++k) circshift2(dgrad + k * N, f, s[k][0], s[k][1], Lx, Ly, threads);
}

/* src/pressure.jl:363-369 */
static inline double power_broad(double arg, int n) {
  double temp = 1.0;
  for (int i = 0; i < n; ++i) temp *= arg;
  return temp;
}
static inline double power_3(double a) { return a * a * a; }
static inline double power_2(double a) { return a * a; }
static inline double fast_93(double a) { double t = power_3(a); return power_3(t) - t; }
static inline double fast_32(double a) { return power_3(a) - power_2(a); }

/* filmpressure!: variant 0 = state form (power_broad) src/pressure.jl:119-155,
 *                variant 1 = array form (fast_93 / fast_32) src/pressure.jl:72-115.
 * cospi_theta_field may be NULL (then the scalar cospi_theta is used).  Returns 1 on DomainError. */
int oracle_filmpressure(double *out, const double *f, double *dgrad, double gamma, double cospi_theta,
                        const double *cospi_theta_field, int n, int m, double hmin, double hcrit, int variant,
                        int Lx, int Ly, int threads) {
  size_t N = (size_t)Lx * Ly;
  int mode;
  if (variant == 1) {
    if (n == 9 && m == 3) mode = 93;
    else if (n == 3 && m == 2) mode = 32;
    else return 1;
  } else mode = 0;
  shift8(dgrad, f, Lx, Ly, threads);
  const double *hip = dgrad, *hjp = dgrad + N, *him = dgrad + 2 * N, *hjm = dgrad + 3 * N, *hipjp = dgrad + 4 * N,
               *himjp = dgrad + 5 * N, *himjm = dgrad + 6 * N, *hipjm = dgrad + 7 * N;
  const double nm1 = (double)(n - 1), mm1 = (double)(m - 1), den = (double)(n - m) * hmin;
  PARFOR
  for (i64 c = 0; c < (i64)N; ++c) {
    double x = hmin / (f[c] + hcrit);
    double pw = mode == 93 ? fast_93(x) : mode == 32 ? fast_32(x) : power_broad(x, n) - power_broad(x, m);
    double ct = cospi_theta_field ? cospi_theta_field[c] : cospi_theta;
    out[c] = -gamma * ((1 - ct) * nm1 * mm1 / den * pw);
  }
  PARFOR
  for (i64 c = 0; c < (i64)N; ++c) {
    out[c] = out[c] - gamma * (2.0 / 3.0 * (hjp[c] + hip[c] + him[c] + hjm[c]) +
                               1.0 / 6.0 * (hipjp[c] + himjp[c] + himjm[c] + hipjm[c]) - 10.0 / 3.0 * f[c]);
  }
  return 0;
}

/* ∇²f!  src/differences.jl:57-75 (dgrad: caller scratch standing in for the allocating circshift) */
void oracle_lap9(double *out, const double *f, double *dgrad, double gamma, int Lx, int Ly, int threads) {
  size_t N = (size_t)Lx * Ly;
  shift8(dgrad, f, Lx, Ly, threads);
  const double *hip = dgrad, *hjp = dgrad + N, *him = dgrad + 2 * N, *hjm = dgrad + 3 * N, *hipjp = dgrad + 4 * N,
               *himjp = dgrad + 5 * N, *himjm = dgrad + 6 * N, *hipjm = dgrad + 7 * N;
  PARFOR
  for (i64 c = 0; c < (i64)N; ++c)
    out[c] = gamma * (2.0 / 3.0 * (hjp[c] + hip[c] + him[c] + hjm[c]) +
                      1.0 / 6.0 * (hipjp[c] + himjp[c] + himjm[c] + hipjm[c]) - 10.0 / 3.0 * f[c]);
}

/* ∇f! (a == NULL: 3-arg form src/differences.jl:153-169; else 4/5-arg :171-206) and h∇p! src/forcing.jl:168-187 */
void oracle_grad9(double *ox, double *oy, const double *f, double *dgrad, const double *a, int Lx, int Ly,
                  int threads) {
  size_t N = (size_t)Lx * Ly;
  shift8(dgrad, f, Lx, Ly, threads);
  const double *fip = dgrad, *fjp = dgrad + N, *fim = dgrad + 2 * N, *fjm = dgrad + 3 * N, *fipjp = dgrad + 4 * N,
               *fimjp = dgrad + 5 * N, *fimjm = dgrad + 6 * N, *fipjm = dgrad + 7 * N;
  PARFOR
  for (i64 c = 0; c < (i64)N; ++c) {
    double gx = -1.0 / 3.0 * (fip[c] - fim[c]) - 1.0 / 12.0 * (fipjp[c] - fimjp[c] - fimjm[c] + fipjm[c]);
    ox[c] = a ? a[c] * gx : gx;
  }
  PARFOR
  for (i64 c = 0; c < (i64)N; ++c) {
    double gy = -1.0 / 3.0 * (fjp[c] - fjm[c]) - 1.0 / 12.0 * (fipjp[c] + fimjp[c] - fimjm[c] - fipjm[c]);
    oy[c] = a ? a[c] * gy : gy;
  }
}

/* slippage! (0) src/forcing.jl:42-46, slippage2! (1) :85-99, slippage_ring_riv! (2) :107-111 */
void oracle_slippage(double *sx, double *sy, const double *h, const double *ux, const double *uy, double delta,
                     double mu, double hcrit, int variant, int Lx, int Ly, int threads) {
  i64 N = (i64)Lx * Ly;
  for (int comp = 0; comp < 2; ++comp) { /* two broadcast lines -> two sweeps */
    double *s = comp ? sy : sx;
    const double *u = comp ? uy : ux;
    PARFOR
    for (i64 c = 0; c < N; ++c) {
      double hh = h[c];
      if (variant == 0) s[c] = (6 * mu * hh * u[c]) / (2 * (hh * hh) + 6 * delta * hh + 3 * (delta * delta));
      else if (variant == 1) {
        double hc = hh + hcrit;
        s[c] = (6 * mu * hc * u[c]) / (2 * (hc * hc) + 6 * delta * hc + 3 * (delta * delta));
      } else s[c] = (6 * mu * hh * u[c]) / (2 * (hh * hh) + 6 * delta * (hh + hcrit));
    }
  }
}

/* deterministic part of thermal!  src/forcing.jl:297-311: k = normal * sqrt(2 kbt mu 6 h / (2hh + 6hδ + 3δδ)) */
void oracle_thermal(double *kx, double *ky, const double *h, double kbt, double mu, double delta, const double *nx,
                    const double *ny, int Lx, int Ly, int threads) {
  i64 N = (i64)Lx * Ly;
  PARFOR
  for (i64 c = 0; c < N; ++c) {
    double hh = h[c];
    double amp = sqrt(2 * kbt * mu * 6 * hh / (2 * hh * hh + 6 * hh * delta + 3 * delta * delta));
    kx[c] = nx[c] * amp;
    ky[c] = ny[c] * amp;
  }
}

/* force sum ("update!") src/simulate.jl:18-19; thermal variant scripts/Rivulet_stability.jl:123-124 */
void oracle_force_sum(double *Fx, double *Fy, const double *gx, const double *gy, const double *sx, const double *sy,
                      const double *kx, const double *ky, int Lx, int Ly, int threads) {
  i64 N = (i64)Lx * Ly;
  PARFOR
  for (i64 c = 0; c < N; ++c) Fx[c] = kx ? -gx[c] - sx[c] - kx[c] : -gx[c] - sx[c];
  PARFOR
  for (i64 c = 0; c < N; ++c) Fy[c] = ky ? -gy[c] - sy[c] - ky[c] : -gy[c] - sy[c];
}

/* inclination!  src/forcing.jl:363-368; factor = 0.5 + 0.5*tanh((t-tstart)/tsmooth) evaluated by the caller */
void oracle_inclination(double *Fx, double *Fy, const double *h, double ax, double ay, double factor, int Lx, int Ly,
                        int threads) {
  i64 N = (i64)Lx * Ly;
  PARFOR
  for (i64 c = 0; c < N; ++c) Fx[c] = Fx[c] + h[c] * ax * factor;
  PARFOR
  for (i64 c = 0; c < N; ++c) Fy[c] = Fy[c] + h[c] * ay * factor;
}

/* equilibrium!  src/equilibrium.jl:63-116 -- ten sweeps */
void oracle_equilibrium(double *feq, const double *h, const double *ux, const double *uy, double *vsq, double g,
                        int Lx, int Ly, int threads) {
  i64 N = (i64)Lx * Ly;
  double *f0 = feq, *f1 = feq + N, *f2 = feq + 2 * N, *f3 = feq + 3 * N, *f4 = feq + 4 * N, *f5 = feq + 5 * N,
         *f6 = feq + 6 * N, *f7 = feq + 7 * N, *f8 = feq + 8 * N;
  const double g0 = 1.5 * g, w1 = 1.0 / 9.0, w5 = 1.0 / 36.0;
  PARFOR
  for (i64 c = 0; c < N; ++c) vsq[c] = ux[c] * ux[c] + uy[c] * uy[c];
  PARFOR
  for (i64 c = 0; c < N; ++c) f0[c] = h[c] * (1 - 5.0 / 6.0 * g * h[c] - 2.0 / 3.0 * vsq[c]);
  PARFOR
  for (i64 c = 0; c < N; ++c) f1[c] = w1 * h[c] * (g0 * h[c] + 3 * ux[c] + 4.5 * (ux[c] * ux[c]) - 1.5 * vsq[c]);
  PARFOR
  for (i64 c = 0; c < N; ++c) f2[c] = w1 * h[c] * (g0 * h[c] + 3 * uy[c] + 4.5 * (uy[c] * uy[c]) - 1.5 * vsq[c]);
  PARFOR
  for (i64 c = 0; c < N; ++c) f3[c] = w1 * h[c] * (g0 * h[c] - 3 * ux[c] + 4.5 * (ux[c] * ux[c]) - 1.5 * vsq[c]);
  PARFOR
  for (i64 c = 0; c < N; ++c) f4[c] = w1 * h[c] * (g0 * h[c] - 3 * uy[c] + 4.5 * (uy[c] * uy[c]) - 1.5 * vsq[c]);
  PARFOR
  for (i64 c = 0; c < N; ++c) {
    double s = ux[c] + uy[c];
    f5[c] = w5 * h[c] * (g0 * h[c] + 3 * s + 4.5 * (s * s) - 1.5 * vsq[c]);
  }
  PARFOR
  for (i64 c = 0; c < N; ++c) {
    double d = uy[c] - ux[c];
    f6[c] = w5 * h[c] * (g0 * h[c] + 3 * d + 4.5 * (d * d) - 1.5 * vsq[c]);
  }
  PARFOR
  for (i64 c = 0; c < N; ++c) {
    double s = ux[c] + uy[c];
    f7[c] = w5 * h[c] * (g0 * h[c] - 3 * s + 4.5 * (s * s) - 1.5 * vsq[c]);
  }
  PARFOR
  for (i64 c = 0; c < N; ++c) {
    double e = ux[c] - uy[c];
    f8[c] = w5 * h[c] * (g0 * h[c] + 3 * e + 4.5 * (e * e) - 1.5 * vsq[c]);
  }
}

/* BGKandStream!  src/collide.jl:70-105 -- nine collision sweeps into fout, nine circshift! into ftemp, copy */
void oracle_bgk_stream(double *fout, const double *feq, double *ftemp, const double *Fx, const double *Fy, double tau,
                       int Lx, int Ly, int threads) {
  i64 N = (i64)Lx * Ly;
  const double omeg = 1 - 1 / tau, it = 1 / tau;
  for (int k = 0; k < 9; ++k) {
    double *fo = fout + k * N;
    const double *ft = ftemp + k * N, *fe = feq + k * N;
    PARFOR
    for (i64 c = 0; c < N; ++c) {
      double b = omeg * ft[c] + it * fe[c];
      switch (k) {
        case 0: fo[c] = b; break;
        case 1: fo[c] = b + 1.0 / 3.0 * Fx[c]; break;
        case 2: fo[c] = b + 1.0 / 3.0 * Fy[c]; break;
        case 3: fo[c] = b - 1.0 / 3.0 * Fx[c]; break;
        case 4: fo[c] = b - 1.0 / 3.0 * Fy[c]; break;
        case 5: fo[c] = b + 1.0 / 24.0 * (Fx[c] + Fy[c]); break;
        case 6: fo[c] = b + 1.0 / 24.0 * (Fy[c] - Fx[c]); break;
        case 7: fo[c] = b - 1.0 / 24.0 * (Fx[c] + Fy[c]); break;
        default: fo[c] = b + 1.0 / 24.0 * (Fx[c] - Fy[c]); break;
      }
    }
  }
  static const int s[9][2] = {{0, 0}, {1, 0}, {0, 1}, {-1, 0}, {0, -1}, {1, 1}, {-1, 1}, {-1, -1}, {1, -1}};
  for (int k = 0; k < 9; ++k) circshift2(ftemp + k * N, fout + k * N, s[k][0], s[k][1], Lx, Ly, threads);
  PARFOR
  for (i64 c = 0; c < 9 * N; ++c) fout[c] = ftemp[c];
}

/* moments!  src/moments.jl:43-52 -- sum! accumulates the nine planes in order onto zero */
void oracle_moments(double *h, double *ux, double *uy, const double *fout, int Lx, int Ly, int threads) {
  i64 N = (i64)Lx * Ly;
  const double *f1 = fout + N, *f2 = fout + 2 * N, *f3 = fout + 3 * N, *f4 = fout + 4 * N, *f5 = fout + 5 * N,
               *f6 = fout + 6 * N, *f7 = fout + 7 * N, *f8 = fout + 8 * N;
  PARFOR
  for (i64 c = 0; c < N; ++c) h[c] = 0.0;
  for (int k = 0; k < 9; ++k) {
    const double *fk = fout + k * N;
    PARFOR
    for (i64 c = 0; c < N; ++c) h[c] = h[c] + fk[c];
  }
  PARFOR
  for (i64 c = 0; c < N; ++c) ux[c] = (f1[c] - f3[c] + f5[c] - f6[c] - f7[c] + f8[c]) / h[c];
  PARFOR
  for (i64 c = 0; c < N; ++c) uy[c] = (f2[c] - f4[c] + f5[c] + f6[c] - f7[c] - f8[c]) / h[c];
}

/* ---- whole step / time loop ------------------------------------------------------------------ */

typedef struct {
  double *fout, *ftemp, *feq, *height, *velx, *vely, *vsq, *pressure, *Fx, *Fy, *slipx, *slipy, *hgradpx, *hgradpy,
      *dgrad, *kbtx, *kbty;
} oracle_state;

typedef struct {
  double tau, mu, delta, kbt, gamma, hmin, hcrit, g;
  int n, m;
  double cospi_theta;              /* cospi(theta), host-evaluated */
  const double *cospi_theta_field; /* or an Lx*Ly field (NULL -> scalar) */
  int pressure_variant;            /* 0 power_broad (state form), 1 fast (array form) */
  int slip_variant;                /* 0 slippage!, 1 slippage2!, 2 slippage_ring_riv! */
  int use_inclination;
  double incl_ax, incl_ay, incl_factor;
} oracle_params;

/* one iteration of time_loop  src/simulate.jl:15-22 (no thermal: normals are not reproducible) */
int oracle_step(oracle_state *s, const oracle_params *p, int Lx, int Ly, int threads) {
  int rc = oracle_filmpressure(s->pressure, s->height, s->dgrad, p->gamma, p->cospi_theta, p->cospi_theta_field, p->n,
                               p->m, p->hmin, p->hcrit, p->pressure_variant, Lx, Ly, threads);
  if (rc) return rc;
  oracle_grad9(s->hgradpx, s->hgradpy, s->pressure, s->dgrad, s->height, Lx, Ly, threads);
  oracle_slippage(s->slipx, s->slipy, s->height, s->velx, s->vely, p->delta, p->mu, p->hcrit, p->slip_variant, Lx, Ly,
                  threads);
  oracle_force_sum(s->Fx, s->Fy, s->hgradpx, s->hgradpy, s->slipx, s->slipy, NULL, NULL, Lx, Ly, threads);
  if (p->use_inclination)
    oracle_inclination(s->Fx, s->Fy, s->height, p->incl_ax, p->incl_ay, p->incl_factor, Lx, Ly, threads);
  oracle_equilibrium(s->feq, s->height, s->velx, s->vely, s->vsq, p->g, Lx, Ly, threads);
  oracle_bgk_stream(s->fout, s->feq, s->ftemp, s->Fx, s->Fy, p->tau, Lx, Ly, threads);
  oracle_moments(s->height, s->velx, s->vely, s->fout, Lx, Ly, threads);
  return 0;
}

/* nsteps iterations; if dh != NULL logs max(h)-min(h) before each step (src/simulate.jl:56) and, if
 * wetted != NULL, the count of h > hthresh in the callback slot (src/simulate.jl:89, measures.jl:13-17;
 * height is not modified between the two places, so one sweep serves both). */
int oracle_time_loop(oracle_state *s, const oracle_params *p, int Lx, int Ly, int nsteps, double *dh, long long *wetted,
                     double hthresh, int threads) {
  i64 N = (i64)Lx * Ly;
  for (int t = 0; t < nsteps; ++t) {
    if (dh || wetted) {
      double mx = -INFINITY, mn = INFINITY;
      long long cnt = 0;
      for (i64 c = 0; c < N; ++c) {
        double v = s->height[c];
        mx = v > mx ? v : mx;
        mn = v < mn ? v : mn;
        cnt += v > hthresh;
      }
      if (dh) dh[t] = mx - mn;
      if (wetted) wetted[t] = cnt;
    }
    int rc = oracle_step(s, p, Lx, Ly, threads);
    if (rc) return rc;
  }
  return 0;
}

/* ---- low-memory fused restatement ------------------------------------------------------------------
 * The same step with the same per-site expressions (copied from the sweeps above, same association), but
 * organised as three passes over 22 planes instead of ~100 sweeps over 46: (1) film pressure from height,
 * (2) per site: h∇p, slip, force, equilibrium, collision IN PLACE into f, (3) streaming f -> g plus
 * moments.  It exists so that the BASELINE-size parity tests (4096^2, 8192^2) fit in host memory and finish in
 * seconds; tests/test_oracle_golden.py pins it bit for bit to oracle_time_loop above on every variant.
 * f holds the current populations (the reference's ftemp == fout), g is scratch of the same size; on return
 * f holds the streamed populations of the last step and `pressure` the pressure that step computed. */
typedef struct {
  double *height, *velx, *vely, *pressure, *f, *g;
} oracle_lowmem_state;

int oracle_time_loop_lowmem(oracle_lowmem_state *s, const oracle_params *p, int Lx, int Ly, int nsteps, int threads) {
  const size_t N = (size_t)Lx * Ly;
  int mode;
  if (p->pressure_variant == 1) {
    if (p->n == 9 && p->m == 3) mode = 93;
    else if (p->n == 3 && p->m == 2) mode = 32;
    else return 1;
  } else mode = 0;
  const double gamma = p->gamma, hmin = p->hmin, hcrit = p->hcrit, delta = p->delta, mu = p->mu, g = p->g, tau = p->tau;
  const int n = p->n, m = p->m, variant = p->slip_variant;
  const double nm1 = (double)(n - 1), mm1 = (double)(m - 1), den = (double)(n - m) * hmin;
  const double g0 = 1.5 * g, w1 = 1.0 / 9.0, w5 = 1.0 / 36.0;
  const double omeg = 1 - 1 / tau, it = 1 / tau;
  static const int cs[9][2] = {{0, 0}, {1, 0}, {0, 1}, {-1, 0}, {0, -1}, {1, 1}, {-1, 1}, {-1, -1}, {1, -1}};
  for (int step = 0; step < nsteps; ++step) {
    const double *h = s->height;
    double *pr = s->pressure, *f = s->f, *gg = s->g;
    /* (1) filmpressure!  src/pressure.jl:141-153 (:89-113 for the array form) */
    PARFOR
    for (int j = 0; j < Ly; ++j) {
      const int jp = j ? j - 1 : Ly - 1, jm = j + 1 < Ly ? j + 1 : 0;
      for (int i = 0; i < Lx; ++i) {
        const int ip = i ? i - 1 : Lx - 1, im = i + 1 < Lx ? i + 1 : 0;
        const size_t c = IDX(i, j);
        const double hip = h[IDX(ip, j)], hjp = h[IDX(i, jp)], him = h[IDX(im, j)], hjm = h[IDX(i, jm)],
                     hipjp = h[IDX(ip, jp)], himjp = h[IDX(im, jp)], himjm = h[IDX(im, jm)], hipjm = h[IDX(ip, jm)];
        double x = hmin / (h[c] + hcrit);
        double pw = mode == 93 ? fast_93(x) : mode == 32 ? fast_32(x) : power_broad(x, n) - power_broad(x, m);
        double ct = p->cospi_theta_field ? p->cospi_theta_field[c] : p->cospi_theta;
        double o = -gamma * ((1 - ct) * nm1 * mm1 / den * pw);
        pr[c] = o - gamma * (2.0 / 3.0 * (hjp + hip + him + hjm) + 1.0 / 6.0 * (hipjp + himjp + himjm + hipjm) -
                             10.0 / 3.0 * h[c]);
      }
    }
    /* (2) h∇p! forcing.jl:181-184, slippage! :43-44 (+variants), force sum simulate.jl:18-19 (+inclination!
     * forcing.jl:364), equilibrium! equilibrium.jl:67-114, collision collide.jl:76-89 -- in place into f */
    PARFOR
    for (int j = 0; j < Ly; ++j) {
      const int jp = j ? j - 1 : Ly - 1, jm = j + 1 < Ly ? j + 1 : 0;
      for (int i = 0; i < Lx; ++i) {
        const int ip = i ? i - 1 : Lx - 1, im = i + 1 < Lx ? i + 1 : 0;
        const size_t c = IDX(i, j);
        const double fip = pr[IDX(ip, j)], fjp = pr[IDX(i, jp)], fim = pr[IDX(im, j)], fjm = pr[IDX(i, jm)],
                     fipjp = pr[IDX(ip, jp)], fimjp = pr[IDX(im, jp)], fimjm = pr[IDX(im, jm)], fipjm = pr[IDX(ip, jm)];
        const double hh = h[c], ux = s->velx[c], uy = s->vely[c];
        double gx = -1.0 / 3.0 * (fip - fim) - 1.0 / 12.0 * (fipjp - fimjp - fimjm + fipjm);
        double gy = -1.0 / 3.0 * (fjp - fjm) - 1.0 / 12.0 * (fipjp + fimjp - fimjm - fipjm);
        double hgx = hh * gx, hgy = hh * gy;
        double sx, sy;
        if (variant == 0) {
          sx = (6 * mu * hh * ux) / (2 * (hh * hh) + 6 * delta * hh + 3 * (delta * delta));
          sy = (6 * mu * hh * uy) / (2 * (hh * hh) + 6 * delta * hh + 3 * (delta * delta));
        } else if (variant == 1) {
          double hc = hh + hcrit;
          sx = (6 * mu * hc * ux) / (2 * (hc * hc) + 6 * delta * hc + 3 * (delta * delta));
          sy = (6 * mu * hc * uy) / (2 * (hc * hc) + 6 * delta * hc + 3 * (delta * delta));
        } else {
          sx = (6 * mu * hh * ux) / (2 * (hh * hh) + 6 * delta * (hh + hcrit));
          sy = (6 * mu * hh * uy) / (2 * (hh * hh) + 6 * delta * (hh + hcrit));
        }
        double Fx = -hgx - sx, Fy = -hgy - sy;
        if (p->use_inclination) {
          Fx = Fx + hh * p->incl_ax * p->incl_factor;
          Fy = Fy + hh * p->incl_ay * p->incl_factor;
        }
        double vsq = ux * ux + uy * uy;
        double fe[9];
        fe[0] = hh * (1 - 5.0 / 6.0 * g * hh - 2.0 / 3.0 * vsq);
        fe[1] = w1 * hh * (g0 * hh + 3 * ux + 4.5 * (ux * ux) - 1.5 * vsq);
        fe[2] = w1 * hh * (g0 * hh + 3 * uy + 4.5 * (uy * uy) - 1.5 * vsq);
        fe[3] = w1 * hh * (g0 * hh - 3 * ux + 4.5 * (ux * ux) - 1.5 * vsq);
        fe[4] = w1 * hh * (g0 * hh - 3 * uy + 4.5 * (uy * uy) - 1.5 * vsq);
        {
          double sm = ux + uy;
          fe[5] = w5 * hh * (g0 * hh + 3 * sm + 4.5 * (sm * sm) - 1.5 * vsq);
          fe[7] = w5 * hh * (g0 * hh - 3 * sm + 4.5 * (sm * sm) - 1.5 * vsq);
          double d = uy - ux;
          fe[6] = w5 * hh * (g0 * hh + 3 * d + 4.5 * (d * d) - 1.5 * vsq);
          double e = ux - uy;
          fe[8] = w5 * hh * (g0 * hh + 3 * e + 4.5 * (e * e) - 1.5 * vsq);
        }
        for (int k = 0; k < 9; ++k) {
          double b = omeg * f[c + k * N] + it * fe[k];
          switch (k) {
            case 0: break;
            case 1: b = b + 1.0 / 3.0 * Fx; break;
            case 2: b = b + 1.0 / 3.0 * Fy; break;
            case 3: b = b - 1.0 / 3.0 * Fx; break;
            case 4: b = b - 1.0 / 3.0 * Fy; break;
            case 5: b = b + 1.0 / 24.0 * (Fx + Fy); break;
            case 6: b = b + 1.0 / 24.0 * (Fy - Fx); break;
            case 7: b = b - 1.0 / 24.0 * (Fx + Fy); break;
            default: b = b + 1.0 / 24.0 * (Fx - Fy); break;
          }
          f[c + k * N] = b;
        }
      }
    }
    /* (3) streaming collide.jl:92-103 into g, moments! moments.jl:47-50 */
    PARFOR
    for (int j = 0; j < Ly; ++j) {
      const int jn[3] = {j ? j - 1 : Ly - 1, j, j + 1 < Ly ? j + 1 : 0};  /* j - cy for cy = 1, 0, -1 */
      for (int i = 0; i < Lx; ++i) {
        const int in[3] = {i ? i - 1 : Lx - 1, i, i + 1 < Lx ? i + 1 : 0};
        const size_t c = IDX(i, j);
        double fk[9];
        for (int k = 0; k < 9; ++k) {
          fk[k] = f[IDX(in[1 - cs[k][0]], jn[1 - cs[k][1]]) + k * N];
          gg[c + k * N] = fk[k];
        }
        double hn = 0.0;
        for (int k = 0; k < 9; ++k) hn = hn + fk[k];
        s->height[c] = hn;
        s->velx[c] = (fk[1] - fk[3] + fk[5] - fk[6] - fk[7] + fk[8]) / hn;
        s->vely[c] = (fk[2] - fk[4] + fk[5] + fk[6] - fk[7] - fk[8]) / hn;
      }
    }
    s->f = gg;
    s->g = f;
  }
  return 0;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
