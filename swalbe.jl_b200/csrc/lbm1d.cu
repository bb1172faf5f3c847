// The 1-D (D1Q3) thin-film family of Swalbe.jl on the device: State_1D / SysConst_1D,
// src/initialize.jl:587-598; loop body src/simulate.jl:98-116.  The reference runs it on the CPU only (no device string
// in its 1-D allocator or drivers) at L ~ 1e3 sites, so there is no HBM case here: one step is ~30 FP64 operations per
// site and pure latency.  Design: lattices that fit the shared memory of one CTA (L <= ~4800 at tau == 1) run ALL the
// steps of a call inside one persistent launch, three phases per step over shared-memory arrays; larger ones take one
// launch per step on global memory, every thread recomputing the two neighbour sites it pulls from.
// Arithmetic: the reference's expressions in its evaluation order, no FMA contraction.
#include <math.h>

#include "common.cuh"

namespace swalbe {
namespace {

struct Consts1D {
  PressureConsts pc;
  SlipConsts sc;
  double g05, g025;      // 0.5*g, 0.25*g
  double omega, invtau;
  int tau1, array_form;  // array form: fast_93 / fast_32 (src/pressure.jl:196-227); state form: power_broad (:230-256)
  const double *ct_field;
  // State_gamma_1D loops (run_gamma, src/simulate.jl:518-560): surface tension per site in the pressure, and a force
  // subtracted after the slip (the surface-tension gradient);  time_loop(sys, state, inclination!, α) (:159-179): body force
  const double *gamma_field;  // NULL: the scalar pc.gamma
  const double *force_extra;  // NULL: none
  int use_incl;
  double incl_a, incl_factor;
};

__device__ __forceinline__ int wrap1(int i, int L) { return i < 0 ? i + L : (i >= L ? i - L : i); }

// film pressure at site j   src/pressure.jl:196-227 / :230-256:  -γ (K pw) - γ ((hip - 2h) + him)
__device__ __forceinline__ double pressure_1d(const double *h, int j, int L, const Consts1D &c) {
  const double hc = h[j], hip = h[wrap1(j - 1, L)], him = h[wrap1(j + 1, L)];
  const double x = div_exact(c.pc.hmin, hc + c.pc.hcrit);
  const double pw = disjoining_powers(x, c.pc.pmode, c.pc.n, c.pc.m);
  const double kappa = c.ct_field ? kappa_from_field(c.ct_field[j], c.pc) : c.pc.kappa;
  const double gam = c.gamma_field ? c.gamma_field[j] : c.pc.gamma;  // filmpressure!(state::State_gamma_1D, sys; γ = field)  :284-315
  return (-gam * (kappa * pw)) - gam * ((hip - 2.0 * hc) + him);
}
// F = -h∇p - slip [- ∇γ] [+ h α s(t)]   src/simulate.jl:112, :544 (run_gamma), src/forcing.jl:379-383 (inclination!)
__device__ __forceinline__ double force_1d(double hgp, double slip, double hc, int j, const Consts1D &c) {
  double F = (-hgp) - slip;
  if (c.force_extra) F = F - c.force_extra[j];
  if (c.use_incl) F = F + (hc * c.incl_a) * c.incl_factor;
  return F;
}

struct Site1D {
  double p, hgp, slip, F, fe[3], fs[3];
};

// everything the loop body computes at site j up to the post-collision populations
__device__ __forceinline__ void collide_1d(const double *h, const double *v, const double *ft, size_t fstride, int j, int L,
                                           const Consts1D &c, Site1D &s) {
  const int jm = wrap1(j - 1, L), jp = wrap1(j + 1, L);
  s.p = pressure_1d(h, j, L, c);
  const double pip = pressure_1d(h, jm, L, c), pim = pressure_1d(h, jp, L, c);  // circshift(p, 1)[j] = p[j-1], (p, -1)[j] = p[j+1]
  const double hc = h[j], vc = v[j];
  s.hgp = (hc * -0.5) * (pip - pim);                                        // src/forcing.jl:189-198
  const double den = ((2.0 * (hc * hc)) + c.sc.delta6 * hc) + c.sc.delta3s;  // src/forcing.jl:68-71
  s.slip = div_exact((c.sc.mu6 * hc) * vc, den);
  s.F = force_1d(s.hgp, s.slip, hc, j, c);                                  // src/simulate.jl:112
  const double vv = vc * vc;                                                // src/equilibrium.jl:173-178
  s.fe[0] = hc * ((1.0 - c.g05 * hc) - vv);
  s.fe[1] = hc * ((c.g025 * hc + 0.5 * vc) + 0.5 * vv);
  s.fe[2] = hc * ((c.g025 * hc - 0.5 * vc) + 0.5 * vv);
  const double hf = 0.5 * s.F;                                              // src/collide.jl:186-191
  if (c.tau1) {
    s.fs[0] = s.fe[0]; s.fs[1] = s.fe[1] + hf; s.fs[2] = s.fe[2] - hf;
  } else {
    s.fs[0] = c.omega * ft[j] + c.invtau * s.fe[0];
    s.fs[1] = (c.omega * ft[j + fstride] + c.invtau * s.fe[1]) + hf;
    s.fs[2] = (c.omega * ft[j + 2 * fstride] + c.invtau * s.fe[2]) - hf;
  }
}

struct Out1D {
  double h, v, f[3];
};

// one site of one step: pull f1 from i-1, f2 from i+1 (src/collide.jl:194-196), moments (src/moments.jl:54-62)
__device__ __forceinline__ void step_site_1d(const double *h, const double *v, const double *ft, size_t fstride, int i, int L,
                                             const Consts1D &c, Out1D &o, Site1D *own) {
  Site1D a, b, m;
  collide_1d(h, v, ft, fstride, i, L, c, m);
  collide_1d(h, v, ft, fstride, wrap1(i - 1, L), L, c, a);
  collide_1d(h, v, ft, fstride, wrap1(i + 1, L), L, c, b);
  o.f[0] = m.fs[0]; o.f[1] = a.fs[1]; o.f[2] = b.fs[2];
  o.h = ((0.0 + o.f[0]) + o.f[1]) + o.f[2];
  o.v = div_exact(o.f[1] - o.f[2], o.h);
  if (own) *own = m;
}

struct Loop1DArgs {
  int L, nsteps;
  Consts1D c;
  double *height, *vel, *fout, *ftemp;  // state planes (fout / ftemp: L x 3)
  double *log_min, *log_max;            // nsteps slots or NULL
};

constexpr int T1D = 1024;

// Persistent loop: the whole lattice in the shared memory of one CTA for all the steps of the call; the state planes
// are read once and written once.  With ONE SM doing all the work the FP64 pipe is what a step costs, so nothing is
// recomputed here: three phases per step (pressure | forces, equilibrium, collision | pull + moments) over
// shared-memory arrays h, v, p, f*[3] (+ the old populations at tau != 1), a barrier after each.
__global__ void __launch_bounds__(T1D, 1) k_loop_1d(const __grid_constant__ Loop1DArgs a) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double r_min[2][T1D / 32], r_max[2][T1D / 32];  // (by step parity: thread 0 folds step s while the others run s+1)
  const int L = a.L, tid = threadIdx.x;
  const Consts1D &c = a.c;
  double *sh = sm, *sv = sm + (size_t)L, *sp = sm + 2 * (size_t)L, *sfs = sm + 3 * (size_t)L, *sft = sm + 6 * (size_t)L;
  const bool pops = !c.tau1;
  for (int i = tid; i < L; i += T1D) {
    sh[i] = a.height[i]; sv[i] = a.vel[i];
    if (pops) { sft[i] = a.ftemp[i]; sft[L + i] = a.ftemp[L + i]; sft[2 * L + i] = a.ftemp[2 * (size_t)L + i]; }
  }
  __syncthreads();
  for (int s = 0; s < a.nsteps; ++s) {
    const int par = s & 1;
    const bool last = s == a.nsteps - 1;
    double d_min = INFINITY, d_max = -INFINITY;
    for (int i = tid; i < L; i += T1D) {  // filmpressure!; max - min of the pre-step height (src/simulate.jl:147)
      sp[i] = pressure_1d(sh, i, L, c);
      if (a.log_min) { d_min = fmin(d_min, sh[i]); d_max = fmax(d_max, sh[i]); }
    }
    if (a.log_min) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        d_min = fmin(d_min, __shfl_down_sync(0xffffffffu, d_min, o));
        d_max = fmax(d_max, __shfl_down_sync(0xffffffffu, d_max, o));
      }
      if ((tid & 31) == 0) { r_min[par][tid >> 5] = d_min; r_max[par][tid >> 5] = d_max; }
    }
    __syncthreads();
    if (a.log_min && tid == 0) {
      for (int w = 1; w < T1D / 32; ++w) { d_min = fmin(d_min, r_min[par][w]); d_max = fmax(d_max, r_max[par][w]); }
      a.log_min[s] = d_min; a.log_max[s] = d_max;
    }
    for (int i = tid; i < L; i += T1D) {  // h∇p!, slippage!, F, equilibrium!, collision (same expressions as collide_1d)
      const double hc = sh[i], vc = sv[i];
      const double hgp = (hc * -0.5) * (sp[wrap1(i - 1, L)] - sp[wrap1(i + 1, L)]);
      const double den = ((2.0 * (hc * hc)) + c.sc.delta6 * hc) + c.sc.delta3s;
      const double F = force_1d(hgp, div_exact((c.sc.mu6 * hc) * vc, den), hc, i, c);
      const double vv = vc * vc, hf = 0.5 * F;
      const double fe0 = hc * ((1.0 - c.g05 * hc) - vv);
      const double fe1 = hc * ((c.g025 * hc + 0.5 * vc) + 0.5 * vv);
      const double fe2 = hc * ((c.g025 * hc - 0.5 * vc) + 0.5 * vv);
      if (c.tau1) {
        sfs[i] = fe0; sfs[L + i] = fe1 + hf; sfs[2 * L + i] = fe2 - hf;
      } else {
        sfs[i] = c.omega * sft[i] + c.invtau * fe0;
        sfs[L + i] = (c.omega * sft[L + i] + c.invtau * fe1) + hf;
        sfs[2 * L + i] = (c.omega * sft[2 * L + i] + c.invtau * fe2) - hf;
      }
    }
    __syncthreads();
    for (int i = tid; i < L; i += T1D) {  // streaming (pull) + moments!
      const double f0 = sfs[i], f1 = sfs[L + wrap1(i - 1, L)], f2 = sfs[2 * L + wrap1(i + 1, L)];
      const double hn = ((0.0 + f0) + f1) + f2;
      const double vn = div_exact(f1 - f2, hn);
      sh[i] = hn; sv[i] = vn;
      if (pops) { sft[i] = f0; sft[L + i] = f1; sft[2 * L + i] = f2; }
      if (last) {
        a.height[i] = hn; a.vel[i] = vn;
        a.fout[i] = f0; a.fout[(size_t)L + i] = f1; a.fout[2 * (size_t)L + i] = f2;
        a.ftemp[i] = f0; a.ftemp[(size_t)L + i] = f1; a.ftemp[2 * (size_t)L + i] = f2;
      }
    }
    __syncthreads();
  }
}

// One step on global memory (any L); `aux`: also materialise pressure / h∇p / slip / F / feq like the reference's state
__global__ void __launch_bounds__(256) k_step_1d(int L, Consts1D c, const double *__restrict__ h, const double *__restrict__ v,
                                                  const double *__restrict__ ft, double *__restrict__ hn, double *__restrict__ vn,
                                                  double *__restrict__ f_out, double *__restrict__ f_out2, double *pressure,
                                                  double *hgp, double *slip, double *F, double *feq, double *log_min,
                                                  double *log_max) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double d_min = INFINITY, d_max = -INFINITY;
  if (i < L) {
    Out1D o;
    Site1D m;
    step_site_1d(h, v, ft, (size_t)L, i, L, c, o, &m);
    d_min = d_max = h[i];
    hn[i] = o.h; vn[i] = o.v;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      f_out[(size_t)k * L + i] = o.f[k];
      if (f_out2) f_out2[(size_t)k * L + i] = o.f[k];
    }
    if (pressure) {
      pressure[i] = m.p; hgp[i] = m.hgp; slip[i] = m.slip; F[i] = m.F;
#pragma unroll
      for (int k = 0; k < 3; ++k) feq[(size_t)k * L + i] = m.fe[k];
    }
  }
  if (log_min) {
    __shared__ double r_min[8], r_max[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      d_min = fmin(d_min, __shfl_down_sync(0xffffffffu, d_min, o));
      d_max = fmax(d_max, __shfl_down_sync(0xffffffffu, d_max, o));
    }
    if ((threadIdx.x & 31) == 0) { r_min[threadIdx.x >> 5] = d_min; r_max[threadIdx.x >> 5] = d_max; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w) { d_min = fmin(d_min, r_min[w]); d_max = fmax(d_max, r_max[w]); }
      unsigned long long *pmn = (unsigned long long *)log_min, *pmx = (unsigned long long *)log_max;
      unsigned long long old = *pmn;
      while (d_min < __longlong_as_double((long long)old)) {
        const unsigned long long seen = atomicCAS(pmn, old, (unsigned long long)__double_as_longlong(d_min));
        if (seen == old) break;
        old = seen;
      }
      old = *pmx;
      while (d_max > __longlong_as_double((long long)old)) {
        const unsigned long long seen = atomicCAS(pmx, old, (unsigned long long)__double_as_longlong(d_max));
        if (seen == old) break;
        old = seen;
      }
    }
  }
}

// ---- per-operator array forms (for user-written 1-D loops) ------------------------------------------------------
__global__ void k_eq_1d(double *feq, const double *h, const double *v, double g05, double g025, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  const double hc = h[i], vc = v[i], vv = vc * vc;
  feq[i] = hc * ((1.0 - g05 * hc) - vv);
  feq[(size_t)L + i] = hc * ((g025 * hc + 0.5 * vc) + 0.5 * vv);
  feq[2 * (size_t)L + i] = hc * ((g025 * hc - 0.5 * vc) + 0.5 * vv);
}
__global__ void k_bgk_1d(double *fout, const double *feq, double *ftemp, const double *F, double omega, double invtau, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // pull form of collide + circshift! + copy: reads ftemp/feq/F at i-1, i, i+1
  if (i >= L) return;
  const int im = wrap1(i - 1, L), ip = wrap1(i + 1, L);
  const size_t N = (size_t)L;
  fout[i] = omega * ftemp[i] + invtau * feq[i];
  fout[N + i] = (omega * ftemp[N + im] + invtau * feq[N + im]) + 0.5 * F[im];
  fout[2 * N + i] = (omega * ftemp[2 * N + ip] + invtau * feq[2 * N + ip]) - 0.5 * F[ip];
}
__global__ void k_copy3_1d(double *dst, const double *src, int L) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 3 * (size_t)L) dst[i] = src[i];
}
__global__ void k_moments_1d(double *h, double *v, const double *f, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  const double f0 = f[i], f1 = f[(size_t)L + i], f2 = f[2 * (size_t)L + i];
  const double hn = ((0.0 + f0) + f1) + f2;
  h[i] = hn;
  v[i] = div_exact(f1 - f2, hn);
}
__global__ void k_pressure_1d(double *p, const double *h, Consts1D c, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < L) p[i] = pressure_1d(h, i, L, c);
}
__global__ void k_grad_1d(double *out, const double *f, const double *a, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  const double d = f[wrap1(i - 1, L)] - f[wrap1(i + 1, L)];
  out[i] = a ? (a[i] * -0.5) * d : -0.5 * d;  // src/differences.jl:208-230
}
__global__ void k_lap_1d(double *out, const double *f, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < L) out[i] = (f[wrap1(i - 1, L)] - 2.0 * f[i]) + f[wrap1(i + 1, L)];  // src/differences.jl:77-85
}
__global__ void k_slip_1d(double *slip, const double *h, const double *v, SlipConsts sc, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  const double hc = h[i];
  slip[i] = div_exact((sc.mu6 * hc) * v[i], ((2.0 * (hc * hc)) + sc.delta6 * hc) + sc.delta3s);
}
__global__ void k_force_1d(double *F, const double *hgp, const double *slip, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < L) F[i] = (-hgp[i]) - slip[i];
}
__global__ void k_init_logs_1d(double *mn, double *mx, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { mn[i] = INFINITY; mx[i] = -INFINITY; }
}

// thermal!(fluc, height, kbt, mu, delta)   src/forcing.jl:322-333: one normal per site from the generator of the 2-D
// operator (counter = (site, step), the second normal of the pair is dropped)
__global__ void k_thermal_1d(double *fluc, const double *h, ThermalConsts tc, PhiloxKey key, unsigned long long step, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  double a, b;
  thermal_pair(h[i], tc, key, step, L, 0ll, i, a, b);
  fluc[i] = a;
}
__global__ void k_incl_1d(double *F, const double *h, double alpha, double factor, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < L) F[i] = F[i] + (h[i] * alpha) * factor;  // src/forcing.jl:379-383
}
// ∇γ!(state)        src/forcing.jl:423-432:  -3/2 * ((γ[i-1] - γ[i+1]) / 2)
// ∇γ!(state, sys)   src/forcing.jl:434-447:  (2h² + 6δh + 3δ²) / (6h) * h / 2 * ((γ[i-1] - γ[i+1]) / 2)
__global__ void k_gradgamma_1d(double *out, const double *gam, const double *h, double delta6, double delta3s, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  const double d = (gam[wrap1(i - 1, L)] - gam[wrap1(i + 1, L)]) / 2.0;
  if (!h) { out[i] = -1.5 * d; return; }
  const double hc = h[i];
  out[i] = ((((2.0 * (hc * hc) + delta6 * hc) + delta3s) / (6.0 * hc)) * hc / 2.0) * d;
}
// filmpressure!(state::State_gamma_1D, sys; γ)   src/pressure.jl:284-315 (power_broad; γ scalar or per site; the two
// contributions are also parked in columns 1 and 2 of ftemp) and the active-matter array form
// filmpressure!(output, f, dgrad, rho, γ, θ, n, m, hmin, hcrit; Gamma)   :318-338 with the tension γ + Gamma*rho
__global__ void k_pressure_gamma_1d(double *p, const double *h, Consts1D c, double gamma, const double *rho, double Gamma,
                                    double *ft1, double *ft2, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  const double hc = h[i], hip = h[wrap1(i - 1, L)], him = h[wrap1(i + 1, L)];
  const double x = div_exact(c.pc.hmin, hc + c.pc.hcrit);
  const double pw = disjoining_powers(x, c.pc.pmode, c.pc.n, c.pc.m);
  const double kappa = c.ct_field ? kappa_from_field(c.ct_field[i], c.pc) : c.pc.kappa;
  double gam = c.gamma_field ? c.gamma_field[i] : gamma;
  if (rho) gam = gam + Gamma * rho[i];
  const double disj = -gam * (kappa * pw);
  const double lap = rho ? ((hip - 2.0 * hc) + him) : ((hip - 2.0 * hc) + him);
  if (ft1) { ft1[i] = disj; ft2[i] = -gam * lap; }
  p[i] = disj - gam * lap;
}
// BGKandStream!(state::StateWithBound_1D, sys::SysConstWithBound_1D)   src/collide.jl:214-249: collision, the populations
// that would enter a wall node (border masks) are held back, streamed, and returned to the opposite direction
__global__ void k_bgk_bound_1d(double *fout, const double *feq, double *ftemp_out, const double *ftemp, double *fbound,
                               const double *F, const double *b0, const double *b1, double omega, double invtau, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  const size_t N = (size_t)L;
  auto post1 = [&](int j) { return (omega * ftemp[N + j] + invtau * feq[N + j]) + 0.5 * F[j]; };
  auto post2 = [&](int j) { return (omega * ftemp[2 * N + j] + invtau * feq[2 * N + j]) - 0.5 * F[j]; };
  const int im = wrap1(i - 1, L), ip = wrap1(i + 1, L);
  const double fo1 = post1(i), fo2 = post2(i);
  const double fb1 = fo1 * b0[i], fb2 = fo2 * b1[i];
  fbound[N + i] = fb1; fbound[2 * N + i] = fb2;
  const double fo1m = post1(im), fo2p = post2(ip);
  const double s1 = fo1m - fo1m * b0[im];  // (fo1 - fb1) streamed from i-1
  const double s2 = fo2p - fo2p * b1[ip];  // (fo2 - fb2) streamed from i+1
  const double n0 = omega * ftemp[i] + invtau * feq[i];
  const double n1 = s1 + fb2, n2 = s2 + fb1;
  fout[i] = n0; fout[N + i] = n1; fout[2 * N + i] = n2;
  ftemp_out[i] = n0; ftemp_out[N + i] = n1; ftemp_out[2 * N + i] = n2;
}
// update_rho!(rho, rho_int, height, dgrad, differentials; D, M)   src/forcing.jl:399-417
__global__ void k_update_rho_1d(double *rho_out, double *rho_int, const double *rho, const double *h, double *diff, double D,
                                double M, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  const size_t N = (size_t)L;
  const int im = wrap1(i - 1, L), ip = wrap1(i + 1, L);
  const double r = rho[i], hc = h[i];
  const double lap_rho = (rho[im] - 2.0 * r) + rho[ip];   // ∇²f!  src/differences.jl:77-85
  const double grad_rho = -0.5 * (rho[im] - rho[ip]);      // ∇f!   src/differences.jl:220-230
  const double lap_h = (h[im] - 2.0 * hc) + h[ip];
  const double grad_h = -0.5 * (h[im] - h[ip]);
  diff[i] = lap_rho; diff[N + i] = grad_rho; diff[2 * N + i] = lap_h; diff[3 * N + i] = grad_h;
  const double gh = grad_h / hc;
  const double ri = (D * lap_rho - M * (grad_rho * grad_rho + r * lap_rho)) -
                    D * ((grad_rho * grad_h) / hc + r * (lap_h / hc - gh * gh));
  rho_int[i] = ri;
  rho_out[i] = r + ri;
}
__global__ void k_force3_1d(double *F, const double *hgp, const double *slip, const double *extra, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < L) F[i] = ((-hgp[i]) - slip[i]) - extra[i];
}

int check_len(int L) {
  if (L < 1) return set_error(SWALBE_ERR_EXTENT, "L = %d: the lattice needs at least one site", L);
  return 0;
}

int fill_consts_1d(Consts1D &c, const swalbe_params &p) {
  if (int e = resolve_pmode(p.pressure_variant, p.n, p.m, &c.pc.pmode)) return e;
  if (!(p.tau > 0.0)) return set_error(SWALBE_ERR_ARG, "tau must be positive");
  c.pc.gamma = p.gamma;
  c.pc.kappa = host_kappa(p.cospi_theta, p.n, p.m, p.hmin);
  c.pc.nm1 = (double)(p.n - 1); c.pc.mm1 = (double)(p.m - 1); c.pc.kden = (double)(p.n - p.m) * p.hmin;
  c.pc.hmin = p.hmin; c.pc.hcrit = p.hcrit; c.pc.n = p.n; c.pc.m = p.m;
  c.sc = make_slip(p.delta, p.mu, p.hcrit, SWALBE_SLIP_STANDARD);
  volatile double a = 0.5 * p.g, b = 0.25 * p.g;
  c.g05 = a; c.g025 = b;
  volatile double it = 1.0 / p.tau;
  volatile double om = 1.0 - it;
  c.invtau = it; c.omega = om; c.tau1 = p.tau == 1.0;
  c.array_form = p.pressure_variant == SWALBE_PRESSURE_FAST;
  c.ct_field = p.cospi_theta_field;
  c.gamma_field = nullptr; c.force_extra = nullptr;
  c.use_incl = p.use_inclination; c.incl_a = p.incl_ax; c.incl_factor = p.incl_factor;
  return 0;
}

#define GRID1(L) ((unsigned)(((L) + 255) / 256)), 256, 0, (cudaStream_t)stream

}  // namespace
}  // namespace swalbe

using namespace swalbe;

extern "C" {

int swalbe_equilibrium_d1q3(double *feq, const double *height, const double *vel, double g, int L, void *stream) {
  if (int e = check_len(L)) return e;
  volatile double a = 0.5 * g, b = 0.25 * g;
  k_eq_1d<<<GRID1(L)>>>(feq, height, vel, a, b, L);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_bgk_stream_d1q3(double *fout, const double *feq, double *ftemp, const double *F, double tau, int L, void *stream) {
  if (int e = check_len(L)) return e;
  if (!(tau > 0.0)) return set_error(SWALBE_ERR_ARG, "tau must be positive");
  volatile double it = 1.0 / tau;
  volatile double om = 1.0 - it;
  k_bgk_1d<<<GRID1(L)>>>(fout, feq, ftemp, F, om, it, L);
  SW_LAUNCH_CHECK();
  k_copy3_1d<<<GRID1(3 * (size_t)L)>>>(ftemp, fout, L);  // fout == ftemp on return (src/collide.jl:199)
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_moments_d1q3(double *height, double *vel, const double *fout, int L, void *stream) {
  if (int e = check_len(L)) return e;
  k_moments_1d<<<GRID1(L)>>>(height, vel, fout, L);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_filmpressure_1d(double *pressure, const double *height, double *dgrad, double gamma, double cospi_theta,
                           const double *cospi_theta_field, int n, int m, double hmin, double hcrit, int pressure_variant,
                           int L, void *stream) {
  (void)dgrad;
  if (int e = check_len(L)) return e;
  swalbe_params p = {};
  p.tau = 1.0; p.gamma = gamma; p.cospi_theta = cospi_theta; p.cospi_theta_field = cospi_theta_field;
  p.n = n; p.m = m; p.hmin = hmin; p.hcrit = hcrit; p.pressure_variant = pressure_variant;
  Consts1D c = {};
  if (int e = fill_consts_1d(c, p)) return e;
  k_pressure_1d<<<GRID1(L)>>>(pressure, height, c, L);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_grad_1d(double *output, const double *f, const double *a, int L, void *stream) {
  if (int e = check_len(L)) return e;
  k_grad_1d<<<GRID1(L)>>>(output, f, a, L);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_lap_1d(double *output, const double *f, int L, void *stream) {
  if (int e = check_len(L)) return e;
  k_lap_1d<<<GRID1(L)>>>(output, f, L);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_slippage_1d(double *slip, const double *height, const double *vel, double delta, double mu, int L, void *stream) {
  if (int e = check_len(L)) return e;
  k_slip_1d<<<GRID1(L)>>>(slip, height, vel, make_slip(delta, mu, 0.0, SWALBE_SLIP_STANDARD), L);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_force_sum_1d(double *F, const double *hgradp, const double *slip, int L, void *stream) {
  if (int e = check_len(L)) return e;
  k_force_1d<<<GRID1(L)>>>(F, hgradp, slip, L);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_force_sum3_1d(double *F, const double *hgradp, const double *slip, const double *extra, int L, void *stream) {
  if (int e = check_len(L)) return e;
  if (!F || !hgradp || !slip || !extra) return set_error(SWALBE_ERR_ARG, "swalbe_force_sum3_1d: NULL argument");
  k_force3_1d<<<GRID1(L)>>>(F, hgradp, slip, extra, L);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_thermal_1d(double *fluc, const double *height, double kbt, double mu, double delta, unsigned long long seed,
                      unsigned long long step, int L, void *stream) {
  if (int e = check_len(L)) return e;
  if (!fluc || !height) return set_error(SWALBE_ERR_ARG, "swalbe_thermal_1d: NULL argument");
  k_thermal_1d<<<GRID1(L)>>>(fluc, height, make_thermal(kbt, mu, delta), make_philox_key(seed), step, L);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_inclination_1d(double *F, const double *height, double alpha, double factor, int L, void *stream) {
  if (int e = check_len(L)) return e;
  if (!F || !height) return set_error(SWALBE_ERR_ARG, "swalbe_inclination_1d: NULL argument");
  k_incl_1d<<<GRID1(L)>>>(F, height, alpha, factor, L);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_gradgamma_1d(double *dgamma, const double *gamma, const double *height, double delta, int L, void *stream) {
  if (int e = check_len(L)) return e;
  if (!dgamma || !gamma) return set_error(SWALBE_ERR_ARG, "swalbe_gradgamma_1d: NULL argument");
  volatile double d6 = 6.0 * delta, d2 = delta * delta;
  volatile double d3s = 3.0 * d2;
  k_gradgamma_1d<<<GRID1(L)>>>(dgamma, gamma, height, d6, d3s, L);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_filmpressure_gamma_1d(double *pressure, const double *height, double gamma, const double *gamma_field,
                                 const double *rho, double Gamma, double cospi_theta, const double *cospi_theta_field, int n,
                                 int m, double hmin, double hcrit, double *ftemp, int L, void *stream) {
  if (int e = check_len(L)) return e;
  if (!pressure || !height) return set_error(SWALBE_ERR_ARG, "swalbe_filmpressure_gamma_1d: NULL argument");
  swalbe_params p = {};
  p.tau = 1.0; p.gamma = gamma; p.cospi_theta = cospi_theta; p.cospi_theta_field = cospi_theta_field;
  p.n = n; p.m = m; p.hmin = hmin; p.hcrit = hcrit; p.pressure_variant = SWALBE_PRESSURE_POWER_BROAD;
  Consts1D c = {};
  if (int e = fill_consts_1d(c, p)) return e;
  c.gamma_field = gamma_field;
  k_pressure_gamma_1d<<<GRID1(L)>>>(pressure, height, c, gamma, rho, Gamma, ftemp ? ftemp + (size_t)L : nullptr,
                                    ftemp ? ftemp + 2 * (size_t)L : nullptr, L);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_bgk_stream_bound_d1q3(double *fout, const double *feq, double *ftemp, double *fbound, const double *F,
                                 const double *border0, const double *border1, double tau, int L, void *stream) {
  if (int e = check_len(L)) return e;
  if (!fout || !feq || !ftemp || !fbound || !F || !border0 || !border1)
    return set_error(SWALBE_ERR_ARG, "swalbe_bgk_stream_bound_d1q3: NULL argument");
  if (!(tau > 0.0)) return set_error(SWALBE_ERR_ARG, "tau must be positive");
  volatile double it = 1.0 / tau;
  volatile double om = 1.0 - it;
  // ftemp is read at i-1, i, i+1 and written at i: the new populations go to fout first, then fout -> ftemp
  double *tmp = nullptr;
  SW_CUDA(cudaMallocAsync((void **)&tmp, 3 * (size_t)L * sizeof(double), (cudaStream_t)stream));
  k_bgk_bound_1d<<<GRID1(L)>>>(fout, feq, tmp, ftemp, fbound, F, border0, border1, om, it, L);
  SW_LAUNCH_CHECK();
  SW_CUDA(cudaMemcpyAsync(ftemp, tmp, 3 * (size_t)L * sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  SW_CUDA(cudaFreeAsync(tmp, (cudaStream_t)stream));
  return 0;
}

int swalbe_update_rho_1d(double *rho, double *rho_int, const double *height, double *differentials, double D, double M, int L,
                         void *stream) {
  if (int e = check_len(L)) return e;
  if (!rho || !rho_int || !height || !differentials) return set_error(SWALBE_ERR_ARG, "swalbe_update_rho_1d: NULL argument");
  double *tmp = nullptr;  // rho is read at i-1, i, i+1 and written at i
  SW_CUDA(cudaMallocAsync((void **)&tmp, (size_t)L * sizeof(double), (cudaStream_t)stream));
  k_update_rho_1d<<<GRID1(L)>>>(tmp, rho_int, rho, height, differentials, D, M, L);
  SW_LAUNCH_CHECK();
  SW_CUDA(cudaMemcpyAsync(rho, tmp, (size_t)L * sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  SW_CUDA(cudaFreeAsync(tmp, (cudaStream_t)stream));
  return 0;
}

int swalbe_time_loop_1d(const swalbe_state_1d *st, const swalbe_params *prm, int L, int nsteps, int flags,
                        const swalbe_loop_logs *logs, void *stream_) {
  if (!st || !prm) return set_error(SWALBE_ERR_ARG, "swalbe_time_loop_1d: NULL state/params");
  if (int e = check_len(L)) return e;
  if (nsteps < 0) return set_error(SWALBE_ERR_ARG, "nsteps < 0");
  if (nsteps == 0) return 0;
#define NEED(f) if (!st->f) return set_error(SWALBE_ERR_ARG, "swalbe_time_loop_1d: state." #f " is NULL")
  NEED(fout); NEED(ftemp); NEED(feq); NEED(height); NEED(vel); NEED(pressure); NEED(F); NEED(slip); NEED(hgradp);
#undef NEED
  cudaStream_t stream = (cudaStream_t)stream_;
  Consts1D c = {};
  if (int e = fill_consts_1d(c, *prm)) return e;
  if (flags & SWALBE_LOOP_GAMMA_FIELD) {
    if (!st->gamma) return set_error(SWALBE_ERR_ARG, "SWALBE_LOOP_GAMMA_FIELD needs state.gamma");
    // the reference's State_gamma_1D pressure parks its two contributions in ftemp ("fine as long as tau = 1",
    // src/pressure.jl:307-312): at tau != 1 that changes the collision, which the fused loop does not imitate
    if (!c.tau1) return set_error(SWALBE_ERR_ARG, "SWALBE_LOOP_GAMMA_FIELD requires tau == 1 (src/pressure.jl:307)");
    c.gamma_field = st->gamma;
  }
  if (flags & SWALBE_LOOP_MARANGONI) {
    if (!st->dgamma) return set_error(SWALBE_ERR_ARG, "SWALBE_LOOP_MARANGONI needs state.dgamma");
    c.force_extra = st->dgamma;
  }
  const bool skip_aux = (flags & SWALBE_LOOP_SKIP_AUX) != 0;
  const bool log_mm = logs && logs->hmin && logs->hmax;
  if (log_mm) {
    k_init_logs_1d<<<(unsigned)((nsteps + 255) / 256), 256, 0, stream>>>(logs->hmin, logs->hmax, nsteps);
    SW_LAUNCH_CHECK();
  }
  // steps [0, npers) inside one persistent launch when the lattice fits one CTA's shared memory; the last step (which
  // materialises pressure / h∇p / slip / F / feq unless SKIP_AUX) and lattices that do not fit go step by step
  int dev = 0, max_optin = 0;
  SW_CUDA(cudaGetDevice(&dev));
  SW_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const size_t smem = (size_t)L * sizeof(double) * (c.tau1 ? 6 : 9);  // h, v, p, f*[3] (+ the old populations)
  int npers = skip_aux ? nsteps : nsteps - 1;
  if (smem + 1024 > (size_t)max_optin || npers < 1) npers = 0;
  if (npers > 0) {
    SW_CUDA(cudaFuncSetAttribute(k_loop_1d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    Loop1DArgs a = {};
    a.L = L; a.nsteps = npers; a.c = c;
    a.height = st->height; a.vel = st->vel; a.fout = st->fout; a.ftemp = st->ftemp;
    a.log_min = log_mm ? logs->hmin : nullptr; a.log_max = log_mm ? logs->hmax : nullptr;
    k_loop_1d<<<1, T1D, smem, stream>>>(a);
    SW_LAUNCH_CHECK();
  }
  if (npers == nsteps) return 0;
  // step-by-step part: h, v ping-pong through a scratch pair (stream-ordered allocation, freed in stream order)
  double *scratch = nullptr;
  SW_CUDA(cudaMallocAsync((void **)&scratch, 2 * (size_t)L * sizeof(double), stream));
  double *hA = st->height, *vA = st->vel, *hB = scratch, *vB = scratch + L;
  bool src_is_A = true;
  const int nrem = nsteps - npers;
  if (nrem & 1) {  // arrange for the last step to land in the caller's planes
    SW_CUDA(cudaMemcpyAsync(hB, hA, (size_t)L * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    SW_CUDA(cudaMemcpyAsync(vB, vA, (size_t)L * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    src_is_A = false;
  }
  bool fsrc_is_ftemp = true;  // tau != 1: the populations alternate between ftemp and fout
  for (int s = npers; s < nsteps; ++s) {
    const bool last = s == nsteps - 1, aux = last && !skip_aux;
    const double *h = src_is_A ? hA : hB, *v = src_is_A ? vA : vB;
    double *hn = src_is_A ? hB : hA, *vn = src_is_A ? vB : vA;
    const double *ft = fsrc_is_ftemp ? st->ftemp : st->fout;
    double *fo = c.tau1 ? st->fout : (fsrc_is_ftemp ? st->fout : st->ftemp);
    k_step_1d<<<GRID1(L)>>>(L, c, h, v, ft, hn, vn, fo, (c.tau1 && last) ? st->ftemp : nullptr, aux ? st->pressure : nullptr,
                            st->hgradp, st->slip, st->F, st->feq, log_mm ? logs->hmin + s : nullptr,
                            log_mm ? logs->hmax + s : nullptr);
    SW_LAUNCH_CHECK();
    src_is_A = !src_is_A;
    fsrc_is_ftemp = !fsrc_is_ftemp;
  }
  if (!c.tau1) {  // fout == ftemp on return: the newest populations are in the array written last
    double *newest = fsrc_is_ftemp ? st->ftemp : st->fout, *other = fsrc_is_ftemp ? st->fout : st->ftemp;
    SW_CUDA(cudaMemcpyAsync(other, newest, 3 * (size_t)L * sizeof(double), cudaMemcpyDeviceToDevice, stream));
  }
  SW_CUDA(cudaFreeAsync(scratch, stream));
  return 0;
}

}  // extern "C"
