// Per-operator kernels + their C-ABI entry points (array forms of the reference's operators).
// These exist so that user-written loops that call filmpressure!, h∇p!, ... one by one keep working
// unchanged; the hot path is the fused kernel in fused.cu.  One thread per lattice site, x contiguous.
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

#include "common.cuh"

namespace swalbe {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

int set_error(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_extent(int Lx, int Ly) {
  if (Lx < 1 || Ly < 1) return set_error(SWALBE_ERR_EXTENT, "bad extents Lx=%d Ly=%d", Lx, Ly);
  if ((long long)Lx * (long long)Ly > (1LL << 40)) return set_error(SWALBE_ERR_EXTENT, "lattice too large");
  return 0;
}

constexpr int BX = 128, BY = 2;
static inline dim3 grid2(int Lx, int Ly) { return dim3((Lx + BX - 1) / BX, (Ly + BY - 1) / BY); }
static inline dim3 block2() { return dim3(BX, BY); }

#define SITE_GUARD()                                   \
  const int i = blockIdx.x * BX + threadIdx.x;         \
  const int j = blockIdx.y * BY + threadIdx.y;         \
  if (i >= Lx || j >= Ly) return;                      \
  const size_t c = (size_t)i + (size_t)Lx * (size_t)j; \
  const size_t N = (size_t)Lx * (size_t)Ly;            \
  (void)N

// neighbour loader with periodic wrap, names as in the reference (src/pressure.jl:131-139)
struct Nb {
  double ip, jp, im, jm, ipjp, imjp, imjm, ipjm;
};
__device__ __forceinline__ Nb load_nb(const double *__restrict__ f, int i, int j, int Lx, int Ly) {
  const int il = i == 0 ? Lx - 1 : i - 1, ir = i == Lx - 1 ? 0 : i + 1;
  const size_t rd = (size_t)Lx * (j == 0 ? Ly - 1 : j - 1), r0 = (size_t)Lx * j, ru = (size_t)Lx * (j == Ly - 1 ? 0 : j + 1);
  Nb n;
  n.ip = f[r0 + il]; n.im = f[r0 + ir];
  n.jp = f[rd + i];  n.jm = f[ru + i];
  n.ipjp = f[rd + il]; n.imjp = f[rd + ir];
  n.imjm = f[ru + ir]; n.ipjm = f[ru + il];
  return n;
}

__global__ void __launch_bounds__(BX *BY) k_equilibrium(double *__restrict__ feq, const double *__restrict__ h,
                                                         const double *__restrict__ ux, const double *__restrict__ uy,
                                                         double *__restrict__ vsq, EqConsts ec, int Lx, int Ly) {
  SITE_GUARD();
  double fe[9], v;
  equilibrium_site(h[c], ux[c], uy[c], ec, fe, v);
  vsq[c] = v;
#pragma unroll
  for (int k = 0; k < 9; ++k) feq[c + k * N] = fe[k];
}

// pull form of collide + stream: new_k[i,j] = f*_k[i - ckx, j - cky]
__global__ void __launch_bounds__(BX *BY) k_bgk_stream(double *__restrict__ fout, const double *__restrict__ feq,
                                                        const double *__restrict__ ftemp, const double *__restrict__ Fx,
                                                        const double *__restrict__ Fy, double omega, double invtau,
                                                        int Lx, int Ly) {
  SITE_GUARD();
  const int il = i == 0 ? Lx - 1 : i - 1, ir = i == Lx - 1 ? 0 : i + 1;
  const int jd = j == 0 ? Ly - 1 : j - 1, ju = j == Ly - 1 ? 0 : j + 1;
  const int sx[9] = {i, il, i, ir, i, il, ir, ir, il};
  const int sy[9] = {j, j, jd, j, ju, jd, jd, ju, ju};
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const size_t s = (size_t)sx[k] + (size_t)Lx * sy[k];
    const double b = omega * ftemp[s + k * N] + invtau * feq[s + k * N];
    double v;
    switch (k) {
      case 0: v = b; break;
      case 1: v = b + SW_1_3 * Fx[s]; break;
      case 2: v = b + SW_1_3 * Fy[s]; break;
      case 3: v = b - SW_1_3 * Fx[s]; break;
      case 4: v = b - SW_1_3 * Fy[s]; break;
      case 5: v = b + SW_1_24 * (Fx[s] + Fy[s]); break;
      case 6: v = b + SW_1_24 * (Fy[s] - Fx[s]); break;
      case 7: v = b - SW_1_24 * (Fx[s] + Fy[s]); break;
      default: v = b + SW_1_24 * (Fx[s] - Fy[s]); break;
    }
    fout[c + k * N] = v;
  }
}

__global__ void __launch_bounds__(BX *BY) k_moments(double *__restrict__ h, double *__restrict__ ux,
                                                     double *__restrict__ uy, const double *__restrict__ f, int Lx,
                                                     int Ly) {
  SITE_GUARD();
  double fl[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) fl[k] = f[c + k * N];
  double hh, vx, vy;
  moments_site(fl, hh, vx, vy);
  h[c] = hh; ux[c] = vx; uy[c] = vy;
}

__global__ void __launch_bounds__(BX *BY) k_filmpressure(double *__restrict__ p, const double *__restrict__ h,
                                                          const double *__restrict__ ct_field, PressureConsts pc, int Lx,
                                                          int Ly) {
  SITE_GUARD();
  const Nb n = load_nb(h, i, j, Lx, Ly);
  const double hc = h[c];
  const double lap = lap9_bracket(hc, n.ip, n.jp, n.im, n.jm, n.ipjp, n.imjp, n.imjm, n.ipjm);
  const double kappa = ct_field ? kappa_from_field(ct_field[c], pc) : pc.kappa;
  p[c] = film_pressure(hc, lap, kappa, pc);
}

__global__ void __launch_bounds__(BX *BY) k_grad9(double *__restrict__ ox, double *__restrict__ oy,
                                                   const double *__restrict__ f, const double *__restrict__ a, int Lx,
                                                   int Ly) {
  SITE_GUARD();
  const Nb n = load_nb(f, i, j, Lx, Ly);
  const double gx = grad9_x(n.ip, n.im, n.ipjp, n.imjp, n.imjm, n.ipjm);
  const double gy = grad9_y(n.jp, n.jm, n.ipjp, n.imjp, n.imjm, n.ipjm);
  if (a) {
    const double ac = a[c];
    ox[c] = ac * gx;
    oy[c] = ac * gy;
  } else {
    ox[c] = gx;
    oy[c] = gy;
  }
}

__global__ void __launch_bounds__(BX *BY) k_lap9(double *__restrict__ out, const double *__restrict__ f, double gamma,
                                                  int Lx, int Ly) {
  SITE_GUARD();
  const Nb n = load_nb(f, i, j, Lx, Ly);
  out[c] = gamma * lap9_bracket(f[c], n.ip, n.jp, n.im, n.jm, n.ipjp, n.imjp, n.imjm, n.ipjm);
}

__global__ void __launch_bounds__(BX *BY) k_slippage(double *__restrict__ sx, double *__restrict__ sy,
                                                      const double *__restrict__ h, const double *__restrict__ ux,
                                                      const double *__restrict__ uy, SlipConsts sc, int Lx, int Ly) {
  SITE_GUARD();
  double a, b;
  slip_terms(h[c], ux[c], uy[c], sc, sc.variant, a, b);
  sx[c] = a; sy[c] = b;
}

__global__ void __launch_bounds__(BX *BY) k_force_sum(double *__restrict__ Fx, double *__restrict__ Fy,
                                                       const double *__restrict__ gx, const double *__restrict__ gy,
                                                       const double *__restrict__ sx, const double *__restrict__ sy,
                                                       const double *__restrict__ kx, const double *__restrict__ ky,
                                                       int Lx, int Ly) {
  SITE_GUARD();
  double fx = (-gx[c]) - sx[c], fy = (-gy[c]) - sy[c];
  if (kx) { fx = fx - kx[c]; fy = fy - ky[c]; }
  Fx[c] = fx; Fy[c] = fy;
}

__global__ void __launch_bounds__(BX *BY) k_thermal(double *__restrict__ kx, double *__restrict__ ky,
                                                     const double *__restrict__ h, ThermalConsts tc,
                                                     PhiloxKey key, unsigned long long step, int Lx, int Ly) {
  SITE_GUARD();
  double a, b;
  thermal_pair(h[c], tc, key, step, Lx, (long long)j, i, a, b);
  kx[c] = a;
  ky[c] = b;
}

__global__ void __launch_bounds__(BX *BY) k_inclination(double *__restrict__ Fx, double *__restrict__ Fy,
                                                         const double *__restrict__ h, double ax, double ay, double factor,
                                                         int Lx, int Ly) {
  SITE_GUARD();
  const double hc = h[c];
  Fx[c] = Fx[c] + (hc * ax) * factor;
  Fy[c] = Fy[c] + (hc * ay) * factor;
}

__global__ void __launch_bounds__(256) k_cospi(double *__restrict__ out, const double *__restrict__ th, size_t n) {
  for (size_t c = (size_t)blockIdx.x * 256 + threadIdx.x; c < n; c += (size_t)gridDim.x * 256) out[c] = cospi(th[c]);
}

// ---- field statistics: fixed-order two-pass reduction (deterministic) --------------------------
constexpr int ST_THREADS = 256;
struct Stat4 { double mn, mx, sm; unsigned long long cnt; };

__device__ __forceinline__ Stat4 stat_merge(Stat4 a, Stat4 b) {
  Stat4 r;
  r.mn = fmin(a.mn, b.mn); r.mx = fmax(a.mx, b.mx); r.sm = a.sm + b.sm; r.cnt = a.cnt + b.cnt;
  return r;
}
__device__ __forceinline__ Stat4 stat_block_reduce(Stat4 v) {
  __shared__ Stat4 sh[ST_THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Stat4 w;
    w.mn = __shfl_down_sync(0xffffffffu, v.mn, o); w.mx = __shfl_down_sync(0xffffffffu, v.mx, o);
    w.sm = __shfl_down_sync(0xffffffffu, v.sm, o); w.cnt = __shfl_down_sync(0xffffffffu, v.cnt, o);
    v = stat_merge(v, w);
  }
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    Stat4 id; id.mn = INFINITY; id.mx = -INFINITY; id.sm = 0.0; id.cnt = 0;
    v = threadIdx.x < ST_THREADS / 32 ? sh[threadIdx.x] : id;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      Stat4 w;
      w.mn = __shfl_down_sync(0xffffffffu, v.mn, o); w.mx = __shfl_down_sync(0xffffffffu, v.mx, o);
      w.sm = __shfl_down_sync(0xffffffffu, v.sm, o); w.cnt = __shfl_down_sync(0xffffffffu, v.cnt, o);
      v = stat_merge(v, w);
    }
  }
  return v;
}
__global__ void __launch_bounds__(ST_THREADS) k_stats_partial(Stat4 *__restrict__ part, const double *__restrict__ f,
                                                               double thresh, size_t N) {
  Stat4 v; v.mn = INFINITY; v.mx = -INFINITY; v.sm = 0.0; v.cnt = 0;
  for (size_t c = (size_t)blockIdx.x * ST_THREADS + threadIdx.x; c < N; c += (size_t)gridDim.x * ST_THREADS) {
    const double x = f[c];
    v.mn = fmin(v.mn, x); v.mx = fmax(v.mx, x); v.sm += x; v.cnt += x > thresh;
  }
  v = stat_block_reduce(v);
  if (threadIdx.x == 0) part[blockIdx.x] = v;
}
__global__ void __launch_bounds__(ST_THREADS) k_stats_final(double *__restrict__ out4, const Stat4 *__restrict__ part,
                                                             int nparts) {
  Stat4 v; v.mn = INFINITY; v.mx = -INFINITY; v.sm = 0.0; v.cnt = 0;
  for (int c = threadIdx.x; c < nparts; c += ST_THREADS) v = stat_merge(v, part[c]);
  v = stat_block_reduce(v);
  if (threadIdx.x == 0) { out4[0] = v.mn; out4[1] = v.mx; out4[2] = v.sm; out4[3] = (double)v.cnt; }
}

// ---- self-test: div_exact / div2_exact against the compiler's IEEE division ------------------------------------
__device__ __forceinline__ bool same_double(double x, double y) {
  if (x != x && y != y) return true;  // NaN == NaN for this purpose
  return __double_as_longlong(x) == __double_as_longlong(y);
}
__global__ void k_selftest_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, PhiloxKey key, unsigned int *out4) {
  uint32_t r[4];
  philox4x32_10(c0, c1, c2, c3, key, r);
  if (threadIdx.x < 4) out4[threadIdx.x] = r[threadIdx.x];
}

__global__ void k_selftest_division(unsigned long long n, unsigned long long seed, unsigned long long *mismatch) {
  unsigned long long bad = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    uint32_t r[4], q[4];
    philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 1u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 2u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), q);
    unsigned long long ba = ((unsigned long long)r[0] << 32) | r[1], bb = ((unsigned long long)r[2] << 32) | r[3];
    unsigned long long bc = ((unsigned long long)q[0] << 32) | q[1];
    const unsigned mode = q[2] & 7u;
    auto squeeze = [](unsigned long long bits, int lo, int span, uint32_t rnd) {  // exponent into [lo, lo+span)
      const unsigned long long e = (unsigned long long)(lo + (int)(rnd % (unsigned)span));
      return (bits & 0x800fffffffffffffull) | (e << 52);
    };
    if (mode == 0) { /* raw bit patterns: every exponent, Inf, NaN, denormals */ }
    else if (mode <= 3) { ba = squeeze(ba, 1023 - 60, 120, q[3]); bb = squeeze(bb, 1023 - 60, 120, q[3] >> 8); bc = squeeze(bc, 1023 - 60, 120, q[3] >> 16); }
    else if (mode == 4) { ba = squeeze(ba, 600, 850, q[3]); bb = squeeze(bb, 600, 850, q[3] >> 7); bc = squeeze(bc, 600, 850, q[3] >> 14); }
    else if (mode == 5) { ba &= 0x8000000000000000ull; bb = squeeze(bb, 900, 250, q[3]); }                    // +-0 numerator
    else if (mode == 6) { ba = squeeze(ba | 0x000fffffffffff00ull, 1000, 50, q[3]); bb = squeeze(bb & 0xfff00000000000ffull, 1000, 50, q[3] >> 9); }
    else { bb = squeeze(bb | 0x000fffffffffffffull, 1010, 30, q[3]); bc &= 0x8000000000000000ull; ba = squeeze(ba, 1010, 30, q[3] >> 5); }
    const double a = __longlong_as_double((long long)ba), b = __longlong_as_double((long long)bb),
                 c = __longlong_as_double((long long)bc);
    double q1, q2;
    div2_exact(a, c, b, q1, q2);
    bad += !same_double(div_exact(a, b), a / b);
    bad += !same_double(q1, a / b);
    bad += !same_double(q2, c / b);
  }
  if (bad) atomicAdd(mismatch, bad);
}

}  // namespace swalbe

using namespace swalbe;

#define REQUIRE(p)                                                                   \
  do {                                                                               \
    if (!(p)) return set_error(SWALBE_ERR_ARG, "%s: required pointer is NULL: %s", __func__, #p); \
  } while (0)

extern "C" {

int swalbe_version(void) { return SWALBE_B200_VERSION; }
const char *swalbe_last_error(void) { return g_err; }
unsigned long long swalbe_launch_count(void) { return g_launches.load(); }

int swalbe_equilibrium_d2q9(double *feq, const double *height, const double *velx, const double *vely, double *vsq,
                            double g, int Lx, int Ly, void *stream) {
  if (int e = check_extent(Lx, Ly)) return e;
  REQUIRE(feq); REQUIRE(height); REQUIRE(velx); REQUIRE(vely); REQUIRE(vsq);
  k_equilibrium<<<grid2(Lx, Ly), block2(), 0, (cudaStream_t)stream>>>(feq, height, velx, vely, vsq, make_eq(g), Lx, Ly);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_bgk_stream_d2q9(double *fout, const double *feq, double *ftemp, const double *Fx, const double *Fy,
                           double tau, int Lx, int Ly, void *stream) {
  if (int e = check_extent(Lx, Ly)) return e;
  REQUIRE(fout); REQUIRE(feq); REQUIRE(ftemp); REQUIRE(Fx); REQUIRE(Fy);
  if (fout == ftemp || fout == feq) return set_error(SWALBE_ERR_ARG, "BGKandStream!: fout must not alias ftemp/feq");
  volatile double it = 1.0 / tau;
  volatile double om = 1.0 - it;  // omeg = 1 - 1/τ   src/collide.jl:76
  k_bgk_stream<<<grid2(Lx, Ly), block2(), 0, (cudaStream_t)stream>>>(fout, feq, ftemp, Fx, Fy, om, it, Lx, Ly);
  SW_LAUNCH_CHECK();
  // ftemp <- streamed populations; afterwards fout == ftemp   (src/collide.jl:92-103)
  SW_CUDA(cudaMemcpyAsync(ftemp, fout, sizeof(double) * 9 * (size_t)Lx * Ly, cudaMemcpyDeviceToDevice,
                          (cudaStream_t)stream));
  return 0;
}

int swalbe_moments_d2q9(double *height, double *velx, double *vely, const double *fout, int Lx, int Ly, void *stream) {
  if (int e = check_extent(Lx, Ly)) return e;
  REQUIRE(height); REQUIRE(velx); REQUIRE(vely); REQUIRE(fout);
  k_moments<<<grid2(Lx, Ly), block2(), 0, (cudaStream_t)stream>>>(height, velx, vely, fout, Lx, Ly);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_filmpressure(double *pressure, const double *height, double *dgrad, double gamma, double cospi_theta,
                        const double *cospi_theta_field, int n, int m, double hmin, double hcrit, int pressure_variant,
                        int Lx, int Ly, void *stream) {
  (void)dgrad;
  if (int e = check_extent(Lx, Ly)) return e;
  REQUIRE(pressure); REQUIRE(height);
  if (pressure == height) return set_error(SWALBE_ERR_ARG, "filmpressure!: output must not alias height");
  PressureConsts pc;
  if (int e = resolve_pmode(pressure_variant, n, m, &pc.pmode)) return e;
  pc.gamma = gamma; pc.kappa = host_kappa(cospi_theta, n, m, hmin);
  pc.nm1 = (double)(n - 1); pc.mm1 = (double)(m - 1); pc.kden = (double)(n - m) * hmin;
  pc.hmin = hmin; pc.hcrit = hcrit; pc.n = n; pc.m = m;
  k_filmpressure<<<grid2(Lx, Ly), block2(), 0, (cudaStream_t)stream>>>(pressure, height, cospi_theta_field, pc, Lx, Ly);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_hgradp(double *hgradpx, double *hgradpy, const double *pressure, const double *height, int Lx, int Ly,
                  void *stream) {
  REQUIRE(height);
  return swalbe_grad9(hgradpx, hgradpy, pressure, height, Lx, Ly, stream);
}

int swalbe_grad9(double *outx, double *outy, const double *f, const double *a, int Lx, int Ly, void *stream) {
  if (int e = check_extent(Lx, Ly)) return e;
  REQUIRE(outx); REQUIRE(outy); REQUIRE(f);
  if (outx == f || outy == f) return set_error(SWALBE_ERR_ARG, "∇f!: outputs must not alias the input field");
  k_grad9<<<grid2(Lx, Ly), block2(), 0, (cudaStream_t)stream>>>(outx, outy, f, a, Lx, Ly);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_lap9(double *output, const double *f, double gamma, int Lx, int Ly, void *stream) {
  if (int e = check_extent(Lx, Ly)) return e;
  REQUIRE(output); REQUIRE(f);
  if (output == f) return set_error(SWALBE_ERR_ARG, "∇²f!: output must not alias the input field");
  k_lap9<<<grid2(Lx, Ly), block2(), 0, (cudaStream_t)stream>>>(output, f, gamma, Lx, Ly);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_slippage(double *slipx, double *slipy, const double *height, const double *velx, const double *vely,
                    double delta, double mu, double hcrit, int slip_variant, int Lx, int Ly, void *stream) {
  if (int e = check_extent(Lx, Ly)) return e;
  REQUIRE(slipx); REQUIRE(slipy); REQUIRE(height); REQUIRE(velx); REQUIRE(vely);
  if (slip_variant < 0 || slip_variant > 2) return set_error(SWALBE_ERR_ARG, "unknown slip_variant %d", slip_variant);
  k_slippage<<<grid2(Lx, Ly), block2(), 0, (cudaStream_t)stream>>>(slipx, slipy, height, velx, vely,
                                                                     make_slip(delta, mu, hcrit, slip_variant), Lx, Ly);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_force_sum(double *Fx, double *Fy, const double *hgradpx, const double *hgradpy, const double *slipx,
                     const double *slipy, const double *kbtx, const double *kbty, int Lx, int Ly, void *stream) {
  if (int e = check_extent(Lx, Ly)) return e;
  REQUIRE(Fx); REQUIRE(Fy); REQUIRE(hgradpx); REQUIRE(hgradpy); REQUIRE(slipx); REQUIRE(slipy);
  if ((kbtx == nullptr) != (kbty == nullptr)) return set_error(SWALBE_ERR_ARG, "kbtx and kbty must both be set or both NULL");
  k_force_sum<<<grid2(Lx, Ly), block2(), 0, (cudaStream_t)stream>>>(Fx, Fy, hgradpx, hgradpy, slipx, slipy, kbtx, kbty, Lx, Ly);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_thermal(double *kbtx, double *kbty, const double *height, double kbt, double mu, double delta,
                   unsigned long long seed, unsigned long long step, int Lx, int Ly, void *stream) {
  if (int e = check_extent(Lx, Ly)) return e;
  REQUIRE(kbtx); REQUIRE(kbty); REQUIRE(height);
  k_thermal<<<grid2(Lx, Ly), block2(), 0, (cudaStream_t)stream>>>(kbtx, kbty, height, make_thermal(kbt, mu, delta), make_philox_key(seed), step, Lx, Ly);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_inclination(double *Fx, double *Fy, const double *height, double alpha_x, double alpha_y, double factor,
                       int Lx, int Ly, void *stream) {
  if (int e = check_extent(Lx, Ly)) return e;
  REQUIRE(Fx); REQUIRE(Fy); REQUIRE(height);
  k_inclination<<<grid2(Lx, Ly), block2(), 0, (cudaStream_t)stream>>>(Fx, Fy, height, alpha_x, alpha_y, factor, Lx, Ly);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_cospi_field(double *out, const double *theta, size_t count, void *stream) {
  REQUIRE(out); REQUIRE(theta);
  if (count == 0) return 0;
  const size_t blocks = (count + 255) / 256;
  k_cospi<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, (cudaStream_t)stream>>>(out, theta, count);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_selftest_philox(const unsigned int ctr[4], const unsigned int key[2], unsigned int *out4, void *stream) {
  REQUIRE(ctr); REQUIRE(key); REQUIRE(out4);
  const unsigned long long seed = (unsigned long long)key[0] | ((unsigned long long)key[1] << 32);
  k_selftest_philox<<<1, 32, 0, (cudaStream_t)stream>>>(ctr[0], ctr[1], ctr[2], ctr[3], make_philox_key(seed), out4);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_selftest_division(unsigned long long n, unsigned long long seed, unsigned long long *mismatches, void *stream) {
  REQUIRE(mismatches);
  SW_CUDA(cudaMemsetAsync(mismatches, 0, sizeof(unsigned long long), (cudaStream_t)stream));
  k_selftest_division<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(n, seed, mismatches);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_field_stats(double *out4, const double *f, double thresh, int Lx, int Ly, void *stream) {
  if (int e = check_extent(Lx, Ly)) return e;
  REQUIRE(out4); REQUIRE(f);
  const size_t N = (size_t)Lx * Ly;
  int nparts = (int)((N + ST_THREADS - 1) / ST_THREADS);
  if (nparts > 1184) nparts = 1184;  // 8 CTAs per SM x 148 SMs
  Stat4 *part = nullptr;
  SW_CUDA(cudaMallocAsync((void **)&part, sizeof(Stat4) * nparts, (cudaStream_t)stream));
  k_stats_partial<<<nparts, ST_THREADS, 0, (cudaStream_t)stream>>>(part, f, thresh, N);
  SW_LAUNCH_CHECK();
  k_stats_final<<<1, ST_THREADS, 0, (cudaStream_t)stream>>>(out4, part, nparts);
  SW_LAUNCH_CHECK();
  SW_CUDA(cudaFreeAsync(part, (cudaStream_t)stream));
  return 0;
}

}  // extern "C"
