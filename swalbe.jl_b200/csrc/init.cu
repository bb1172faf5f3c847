// On-device initial conditions and substrate motion: what the reference builds on one host thread
// and uploads (src/initialvalues.jl, src/simulate.jl:350-353) or does with circshift! on a CuArray
// (scripts/Moving_wettability_structs.jl:139-152).  Every kernel fills a row slab [j_begin, j_begin + Ly_local) of a
// global lattice, so each rank of the slab runtime builds its own rows of a 32768^2 field without a host copy.
// Coordinates are the reference's 1-based (i, j).  Same -fmad=false arithmetic order as the Julia expressions; the
// transcendental functions (sin, cos, asin) are CUDA's, which agree with Julia's openlibm to a few ulp, not bit for bit
// -- an initial condition is input data, and the tests bound the difference at 4 ulp of the drop height.
#include <math.h>

#include "common.cuh"

namespace swalbe {
namespace {

constexpr int IB = 256;

__device__ __forceinline__ bool cell_of(size_t idx, int Lx, int Ly_local, int j_begin, double &i1, double &j1,
                                        unsigned long long &gcell) {
  if (idx >= (size_t)Lx * (size_t)Ly_local) return false;
  const int j = (int)(idx / (size_t)Lx), i = (int)(idx - (size_t)j * Lx);
  i1 = (double)(i + 1);
  j1 = (double)(j_begin + j + 1);
  gcell = (unsigned long long)(j_begin + j) * (unsigned long long)Lx + (unsigned long long)i;
  return true;
}

// one standard normal per (seed, stream id, global cell): Philox4x32-10 + Box-Muller (decomposition independent)
__device__ __forceinline__ double normal_of(const PhiloxKey &key, unsigned long long stream_id, unsigned long long cell,
                                            const NormalTables &T) {
  double z0, z1;
  normal_pair(key, stream_id, cell, T, z0, z1);
  return z0;
}
#define IC_NORMAL_TABLES()                        \
  __shared__ NormalTables s_nt;                   \
  normal_tables_fill(s_nt, threadIdx.x, IB);      \
  __syncthreads()

// singledroplet  src/initialvalues.jl:203-224
__global__ void k_ic_singledroplet(double *h, double radius, double ct, double cx, double cy, double precursor, int Lx,
                                   int Ly_local, int j_begin) {
  double i1, j1; unsigned long long g;
  const size_t idx = (size_t)blockIdx.x * IB + threadIdx.x;
  if (!cell_of(idx, Lx, Ly_local, j_begin, i1, j1, g)) return;
  const double dx = i1 - cx, dy = j1 - cy;
  const double circ = sqrt(dx * dx + dy * dy);
  double v = precursor;
  if (circ <= radius) v = (cos(asin(circ / radius)) - ct) * radius;
  if (v < 0.0) v = precursor;
  h[idx] = v;
}

// torus  src/initialvalues.jl:144-168
__global__ void k_ic_torus(double *h, double r1, double R2, double ct, double cx, double cy, double hmin, double noise,
                           PhiloxKey seed, int Lx, int Ly_local, int j_begin) {
  IC_NORMAL_TABLES();
  double i1, j1; unsigned long long g;
  const size_t idx = (size_t)blockIdx.x * IB + threadIdx.x;
  if (!cell_of(idx, Lx, Ly_local, j_begin, i1, j1, g)) return;
  const double dx = i1 - cx, dy = j1 - cy;
  const double coord = sqrt(dx * dx + dy * dy);
  const double half = r1 * r1 - (coord - R2) * (coord - R2);
  double v = half <= 0.0 ? hmin : sqrt(half);
  const double corr = v - r1 * ct;
  if (corr < hmin) v = hmin;
  else v = noise != 0.0 ? corr + normal_of(seed, 0x746f727573ull, g, s_nt) * noise : corr;
  h[idx] = v;
}

// rivulet  src/initialvalues.jl:69-104   (orientation 0 = :y, the ridge runs along j; 1 = :x)
__global__ void k_ic_rivulet(double *h, double radius, double ct, int orientation, double center, double hmin, double noise,
                             PhiloxKey seed, int Lx, int Ly_local, int j_begin) {
  IC_NORMAL_TABLES();
  double i1, j1; unsigned long long g;
  const size_t idx = (size_t)blockIdx.x * IB + threadIdx.x;
  if (!cell_of(idx, Lx, Ly_local, j_begin, i1, j1, g)) return;
  const double d = (orientation == 0 ? i1 : j1) - center;
  const double circ = sqrt(d * d);
  double v = hmin;
  if (circ <= radius) {
    v = (cos(asin(circ / radius)) - ct) * radius;
    if (noise != 0.0) v = v + normal_of(seed, 0x726976756c6574ull, g, s_nt) * noise;
  }
  if (v <= hmin) v = hmin;
  h[idx] = v;
}

// Rayleigh-Taylor / sine initial condition  src/simulate.jl:350-353:
//   h0 * (1 + eps * sin(2pi*kx*i/(Lx-1)) * sin(2pi*ky*j/(Ly-1)))    (divides by L-1 like the reference)
__global__ void k_ic_sine(double *h, double h0, double eps, double kx, double ky, double denx, double deny, int Lx,
                          int Ly_local, int j_begin) {
  double i1, j1; unsigned long long g;
  const size_t idx = (size_t)blockIdx.x * IB + threadIdx.x;
  if (!cell_of(idx, Lx, Ly_local, j_begin, i1, j1, g)) return;
  const double twopi = 6.283185307179586;  // 2π == 2*Float64(π)
  const double sx = sin(((twopi * kx) * i1) / denx), sy = sin(((twopi * ky) * j1) / deny);
  h[idx] = h0 * (1.0 + (eps * sx) * sy);
}

// randinterface!  src/initialvalues.jl:23-33 with counter-based normals in place of Julia's unseeded randn!
__global__ void k_ic_rand(double *h, double h0, double eps, PhiloxKey seed, int Lx, int Ly_local, int j_begin) {
  IC_NORMAL_TABLES();
  double i1, j1; unsigned long long g;
  const size_t idx = (size_t)blockIdx.x * IB + threadIdx.x;
  if (!cell_of(idx, Lx, Ly_local, j_begin, i1, j1, g)) return;
  h[idx] = h0 * (1.0 + eps * normal_of(seed, 0x72616e64ull, g, s_nt));
}

// circshift!(dst, src, (sx, sy)):  dst[i, j] = src[i - sx, j - sy]  (periodic)
__global__ void k_circshift(double *__restrict__ dst, const double *__restrict__ src, int sx, int sy, int Lx, int Ly) {
  const size_t idx = (size_t)blockIdx.x * IB + threadIdx.x;
  if (idx >= (size_t)Lx * (size_t)Ly) return;
  const int j = (int)(idx / (size_t)Lx), i = (int)(idx - (size_t)j * Lx);
  int is = i - sx, js = j - sy;
  is += is < 0 ? Lx : 0;
  js += js < 0 ? Ly : 0;
  dst[idx] = src[(size_t)js * Lx + is];
}

inline unsigned blocks_for(int Lx, int Ly) { return (unsigned)(((size_t)Lx * Ly + IB - 1) / IB); }

int check_slab(int Lx, int Ly_local, int j_begin) {
  if (int e = check_extent(Lx, Ly_local)) return e;
  if (j_begin < 0) return set_error(SWALBE_ERR_ARG, "j_begin must be >= 0 (got %d)", j_begin);
  return 0;
}

}  // namespace
}  // namespace swalbe

using namespace swalbe;

#define REQUIRE(p)                                                                                 \
  do {                                                                                             \
    if (!(p)) return set_error(SWALBE_ERR_ARG, "%s: required pointer is NULL: %s", __func__, #p); \
  } while (0)

extern "C" {

int swalbe_ic_singledroplet(double *height, double radius, double cospi_theta, double cx, double cy, double precursor,
                            int Lx, int Ly_local, int j_begin, void *stream) {
  if (int e = check_slab(Lx, Ly_local, j_begin)) return e;
  REQUIRE(height);
  k_ic_singledroplet<<<blocks_for(Lx, Ly_local), IB, 0, (cudaStream_t)stream>>>(height, radius, cospi_theta, cx, cy,
                                                                                 precursor, Lx, Ly_local, j_begin);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_ic_torus(double *height, double r1, double R2, double cospi_theta, double cx, double cy, double hmin,
                    double noise, unsigned long long seed, int Lx, int Ly_local, int j_begin, void *stream) {
  if (int e = check_slab(Lx, Ly_local, j_begin)) return e;
  REQUIRE(height);
  k_ic_torus<<<blocks_for(Lx, Ly_local), IB, 0, (cudaStream_t)stream>>>(height, r1, R2, cospi_theta, cx, cy, hmin, noise,
                                                                         make_philox_key(seed), Lx, Ly_local, j_begin);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_ic_rivulet(double *height, double radius, double cospi_theta, int orientation, double center, double hmin,
                      double noise, unsigned long long seed, int Lx, int Ly_local, int j_begin, void *stream) {
  if (int e = check_slab(Lx, Ly_local, j_begin)) return e;
  REQUIRE(height);
  if (orientation != 0 && orientation != 1)
    return set_error(SWALBE_ERR_ARG, "rivulet: orientation must be 0 (:y) or 1 (:x), got %d", orientation);
  k_ic_rivulet<<<blocks_for(Lx, Ly_local), IB, 0, (cudaStream_t)stream>>>(height, radius, cospi_theta, orientation, center,
                                                                           hmin, noise, make_philox_key(seed), Lx, Ly_local, j_begin);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_ic_sinewave2d(double *height, double h0, double eps, double kx, double ky, int Lx, int Ly, int Ly_local,
                         int j_begin, void *stream) {
  if (int e = check_slab(Lx, Ly_local, j_begin)) return e;
  REQUIRE(height);
  if (Ly < j_begin + Ly_local) return set_error(SWALBE_ERR_ARG, "sinewave2d: slab [%d,%d) outside Ly=%d", j_begin,
                                                j_begin + Ly_local, Ly);
  k_ic_sine<<<blocks_for(Lx, Ly_local), IB, 0, (cudaStream_t)stream>>>(height, h0, eps, kx, ky, (double)(Lx - 1),
                                                                        (double)(Ly - 1), Lx, Ly_local, j_begin);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_ic_randinterface(double *height, double h0, double eps, unsigned long long seed, int Lx, int Ly_local,
                            int j_begin, void *stream) {
  if (int e = check_slab(Lx, Ly_local, j_begin)) return e;
  REQUIRE(height);
  k_ic_rand<<<blocks_for(Lx, Ly_local), IB, 0, (cudaStream_t)stream>>>(height, h0, eps, make_philox_key(seed), Lx, Ly_local, j_begin);
  SW_LAUNCH_CHECK();
  return 0;
}

int swalbe_circshift(double *dst, const double *src, int sx, int sy, int Lx, int Ly, void *stream) {
  if (int e = check_extent(Lx, Ly)) return e;
  REQUIRE(dst); REQUIRE(src);
  if (dst == src) return set_error(SWALBE_ERR_ARG, "circshift!: dst must not alias src");
  sx %= Lx; sy %= Ly;
  if (sx < 0) sx += Lx;
  if (sy < 0) sy += Ly;
  k_circshift<<<blocks_for(Lx, Ly), IB, 0, (cudaStream_t)stream>>>(dst, src, sx, sy, Lx, Ly);
  SW_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
