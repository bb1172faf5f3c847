// Table-driven Box-Muller: 128 random bits -> (-2 ln u1, cos φ, sin φ).
//
// The thermal noise is compared with the reference statistically only (Julia's randn! stream cannot be reproduced), so
// the normals need to be good, not bit-compatible with anything.  libdevice's log + sincospi cost ~110 FP64-pipe
// instructions per pair; here both are one 16-byte shared-memory table lookup plus a short polynomial (~35
// instructions) with absolute errors < 1e-15.  This header also compiles as plain C++ (no CUDA), which is how
// tools/normal_host_test.cpp checks every output against long double and the moments / tails of 2e8 deviates.
#pragma once
#include <math.h>
#include <stdint.h>
#ifdef __CUDACC__
#define SW_HD __host__ __device__ __forceinline__
#else
#define SW_HD inline
#ifndef __VECTOR_TYPES_H__  // (plain C++ without the CUDA headers: tools/normal_host_test.cpp)
struct double2 {
  double x, y;
};
#endif
#endif

namespace swalbe {

constexpr int NRM_LOG_BITS = 7, NRM_LOG_N = 1 << NRM_LOG_BITS;  // log table: 128 intervals of the mantissa [1,2)
constexpr int NRM_ANG_BITS = 6, NRM_ANG_N = 1 << NRM_ANG_BITS;  // angle table: 64 sectors of the circle
struct NormalTables {
  double2 lg[NRM_LOG_N];  // (1/c_k rounded to float, log(c_k)),  c_k ~ 1 + (k+1/2)/128
  double2 sc[NRM_ANG_N];  // (cos, sin) of the sector centres 2π(k+1/2)/64
};

// polynomial coefficients: in device code they sit in the constant bank (operands of the FP64 instructions, no moves)
struct NrmConsts {
  double l6, l5, l3, mln2, dsc, s7, s5, s3, c8, c6, c4;
};
#define NRM_CONSTS_INIT                                                                                              \
  {-1.0 / 6.0,    0.2,         1.0 / 3.0,  -0.6931471805599453, 6.283185307179586 / 4294967296.0, -1.0 / 5040.0, \
   1.0 / 120.0,   -1.0 / 6.0,  1.0 / 40320.0, -1.0 / 720.0,     1.0 / 24.0}
#ifdef __CUDACC__
static __constant__ NrmConsts nrm_kd = NRM_CONSTS_INIT;
#endif
static const NrmConsts nrm_kh = NRM_CONSTS_INIT;
#if defined(__CUDA_ARCH__) && !defined(NRM_LITERAL_CONSTS)
#define NRM_K ::swalbe::nrm_kd
#else
#define NRM_K ::swalbe::nrm_kh
#endif

SW_HD void normal_table_entry(NormalTables &T, int k) {
  if (k < NRM_LOG_N) {
    const float invc = 1.0f / (1.0f + ((float)k + 0.5f) * (1.0f / NRM_LOG_N));  // any float near 1/c_k will do
    T.lg[k].x = (double)invc;
    T.lg[k].y = -log((double)invc);
  } else {
    const double a = ((double)(k - NRM_LOG_N) + 0.5) * (2.0 / NRM_ANG_N);  // angle / π
    double sn, cs;
#ifdef __CUDA_ARCH__
    sincospi(a, &sn, &cs);
#else
    sn = sin(a * 3.141592653589793);
    cs = cos(a * 3.141592653589793);
#endif
    T.sc[k - NRM_LOG_N].x = cs;
    T.sc[k - NRM_LOG_N].y = sn;
  }
}
// every thread of the CTA calls this once; __syncthreads() before the first use
SW_HD void normal_tables_fill(NormalTables &T, int tid, int nthreads) {
  for (int k = tid; k < NRM_LOG_N + NRM_ANG_N; k += nthreads) normal_table_entry(T, k);
}

SW_HD int nrm_clz32(uint32_t x) {
#ifdef __CUDA_ARCH__
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}
SW_HD double nrm_hilo(uint32_t hi, uint32_t lo) {
#ifdef __CUDA_ARCH__
  return __hiloint2double((int)hi, (int)lo);
#else
  union {
    uint64_t u;
    double d;
  } v;
  v.u = ((uint64_t)hi << 32) | lo;
  return v.d;
#endif
}

// u1 = 2^-(e+1) m is uniform on (0,1) with a full 52-bit mantissa m at every magnitude down to 2^-45 (|z| up to 7.9):
// e is geometric (leading zeros of r0, extended by 12 bits of r2 when r0 == 0).  φ = 2π (r3 + 1/2) / 2^32.
SW_HD void normal_polar_from_bits(const uint32_t r[4], const NormalTables &T, double &m2lnu, double &c, double &s) {
  int e = nrm_clz32(r[0]);
  if (e == 32) e += nrm_clz32(((r[2] & 0xfffu) << 20) | 0x80000u);
  const double m = nrm_hilo(0x3ff00000u | (r[1] >> 12), (r[1] << 20) | (r[2] >> 12));
  const double2 lt = T.lg[r[1] >> (32 - NRM_LOG_BITS)];
  const double x = fma(m, lt.x, -1.0);  // m/c - 1, |x| < 2^-7.9;  log1p(x) to x^6
  double p = fma(x, NRM_K.l6, NRM_K.l5);
  p = fma(x, p, -0.25);
  p = fma(x, p, NRM_K.l3);
  p = fma(x, p, -0.5);
  const double ln_m = lt.y + fma(x * x, p, x);
  m2lnu = -2.0 * fma((double)(e + 1), NRM_K.mln2, ln_m);
  const double2 cs = T.sc[r[3] >> (32 - NRM_ANG_BITS)];
  const int rho = (int)(r[3] & ((1u << (32 - NRM_ANG_BITS)) - 1u)) - (1 << (31 - NRM_ANG_BITS));
  const double d = ((double)rho + 0.5) * NRM_K.dsc;  // offset from the sector centre
  const double d2 = d * d;
  double ps = fma(d2, NRM_K.s7, NRM_K.s5);
  ps = fma(d2, ps, NRM_K.s3);
  const double sd = fma(d * d2, ps, d);  // sin δ
  double pc = fma(d2, NRM_K.c8, NRM_K.c6);
  pc = fma(d2, pc, NRM_K.c4);
  pc = fma(d2, pc, -0.5);
  const double cm1 = d2 * pc;  // cos δ - 1
  c = fma(-cs.y, sd, fma(cs.x, cm1, cs.x));
  s = fma(cs.x, sd, fma(cs.y, cm1, cs.y));
}

}  // namespace swalbe
