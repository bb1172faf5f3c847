// The fused thin-film LBM step kernel: pressure -> h∇p -> slip/thermal/force -> equilibrium -> BGK collide
// -> stream -> moments in ONE pass over HBM  (replaces the ~100 broadcast/circshift! launches of
// src/simulate.jl:15-22).
//
// Data movement per lattice update (tau = 1): read h,ux,uy (24 B), write 9 populations + h,ux,uy (96 B).
// For tau != 1 the nine old populations are read as well (+72 B).
//
// Mapping.  x (Julia's first index) is contiguous, so a CTA is a 1-D line of NT threads covering NT consecutive
// columns [s0-4, s0-4+NT) (periodic in x); each thread owns one column and the CTA MARCHES along y over
// `rows_per_cta` rows.  The three dependent stencils are software-pipelined over rows through shared-memory
// rings, one __syncthreads per row:
//
//   iteration t:  prefetch  h row N(t+D), u rows F(t+D)   global -> smem with cp.async (LDGSTS), D rows ahead,
//                                                          so no thread ever waits on a global load
//                 B: film pressure   row P = j0-5+t        reads h rows P-1..P+1 from the h ring     -> p ring
//                 C: forces, feq, f* row F = j0-7+t        reads p rows F-1..F+1, h and u of row F   -> f* rings
//                 D: pull + moments  row O = j0-9+t        reads f* rows O-1..O+1 (x-shifted)        -> HBM
//
// B, C and D of one iteration read only rows written in EARLIER iterations, so the three instruction streams are
// independent and interleave freely (ILP); nothing but addresses lives in registers across iterations except the
// three own-column populations f*0, f*2, f*4.  Redundant work: the 8 halo columns per CTA and the 9-row pipeline
// fill per chunk of rows; nothing is recomputed in y inside a chunk.
//
// All arithmetic comes from common.cuh (reference evaluation order, no FMA contraction).
#pragma once
#include <limits.h>

#include "common.cuh"

namespace swalbe {

struct FusedArgs {
  // geometry
  int Lx, Ly;        // extents of the (local) lattice; Lx contiguous
  int jbeg, jend;    // rows produced by this launch: [jbeg, jend)
  int rows_per_cta;  // rows marched by one CTA
  int W;             // output columns per CTA (<= NT-8)
  int wrap_y;        // 1: rows are periodic modulo Ly (caller-owned, un-padded arrays)
                     // 0: ghost rows; the pointers below already point at logical row 0
  long long jglobal0;  // global index of local row 0 (thermal-noise counter)
  long long Ly_global;
  size_t fstride_in, fstride_out, fstride_out2;  // plane strides (elements) of the population arrays
  // inputs
  const double *h_in, *ux_in, *uy_in, *f_in, *ct_field;
  // outputs
  double *h_out, *ux_out, *uy_out, *f_out, *f_out2;
  // optional materialisation of the reference's intermediate fields (un-padded Lx*Ly, NULL = skip)
  double *pressure, *hgx, *hgy, *slipx, *slipy, *Fx, *Fy, *feq, *vsq, *kbtx, *kbty;
  // constants
  PressureConsts pc;
  SlipConsts sc;
  EqConsts ec;
  ThermalConsts tc;
  double omega, invtau;
  double incl_ax, incl_ay, incl_factor;
  int use_incl;
  unsigned long long seed, step;
  // per-step logs (slot pointers for THIS step, NULL = off)
  double *log_min, *log_max;
  unsigned long long *log_wet;
  double hthresh;
};

__device__ __forceinline__ void atomic_min_double(double *addr, double v) {
  unsigned long long *p = (unsigned long long *)addr;
  unsigned long long old = *p;
  while (v < __longlong_as_double((long long)old)) {
    unsigned long long assumed = old;
    old = atomicCAS(p, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
}
__device__ __forceinline__ void atomic_max_double(double *addr, double v) {
  unsigned long long *p = (unsigned long long *)addr;
  unsigned long long old = *p;
  while (v > __longlong_as_double((long long)old)) {
    unsigned long long assumed = old;
    old = atomicCAS(p, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
}

// 8-byte asynchronous global -> shared copy (LDGSTS); completion is tracked per thread by commit/wait groups,
// not by the register scoreboard.
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// A row cursor: logical row -> element offset of its first column, advanced one row per iteration with the
// periodic wrap folded in (no integer division in the loop).  All of it is CTA-uniform.
struct RowCursor {
  int r;          // physical row
  long long off;  // r * Lx
  __device__ __forceinline__ void init(int logical, int Lx, int Ly, int wrap_y) {
    r = wrap_y ? wrapi(logical, Ly) : logical;
    off = (long long)r * Lx;
  }
  __device__ __forceinline__ void advance(int Lx, int wrapLy) {
    ++r;
    off += Lx;
    if (r == wrapLy) { r = 0; off = 0; }
  }
};

constexpr int FUSED_D = 3;   // cp.async prefetch distance in rows
constexpr int FUSED_SH = 8;  // h ring slots (rows N(t-3) .. N(t+D) are live: D+4 <= 8)
constexpr int FUSED_SU = 4;  // u ring slots (rows F(t) .. F(t+D): D+1 <= 4)

constexpr size_t fused_smem_doubles(int NT) {
  // h ring + p ring + (f1,f3)[4] + (f5,f6)[4] + (f7,f8)[2] + u ring
  return (size_t)(FUSED_SH + 4 + 8 + 8 + 4) * (NT + 2) + (size_t)FUSED_SU * 2 * NT;
}

template <int NT, int MINB, bool TAU1, bool THERMAL>
__global__ void __launch_bounds__(NT, MINB) k_fused_step(const __grid_constant__ FusedArgs a) {
  extern __shared__ __align__(16) double smem[];
  constexpr int D = FUSED_D, SH = FUSED_SH, SU = FUSED_SU, LW = NT + 2;
  double *const s_h = smem;               // [SH][LW]   row N(t)  <-> slot t & 7
  double *const s_p = s_h + SH * LW;      // [4][LW]    row P(t)  <-> slot t & 3
  double *const s_f13 = s_p + 4 * LW;     // [4][2][LW] f*1,f*3 of row F(t) <-> slot t & 3   (read at t+2)
  double *const s_f56 = s_f13 + 8 * LW;   // [4][2][LW] f*5,f*6             <-> slot t & 3   (read at t+3)
  double *const s_f78 = s_f56 + 8 * LW;   // [2][2][LW] f*7,f*8             <-> slot t & 1   (read at t+1)
  double *const s_u = s_f78 + 4 * LW;     // [SU][2][NT] ux,uy of row F(t)  <-> slot t & 3

  const int tid = threadIdx.x;
  const int sm = tid + 1;
  const int s0 = blockIdx.x * a.W;
  const int Lx = a.Lx;
  const int ci = wrapi(s0 - 4 + tid, Lx);
  const bool col_out = tid >= 4 && tid < 4 + a.W && (s0 + tid - 4) < Lx;
  const int j0 = a.jbeg + blockIdx.y * a.rows_per_cta;
  const int R = min(a.rows_per_cta, a.jend - j0);
  const int wrapLy = a.wrap_y ? a.Ly : INT_MAX;

  if (tid < 2) {  // the two pad cells of every line are never written by the pipeline; keep them finite
    const int e = tid ? NT + 1 : 0;
    for (int q = 0; q < SH + 4 + 8 + 8 + 4; ++q) smem[q * LW + e] = 0.0;
  }

  RowCursor cN, cU, cO, cC, cT;
  cN.init(j0 - 4 - D + D, Lx, a.Ly, a.wrap_y);      // first prefetched h row: N(-D + D) = j0-4
  cU.init(j0 - 7 - D + D, Lx, a.Ly, a.wrap_y);      // first prefetched u row: F(-D + D) = j0-7
  cO.init(j0 - 9, Lx, a.Ly, a.wrap_y);              // output row O(0)
  cC.init(j0 - 5, Lx, a.Ly, a.wrap_y);              // cospi(theta) field row P(0)
  cT.init(j0 - 6, Lx, a.Ly, a.wrap_y);              // old populations (tau != 1): row F(1)

  // own-column populations that move along y only
  double f0_a = 0, f0_b = 0, f2_a = 0, f2_b = 0, f2_c = 0, f4_a = 0;
  double ft_c[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) ft_c[k] = 0.0;

  double d_min = INFINITY, d_max = -INFINITY;
  unsigned int d_wet = 0;
  const bool logging = a.log_min != nullptr || a.log_wet != nullptr;
  const bool theta_field = a.ct_field != nullptr;

  const int t_end = R + 8;
  for (int t = -D; t <= t_end; ++t) {
    // ---- asynchronous prefetch, D rows ahead -------------------------------------------------------
    {
      const int tn = t + D;               // h row N(tn) = j0-4+tn is needed for tn in [1, R+6]
      if (tn >= 1 && tn <= R + 6) cp_async8(s_h + (tn & (SH - 1)) * LW + sm, a.h_in + cN.off + ci);
      cN.advance(Lx, wrapLy);
      if (tn >= 6 && tn <= R + 7) {       // u rows F(tn) = j0-7+tn in [j0-1, j0+R]
        double *dst = s_u + (tn & (SU - 1)) * 2 * NT + tid;
        cp_async8(dst, a.ux_in + cU.off + ci);
        cp_async8(dst + NT, a.uy_in + cU.off + ci);
      }
      cU.advance(Lx, wrapLy);
      cp_async_commit();
    }
    if (t >= 0) {
      // ---- register prefetch of data that is not staged through smem ----------------------------------
      double ct_c = 0.0;
      const bool doB = t >= 3 && t <= R + 6;   // row P(t) = j0-5+t in [j0-2, j0+R+1]
      if (theta_field && doB) ct_c = __ldg(a.ct_field + cC.off + ci);
      cC.advance(Lx, wrapLy);
      double ft_n[9];
      if (!TAU1) {
        if (t >= 5 && t <= R + 6) {  // row F(t+1) = j0-6+t in [j0-1, j0+R]
#pragma unroll
          for (int k = 0; k < 9; ++k) ft_n[k] = __ldg(a.f_in + k * a.fstride_in + cT.off + ci);
        }
        cT.advance(Lx, wrapLy);
      }

      // ---- stage B: film pressure at row P(t) ----------------------------------------------------------
      if (doB) {
        const double *r0 = s_h + ((t - 2) & (SH - 1)) * LW + sm;  // row P-1
        const double *r1 = s_h + ((t - 1) & (SH - 1)) * LW + sm;  // row P
        const double *r2 = s_h + (t & (SH - 1)) * LW + sm;        // row P+1
        const double hc = r1[0];
        const double lap = lap9_bracket(hc, r1[-1], r0[0], r1[1], r2[0], r0[-1], r0[1], r2[1], r2[-1]);
        const double kappa = theta_field ? kappa_from_field(ct_c, a.pc) : a.pc.kappa;
        s_p[(t & 3) * LW + sm] = film_pressure(hc, lap, kappa, a.pc);
      }

      // ---- stage C: forces, equilibrium, collision at row F(t) = j0-7+t -----------------------------------
      double fs0 = 0.0, fs2 = 0.0, fs4 = 0.0;
      if (t >= 6 && t <= R + 7) {
        const double *q0 = s_p + ((t - 3) & 3) * LW + sm;  // row F-1
        const double *q1 = s_p + ((t - 2) & 3) * LW + sm;  // row F
        const double *q2 = s_p + ((t - 1) & 3) * LW + sm;  // row F+1
        const double hc = s_h[((t - 3) & (SH - 1)) * LW + sm];
        const double ux_c = s_u[(t & (SU - 1)) * 2 * NT + tid];
        const double uy_c = s_u[(t & (SU - 1)) * 2 * NT + NT + tid];
        const double pipjp = q0[-1], pimjp = q0[1], pimjm = q2[1], pipjm = q2[-1];
        const double gx = grad9_x(q1[-1], q1[1], pipjp, pimjp, pimjm, pipjm);
        const double gy = grad9_y(q0[0], q2[0], pipjp, pimjp, pimjm, pipjm);
        const double hgx = hc * gx, hgy = hc * gy;
        double sx, sy;
        slip_terms(hc, ux_c, uy_c, a.sc, sx, sy);
        double Fx = (-hgx) - sx, Fy = (-hgy) - sy;
        double kx = 0.0, ky = 0.0;
        const int rF = j0 - 7 + t;
        if (THERMAL) {
          long long jg = (a.jglobal0 + rF) % a.Ly_global;
          if (jg < 0) jg += a.Ly_global;
          double n1, n2;
          normal_pair(a.seed, a.step, (unsigned long long)ci + (unsigned long long)Lx * (unsigned long long)jg, n1, n2);
          const double amp = thermal_amplitude(hc, a.tc);
          kx = n1 * amp;
          ky = n2 * amp;
          Fx = Fx - kx;
          Fy = Fy - ky;
        }
        if (a.use_incl) {
          Fx = Fx + (hc * a.incl_ax) * a.incl_factor;
          Fy = Fy + (hc * a.incl_ay) * a.incl_factor;
        }
        double fe[9], vsq, fs[9];
        equilibrium_site(hc, ux_c, uy_c, a.ec, fe, vsq);
        if (TAU1) collide_site_tau1(fe, Fx, Fy, fs);
        else collide_site(ft_c, fe, Fx, Fy, a.omega, a.invtau, fs);
        double *w13 = s_f13 + (t & 3) * 2 * LW + sm, *w56 = s_f56 + (t & 3) * 2 * LW + sm;
        double *w78 = s_f78 + (t & 1) * 2 * LW + sm;
        w13[0] = fs[1]; w13[LW] = fs[3];
        w56[0] = fs[5]; w56[LW] = fs[6];
        w78[0] = fs[7]; w78[LW] = fs[8];
        fs0 = fs[0]; fs2 = fs[2]; fs4 = fs[4];

        const bool own = col_out && t >= 7 && t <= R + 6;  // row F(t) in [j0, j0+R-1]
        if (own) {
          if (logging) {
            d_min = fmin(d_min, hc);
            d_max = fmax(d_max, hc);
            d_wet += hc > a.hthresh;
          }
          if (a.pressure != nullptr) {  // materialise the reference's intermediate fields (last step of a call)
            const size_t o = (size_t)rF * Lx + ci;
            const size_t N = (size_t)Lx * a.Ly;
            a.pressure[o] = q1[0];
            a.hgx[o] = hgx; a.hgy[o] = hgy;
            a.slipx[o] = sx; a.slipy[o] = sy;
            a.Fx[o] = Fx; a.Fy[o] = Fy;
            a.vsq[o] = vsq;
#pragma unroll
            for (int k = 0; k < 9; ++k) a.feq[o + k * N] = fe[k];
            if (THERMAL && a.kbtx != nullptr) { a.kbtx[o] = kx; a.kbty[o] = ky; }
          }
        }
      }

      // ---- stage D: pull-stream + moments at row O(t) = j0-9+t ---------------------------------------------
      if (t >= 9 && col_out) {
        const double *m = s_f13 + ((t - 2) & 3) * 2 * LW + sm;  // row O
        const double *b = s_f56 + ((t - 3) & 3) * 2 * LW + sm;  // row O-1 (moving +y)
        const double *u = s_f78 + ((t - 1) & 1) * 2 * LW + sm;  // row O+1 (moving -y)
        double fn[9];
        fn[0] = f0_b;      fn[1] = m[-1];      fn[3] = m[LW + 1];
        fn[2] = f2_c;      fn[5] = b[-1];      fn[6] = b[LW + 1];
        fn[4] = f4_a;      fn[7] = u[1];       fn[8] = u[LW - 1];
        double hn, uxn, uyn;
        moments_site(fn, hn, uxn, uyn);
        const long long o = cO.off + ci;
        a.h_out[o] = hn; a.ux_out[o] = uxn; a.uy_out[o] = uyn;
        if (a.f_out != nullptr) {
#pragma unroll
          for (int k = 0; k < 9; ++k) a.f_out[(long long)(k * a.fstride_out) + o] = fn[k];
          if (a.f_out2 != nullptr) {
#pragma unroll
            for (int k = 0; k < 9; ++k) a.f_out2[(long long)(k * a.fstride_out2) + o] = fn[k];
          }
        }
      }
      cO.advance(Lx, wrapLy);
      f0_b = f0_a; f0_a = fs0;
      f2_c = f2_b; f2_b = f2_a; f2_a = fs2;
      f4_a = fs4;
      if (!TAU1) {
#pragma unroll
        for (int k = 0; k < 9; ++k) ft_c[k] = ft_n[k];
      }
    }
    cp_async_wait<D - 1>();  // the group issued D-1 iterations ago (h row N(t+1), u rows F(t+1)) has landed
    __syncthreads();
  }

  if (logging) {  // CTA reduction of the pre-step height statistics, one atomic per CTA
    __shared__ double r_min[NT / 32], r_max[NT / 32];
    __shared__ unsigned int r_wet[NT / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      d_min = fmin(d_min, __shfl_down_sync(0xffffffffu, d_min, o));
      d_max = fmax(d_max, __shfl_down_sync(0xffffffffu, d_max, o));
      d_wet += __shfl_down_sync(0xffffffffu, d_wet, o);
    }
    if ((tid & 31) == 0) { r_min[tid >> 5] = d_min; r_max[tid >> 5] = d_max; r_wet[tid >> 5] = d_wet; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < NT / 32; ++w) { d_min = fmin(d_min, r_min[w]); d_max = fmax(d_max, r_max[w]); d_wet += r_wet[w]; }
      if (a.log_min != nullptr) { atomic_min_double(a.log_min, d_min); atomic_max_double(a.log_max, d_max); }
      if (a.log_wet != nullptr) atomicAdd(a.log_wet, (unsigned long long)d_wet);
    }
  }
}

}  // namespace swalbe
