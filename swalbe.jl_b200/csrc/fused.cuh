// The fused thin-film LBM step kernel: pressure -> h∇p -> slip/thermal/force -> equilibrium -> BGK collide
// -> stream -> moments in ONE pass over HBM  (replaces the ~100 broadcast/circshift! launches of
// src/simulate.jl:15-22).
//
// Data movement per lattice update (tau = 1): read h,ux,uy (24 B), write 9 populations + h,ux,uy (96 B).
// For tau != 1 the nine old populations are read as well (+72 B).
//
// Mapping.  x (Julia's first index) is contiguous, so a CTA is a 1-D line of NT threads covering NT
// consecutive columns [s0-4, s0-4+NT) (periodic in x) and it MARCHES along y over `rows_per_cta` rows.
// Each thread owns one column.  The three dependent stencils are software-pipelined over rows:
//
//   iteration it:   load  h        row L = j0-3+it   (global -> register -> smem at end of iteration)
//                   B:    pressure row P = L-2       (needs h rows P-1..P+1, x-neighbours through smem)
//                   C:    F, feq, f*  row F = P-2    (needs p rows F-1..F+1, x-neighbours through smem)
//                   D:    pull + moments row O = F-2 (needs f* rows O-1..O+1, x-shifted ones through smem)
//
// Every value a thread needs from its own column stays in registers (3x3 windows of h and p, the parked
// populations); only x-neighbour values cross threads, through double-buffered one-row smem lines, so there
// is ONE __syncthreads per row and 8 smem stores + 10 smem loads per lattice update.  Redundant work is
// limited to the 8 halo columns per CTA (NT-8 of NT threads produce output) and the 9-row pipeline fill per
// chunk of rows; nothing is recomputed in y inside a chunk.
//
// All arithmetic comes from common.cuh (reference evaluation order, no FMA contraction).
#pragma once
#include "common.cuh"

namespace swalbe {

struct FusedArgs {
  // geometry
  int Lx, Ly;        // extents of the (local) lattice; Lx contiguous
  int jbeg, jend;    // rows produced by this launch: [jbeg, jend)
  int rows_per_cta;  // rows marched by one CTA
  int W;             // output columns per CTA (<= NT-8)
  int wrap_y;        // 1: rows are periodic modulo Ly (caller-owned, un-padded arrays); 0: ghost rows
  int gh_m, gh_f;    // ghost rows below row 0 in the moment planes / population planes (wrap_y == 0)
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

template <int NT, int MINB, bool TAU1, bool THERMAL>
__global__ void __launch_bounds__(NT, MINB) k_fused_step(const __grid_constant__ FusedArgs a) {
  __shared__ double s_h[2][NT + 2];
  __shared__ double s_p[2][NT + 2];
  __shared__ double s_f[2][6][NT + 2];  // f*1 f*5 f*8 (moving +x), f*3 f*6 f*7 (moving -x)

  const int tid = threadIdx.x;
  const int sm = tid + 1;
  const int s0 = blockIdx.x * a.W;
  const int ci = wrapi(s0 - 4 + tid, a.Lx);
  const bool col_out = tid >= 4 && tid < 4 + a.W && (s0 + tid - 4) < a.Lx;
  const int j0 = a.jbeg + blockIdx.y * a.rows_per_cta;
  const int R = min(a.rows_per_cta, a.jend - j0);
  const int Lx = a.Lx;

  if (tid < 2) {  // the two pad cells of every line are never written by the pipeline
    const int e = tid ? NT + 1 : 0;
    s_h[0][e] = s_h[1][e] = s_p[0][e] = s_p[1][e] = 0.0;
#pragma unroll
    for (int q = 0; q < 6; ++q) s_f[0][q][e] = s_f[1][q][e] = 0.0;
  }
  s_h[0][sm] = s_h[1][sm] = s_p[0][sm] = s_p[1][sm] = 0.0;
#pragma unroll
  for (int q = 0; q < 6; ++q) s_f[0][q][sm] = s_f[1][q][sm] = 0.0;
  __syncthreads();

  // physical row offsets (in elements) of logical row r
  auto mrow = [&](int r) -> size_t { return (size_t)(a.wrap_y ? wrapi(r, a.Ly) : r + a.gh_m) * (size_t)Lx; };
  auto frow = [&](int r) -> size_t { return (size_t)(a.wrap_y ? wrapi(r, a.Ly) : r + a.gh_f) * (size_t)Lx; };

  // 3x3 register windows: index [row][col], row 0 = j-1 (older), 2 = j+1 (newer); col 0 = i-1, 2 = i+1
  double h00 = 0, h01 = 0, h02 = 0, h10 = 0, h11 = 0, h12 = 0, h20 = 0, h21 = 0, h22 = 0;
  double p00 = 0, p01 = 0, p02 = 0, p10 = 0, p11 = 0, p12 = 0, p20 = 0, p21 = 0, p22 = 0;
  double h_old = 0;            // own-column h one row below the window (the row stage C works on)
  double hnew = 0, pnew = 0;   // own-column values produced in the previous iteration (now in smem)
  double pf0 = 0, pf2 = 0, pf4 = 0;  // own-column f*0, f*2, f*4 of the previous iteration
  // parked arrivals: B1 = (f1,f0,f3) of row q-1 ; A1/A2 = (f5,f2,f6) of rows q-1 / q-2
  double b1_1 = 0, b1_0 = 0, b1_3 = 0, a1_5 = 0, a1_2 = 0, a1_6 = 0, a2_5 = 0, a2_2 = 0, a2_6 = 0;
  double ux_c = 0, uy_c = 0;   // velocities of the row stage C works on (prefetched one iteration ahead)
  double ft_c[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) ft_c[k] = 0.0;

  double d_min = INFINITY, d_max = -INFINITY;
  unsigned int d_wet = 0;
  const bool logging = a.log_min != nullptr || a.log_wet != nullptr;

  const int n_it = R + 9;
  for (int it = 0; it < n_it; ++it) {
    const int pb = (it + 1) & 1, cb = it & 1;  // previous / current smem line
    // ---- global loads issued first (consumed at the end of this iteration / in the next one) -------
    double h_ld = 0.0, ux_n = 0.0, uy_n = 0.0, ct_n = 0.0;
    double ft_n[9];
    if (it <= R + 5) h_ld = __ldg(a.h_in + mrow(j0 - 3 + it) + ci);
    if (it >= 5 && it <= R + 6) {  // row F(it+1) = j0-6+it in [j0-1, j0+R]
      const size_t o = mrow(j0 - 6 + it) + ci;
      ux_n = __ldg(a.ux_in + o);
      uy_n = __ldg(a.uy_in + o);
      if (!TAU1) {
        const size_t of = frow(j0 - 6 + it) + ci;
#pragma unroll
        for (int k = 0; k < 9; ++k) ft_n[k] = __ldg(a.f_in + k * a.fstride_in + of);
      }
    }
    const bool doB = it >= 3 && it <= R + 6;   // row P(it) = j0-5+it in [j0-2, j0+R+1]
    if (a.ct_field != nullptr && doB) ct_n = __ldg(a.ct_field + mrow(j0 - 5 + it) + ci);

    // ---- h window <- row L(it-1) ------------------------------------------------------------------
    {
      const double hl = s_h[pb][sm - 1], hr = s_h[pb][sm + 1];
      h_old = h01;
      h00 = h10; h01 = h11; h02 = h12;
      h10 = h20; h11 = h21; h12 = h22;
      h20 = hl; h21 = hnew; h22 = hr;
    }
    // ---- stage B: film pressure at row P(it) (window centre) -------------------------------------
    double p_cur = 0.0;
    if (doB) {
      const double lap = lap9_bracket(h11, h10, h01, h12, h21, h00, h02, h22, h20);
      const double kappa = a.ct_field != nullptr ? kappa_from_field(ct_n, a.pc) : a.pc.kappa;
      p_cur = film_pressure(h11, lap, kappa, a.pc);
    }
    // ---- p window <- row P(it-1) ------------------------------------------------------------------
    {
      const double pl = s_p[pb][sm - 1], pr = s_p[pb][sm + 1];
      p00 = p10; p01 = p11; p02 = p12;
      p10 = p20; p11 = p21; p12 = p22;
      p20 = pl; p21 = pnew; p22 = pr;
    }
    // ---- stage C: forces, equilibrium, collision at row F(it) = j0-7+it ---------------------------
    double fs[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) fs[k] = 0.0;
    if (it >= 6 && it <= R + 7) {
      const double hc = h_old;
      const double gx = grad9_x(p10, p12, p00, p02, p22, p20);
      const double gy = grad9_y(p01, p21, p00, p02, p22, p20);
      const double hgx = hc * gx, hgy = hc * gy;
      double sx, sy;
      slip_terms(hc, ux_c, uy_c, a.sc, sx, sy);
      double Fx = (-hgx) - sx, Fy = (-hgy) - sy;
      double kx = 0.0, ky = 0.0;
      const int rF = j0 - 7 + it;
      if (THERMAL) {
        long long jg = a.jglobal0 + rF;
        jg %= a.Ly_global;
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
      double fe[9], vsq;
      equilibrium_site(hc, ux_c, uy_c, a.ec, fe, vsq);
      if (TAU1) collide_site_tau1(fe, Fx, Fy, fs);
      else collide_site(ft_c, fe, Fx, Fy, a.omega, a.invtau, fs);

      const bool own = col_out && it >= 7 && it <= R + 6;  // row F(it) in [j0, j0+R-1]
      if (own) {
        if (logging) {
          d_min = fmin(d_min, hc);
          d_max = fmax(d_max, hc);
          d_wet += hc > a.hthresh;
        }
        if (a.pressure != nullptr) {  // materialise the reference's intermediate fields (last step of a call)
          const size_t o = (size_t)rF * Lx + ci;
          const size_t N = (size_t)Lx * a.Ly;
          a.pressure[o] = p11;
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
    // ---- stage D: pull-stream + moments at row O(it) = j0-9+it ------------------------------------
    {
      // x-shifted arrivals of row q = F(it-1)
      const double r1 = s_f[pb][0][sm - 1], r5 = s_f[pb][1][sm - 1], r8 = s_f[pb][2][sm - 1];
      const double r3 = s_f[pb][3][sm + 1], r6 = s_f[pb][4][sm + 1], r7 = s_f[pb][5][sm + 1];
      if (it >= 9 && col_out) {
        double fn[9];
        fn[0] = b1_0; fn[1] = b1_1; fn[3] = b1_3;   // row q-1 (same row as the output)
        fn[5] = a2_5; fn[2] = a2_2; fn[6] = a2_6;   // row q-2 (moving +y)
        fn[7] = r7;   fn[4] = pf4;  fn[8] = r8;     // row q   (moving -y)
        double hn, uxn, uyn;
        moments_site(fn, hn, uxn, uyn);
        const int rO = j0 - 9 + it;
        const size_t om = mrow(rO) + ci;
        a.h_out[om] = hn; a.ux_out[om] = uxn; a.uy_out[om] = uyn;
        if (a.f_out != nullptr) {
          const size_t of = frow(rO) + ci;
#pragma unroll
          for (int k = 0; k < 9; ++k) a.f_out[k * a.fstride_out + of] = fn[k];
          if (a.f_out2 != nullptr) {
#pragma unroll
            for (int k = 0; k < 9; ++k) a.f_out2[k * a.fstride_out2 + of] = fn[k];
          }
        }
      }
      a2_5 = a1_5; a2_2 = a1_2; a2_6 = a1_6;
      a1_5 = r5; a1_2 = pf2; a1_6 = r6;
      b1_1 = r1; b1_0 = pf0; b1_3 = r3;
    }
    // ---- publish this iteration's products --------------------------------------------------------
    s_h[cb][sm] = h_ld;
    s_p[cb][sm] = p_cur;
    s_f[cb][0][sm] = fs[1]; s_f[cb][1][sm] = fs[5]; s_f[cb][2][sm] = fs[8];
    s_f[cb][3][sm] = fs[3]; s_f[cb][4][sm] = fs[6]; s_f[cb][5][sm] = fs[7];
    pf0 = fs[0]; pf2 = fs[2]; pf4 = fs[4];
    hnew = h_ld; pnew = p_cur;
    ux_c = ux_n; uy_c = uy_n;
    if (!TAU1) {
#pragma unroll
      for (int k = 0; k < 9; ++k) ft_c[k] = ft_n[k];
    }
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
