// Persistent multi-step kernel for lattices that fit the shared memory of ONE thread-block cluster (<= ~128 x 128 sites;
// the reference's README example is 100 x 100, src/simulate.jl:338-358).
//
// At that size a time step is pure latency: one launch per step costs ~4.4 us even for a single tile (kernel ramp-up,
// global-memory round trips, launch gap), against < 1 us of dependent arithmetic.  Here the whole lattice lives in the
// distributed shared memory of a cluster of C CTAs for ALL the steps of a swalbe_time_loop call:
//
//   * CTA r of the cluster owns a row slab of the lattice (the multi-GPU decomposition in miniature);
//   * per step: fetch the 3 halo rows of h and 1 of u from the two neighbour CTAs' shared memory (DSMEM), then the
//     three dependent stencils as phases over the slab plus its halo -- pressure on rows+4, forces / equilibrium /
//     collision on rows+2, pull + moments on the slab -- with a __syncthreads between phases and ONE cluster barrier per
//     step; h and u are double-buffered so that a neighbour may still read step s while step s+1 is being written;
//   * global memory is touched at the first step (load), by the populations (write-only at tau == 1; every step, or
//     the last one only with SWALBE_LOOP_LAZY_POPULATIONS) and at the last step (h, u); per-step logs (min / max /
//     wetted count of the pre-step height: time_loop(sys, state, Δh), the wetted! callback) leave as per-CTA partial
//     results by plain stores and are folded by a one-block kernel after the loop (k_cluster_logs).
//
// Same site functions as every other kernel (common.cuh), so the fields are bit-identical to the per-step kernels.
// Strict lean only: tau == 1, scalar theta, standard slip, no noise, no inclination.
#pragma once
#include "fused.cuh"

#ifndef SW_HOST_EMULATION
#include <cooperative_groups.h>
#endif

namespace swalbe {

struct ClusterArgs {
  FusedArgs a;        // lattice, constants, h_in/ux_in/uy_in (read at step 0), h_out/ux_out/uy_out (written by the last step)
  int nsteps;
  int lazy;           // populations: 1 = written by the last step only
  int rows_max;       // ceil(Ly / cluster size): the slab height the shared-memory layout is sized for
  double *log_part;   // per-step logs (NULL = off): nsteps x cluster size x 3 partial results (min, max, count) of the
                      // pre-step height, written with plain stores; k_cluster_logs folds them into the caller's arrays
};

// rows [cl_row0(r), cl_row0(r+1)) of the lattice belong to CTA r
__host__ __device__ inline int cl_row0(int r, int C, int Ly) { return (int)(((long long)r * Ly) / C); }

// doubles of dynamic shared memory per CTA for an Lx-wide lattice whose tallest slab has R rows
constexpr int CLUSTER_RED = 96;  // reduction scratch of the per-step logs (3 x 32 warps at most)
constexpr size_t cluster_smem_doubles(int Lx, int R) {
  return (size_t)Lx * (2 * (R + 6) + 4 * (R + 2) + (R + 4) + 9 * (R + 2)) + CLUSTER_RED;
}

#ifndef SW_HOST_EMULATION  // (tests/simt_emulation.cpp supplies host versions of these four)
__device__ __forceinline__ unsigned cl_rank() { return cooperative_groups::this_cluster().block_rank(); }
__device__ __forceinline__ unsigned cl_size() { return cooperative_groups::this_cluster().num_blocks(); }
__device__ __forceinline__ const double *cl_map(const double *p, unsigned rank) {
  return cooperative_groups::this_cluster().map_shared_rank(const_cast<double *>(p), rank);
}
__device__ __forceinline__ void cl_sync() { cooperative_groups::this_cluster().sync(); }
#endif

template <int NT, int PM, bool GZ>
__global__ void __launch_bounds__(NT, 1) k_cluster_steps(const __grid_constant__ ClusterArgs ca) {
#ifdef SW_HOST_EMULATION
  double *const smem = emul_dynamic_smem();
#else
  extern __shared__ __align__(16) double smem[];
#endif
  const FusedArgs &a = ca.a;
  const int tid = threadIdx.x;
  const int Lx = a.Lx, Ly = a.Ly, R = ca.rows_max;
  const int C = (int)cl_size(), rank = (int)cl_rank();
  const int j0 = cl_row0(rank, C, Ly), rows = cl_row0(rank + 1, C, Ly) - j0;
  const int dn = (rank + C - 1) % C, up = (rank + 1) % C;  // owners of the rows below / above this slab
  const int rows_dn = cl_row0(dn + 1, C, Ly) - cl_row0(dn, C, Ly);

  // layout (every CTA the same, sized for R rows): h[2][(R+6) Lx] | ux[2][(R+2) Lx] | uy[2][(R+2) Lx] | p[(R+4) Lx] | f*[9][(R+2) Lx]
  const size_t nh = (size_t)(R + 6) * Lx, nu = (size_t)(R + 2) * Lx, np = (size_t)(R + 4) * Lx;
  double *const sh = smem, *const sux = sh + 2 * nh, *const suy = sux + 2 * nu, *const sp = suy + 2 * nu, *const sf = sp + np;
  double *const r_min = sf + 9 * nu, *const r_max = r_min + 32;
  unsigned int *const r_wet = reinterpret_cast<unsigned int *>(r_max + 32);
  // row `l` of the slab (l = -3 .. rows+2 for h, -1 .. rows for u) lives at index (l + 3) resp. (l + 1)
  auto H = [&](int buf, int l) { return sh + buf * nh + (size_t)(l + 3) * Lx; };
  auto UX = [&](int buf, int l) { return sux + buf * nu + (size_t)(l + 1) * Lx; };
  auto UY = [&](int buf, int l) { return suy + buf * nu + (size_t)(l + 1) * Lx; };

  // step 0: slab + halo straight from global memory (periodic rows)
  for (int idx = tid; idx < (rows + 6) * Lx; idx += NT) {
    const int l = idx / Lx - 3, i = idx - (l + 3) * Lx;
    H(0, l)[i] = a.h_in[(size_t)wrapi(j0 + l, Ly) * Lx + i];
  }
  for (int idx = tid; idx < (rows + 2) * Lx; idx += NT) {
    const int l = idx / Lx - 1, i = idx - (l + 1) * Lx;
    const size_t g = (size_t)wrapi(j0 + l, Ly) * Lx + i;
    UX(0, l)[i] = a.ux_in[g];
    UY(0, l)[i] = a.uy_in[g];
  }
  __syncthreads();

  const bool logging = ca.log_part != nullptr;
  // site index -> (row, column) without a division in the step loop: thread tid starts every phase at (tid / Lx, tid % Lx)
  // of the phase's row range and advances by NT sites = (NT / Lx) rows + (NT % Lx) columns with one carry
  const int l_first = tid / Lx, i_first = tid - l_first * Lx, dl = NT / Lx, di = NT - dl * Lx;
  const int nhi = (int)nh, nui = (int)nu;  // (all shared-memory offsets fit 32 bits)
  for (int s = 0; s < ca.nsteps; ++s) {
    const int cur = s & 1, nxt = cur ^ 1;
    const bool last = s == ca.nsteps - 1;
    if (s > 0) {
      // halo rows of this step's input from the neighbours' shared memory: what they wrote in phase D of step s-1
      cl_sync();
      const double *hd = cl_map(H(cur, rows_dn - 3), dn), *hu = cl_map(H(cur, 0), up);  // (same layout in every CTA)
      for (int idx = tid; idx < 3 * Lx; idx += NT) {
        H(cur, -3)[idx] = hd[idx];
        H(cur, rows)[idx] = hu[idx];
      }
      const double *xd = cl_map(UX(cur, rows_dn - 1), dn), *xu = cl_map(UX(cur, 0), up);
      const double *yd = cl_map(UY(cur, rows_dn - 1), dn), *yu = cl_map(UY(cur, 0), up);
      for (int idx = tid; idx < Lx; idx += NT) {
        UX(cur, -1)[idx] = xd[idx]; UX(cur, rows)[idx] = xu[idx];
        UY(cur, -1)[idx] = yd[idx]; UY(cur, rows)[idx] = yu[idx];
      }
      __syncthreads();
    }

    // phase B: film pressure on rows -2 .. rows+1   (src/pressure.jl:141-153; fused.cuh stage B)
    const double *const hb = sh + cur * nhi;  // this step's h buffer; row l at (l + 3) * Lx
    for (int idx = tid, lr = l_first, i = i_first; idx < (rows + 4) * Lx; idx += NT) {  // lr = l + 2
      const int im = i ? i - 1 : Lx - 1, ip = i + 1 < Lx ? i + 1 : 0;  // columns i-1 / i+1 (periodic)
      const double *r1 = hb + (lr + 1) * Lx, *r0 = r1 - Lx, *r2 = r1 + Lx;
      const double hc = r1[i];
      const double lap = lap9_bracket(hc, r1[im], r0[i], r1[ip], r2[i], r0[im], r0[ip], r2[ip], r2[im]);
      const double x = div_exact(a.pc.hmin, hc + a.pc.hcrit);
      const double pw = disjoining_powers(x, PM, a.pc.n, a.pc.m);
      sp[lr * Lx + i] = (-a.pc.gamma * (a.pc.kappa * pw)) - a.pc.gamma * lap;
      i += di; lr += dl;
      if (i >= Lx) { i -= Lx; ++lr; }
    }
    __syncthreads();

    // phase C: forces, equilibrium, collision on rows -1 .. rows   (fused.cuh stage C); logs of the pre-step height
    double d_min = INFINITY, d_max = -INFINITY;
    unsigned int d_wet = 0;
    const double *const uxb = sux + cur * nui, *const uyb = suy + cur * nui;  // row l at (l + 1) * Lx
    for (int idx = tid, lr = l_first, i = i_first; idx < (rows + 2) * Lx; idx += NT) {  // lr = l + 1
      const int l = lr - 1;
      const int im = i ? i - 1 : Lx - 1, ip = i + 1 < Lx ? i + 1 : 0;
      const double *q0 = sp + lr * Lx, *q1 = q0 + Lx, *q2 = q1 + Lx;  // p rows l-1, l, l+1
      const double hc = hb[(lr + 2) * Lx + i];
      const double ux = uxb[lr * Lx + i], uy = uyb[lr * Lx + i];
      const double pipjp = q0[im], pimjp = q0[ip], pimjm = q2[ip], pipjm = q2[im];
      const double gx = grad9_x(q1[im], q1[ip], pipjp, pimjp, pimjm, pipjm);
      const double gy = grad9_y(q0[i], q2[i], pipjp, pimjp, pimjm, pipjm);
      const double hgx = hc * gx, hgy = hc * gy;
      double sx, sy;
      slip_terms(hc, ux, uy, a.sc, SWALBE_SLIP_STANDARD, sx, sy);
      const double Fx = (-hgx) - sx, Fy = (-hgy) - sy;
      double fe[9], vsq, fs[9];
      equilibrium_site<GZ>(hc, ux, uy, a.ec, fe, vsq);
      collide_site_tau1(fe, Fx, Fy, fs);
#pragma unroll
      for (int k = 0; k < 9; ++k) sf[k * nui + lr * Lx + i] = fs[k];
      if (logging && l >= 0 && l < rows) {
        d_min = fmin(d_min, hc);
        d_max = fmax(d_max, hc);
        d_wet += hc > a.hthresh;
      }
      i += di; lr += dl;
      if (i >= Lx) { i -= Lx; ++lr; }
    }
    if (logging) {  // CTA reduction, one atomic per CTA and step (as the marching kernel does)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        d_min = fmin(d_min, __shfl_down_sync(0xffffffffu, d_min, o));
        d_max = fmax(d_max, __shfl_down_sync(0xffffffffu, d_max, o));
        d_wet += __shfl_down_sync(0xffffffffu, d_wet, o);
      }
      if ((tid & 31) == 0) { r_min[tid >> 5] = d_min; r_max[tid >> 5] = d_max; r_wet[tid >> 5] = d_wet; }
      __syncthreads();
      if (tid == 0) {  // (plain stores: nothing in the step loop waits on global memory)
        for (int w = 1; w < NT / 32; ++w) { d_min = fmin(d_min, r_min[w]); d_max = fmax(d_max, r_max[w]); d_wet += r_wet[w]; }
        double *o = ca.log_part + ((size_t)s * C + rank) * 3;
        o[0] = d_min; o[1] = d_max; o[2] = (double)d_wet;
      }
    } else {
      __syncthreads();
    }

    // phase D: pull-stream + moments of the slab   (fused.cuh stage D) -> the other h / u buffer
    if (ca.nsteps == 1) cl_sync();  // in-place calls: nobody may still be loading step 0 from the planes written below
    const bool write_f = a.f_out != nullptr && (last || !ca.lazy);
    double *const hnb = sh + nxt * nhi, *const uxnb = sux + nxt * nui, *const uynb = suy + nxt * nui;
    for (int idx = tid, l = l_first, i = i_first; idx < rows * Lx; idx += NT) {
      const int im = i ? i - 1 : Lx - 1, ip = i + 1 < Lx ? i + 1 : 0;
      const int c0 = l * Lx, c1 = c0 + Lx, c2 = c1 + Lx;  // f* rows l-1, l, l+1
      double fn[9];
      fn[0] = sf[0 * nui + c1 + i];
      fn[1] = sf[1 * nui + c1 + im]; fn[3] = sf[3 * nui + c1 + ip];
      fn[2] = sf[2 * nui + c0 + i];  fn[4] = sf[4 * nui + c2 + i];
      fn[5] = sf[5 * nui + c0 + im]; fn[6] = sf[6 * nui + c0 + ip];
      fn[7] = sf[7 * nui + c2 + ip]; fn[8] = sf[8 * nui + c2 + im];
      double hn, uxn, uyn;
      moments_site(fn, hn, uxn, uyn);
      hnb[(l + 3) * Lx + i] = hn; uxnb[(l + 1) * Lx + i] = uxn; uynb[(l + 1) * Lx + i] = uyn;
      const size_t o = (size_t)(j0 + l) * Lx + i;
      if (last) { a.h_out[o] = hn; a.ux_out[o] = uxn; a.uy_out[o] = uyn; }
      if (write_f) {
#pragma unroll
        for (int k = 0; k < 9; ++k) a.f_out[o + k * a.fstride_out] = fn[k];
        if (last && a.f_out2 != nullptr) {
#pragma unroll
          for (int k = 0; k < 9; ++k) a.f_out2[o + k * a.fstride_out2] = fn[k];
        }
      }
      i += di; l += dl;
      if (i >= Lx) { i -= Lx; ++l; }
    }
    __syncthreads();
  }
  cl_sync();  // a CTA must not exit while a neighbour may still read its shared memory
}

// folds the per-CTA partial logs of k_cluster_steps into the caller's per-step arrays
static __global__ void k_cluster_logs(const double *__restrict__ part, int C, int nsteps, double *__restrict__ log_min,
                               double *__restrict__ log_max, unsigned long long *__restrict__ log_wet) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nsteps) return;
  double mn = INFINITY, mx = -INFINITY, cnt = 0.0;
  for (int r = 0; r < C; ++r) {
    const double *o = part + ((size_t)s * C + r) * 3;
    mn = fmin(mn, o[0]); mx = fmax(mx, o[1]); cnt += o[2];
  }
  if (log_min != nullptr) { log_min[s] = mn; log_max[s] = mx; }
  if (log_wet != nullptr) log_wet[s] = (unsigned long long)cnt;
}

}  // namespace swalbe
