// Small-lattice flavour of the fused step (strict lean: tau == 1, scalar theta, standard slip, no noise / logs).
//
// The marching kernel (fused.cuh) pipelines rows through a CTA: R + 11 dependent iterations with a barrier each, which
// is what a lattice that fits one wave pays as pure latency (100^2: 6.2 us per step).  Here a CTA owns a 32 x 8 tile and
// runs the three dependent stencils as three phases over the tile plus its halo (h: 3 cells, p: 2, f*: 1) in shared
// memory -- three barriers per step instead of R + 11, at the price of recomputing the halo (1.7x pressures, 1.3x
// collisions).  Same site functions, same operation order: bit-identical to the marching kernel.  Used below
// SWALBE_TILE_MAX lattice sites (fused.cu), where latency, not HBM, bounds the step.
#pragma once
#include "fused.cuh"

namespace swalbe {
namespace {

constexpr int TX = 32, TY = 8, TT = TX * TY;

__device__ __forceinline__ int wrapm(int v, int L) {
  v %= L;
  return v < 0 ? v + L : v;
}

// TF: compiled for a contact-angle FIELD (a.ct_field != NULL): cospi(theta) is staged on the tile + 2 next to h
template <int PM, bool GZ, bool TF = false>
__global__ void __launch_bounds__(TT) k_tile_step(const __grid_constant__ FusedArgs a) {
  __shared__ double sh[TY + 6][TX + 6];
  __shared__ double sux[TY + 2][TX + 2], suy[TY + 2][TX + 2];
  __shared__ double sp[TY + 4][TX + 4];
  __shared__ double sf[9][TY + 2][TX + 2];
  __shared__ double sct[TF ? TY + 4 : 1][TF ? TX + 4 : 1];
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int Lx = a.Lx, Ly = a.Ly;

  // phase 1: h on the tile + 3, u on the tile + 1 (periodic)
  for (int idx = tid; idx < (TY + 6) * (TX + 6); idx += TT) {
    const int ly = idx / (TX + 6), lx = idx - ly * (TX + 6);
    sh[ly][lx] = a.h_in[(size_t)wrapm(y0 - 3 + ly, Ly) * Lx + wrapm(x0 - 3 + lx, Lx)];
  }
  for (int idx = tid; idx < (TY + 2) * (TX + 2); idx += TT) {
    const int ly = idx / (TX + 2), lx = idx - ly * (TX + 2);
    const size_t g = (size_t)wrapm(y0 - 1 + ly, Ly) * Lx + wrapm(x0 - 1 + lx, Lx);
    sux[ly][lx] = a.ux_in[g];
    suy[ly][lx] = a.uy_in[g];
  }
  if (TF) {
    for (int idx = tid; idx < (TY + 4) * (TX + 4); idx += TT) {
      const int ly = idx / (TX + 4), lx = idx - ly * (TX + 4);
      sct[ly][lx] = a.ct_field[(size_t)wrapm(y0 - 2 + ly, Ly) * Lx + wrapm(x0 - 2 + lx, Lx)];
    }
  }
  __syncthreads();

  // phase 2: film pressure on the tile + 2   (src/pressure.jl:141-153; same expression as fused.cuh stage B)
  for (int idx = tid; idx < (TY + 4) * (TX + 4); idx += TT) {
    const int ly = idx / (TX + 4), lx = idx - ly * (TX + 4);
    const double hc = sh[ly + 1][lx + 1];
    const double lap = lap9_bracket(hc, sh[ly + 1][lx], sh[ly][lx + 1], sh[ly + 1][lx + 2], sh[ly + 2][lx + 1], sh[ly][lx],
                                    sh[ly][lx + 2], sh[ly + 2][lx + 2], sh[ly + 2][lx]);
    const double x = div_exact(a.pc.hmin, hc + a.pc.hcrit);
    const double pw = disjoining_powers(x, PM, a.pc.n, a.pc.m);
    const double kappa = TF ? kappa_from_field(sct[ly][lx], a.pc) : a.pc.kappa;
    sp[ly][lx] = (-a.pc.gamma * (kappa * pw)) - a.pc.gamma * lap;
  }
  __syncthreads();

  // phase 3: forces, equilibrium, collision on the tile + 1   (fused.cuh stage C)
  for (int idx = tid; idx < (TY + 2) * (TX + 2); idx += TT) {
    const int ly = idx / (TX + 2), lx = idx - ly * (TX + 2);
    const double hc = sh[ly + 2][lx + 2];
    const double pipjp = sp[ly][lx], pimjp = sp[ly][lx + 2], pimjm = sp[ly + 2][lx + 2], pipjm = sp[ly + 2][lx];
    const double gx = grad9_x(sp[ly + 1][lx], sp[ly + 1][lx + 2], pipjp, pimjp, pimjm, pipjm);
    const double gy = grad9_y(sp[ly][lx + 1], sp[ly + 2][lx + 1], pipjp, pimjp, pimjm, pipjm);
    const double hgx = hc * gx, hgy = hc * gy;
    const double ux = sux[ly][lx], uy = suy[ly][lx];
    double sx, sy;
    slip_terms(hc, ux, uy, a.sc, SWALBE_SLIP_STANDARD, sx, sy);
    const double Fx = (-hgx) - sx, Fy = (-hgy) - sy;
    double fe[9], vsq, fs[9];
    equilibrium_site<GZ>(hc, ux, uy, a.ec, fe, vsq);
    collide_site_tau1(fe, Fx, Fy, fs);
#pragma unroll
    for (int k = 0; k < 9; ++k) sf[k][ly][lx] = fs[k];
  }
  __syncthreads();

  // phase 4: pull-stream + moments of the tile   (fused.cuh stage D)
  const int ty = tid / TX, tx = tid - ty * TX;
  const int gx = x0 + tx, gy = y0 + ty;
  if (gx >= Lx || gy >= Ly) return;
  double fn[9];
  fn[0] = sf[0][ty + 1][tx + 1];
  fn[1] = sf[1][ty + 1][tx];      fn[3] = sf[3][ty + 1][tx + 2];
  fn[2] = sf[2][ty][tx + 1];      fn[4] = sf[4][ty + 2][tx + 1];
  fn[5] = sf[5][ty][tx];          fn[6] = sf[6][ty][tx + 2];
  fn[7] = sf[7][ty + 2][tx + 2];  fn[8] = sf[8][ty + 2][tx];
  double hn, uxn, uyn;
  moments_site(fn, hn, uxn, uyn);
  const size_t o = (size_t)gy * Lx + gx;
  a.h_out[o] = hn; a.ux_out[o] = uxn; a.uy_out[o] = uyn;
  if (a.f_out != nullptr) {
#pragma unroll
    for (int k = 0; k < 9; ++k) a.f_out[o + k * a.fstride_out] = fn[k];
    if (a.f_out2 != nullptr) {
#pragma unroll
      for (int k = 0; k < 9; ++k) a.f_out2[o + k * a.fstride_out2] = fn[k];
    }
  }
}

}  // namespace
}  // namespace swalbe
