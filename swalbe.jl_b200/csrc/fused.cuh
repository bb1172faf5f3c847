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
//   iteration t:  prefetch  h row N(t+D), u rows F(t+D)   global -> smem D rows ahead, asynchronously: per-thread
//                                                          cp.async (LDGSTS), or -- BULK flavour, large even-Lx
//                                                          lattices -- cp.async.bulk through the TMA unit, one
//                                                          thread per row segment, completion on an mbarrier;
//                                                          no thread ever waits on a global load
//                 B: film pressure   row P = j0-5+t        reads h rows P-1..P+1 from the h ring     -> p ring
//                 C: forces, feq, f* row F = j0-7+t        reads p rows F-1..F+1, h and u of row F   -> f* rings
//                 D: pull + moments  row O = j0-9+t        reads f* rows O-1..O+1 (x-shifted)        -> HBM
//
// B, C and D of one iteration read only rows written in EARLIER iterations, so in the steady state (rows 9..R+4
// of a chunk, instantiated without any stage predicate) the three instruction streams form one basic block and
// interleave freely (ILP); only addresses and the three own-column populations f*0, f*2, f*4 live in registers across
// iterations.  Redundant work: the 8 halo columns per CTA and the 9-row pipeline fill per chunk of rows; nothing is
// recomputed in y inside a chunk.
//
// Two flavours per (NT, tau==1, thermal): PM >= 0 is the LEAN kernel for the steps in the middle of a
// swalbe_time_loop call (no materialisation, pressure mode PM fixed at compile time; strict: scalar theta, standard
// slip, no inclination, no logs, optionally GZ: gravity == 0 folded in and BULK; OPTS: theta field, slip variant,
// inclination and per-step logs as run-time options -- they cost ~8 registers, which the strict kernels cannot spare); PM == -1 is the FULL kernel with
// every option decided at run time (used for the last step of a call, which materialises the reference's
// intermediate fields, and for all uncommon options).  Instantiated per CTA width in fused_v*.cu (variants.h).
//
// FM flavour (tau != 1 only, "from moments"): the height and velocity of a site ARE the moments of the populations the
// previous step streamed there (src/moments.jl:47-50), so a step in the middle of a swalbe_time_loop call does not read
// the h / ux / uy planes at all and does not write them either: each iteration loads the nine old populations of row
// N(t+1) (L2-prefetched a few rows ahead), folds them into h for the h ring, and stage C derives u from the populations
// it needs for the collision anyway.  HBM traffic: 72 B read + 72 B written per lattice update -- the D2Q9 figure.
//
// NS flavour ("neighbour sync"): the per-row __syncthreads is replaced by warp-to-warp hand-shakes.  Only x-shifted values
// cross threads, so a warp that starts iteration t+1 needs exactly its two neighbour warps to have finished iteration t
// (their ring writes visible: RAW; their reads of the slots it is about to overwrite done: WAR -- every slot written in
// iteration t+1 was last read in iteration t or earlier).  Each warp owns two mbarriers (even / odd iterations, 32
// arrivals); a warp arrives on its own and waits on its neighbours'.  Neighbours are never more than one iteration
// apart, so the two-deep ring cannot alias a phase.  Rows are prefetched per warp: per-thread LDGSTS as before, or --
// BULK -- lane 0 of every warp moves its warp's 32 columns through the TMA unit onto a per-warp mbarrier.
//
// All arithmetic comes from common.cuh (reference evaluation order, no FMA contraction).
#pragma once
#include <limits.h>

#include <type_traits>

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
  PhiloxKey pk;  // round keys of the noise generator (seed expanded on the host)
  unsigned long long step;
  // per-step logs (slot pointers for THIS step, NULL = off)
  double *log_min, *log_max;
  unsigned long long *log_wet;
  double hthresh;
  int fm_prefetch;  // FM kernels: rows ahead of the population loads that are prefetched into L2 (0 = off)
  int fm_hints;     // FM kernels, L2 residency: bit 0 = streaming (evict-first) stores of the new populations, bit 1 = the
                    // first read of a population row is kept (evict_last) for its second read three rows later, which
                    // is marked evict_first; bit 2 = the L2 prefetch (fm_prefetch rows ahead) carries the evict_last priority
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

#ifndef SW_HOST_EMULATION  // (tests/simt_emulation.cpp supplies host versions of these eighteen helpers)
// 8-byte asynchronous global -> shared copy (LDGSTS); completion is tracked per thread by commit/wait groups,
// not by the register scoreboard.
__device__ __forceinline__ void cp_async8(double *smem_dst, const void *gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// TMA-style bulk asynchronous copy (cp.async.bulk, SASS UBLKCP): ONE thread moves a whole contiguous row segment
// global -> shared; completion is signalled on an mbarrier by byte count (complete_tx), not on any scoreboard.
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(b), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(double *smem_dst, const void *gsrc, unsigned bytes, unsigned long long *bar) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(d),
               "l"(gsrc), "r"(bytes), "r"(b)
               : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *gsrc) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(gsrc)); }
__device__ __forceinline__ void prefetch_l2_keep(const void *gsrc) { asm volatile("prefetch.global.L2::evict_last [%0];\n" ::"l"(gsrc)); }
// L2 residency hints of the FM flavour: a population row is read twice, three iterations apart; in between the CTAs
// of the whole grid stream ~5 MB of new populations per row through the same L2
__device__ __forceinline__ unsigned long long l2_policy_keep() {
  unsigned long long pol;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ unsigned long long l2_policy_drop() {
  unsigned long long pol;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ double ldg_hint(const double *p, unsigned long long pol) {
  double v;
  asm("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;\n" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_stream(double *p, double v) { asm volatile("st.global.cs.f64 [%0], %1;\n" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity);
// warp hand-shake of the NS flavour (emulated with real atomics on the CPU, unlike the data barriers above)
__device__ __forceinline__ void ns_init(unsigned long long *bar, unsigned count) { mbar_init(bar, count); }
__device__ __forceinline__ void ns_arrive(unsigned long long *bar) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(b) : "memory");  // (release at CTA scope)
}
__device__ __forceinline__ void ns_wait(unsigned long long *bar, unsigned parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void ns_syncwarp() { __syncwarp(); }
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done = 0;
  for (unsigned spin = 0; !done; ++spin) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"(b), "r"(parity)
                 : "memory");
    if (spin > (1u << 24)) __trap();  // a lost copy must fail loudly instead of hanging the device
  }
}
#endif  // SW_HOST_EMULATION

// A row cursor: BYTE offset of (row, own column) inside a plane, advanced one row per iteration with the periodic
// wrap folded in (no integer division and no index->byte scaling inside the loop).
struct RowCursor {
  int r;          // physical row (CTA-uniform)
  long long off;  // (r * Lx + ci) * 8
  __device__ __forceinline__ void init(int logical, int Lx, int Ly, int wrap_y, int ci) {
    r = wrap_y ? wrapi(logical, Ly) : logical;
    off = ((long long)r * Lx + ci) * 8;
  }
  __device__ __forceinline__ void advance(long long row_bytes, int wrapLy, long long col_bytes) {
    ++r;
    off += row_bytes;
    if (r == wrapLy) { r = 0; off = col_bytes; }
  }
};
__device__ __forceinline__ const double *at(const double *base, long long byte_off) {
  return (const double *)((const char *)base + byte_off);
}
__device__ __forceinline__ double *at(double *base, long long byte_off) { return (double *)((char *)base + byte_off); }

// shared-memory layout (in lines of LW = NT+4 doubles; every line has two pad cells on either side, so that column 0 of a
// line is 16-byte aligned for bulk copies; only the inner pad cell is ever read)
//   h ring   : 8 slots x 1 line               row N(t) <-> slot t & 7   (rows N(t-3) .. N(t+D) are live)
//   ring-4   : 4 slots x 5 lines  P F1 F3 F5 F6   row P(t) / F(t) <-> slot t & 3
//   ring-2   : 2 slots x 2 lines  F7 F8           row F(t) <-> slot t & 1
//   u ring   : 4 slots x 2 lines  UX UY           row F(t) <-> slot t & 3   (rows F(t) .. F(t+D) are live)
// f*0, f*2, f*4 never leave their column and are parked in registers instead.
constexpr int FUSED_D = 2;  // cp.async prefetch distance in rows (D+4 <= 8 h slots, D+1 <= 4 u slots)
constexpr int R4_P = 0, R4_F1 = 1, R4_F3 = 2, R4_F5 = 3, R4_F6 = 4, R4_LINES = 5;
constexpr int R2_F7 = 0, R2_F8 = 1, R2_LINES = 2;
constexpr int RU_UX = 0, RU_UY = 1, RU_LINES = 2;
constexpr int FUSED_H_SLOTS = 8;
constexpr int FUSED_LINES = FUSED_H_SLOTS + 4 * R4_LINES + 2 * R2_LINES + 4 * RU_LINES;
constexpr int FUSED_PAD = 2;
constexpr size_t fused_smem_doubles(int NT) { return (size_t)FUSED_LINES * (NT + 2 * FUSED_PAD); }

template <int NT, int MINB, bool TAU1, bool THERMAL, int PM, bool BULK, bool GZ, bool OPTS, bool FM = false, bool NS = false>
__global__ void __launch_bounds__(NT, MINB) k_fused_step(const __grid_constant__ FusedArgs a) {
  static_assert(!FM || (!TAU1 && !BULK), "the from-moments flavour exists for tau != 1 (at tau == 1 no population is read)");
  static_assert(!NS || (NT % 32 == 0 && NT / 32 <= 8 && !FM), "neighbour sync: whole warps, at most 8 per CTA");
#ifdef SW_HOST_EMULATION
  double *const smem = emul_dynamic_smem();
#else
  extern __shared__ __align__(16) double smem[];
#endif
  constexpr bool LEAN = PM >= 0;
  constexpr int LW = NT + 2 * FUSED_PAD;
  constexpr int D = FUSED_D;
  constexpr int R4S = R4_LINES * LW, R2S = R2_LINES * LW, RUS = RU_LINES * LW;  // slot strides in doubles

  const int tid = threadIdx.x;
  const int s0 = blockIdx.x * a.W;
  const int Lx = a.Lx;
  const int ci = wrapi(s0 - 4 + tid, Lx);
  const bool col_out = tid >= 4 && tid < 4 + a.W && (s0 + tid - 4) < Lx;
  const int j0 = a.jbeg + blockIdx.y * a.rows_per_cta;
  const int R = min(a.rows_per_cta, a.jend - j0);
  const int wrapLy = a.wrap_y ? a.Ly : INT_MAX;
  const long long row_bytes = (long long)Lx * 8, col_bytes = (long long)ci * 8;

  double *const sh = smem + tid + FUSED_PAD;       // own column, h ring slot 0
  double *const s4 = sh + FUSED_H_SLOTS * LW;      // own column, ring-4 slot 0, line 0
  double *const s2 = s4 + 4 * R4S;                 // own column, ring-2 slot 0, line 0
  double *const su = s2 + 2 * R2S;                 // own column, u ring slot 0, line 0

  // Programmatic dependent launch: let the next step's grid start filling SMs as ours drains ...
#ifndef SW_HOST_EMULATION
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
  if (tid < 2 * FUSED_PAD) {  // the pad cells of every line are never written by the pipeline; keep them finite
    const int e = tid < FUSED_PAD ? tid : NT + tid;
    for (int q = 0; q < FUSED_LINES; ++q) smem[q * LW + e] = 0.0;
  }
  // bulk-copy variant: one mbarrier per in-flight prefetch group (4 >= D+1), armed and fed by thread 0 only.
  // The strip's NT columns are one contiguous run of the row, or two when the strip crosses the periodic x boundary.
  __shared__ unsigned long long s_bar[4];
  int seg_a = 0;  // columns [c_start, c_start + seg_a) then [0, NT - seg_a)
  // NS: two hand-shake barriers per warp (even / odd iterations) and, with BULK, four data barriers per warp
  constexpr int NW = NT / 32;
  __shared__ unsigned long long s_done[NS ? 2 : 1][8];
  __shared__ unsigned long long s_wbar[NS && BULK ? 8 : 1][4];
  const int wid = tid >> 5, lane = tid & 31;
  if (NS) {
    seg_a = min(32, Lx - ci);  // (lane 0 of a warp: ci is the warp's first column)
    if (tid < 32) {
      if (tid < 16) ns_init(&s_done[tid >> 3][tid & 7], 32);
      if (BULK) mbar_init(&s_wbar[tid >> 2][tid & 3], 1);
      mbar_fence_init();
    }
    __syncthreads();
  } else if (BULK) {
    seg_a = min(NT, Lx - ci);  // (thread 0: ci == c_start; only thread 0 uses it)
    if (tid == 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q) mbar_init(&s_bar[q], 1);
      mbar_fence_init();
    }
    __syncthreads();
  }

  RowCursor cN, cU, cO, cC, cT, cPf;
  // h row N(0); prefetched D iterations ahead, starting at t = -D.  FM: old populations of row N(t+1) at iteration t
  cN.init(FM ? j0 - 3 - D : j0 - 4, Lx, a.Ly, a.wrap_y, ci);
  if (FM) cPf.init(j0 - 3 - D + a.fm_prefetch, Lx, a.Ly, a.wrap_y, ci);  // L2 prefetch runs fm_prefetch rows ahead of cN
  cU.init(j0 - 7, Lx, a.Ly, a.wrap_y, ci);  // u row F(0)
  cO.init(j0 - 9, Lx, a.Ly, a.wrap_y, ci);  // output row O(0)
  cC.init(j0 - 4, Lx, a.Ly, a.wrap_y, ci);  // cospi(theta) field: row P(1), loaded one iteration ahead
  cT.init(j0 - 6, Lx, a.Ly, a.wrap_y, ci);  // old populations (tau != 1): row F(1)

  // thermal noise: global row of F(t), advanced one row per iteration from t = 0; the Philox block of a row pair
  // (2k, 2k+1) is drawn at the even row (or at the first row this CTA computes) and its second half carried over
  long long g_row = 0;
  uint32_t nz_w2 = 0, nz_w3 = 0;
  bool nz_have = false;
  if (THERMAL) {
    g_row = (a.jglobal0 + (long long)(j0 - 7)) % a.Ly_global;
    if (g_row < 0) g_row += a.Ly_global;
  }

  // own-column populations that move along y only: f*0 of rows F(t-1), F(t-2); f*2 of F(t-1..t-3); f*4 of F(t-1)
  double f0_a = 0, f0_b = 0, f2_a = 0, f2_b = 0, f2_c = 0, f4_a = 0;
  double ft_c[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) ft_c[k] = 0.0;
  double ct_c = 0.0;  // cospi(theta) of the row stage B works on (register prefetch, one iteration ahead)

  double d_min = INFINITY, d_max = -INFINITY;
  unsigned int d_wet = 0;
  // OPTS kernels take the uncommon options at run time; the strict lean kernels (OPTS == false) have them compiled out
  const bool logging = OPTS && (a.log_min != nullptr || a.log_wet != nullptr);
  const bool theta_field = OPTS && a.ct_field != nullptr;
  const bool aux = !LEAN && a.pressure != nullptr;
  const int pmode = LEAN ? PM : a.pc.pmode;
  const int slipv = OPTS ? a.sc.variant : SWALBE_SLIP_STANDARD;
  const long long fs_in8 = (long long)a.fstride_in * 8, fs_out8 = (long long)a.fstride_out * 8,
                  fs_out2_8 = (long long)a.fstride_out2 * 8;

  const unsigned long long pol_keep = FM ? l2_policy_keep() : 0ull, pol_drop = FM ? l2_policy_drop() : 0ull;

  // one pipeline iteration; `steady` is a compile-time tag: std::true_type drops every stage predicate
  auto iter = [&](const int t, auto steady) {
    constexpr bool S = decltype(steady)::value;
    // ---- asynchronous prefetch of the next row ----------------------------------------------------------
    const long long offN = cN.off;  // FM: row N(t+1)
    if (FM) {
      if (a.fm_prefetch > 0) {  // (FM launches are periodic in y, wrap_y == 1: every row the cursor reaches exists)
        if (a.fm_hints & 4) {
#pragma unroll
          for (int k = 0; k < 9; ++k) prefetch_l2_keep(at(a.f_in, cPf.off + k * fs_in8));
        } else {
#pragma unroll
          for (int k = 0; k < 9; ++k) prefetch_l2(at(a.f_in, cPf.off + k * fs_in8));
        }
      }
      cPf.advance(row_bytes, wrapLy, col_bytes);
      cN.advance(row_bytes, wrapLy, col_bytes);
    } else {
      const int tn = t + D;  // h row N(tn) = j0-4+tn is needed for tn in [1, R+6]; u rows F(tn) for tn in [6, R+7]
      const bool need_h = S || (tn >= 1 && tn <= R + 6), need_u = S || (tn >= 6 && tn <= R + 7);
      if (BULK) {
        if (NS ? lane == 0 : tid == 0) {
          unsigned long long *bar = NS ? &s_wbar[wid][tn & 3] : &s_bar[tn & 3];
          const unsigned nb = (unsigned)(NS ? 32 : NT) * 8u;
          mbar_arrive_expect_tx(bar, (need_h ? nb : 0u) + (need_u ? 2u * nb : 0u));  // 0 bytes: completes at once
          const unsigned ba = (unsigned)seg_a * 8u, bb = nb - ba;
          if (need_h) {
            double *dst = sh + (tn & 7) * LW;
            bulk_g2s(dst, at(a.h_in, cN.off), ba, bar);
            if (bb) bulk_g2s(dst + seg_a, at(a.h_in, cN.off - col_bytes), bb, bar);
          }
          if (need_u) {
            double *dst = su + (tn & 3) * RUS;
            bulk_g2s(dst + RU_UX * LW, at(a.ux_in, cU.off), ba, bar);
            bulk_g2s(dst + RU_UY * LW, at(a.uy_in, cU.off), ba, bar);
            if (bb) {
              bulk_g2s(dst + RU_UX * LW + seg_a, at(a.ux_in, cU.off - col_bytes), bb, bar);
              bulk_g2s(dst + RU_UY * LW + seg_a, at(a.uy_in, cU.off - col_bytes), bb, bar);
            }
          }
        }
      } else {
        if (need_h) cp_async8(sh + (tn & 7) * LW, at(a.h_in, cN.off));
        if (need_u) {
          double *dst = su + (tn & 3) * RUS;
          cp_async8(dst + RU_UX * LW, at(a.ux_in, cU.off));
          cp_async8(dst + RU_UY * LW, at(a.uy_in, cU.off));
        }
      }
      cN.advance(row_bytes, wrapLy, col_bytes);
      cU.advance(row_bytes, wrapLy, col_bytes);
      if (!BULK) cp_async_commit();
    }
    if (S || t >= 0) {
      // ring slots of this iteration
      double *const w4 = s4 + (t & 3) * R4S;               // written now: P(t), F(t)
      const double *const a1 = s4 + ((t - 1) & 3) * R4S;   // age 1
      const double *const a2 = s4 + ((t - 2) & 3) * R4S;   // age 2
      const double *const a3 = s4 + ((t - 3) & 3) * R4S;   // age 3
      double *const w2 = s2 + (t & 1) * R2S;               // written now: F7 F8 of F(t)
      const double *const o2 = s2 + ((t + 1) & 1) * R2S;   // age 1
      const double *const ur = su + (t & 3) * RUS;         // u of row F(t)

      // ---- register prefetch of data that is not staged through smem ------------------------------------
      const bool doB = S || (t >= 3 && t <= R + 6);   // row P(t) = j0-5+t in [j0-2, j0+R+1]
      double ct_n = 0.0;
      if (theta_field) {  // row P(t+1) = j0-4+t, needed for t+1 in [3, R+6]
        if (S || (t >= 2 && t <= R + 5)) ct_n = __ldg(at(a.ct_field, cC.off));
        cC.advance(row_bytes, wrapLy, col_bytes);
      }
      double fm_n[9];
      const bool need_m = FM && (S || t <= R + 5);  // row N(t+1) = j0-3+t is needed for t+1 in [1, R+6]
      if (need_m) {
        if (a.fm_hints & 2) {
#pragma unroll
          for (int k = 0; k < 9; ++k) fm_n[k] = ldg_hint(at(a.f_in, offN + k * fs_in8), pol_keep);
        } else {
#pragma unroll
          for (int k = 0; k < 9; ++k) fm_n[k] = __ldg(at(a.f_in, offN + k * fs_in8));
        }
      }
      double ft_n[9];
      if (!TAU1) {
        if (S || (t >= 5 && t <= R + 6)) {  // row F(t+1) = j0-6+t in [j0-1, j0+R]
          if (FM && (a.fm_hints & 2)) {
#pragma unroll
            for (int k = 0; k < 9; ++k) ft_n[k] = ldg_hint(at(a.f_in, cT.off + k * fs_in8), pol_drop);
          } else {
#pragma unroll
            for (int k = 0; k < 9; ++k) ft_n[k] = __ldg(at(a.f_in, cT.off + k * fs_in8));
          }
        }
        cT.advance(row_bytes, wrapLy, col_bytes);
      }

      // ---- stage B: film pressure at row P(t) ------------------------------------------------------------
      if (doB) {
        const double *r0 = sh + ((t - 2) & 7) * LW;  // row P-1
        const double *r1 = sh + ((t - 1) & 7) * LW;  // row P
        const double *r2 = sh + (t & 7) * LW;        // row P+1
        const double hc = r1[0];
        const double lap = lap9_bracket(hc, r1[-1], r0[0], r1[1], r2[0], r0[-1], r0[1], r2[1], r2[-1]);
        const double kappa = theta_field ? kappa_from_field(ct_c, a.pc) : a.pc.kappa;
        const double x = div_exact(a.pc.hmin, hc + a.pc.hcrit);
        const double pw = disjoining_powers(x, pmode, a.pc.n, a.pc.m);
        w4[R4_P * LW] = (-a.pc.gamma * (kappa * pw)) - a.pc.gamma * lap;  // == film_pressure()
      }

      // ---- stage C: forces, equilibrium, collision at row F(t) = j0-7+t -----------------------------------
      double fs0 = 0.0, fs2 = 0.0, fs4 = 0.0;
      if (S || (t >= 6 && t <= R + 7)) {
        const double *q0 = a3 + R4_P * LW;  // p row F-1
        const double *q1 = a2 + R4_P * LW;  // p row F
        const double *q2 = a1 + R4_P * LW;  // p row F+1
        const double hc = sh[((t - 3) & 7) * LW];
        double ux_c, uy_c;
        if (FM) velocity_site(ft_c, hc, ux_c, uy_c);  // moments! of the previous step (src/moments.jl:49-50), not stored
        else { ux_c = ur[RU_UX * LW]; uy_c = ur[RU_UY * LW]; }
        const double pipjp = q0[-1], pimjp = q0[1], pimjm = q2[1], pipjm = q2[-1];
        const double gx = grad9_x(q1[-1], q1[1], pipjp, pimjp, pimjm, pipjm);
        const double gy = grad9_y(q0[0], q2[0], pipjp, pimjp, pimjm, pipjm);
        const double hgx = hc * gx, hgy = hc * gy;
        double sx, sy;
        slip_terms(hc, ux_c, uy_c, a.sc, slipv, sx, sy);
        double Fx = (-hgx) - sx, Fy = (-hgy) - sy;
        double kx = 0.0, ky = 0.0;
        const int rF = j0 - 7 + t;
        if (THERMAL) {
          const bool odd = (g_row & 1) != 0;  // (CTA-uniform)
          uint32_t w0, w1;
          if (!odd || !nz_have) {
            uint32_t r[4];
            noise_block(a.pk, a.step, noise_pair_index(Lx, g_row, ci), r);
            nz_w2 = r[2]; nz_w3 = r[3]; nz_have = true;
            w0 = odd ? r[2] : r[0]; w1 = odd ? r[3] : r[1];
          } else {
            w0 = nz_w2; w1 = nz_w3;
          }
          thermal_from_words(hc, a.tc, w0, w1, kx, ky);
          Fx = Fx - kx;
          Fy = Fy - ky;
        }
        if (OPTS && a.use_incl) {
          Fx = Fx + (hc * a.incl_ax) * a.incl_factor;
          Fy = Fy + (hc * a.incl_ay) * a.incl_factor;
        }
        double fe[9], vsq, fs[9];
        equilibrium_site<GZ>(hc, ux_c, uy_c, a.ec, fe, vsq);  // GZ instantiations are only chosen for g == 0
        if (TAU1) collide_site_tau1(fe, Fx, Fy, fs);
        else collide_site(ft_c, fe, Fx, Fy, a.omega, a.invtau, fs);
        w4[R4_F1 * LW] = fs[1]; w4[R4_F3 * LW] = fs[3]; w4[R4_F5 * LW] = fs[5]; w4[R4_F6 * LW] = fs[6];
        w2[R2_F7 * LW] = fs[7]; w2[R2_F8 * LW] = fs[8];
        fs0 = fs[0]; fs2 = fs[2]; fs4 = fs[4];

        if (logging || aux) {
          const bool own = col_out && t >= 7 && t <= R + 6;  // row F(t) in [j0, j0+R-1]
          if (own) {
            if (logging) {
              d_min = fmin(d_min, hc);
              d_max = fmax(d_max, hc);
              d_wet += hc > a.hthresh;
            }
            if (aux) {  // materialise the reference's intermediate fields (last step of a call)
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
      }

      // ---- stage D: pull-stream + moments at row O(t) = j0-9+t --------------------------------------------
      if (S || t >= 9) {
        double fn[9];
        fn[0] = f0_b;  fn[1] = a2[R4_F1 * LW - 1];  fn[3] = a2[R4_F3 * LW + 1];  // row O   = F(t-2)
        fn[2] = f2_c;  fn[5] = a3[R4_F5 * LW - 1];  fn[6] = a3[R4_F6 * LW + 1];  // row O-1 = F(t-3), moving +y
        fn[4] = f4_a;  fn[7] = o2[R2_F7 * LW + 1];  fn[8] = o2[R2_F8 * LW - 1];  // row O+1 = F(t-1), moving -y
        const bool want_m = TAU1 || a.h_out != nullptr;  // tau != 1: only the last step of a call stores the moments
        double hn = 0.0, uxn = 0.0, uyn = 0.0;
        if (want_m) moments_site(fn, hn, uxn, uyn);
        if (col_out) {
          if (want_m) { *at(a.h_out, cO.off) = hn; *at(a.ux_out, cO.off) = uxn; *at(a.uy_out, cO.off) = uyn; }
          if (FM && (a.fm_hints & 1)) {
#pragma unroll
            for (int k = 0; k < 9; ++k) st_stream(at(a.f_out, cO.off + k * fs_out8), fn[k]);
          } else if (a.f_out != nullptr) {
#pragma unroll
            for (int k = 0; k < 9; ++k) *at(a.f_out, cO.off + k * fs_out8) = fn[k];
            if (a.f_out2 != nullptr) {
#pragma unroll
              for (int k = 0; k < 9; ++k) *at(a.f_out2, cO.off + k * fs_out2_8) = fn[k];
            }
          }
        }
      }
      cO.advance(row_bytes, wrapLy, col_bytes);
      if (THERMAL) {  // global row of F(t+1)
        if (++g_row == a.Ly_global) g_row = 0;
      }
      ct_c = ct_n;
      f0_b = f0_a; f0_a = fs0;
      f2_c = f2_b; f2_b = f2_a; f2_a = fs2;
      f4_a = fs4;
      if (!TAU1) {
#pragma unroll
        for (int k = 0; k < 9; ++k) ft_c[k] = ft_n[k];
      }
      if (need_m) sh[((t + 1) & 7) * LW] = height_site(fm_n);  // h row N(t+1) into its ring slot (free since iteration t-3)
    }
    // the prefetch issued D-1 iterations ago (h row N(t+1), u rows F(t+1)) must have landed before the next iteration
    if (BULK) mbar_wait(NS ? &s_wbar[wid][(t + 1) & 3] : &s_bar[(t + 1) & 3], (unsigned)((t + 1) >> 2) & 1u);
    else if (!FM) cp_async_wait<D - 1>();
    if (NS) {
      const int tt = t + D;  // >= 0
      unsigned long long *mine = &s_done[tt & 1][wid];
      const unsigned par = (unsigned)(tt >> 1) & 1u;
      ns_arrive(mine);
      if (wid > 0) ns_wait(mine - 1, par);
      if (wid < NW - 1) ns_wait(mine + 1, par);
      ns_syncwarp();  // the lanes of this warp exchange x-neighbours among themselves, too
    } else {
      __syncthreads();
    }
  };

  // ... and do not touch global memory before the previous step's grid has completed and flushed its writes.
#ifndef SW_HOST_EMULATION
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif

  // pipeline fill (predicated), steady state (predicate-free), drain (predicated)
  const int t_end = R + 8;
  int t = -D;
  for (; t < 9 && t <= t_end; ++t) iter(t, std::false_type{});
#ifdef SW_UNROLL8
  // experiment: steady state unrolled by the ring period so that every ring slot is a compile-time constant
  for (; t < 16 && t <= R + 6 - D; ++t) iter(t, std::true_type{});
  if (t == 16) {
    int tt = 16;
#pragma unroll 1
    for (; tt + 7 <= R + 6 - D; tt += 8) {
      iter(tt, std::true_type{}); iter(tt + 1, std::true_type{}); iter(tt + 2, std::true_type{}); iter(tt + 3, std::true_type{});
      iter(tt + 4, std::true_type{}); iter(tt + 5, std::true_type{}); iter(tt + 6, std::true_type{}); iter(tt + 7, std::true_type{});
    }
    t = tt;
  }
#endif
  for (; t <= R + 6 - D; ++t) iter(t, std::true_type{});  // (the last prefetched h row is N(R+6))
  for (; t <= t_end; ++t) iter(t, std::false_type{});

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
