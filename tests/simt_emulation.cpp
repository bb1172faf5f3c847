// Host SIMT emulation of the step KERNELS themselves (swalbe.jl_b200/csrc/fused.cuh, tile.cuh) for
// tests/test_simt_emulation.py -- test infrastructure only, nothing in the product links this.
//   g++ -O1 -ffp-contract=off -DSW_HOST_EMULATION -w -I/usr/local/cuda/include -shared -fPIC -pthread tests/simt_emulation.cpp
// Every CUDA thread of a CTA runs as one OS thread, __syncthreads() is a pthread barrier, __shared__ is static storage,
// CTAs run one after the other; cp.async / cp.async.bulk / mbarrier become immediate copies and no-ops (legal: the
// kernels never read a ring slot in the iteration that prefetches into it).  What is exercised on the CPU is therefore
// the kernels' own index logic -- strips, halo columns, row cursors with periodic wrap or ghost rows, the software
// pipeline and its ring slots, fill/steady/drain instantiations, the tile phases -- against the oracle, bit for bit.
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <sched.h>

#include <atomic>
#include <thread>
#include <vector>

// ---- the CUDA execution model, on the host --------------------------------------------------------------------------
struct Emu3 {
  unsigned x = 0, y = 0, z = 0;
};
static thread_local Emu3 threadIdx, blockIdx, blockDim;
static pthread_barrier_t g_cta_barrier;
static double *g_dynamic_smem = nullptr;
// cluster launches run several CTAs at once: each OS thread then carries its own CTA's barrier / smem / shuffle state
static thread_local pthread_barrier_t *tl_cta_barrier = nullptr, *tl_warp_barrier = nullptr;
static thread_local double *tl_smem = nullptr;
static thread_local unsigned long long (*tl_shfl_slot)[32] = nullptr;
static inline void __syncthreads() { pthread_barrier_wait(tl_cta_barrier ? tl_cta_barrier : &g_cta_barrier); }
static inline double *emul_dynamic_smem() { return tl_smem ? tl_smem : g_dynamic_smem; }
#define __shared__ static
#define __launch_bounds__(...)
#define __grid_constant__
#define __noinline__ __attribute__((noinline))

// ---- intrinsics -----------------------------------------------------------------------------------------------------
static inline int __double2hiint(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(u >> 32); }
static inline double __hiloint2double(int hi, int lo) {
  const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double d; memcpy(&d, &u, 8); return d;
}
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline void sincospi(double a, double *s, double *c) { *s = sin(a * 3.141592653589793); *c = cos(a * 3.141592653589793); }
static inline long long __double_as_longlong(double x) { long long u; memcpy(&u, &x, 8); return u; }
static inline double __longlong_as_double(long long u) { double x; memcpy(&x, &u, 8); return x; }
template <class T> static inline T __ldg(const T *p) { return *p; }
// warp shuffles: the 32 OS threads of a warp meet at a per-warp barrier around a shared slot array
static pthread_barrier_t g_warp_barrier[8];
static unsigned long long g_shfl_slot[8][32];
template <class T> static inline T __shfl_down_sync(unsigned, T v, int delta);
static inline unsigned long long atomicCAS(unsigned long long *p, unsigned long long cmp, unsigned long long v) {
  return __sync_val_compare_and_swap(p, cmp, v);
}
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __sync_fetch_and_add(p, v); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int delta) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  const unsigned w = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  unsigned long long bits = 0;
  memcpy(&bits, &v, sizeof(T));
  unsigned long long(*slot)[32] = tl_shfl_slot ? tl_shfl_slot : g_shfl_slot;
  pthread_barrier_t *wb = tl_warp_barrier ? tl_warp_barrier : g_warp_barrier;
  slot[w][lane] = bits;
  pthread_barrier_wait(&wb[w]);
  const unsigned src = lane + (unsigned)delta;
  const unsigned long long got = src < 32u ? slot[w][src] : bits;
  pthread_barrier_wait(&wb[w]);
  T out;
  memcpy(&out, &got, sizeof(T));
  return out;
}
static inline void __trap() { abort(); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }

// ---- the thirteen asynchronous-copy / cache-hint helpers of fused.cuh, as immediate copies / plain accesses / no-ops -----------------------------------------
static inline void cp_async8(double *smem_dst, const void *gsrc) { *smem_dst = *(const double *)gsrc; }
static inline void cp_async_commit() {}
template <int N> static inline void cp_async_wait() {}
static inline void mbar_init(unsigned long long *, unsigned) {}
static inline void mbar_fence_init() {}
static inline void mbar_arrive_expect_tx(unsigned long long *, unsigned) {}
static inline void bulk_g2s(double *smem_dst, const void *gsrc, unsigned bytes, unsigned long long *) { memcpy(smem_dst, gsrc, bytes); }
static inline void mbar_wait(unsigned long long *, unsigned) {}
static inline void prefetch_l2(const void *) {}
static inline void prefetch_l2_keep(const void *) {}
static inline unsigned long long l2_policy_keep() { return 1; }
static inline unsigned long long l2_policy_drop() { return 2; }
static inline double ldg_hint(const double *p, unsigned long long) { return *p; }
static inline void st_stream(double *p, double v) { *p = v; }
// the warp hand-shake of the NS flavour is emulated FOR REAL (atomics, OS threads that run ahead of each other): this is
// what the flavour's correctness rests on -- a warp may only see its neighbours' ring slots through these barriers
struct EmuNsBar {
  std::atomic<uint32_t> arrived, phase;
};
static_assert(sizeof(EmuNsBar) == 8, "one mbarrier word");
static uint32_t g_ns_expected = 32;
static inline void ns_init(unsigned long long *bar, unsigned count) {
  EmuNsBar *b = reinterpret_cast<EmuNsBar *>(bar);
  b->arrived.store(0); b->phase.store(0);
  g_ns_expected = count;
}
static inline void ns_arrive(unsigned long long *bar) {
  EmuNsBar *b = reinterpret_cast<EmuNsBar *>(bar);
  if (b->arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == g_ns_expected) {
    b->arrived.store(0, std::memory_order_relaxed);
    b->phase.fetch_add(1, std::memory_order_release);
  }
}
static inline void ns_wait(unsigned long long *bar, unsigned parity) {
  EmuNsBar *b = reinterpret_cast<EmuNsBar *>(bar);
  while ((b->phase.load(std::memory_order_acquire) & 1u) == parity) sched_yield();
}
static inline void ns_syncwarp();

static inline void ns_syncwarp() { pthread_barrier_wait(&g_warp_barrier[threadIdx.x >> 5]); }

// ---- thread-block cluster: rank, size, distributed shared memory, cluster barrier ------------------------------------
static unsigned g_cluster_size = 1;
static double *g_cluster_smem[16];
static pthread_barrier_t g_cluster_barrier;
static inline unsigned cl_rank() { return blockIdx.x; }
static inline unsigned cl_size() { return g_cluster_size; }
static inline const double *cl_map(const double *p, unsigned rank) { return g_cluster_smem[rank] + (p - tl_smem); }
static inline void cl_sync() { pthread_barrier_wait(&g_cluster_barrier); }

#include "../swalbe.jl_b200/csrc/cluster.cuh"
#include "../swalbe.jl_b200/csrc/tile.cuh"  // (includes fused.cuh and common.cuh)

using namespace swalbe;
namespace swalbe {
int set_error(int code, const char *, ...) { return code; }
void count_launch(unsigned) {}
}  // namespace swalbe

typedef void (*kernel_fn)(const FusedArgs);

static void launch(kernel_fn k, unsigned gx, unsigned gy, unsigned nthreads, size_t dyn_doubles, const FusedArgs &a) {
  std::vector<double> smem(dyn_doubles + 2, 0.0);
  g_dynamic_smem = (double *)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
  for (unsigned by = 0; by < gy; ++by)
    for (unsigned bx = 0; bx < gx; ++bx) {
      pthread_barrier_init(&g_cta_barrier, nullptr, nthreads);
      for (unsigned w = 0; w < nthreads / 32; ++w) pthread_barrier_init(&g_warp_barrier[w], nullptr, 32);
      std::vector<std::thread> cta;
      cta.reserve(nthreads);
      for (unsigned t = 0; t < nthreads; ++t)
        cta.emplace_back([=]() {
          threadIdx.x = t; blockIdx.x = bx; blockIdx.y = by;
          k(a);
        });
      for (auto &th : cta) th.join();
      pthread_barrier_destroy(&g_cta_barrier);
      for (unsigned w = 0; w < nthreads / 32; ++w) pthread_barrier_destroy(&g_warp_barrier[w]);
    }
}

constexpr int ENT = 128;  // CTA width of the emulated marching kernels

template <bool GZ, bool OPTS, bool BULK, int NT = ENT>
static kernel_fn lean_kernel(int pm) {
  switch (pm) {
    case PM_BROAD_93: return k_fused_step<NT, 3, true, false, PM_BROAD_93, BULK, GZ, OPTS>;
    case PM_BROAD_32: return k_fused_step<NT, 3, true, false, PM_BROAD_32, BULK, GZ, OPTS>;
    case PM_FAST_93: return k_fused_step<NT, 3, true, false, PM_FAST_93, BULK, GZ, OPTS>;
    case PM_FAST_32: return k_fused_step<NT, 3, true, false, PM_FAST_32, BULK, GZ, OPTS>;
    default: return nullptr;
  }
}
template <bool GZ, bool BULK, bool TH = false>
static kernel_fn ns_kernel(int pm) {  // neighbour-sync flavour (no CTA barrier in the row loop)
  switch (pm) {
    case PM_BROAD_93: return k_fused_step<ENT, 3, true, TH, PM_BROAD_93, BULK, GZ, false, false, true>;
    case PM_BROAD_32: return k_fused_step<ENT, 3, true, TH, PM_BROAD_32, BULK, GZ, false, false, true>;
    case PM_FAST_93: return k_fused_step<ENT, 3, true, TH, PM_FAST_93, BULK, GZ, false, false, true>;
    case PM_FAST_32: return k_fused_step<ENT, 3, true, TH, PM_FAST_32, BULK, GZ, false, false, true>;
    default: return nullptr;
  }
}
template <bool GZ>
static kernel_fn thermal_kernel(int pm) {
  switch (pm) {
    case PM_BROAD_93: return k_fused_step<ENT, 5, true, true, PM_BROAD_93, false, GZ, false>;
    case PM_BROAD_32: return k_fused_step<ENT, 5, true, true, PM_BROAD_32, false, GZ, false>;
    case PM_FAST_93: return k_fused_step<ENT, 5, true, true, PM_FAST_93, false, GZ, false>;
    case PM_FAST_32: return k_fused_step<ENT, 5, true, true, PM_FAST_32, false, GZ, false>;
    default: return nullptr;
  }
}
template <bool GZ>
static kernel_fn fm_kernel(int pm) {  // tau != 1, moments derived from the streamed populations
  switch (pm) {
    case PM_BROAD_93: return k_fused_step<ENT, 3, false, false, PM_BROAD_93, false, GZ, false, true>;
    case PM_BROAD_32: return k_fused_step<ENT, 3, false, false, PM_BROAD_32, false, GZ, false, true>;
    case PM_FAST_93: return k_fused_step<ENT, 3, false, false, PM_FAST_93, false, GZ, false, true>;
    case PM_FAST_32: return k_fused_step<ENT, 3, false, false, PM_FAST_32, false, GZ, false, true>;
    default: return nullptr;
  }
}
template <bool GZ, bool TF = false>
static kernel_fn tile_kernel(int pm) {
  switch (pm) {
    case PM_BROAD_93: return k_tile_step<PM_BROAD_93, GZ, TF>;
    case PM_BROAD_32: return k_tile_step<PM_BROAD_32, GZ, TF>;
    case PM_FAST_93: return k_tile_step<PM_FAST_93, GZ, TF>;
    case PM_FAST_32: return k_tile_step<PM_FAST_32, GZ, TF>;
    default: return nullptr;
  }
}

// all CTAs of one cluster at once (C x nthreads OS threads), each with its own barrier, shuffle slots and smem block
typedef void (*cluster_kernel_fn)(const ClusterArgs);
static void launch_cluster_emul(cluster_kernel_fn k, unsigned C, unsigned nthreads, size_t dyn_doubles, const ClusterArgs &ca) {
  std::vector<std::vector<double>> smem(C, std::vector<double>(dyn_doubles + 2, 0.0));
  std::vector<pthread_barrier_t> cta_bar(C), warp_bar(C * 8);
  std::vector<unsigned long long> slots((size_t)C * 8 * 32, 0ull);
  g_cluster_size = C;
  pthread_barrier_init(&g_cluster_barrier, nullptr, C * nthreads);
  for (unsigned r = 0; r < C; ++r) {
    g_cluster_smem[r] = (double *)(((uintptr_t)smem[r].data() + 15) & ~(uintptr_t)15);
    pthread_barrier_init(&cta_bar[r], nullptr, nthreads);
    for (unsigned w = 0; w < nthreads / 32; ++w) pthread_barrier_init(&warp_bar[r * 8 + w], nullptr, 32);
  }
  std::vector<std::thread> th;
  th.reserve((size_t)C * nthreads);
  for (unsigned r = 0; r < C; ++r)
    for (unsigned t = 0; t < nthreads; ++t)
      th.emplace_back([&, r, t]() {
        threadIdx.x = t; blockIdx.x = r; blockIdx.y = 0;
        tl_cta_barrier = &cta_bar[r]; tl_warp_barrier = &warp_bar[r * 8]; tl_smem = g_cluster_smem[r];
        tl_shfl_slot = reinterpret_cast<unsigned long long(*)[32]>(&slots[(size_t)r * 8 * 32]);
        k(ca);
      });
  for (auto &t : th) t.join();
  pthread_barrier_destroy(&g_cluster_barrier);
  for (auto &b : cta_bar) pthread_barrier_destroy(&b);
  for (unsigned r = 0; r < C; ++r)
    for (unsigned w = 0; w < nthreads / 32; ++w) pthread_barrier_destroy(&warp_bar[r * 8 + w]);
}

constexpr int CENT = 128;  // CTA width of the emulated cluster kernel
template <bool GZ>
static cluster_kernel_fn cluster_kernel(int pm) {
  switch (pm) {
    case PM_BROAD_93: return k_cluster_steps<CENT, PM_BROAD_93, GZ>;
    case PM_BROAD_32: return k_cluster_steps<CENT, PM_BROAD_32, GZ>;
    case PM_FAST_93: return k_cluster_steps<CENT, PM_FAST_93, GZ>;
    case PM_FAST_32: return k_cluster_steps<CENT, PM_FAST_32, GZ>;
    default: return nullptr;
  }
}

extern "C" {

struct SimtStep {  // one launch: what swalbe_time_loop / swalbe_dist_time_loop put into FusedArgs
  int flavour;     // 0 strict lean, 1 OPTS lean, 2 FULL, 3 strict lean with bulk row prefetch, 4 tile kernel,
                   // 5 strict lean with in-kernel thermal noise, 6 strict lean with CTAs of 224 threads,
                   // 7 tau != 1 strict lean from-moments (FM), 8 tau != 1 FULL from-moments,
                   // 9 / 10 / 11 neighbour-sync strict lean: LDGSTS rows / per-warp bulk rows / with thermal noise
  int Lx, Ly, jbeg, jend, W, rows_per_cta, wrap_y;
  double tau, mu, delta, gamma, hmin, hcrit, g, cospi_theta;
  int n, m, pressure_variant, slip_variant, use_incl;
  double incl_ax, incl_ay, incl_factor;
  const double *h_in, *ux_in, *uy_in, *f_in, *ct_field;
  double *h_out, *ux_out, *uy_out, *f_out, *f_out2;
  double *pressure, *hgx, *hgy, *slipx, *slipy, *Fx, *Fy, *feq, *vsq;
  size_t fstride;
  double kbt;                      // thermal flavour
  unsigned long long seed, step;
  long long jglobal0, Ly_global;   // slab runs: global index of local row 0, global extent (noise counter)
  double *log_min, *log_max;       // per-step logs (OPTS / FULL flavours): slots of THIS step, pre-set to +-inf / 0
  unsigned long long *log_wet;
  double hthresh;
  int fm_prefetch;                 // FM flavours: L2 prefetch distance (a no-op here, but the cursor logic runs)
};

static int fill_args(const SimtStep *s, FusedArgs &a);

// the persistent cluster kernel: nsteps steps in one launch on a cluster of C CTAs, in place or not; s->log_* point at
// nsteps slots
int simt_cluster_steps(const SimtStep *s, int nsteps, int C, int lazy) {
  ClusterArgs ca = {};
  if (int e = fill_args(s, ca.a)) return e;
  if (s->tau != 1.0 || s->ct_field || C < 1 || C > 16 || s->Ly / C < 3) return -2;
  cluster_kernel_fn k = s->g == 0.0 ? cluster_kernel<true>(ca.a.pc.pmode) : cluster_kernel<false>(ca.a.pc.pmode);
  if (!k) return -1;
  ca.nsteps = nsteps; ca.lazy = lazy; ca.rows_max = (s->Ly + C - 1) / C;
  std::vector<double> part((size_t)nsteps * C * 3, 0.0);
  const bool logs = s->log_min || s->log_wet;
  ca.log_part = logs ? part.data() : nullptr;
  launch_cluster_emul(k, (unsigned)C, CENT, cluster_smem_doubles(s->Lx, ca.rows_max), ca);
  if (logs)
    for (int q = 0; q < nsteps; ++q) {  // k_cluster_logs: one CUDA thread per step
      threadIdx.x = q; blockIdx.x = 0;
      k_cluster_logs(part.data(), C, nsteps, s->log_min, s->log_max, s->log_wet);
    }
  return 0;
}

int simt_step(const SimtStep *s) {
  FusedArgs a = {};
  if (int e = fill_args(s, a)) return e;
  const bool gz = s->g == 0.0, tau1 = s->tau == 1.0;
  const int pm = a.pc.pmode;
  kernel_fn k = nullptr;
  if (s->flavour == 4) {
    if (s->ct_field) k = gz ? tile_kernel<true, true>(pm) : tile_kernel<false, true>(pm);
    else k = gz ? tile_kernel<true>(pm) : tile_kernel<false>(pm);
    if (!k) return -1;
    launch(k, (s->Lx + 31) / 32, (s->Ly + 7) / 8, 256, 0, a);
    return 0;
  }
  if (s->flavour == 0) k = gz ? lean_kernel<true, false, false>(pm) : lean_kernel<false, false, false>(pm);
  else if (s->flavour == 1) k = lean_kernel<false, true, false>(pm);
  else if (s->flavour == 3) k = gz ? lean_kernel<true, false, true>(pm) : lean_kernel<false, false, true>(pm);
  else if (s->flavour == 5) k = gz ? thermal_kernel<true>(pm) : thermal_kernel<false>(pm);
  else if (s->flavour == 6) {
    k = gz ? lean_kernel<true, false, false, 224>(pm) : lean_kernel<false, false, false, 224>(pm);
    if (!k || s->W < 1 || s->W > 224 - 8 || s->rows_per_cta < 1) return -2;
    launch(k, (s->Lx + s->W - 1) / s->W, (s->jend - s->jbeg + s->rows_per_cta - 1) / s->rows_per_cta, 224, fused_smem_doubles(224), a);
    return 0;
  }
  else if (s->flavour == 9) k = gz ? ns_kernel<true, false>(pm) : ns_kernel<false, false>(pm);
  else if (s->flavour == 10) k = gz ? ns_kernel<true, true>(pm) : ns_kernel<false, true>(pm);
  else if (s->flavour == 11) k = gz ? ns_kernel<true, false, true>(pm) : ns_kernel<false, false, true>(pm);
  else if (s->flavour == 7) k = tau1 ? nullptr : gz ? fm_kernel<true>(pm) : fm_kernel<false>(pm);
  else if (s->flavour == 8) k = tau1 ? nullptr : (kernel_fn)k_fused_step<ENT, 3, false, false, -1, false, false, true, true>;
  else if (s->flavour == 2) k = tau1 ? (kernel_fn)k_fused_step<ENT, 5, true, false, -1, false, false, true>
                                     : (kernel_fn)k_fused_step<ENT, 3, false, false, -1, false, false, true>;
  if (!k) return -1;
  if (s->W < 1 || s->W > ENT - 8 || s->rows_per_cta < 1) return -2;
  const int nrows = s->jend - s->jbeg;
  launch(k, (s->Lx + s->W - 1) / s->W, (nrows + s->rows_per_cta - 1) / s->rows_per_cta, ENT, fused_smem_doubles(ENT), a);
  return 0;
}

}  // extern "C"

static int fill_args(const SimtStep *s, FusedArgs &a) {
  if (int e = resolve_pmode(s->pressure_variant, s->n, s->m, &a.pc.pmode)) return e;
  a.pc.gamma = s->gamma; a.pc.kappa = host_kappa(s->cospi_theta, s->n, s->m, s->hmin);
  a.pc.nm1 = (double)(s->n - 1); a.pc.mm1 = (double)(s->m - 1); a.pc.kden = (double)(s->n - s->m) * s->hmin;
  a.pc.hmin = s->hmin; a.pc.hcrit = s->hcrit; a.pc.n = s->n; a.pc.m = s->m;
  a.sc = make_slip(s->delta, s->mu, s->hcrit, s->slip_variant);
  a.ec = make_eq(s->g);
  volatile double it = 1.0 / s->tau;
  volatile double om = 1.0 - it;
  a.invtau = it; a.omega = om;
  a.use_incl = s->use_incl; a.incl_ax = s->incl_ax; a.incl_ay = s->incl_ay; a.incl_factor = s->incl_factor;
  a.Lx = s->Lx; a.Ly = s->Ly; a.jbeg = s->jbeg; a.jend = s->jend; a.W = s->W; a.rows_per_cta = s->rows_per_cta;
  a.wrap_y = s->wrap_y; a.jglobal0 = s->jglobal0; a.Ly_global = s->Ly_global > 0 ? s->Ly_global : s->Ly;
  a.tc = make_thermal(s->kbt, s->mu, s->delta); a.pk = make_philox_key(s->seed); a.step = s->step;
  a.fstride_in = a.fstride_out = a.fstride_out2 = s->fstride;
  a.h_in = s->h_in; a.ux_in = s->ux_in; a.uy_in = s->uy_in; a.f_in = s->f_in; a.ct_field = s->ct_field;
  a.h_out = s->h_out; a.ux_out = s->ux_out; a.uy_out = s->uy_out; a.f_out = s->f_out; a.f_out2 = s->f_out2;
  a.pressure = s->pressure; a.hgx = s->hgx; a.hgy = s->hgy; a.slipx = s->slipx; a.slipy = s->slipy;
  a.Fx = s->Fx; a.Fy = s->Fy; a.feq = s->feq; a.vsq = s->vsq;
  a.log_min = s->log_min; a.log_max = s->log_max; a.log_wet = s->log_wet; a.hthresh = s->hthresh;
  a.fm_prefetch = s->fm_prefetch;
  a.fm_hints = s->fm_prefetch ? 7 : 0;  // (the hinted and the plain code paths both run in the FM test)
  return 0;
}

