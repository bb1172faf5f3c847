// Host side of the fused time loop: plan (scratch + launch geometry) and swalbe_time_loop.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>

#include "cluster.cuh"
#include "launch.h"
#include "sweep.h"
#include "variants.h"

namespace swalbe {

__global__ void k_init_logs(double *mn, double *mx, unsigned long long *wet, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mn) { mn[i] = INFINITY; mx[i] = -INFINITY; }
  if (wet) wet[i] = 0ull;
}

// mass log (logs->hsum): sum of one row of the height in a fixed order, then of the row sums in a fixed order -- the result
// does not depend on how a loop is cut into launches (whole lattice, bands of a host-loop sweep, seam strips)
__global__ void __launch_bounds__(256) k_rowsum(const double *__restrict__ h, int Lx, int j0, double *__restrict__ rowsum) {
  __shared__ double sh[256];
  const int row = j0 + blockIdx.x;
  const double *p = h + (size_t)row * Lx;
  double v = 0.0;
  for (int i = threadIdx.x; i < Lx; i += 256) v += p[i];
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) rowsum[row] = sh[0];
}
__global__ void __launch_bounds__(1024) k_rowsum_final(const double *__restrict__ rowsum, int Ly, double *out) {
  __shared__ double sh[1024];
  double v = 0.0;
  for (int j = threadIdx.x; j < Ly; j += 1024) v += rowsum[j];
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int w = 512; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *(volatile double *)out = sh[0];  // (out may be page-locked host memory that a host thread polls)
    __threadfence_system();
  }
}

// ---- kernel variant table (instantiated in fused_v*.cu) ---------------------------------------------
static const Variant *const g_variants[] = {&g_variant_128, &g_variant_160, &g_variant_192, &g_variant_224, &g_variant_256};
static const int g_nvariants = sizeof(g_variants) / sizeof(g_variants[0]);

static fused_fn pick_kernel(const Variant &var, const KernelKey &k) {
  if (k.fm) return k.lean_pm > 0 ? var.fm_lean[k.lean_pm][k.gz ? 1 : 0] : var.fm_full[k.thermal ? 1 : 0];
  if (k.lean_pm > 0 && k.opts) return var.opts[k.thermal ? 1 : 0][k.lean_pm];
  if (k.lean_pm > 0 && k.bulk) return (k.ns ? var.bulk_ns : var.bulk)[k.lean_pm][k.gz ? 1 : 0];
  if (k.lean_pm > 0) return (k.ns ? var.lean_ns : var.lean)[k.thermal ? 1 : 0][k.lean_pm][k.gz ? 1 : 0];
  return var.full[k.tau1 ? 1 : 0][k.thermal ? 1 : 0];
}

static int env_int(const char *name, int dflt) {
  const char *s = getenv(name);
  return s && *s ? atoi(s) : dflt;
}

// Launch geometry.  Time model calibrated on B200.  A CTA marching R rows runs R + 8 + D pipeline
// iterations; one iteration of every CTA resident on an SM costs
//   t_iter = max(0.87 us, (CTAs on the SM) * NT * c(NT)),   c(NT) = (1.80 + 0.0018 NT) ns per thread-row
// for launches of several waves, and max(0.87 us, 0.56 us + 1.86 ns * resident threads) for a single partial wave
// (0.87 us: dependent-chain latency of a lone CTA, measured on the 100^2 grid; c(NT) fitted to the compute-bound
// moments-only rates at 8192^2: 2.02 / 2.10 / 2.14 / 2.23 ns for NT = 128 / 160 / 192 / 224 -- more, smaller CTAs hide the
// per-row barrier better), x1.2 for the thermal kernels (single-precision noise on shared Philox blocks; x1.75 with the
// double-precision generator of round 1), x1.25 for tau != 1, x1.3 for the run-time-option kernel.  The launch cannot beat
// the HBM floor (120 B per lattice update, 48 B when the populations are not written, at 5.5 TB/s, tools/membench.cu);
// when several widths are HBM-bound the one with the fewest halo columns per CTA that still keeps >= 3 CTAs per SM wins
// (measured at 8192^2: NT = 224 and 160 lead, 256 with 2 CTAs/SM trails by 4 %).
int choose_geometry(int Lx, int nrows, const KernelKey &key, LaunchGeom *g) {
  int dev = 0, nsm = 148;
  SW_CUDA(cudaGetDevice(&dev));
  SW_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  const int force_nt = env_int("SWALBE_NT", 0);
  // rows marched per CTA: the pipeline fill re-reads 6 rows of h (FM: of all nine populations) per chunk
  const int rmax = std::max(1, env_int("SWALBE_RMAX", key.fm ? 256 : 128));
  const double flavour0 = (key.thermal ? 1.2 : 1.0) * (key.lean_pm > 0 ? 1.0 : 1.3) * (key.tau1 ? 1.0 : 1.25);
  double flavour = flavour0;
  auto t_iter = [&](int ctas_on_sm, int nt) {
    return std::max(0.87, ctas_on_sm * nt * (1.80 + 0.0018 * nt) * 1e-3 * flavour);
  };
  // bytes really moved per lattice update: tau == 1 reads 3 planes and writes 12 (3 when the populations stay lazy);
  // tau != 1 reads 12 and writes 12, or 9 and 9 when the moments are derived from the populations (FM)
  const double bytes_lu = key.tau1 ? (key.lazy ? 48.0 : 120.0) : (key.fm ? 144.0 : 192.0);
  const double hbm_floor = (double)Lx * (double)nrows * bytes_lu / 5.5e6;  // us
  double best_bound_eff = -1.0;  // best W/NT among HBM-bound candidates
  const int fill = 8 + FUSED_D;
  double best_cost = 1e300;
  for (int v = 0; v < g_nvariants; ++v) {
    const Variant &var = *g_variants[v];
    if (force_nt && var.nt != force_nt) continue;
    if (key.bulk && Lx < var.nt) continue;  // a strip may cross the periodic x boundary at most once
    fused_fn fn = pick_kernel(var, key);
    int bps = 0;
    const size_t smem = fused_smem_doubles(var.nt) * sizeof(double);
    SW_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, fn, var.nt, smem));
    if (bps < 1) continue;
    cudaFuncAttributes fattr;
    SW_CUDA(cudaFuncGetAttributes(&fattr, fn));
    const bool spills = fattr.localSizeBytes > 0;  // instantiations that spill under their register cap
    flavour = flavour0 * (spills ? 1.12 : 1.0);
    const int wmax = var.nt - 8;
    const int nstrips = (Lx + wmax - 1) / wmax;
    int W = (Lx + nstrips - 1) / nstrips;
    W = std::min(wmax, (W + 3) & ~3);
    const long long slots = (long long)nsm * bps;
    // candidate chunk counts: powers of two of rows, and the counts that fill 1..64 whole waves
    int cand[160], nc = 0;
    for (int c = 1; c <= nrows && nc < 40; c *= 2) cand[nc++] = c;
    cand[nc++] = nrows;
    cand[nc++] = std::min(nrows, 65535);
    for (int w = 1; w <= 64 && nc < 150; ++w) {
      const long long c = (w * slots) / nstrips;
      if (c >= 1 && c <= nrows) cand[nc++] = (int)c;
      const long long c2 = ((long long)w * nsm) / nstrips;  // one CTA per SM and multiples
      if (c2 >= 1 && c2 <= nrows) cand[nc++] = (int)c2;
    }
    for (int q = 0; q < nc; ++q) {
      int R = (nrows + cand[q] - 1) / cand[q];
      if (R > rmax && nrows > rmax) continue;
      const int nchunks = (nrows + R - 1) / R;
      if (nchunks > 65535) continue;  // gridDim.y limit
      const long long ctas = (long long)nstrips * nchunks;
      double cost;
      if (ctas <= slots) {  // a single, possibly partial wave: c CTAs share an SM
        const int c = (int)((ctas + nsm - 1) / nsm);
        // a single, possibly partial wave (small lattices): fit to the 100^2 .. 2048^2 sweeps
        cost = (R + fill) * std::max(0.87, (0.56 + 0.00186 * c * var.nt) * flavour);
      } else {
        const double waves = (double)ctas / (double)slots;  // CTAs are re-issued as slots free up
        cost = std::ceil(waves - 1e-9) * (R + fill) * t_iter(bps, var.nt);
      }
      bool take;
      if (cost <= 0.95 * hbm_floor && bps >= 3 && !spills) {  // HBM-bound: prefer the least redundant strip decomposition
        const double eff = (double)W / var.nt - 1e-6 * cost;
        take = eff > best_bound_eff;
        if (take) best_bound_eff = eff;
        cost = hbm_floor;
      } else {
        cost = std::max(cost, hbm_floor) + 0.01 * cost;
        take = best_bound_eff < 0.0 && cost < best_cost;
      }
      if (take) {
        best_cost = cost;
        g->nt = var.nt; g->variant = v; g->W = W; g->nstrips = nstrips; g->rows_per_cta = R; g->nchunks = nchunks;
        g->blocks_per_sm = bps;
      }
    }
  }
  if (best_cost == 1e300) return set_error(SWALBE_ERR_CUDA, "no fused-kernel variant fits on this device");
  if (int r = env_int("SWALBE_ROWS", 0)) { g->rows_per_cta = r; g->nchunks = (nrows + r - 1) / r; }
  if (env_int("SWALBE_DEBUG", 0))
    fprintf(stderr, "[swalbe] geometry Lx=%d rows=%d lean_pm=%d tau1=%d thermal=%d: NT=%d W=%d strips=%d rows/CTA=%d chunks=%d "
            "CTAs/SM=%d model %.1f us\n", Lx, nrows, key.lean_pm, (int)key.tau1, (int)key.thermal, g->nt, g->W, g->nstrips,
            g->rows_per_cta, g->nchunks, g->blocks_per_sm, best_cost);
  return 0;
}

int launch_fused(const LaunchGeom &g, const FusedArgs &a, const KernelKey &key, cudaStream_t stream) {
  fused_fn fn = pick_kernel(*g_variants[g.variant], key);
  const int nrows = a.jend - a.jbeg;
  if (nrows <= 0) return 0;
  dim3 grid(g.nstrips, (nrows + a.rows_per_cta - 1) / a.rows_per_cta);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(g.nt);
  cfg.dynamicSmemBytes = fused_smem_doubles(g.nt) * sizeof(double);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  // Programmatic dependent launch is wired (griddepcontrol in the kernel) but OFF by default: measured on B200 it is
  // neutral at >= 2048^2 and slower on small grids (100^2: 8.0 vs 6.2 us/step; 1024^2: 45 vs 35 us/step).
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = env_int("SWALBE_PDL", 0) ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SW_CUDA(cudaLaunchKernelEx(&cfg, fn, a));
  SW_LAUNCH_CHECK();
  return 0;
}

// cp.async.bulk row prefetch: even Lx (16-byte aligned row segments), at most one periodic wrap per strip, HBM-bound size
// SWALBE_BULK: 0 = never, 1 (default) = large lattices only, 2 = whenever the alignment rules allow (tests)
bool bulk_eligible(int Lx, size_t ncells) {
  const int mode = env_int("SWALBE_BULK", 1);
  if (mode == 0 || (Lx % 2) != 0 || Lx < 256) return false;
  return mode >= 2 || ncells >= ((size_t)1 << 22);
}

// the lean kernels cover: tau == 1 and a known (n, m) pressure mode: a strict flavour (scalar theta, standard slip, no
// inclination, no logs; further specialised for gravity == 0 and bulk-copy prefetch) and an OPTS flavour that takes
// those options at run time
KernelKey make_key(const swalbe_params &p, int pmode, bool want_lean) {
  KernelKey k;
  k.tau1 = p.tau == 1.0;
  k.thermal = p.use_thermal != 0;
  k.lean_pm = 0;
  k.bulk = false;
  k.gz = p.g == 0.0;
  k.lazy = false;
  k.fm = false;
  k.ns = env_int("SWALBE_NSYNC", 0) != 0;  // (honoured by the strict lean tau == 1 kernels only)
  k.opts = p.cospi_theta_field != nullptr || p.slip_variant != SWALBE_SLIP_STANDARD || p.use_inclination != 0;
  if (want_lean && k.tau1 && pmode != PM_GENERIC && !env_int("SWALBE_NO_LEAN", 0))
    k.lean_pm = pmode;
  return k;
}

int fill_consts(FusedArgs &a, const swalbe_params &p) {
  if (int e = resolve_pmode(p.pressure_variant, p.n, p.m, &a.pc.pmode)) return e;
  if (p.slip_variant < 0 || p.slip_variant > 2) return set_error(SWALBE_ERR_ARG, "unknown slip_variant %d", p.slip_variant);
  if (!(p.tau > 0.0)) return set_error(SWALBE_ERR_ARG, "tau must be positive");
  a.pc.gamma = p.gamma;
  a.pc.kappa = host_kappa(p.cospi_theta, p.n, p.m, p.hmin);
  a.pc.nm1 = (double)(p.n - 1); a.pc.mm1 = (double)(p.m - 1); a.pc.kden = (double)(p.n - p.m) * p.hmin;
  a.pc.hmin = p.hmin; a.pc.hcrit = p.hcrit; a.pc.n = p.n; a.pc.m = p.m;
  a.sc = make_slip(p.delta, p.mu, p.hcrit, p.slip_variant);
  a.ec = make_eq(p.g);
  a.tc = make_thermal(p.kbt, p.mu, p.delta);
  volatile double it = 1.0 / p.tau;
  volatile double om = 1.0 - it;
  a.invtau = it; a.omega = om;
  a.use_incl = p.use_inclination; a.incl_ax = p.incl_ax; a.incl_ay = p.incl_ay; a.incl_factor = p.incl_factor;
  a.pk = make_philox_key(p.seed);
  return 0;
}

}  // namespace swalbe

using namespace swalbe;

// what a captured loop depends on: every pointer and parameter baked into its kernel nodes
struct GraphKey {
  const void *ptr[17];
  double d[13];
  int i[8];
  const void *ct;
};

struct swalbe_plan {
  int Lx, Ly;
  double *scratch;  // 3 moment planes (ping-pong partner of the caller's height/velx/vely)
  double *rowsum;   // row sums of the height (mass log): rowsum_planes x Ly, allocated on first use
  int rowsum_planes;
  double *log_part; // partial per-step logs of the cluster kernel (grown on demand)
  size_t log_part_doubles;
  int graph_launches;  // kernels inside the captured graph (the launch counter advances by this on replay)
  // CUDA graph of the last repeated loop (small, launch-bound lattices): built the second time the same call is seen
  cudaStream_t cap_stream;
  cudaGraphExec_t graph_exec;
  GraphKey graph_key, seen_key;
  bool have_graph, have_seen;
  int graph_nsteps;
  // launch geometry per kernel flavour, chosen the first time the flavour is used
  struct GeomEntry { KernelKey key; int nrows; LaunchGeom geom; };
  GeomEntry geoms[96];
  int ngeoms;
  // swalbe_time_loop_host: copy streams (one per direction: PCIe is full duplex) and their events
  cudaStream_t s_h2d, s_d2h, s_aux;  // s_aux: second compute stream of the sweeps (odd band stages)
  cudaEvent_t ev_user, ev_dn, ev_done, ev_up[64], ev_join, ev_phase, ev_wave[2][33];
  bool have_host_streams;
};

// nrows: rows of one launch (the whole lattice, or a band / the seam strips of the host loop's sweeps)
static int plan_geometry(swalbe_plan *plan, const KernelKey &k, LaunchGeom **g, int nrows = 0) {
  auto same = [](const KernelKey &x, const KernelKey &y) {
    return x.tau1 == y.tau1 && x.thermal == y.thermal && x.lean_pm == y.lean_pm && x.bulk == y.bulk && x.gz == y.gz &&
           x.lazy == y.lazy && x.opts == y.opts && x.fm == y.fm && x.ns == y.ns;
  };
  if (nrows <= 0) nrows = plan->Ly;
  for (int q = 0; q < plan->ngeoms; ++q)
    if (plan->geoms[q].nrows == nrows && same(plan->geoms[q].key, k)) { *g = &plan->geoms[q].geom; return 0; }
  const int slot = plan->ngeoms < 96 ? plan->ngeoms++ : 95;  // (a plan sees a handful of flavours; the last slot is recycled)
  plan->geoms[slot].key = k;
  plan->geoms[slot].nrows = nrows;
  if (int e = choose_geometry(plan->Lx, nrows, k, &plan->geoms[slot].geom)) { plan->ngeoms = slot; return e; }
  *g = &plan->geoms[slot].geom;
  return 0;
}

extern "C" {

int swalbe_plan_create(swalbe_plan **plan, int Lx, int Ly) {
  if (!plan) return set_error(SWALBE_ERR_ARG, "plan is NULL");
  if (int e = check_extent(Lx, Ly)) return e;
  swalbe_plan *p = new swalbe_plan();
  p->Lx = Lx; p->Ly = Ly; p->scratch = nullptr; p->rowsum = nullptr; p->rowsum_planes = 0; p->log_part = nullptr; p->log_part_doubles = 0; p->graph_launches = 0;
  p->cap_stream = nullptr; p->graph_exec = nullptr; p->have_graph = p->have_seen = false; p->graph_nsteps = 0;
  p->ngeoms = 0;
  p->s_h2d = p->s_d2h = nullptr; p->have_host_streams = false;
  cudaError_t e = cudaMalloc((void **)&p->scratch, sizeof(double) * 3 * (size_t)Lx * Ly);
  if (e != cudaSuccess) {
    delete p;
    return set_error(SWALBE_ERR_CUDA, "cudaMalloc(plan scratch) failed: %s", cudaGetErrorString(e));
  }
  *plan = p;
  return 0;
}

int swalbe_plan_destroy(swalbe_plan *plan) {
  if (!plan) return 0;
  if (plan->graph_exec) cudaGraphExecDestroy(plan->graph_exec);
  if (plan->cap_stream) cudaStreamDestroy(plan->cap_stream);
  if (plan->have_host_streams) {  // (copies still in flight are drained by cudaStreamDestroy's deferred release)
    cudaStreamDestroy(plan->s_h2d); cudaStreamDestroy(plan->s_d2h); cudaStreamDestroy(plan->s_aux);
    cudaEventDestroy(plan->ev_user); cudaEventDestroy(plan->ev_dn); cudaEventDestroy(plan->ev_done);
    cudaEventDestroy(plan->ev_join); cudaEventDestroy(plan->ev_phase);
    for (cudaEvent_t ev : plan->ev_up) cudaEventDestroy(ev);
    for (int q = 0; q < 2; ++q)
      for (cudaEvent_t ev : plan->ev_wave[q]) cudaEventDestroy(ev);
  }
  cudaFree(plan->scratch);
  cudaFree(plan->rowsum);
  cudaFree(plan->log_part);
  delete plan;
  return 0;
}

static int enqueue_steps(swalbe_plan *plan, const swalbe_state *st, const swalbe_params *prm, int nsteps,
                         unsigned long long step0, int flags, const swalbe_loop_logs *logs, cudaStream_t stream);
static int enqueue_steps_dumps(swalbe_plan *plan, const swalbe_state *st, const swalbe_params *prm, int nsteps,
                               unsigned long long step0, int flags, const swalbe_loop_logs *logs, cudaStream_t stream);

static GraphKey make_graph_key(const swalbe_state *st, const swalbe_params *p, int nsteps, int flags) {
  GraphKey k;
  memset(&k, 0, sizeof(k));
  const void *ptrs[17] = {st->fout, st->ftemp, st->feq, st->height, st->velx, st->vely, st->vsq, st->pressure, st->Fx, st->Fy,
                          st->slipx, st->slipy, st->hgradpx, st->hgradpy, st->dgrad, st->kbtx, st->kbty};
  memcpy(k.ptr, ptrs, sizeof(ptrs));
  const double d[13] = {p->tau, p->mu, p->delta, p->kbt, p->gamma, p->hmin, p->hcrit, p->g, p->cospi_theta,
                        p->incl_ax, p->incl_ay, p->incl_factor, 0.0};
  memcpy(k.d, d, sizeof(d));
  const int i[8] = {p->n, p->m, p->pressure_variant, p->slip_variant, p->use_inclination, p->use_thermal, nsteps, flags};
  memcpy(k.i, i, sizeof(i));
  k.ct = p->cospi_theta_field;
  return k;
}

// Launch-bound lattices (<= SWALBE_GRAPH_MAX sites, default 1024^2): a loop that is repeated with the same state, parameters
// and length -- the drivers' chunks between two mass prints -- is captured into a CUDA graph the second time it is seen
// and replayed from then on (measured on B200: 100^2 6.2 -> 4.9 us/step, 256^2 8.2 -> 5.6, 512^2 14.4 -> 12.5).
// Not for thermal loops (the step counter is baked into the nodes) or per-step logs (their slots move).
static bool graph_candidate(const swalbe_plan *plan, const swalbe_params *prm, int nsteps, const swalbe_loop_logs *logs,
                            cudaStream_t stream) {
  if (!env_int("SWALBE_GRAPH", 1) || nsteps < 8 || nsteps > 16384 || prm->use_thermal) return false;  // (graph size bound)
  if (logs && (logs->hmin || logs->hmax || logs->wetted || logs->hsum)) return false;
  if ((size_t)plan->Lx * plan->Ly > (size_t)std::max(0, env_int("SWALBE_GRAPH_MAX", 1024 * 1024))) return false;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess) { cudaGetLastError(); return false; }
  return cs == cudaStreamCaptureStatusNone;  // inside somebody else's capture: just enqueue
}

int swalbe_time_loop(swalbe_plan *plan, const swalbe_state *st, const swalbe_params *prm, int nsteps,
                     unsigned long long step0, int flags, const swalbe_loop_logs *logs, void *stream_) {
  if (!plan || !st || !prm) return set_error(SWALBE_ERR_ARG, "swalbe_time_loop: NULL plan/state/params");
  if (nsteps < 0) return set_error(SWALBE_ERR_ARG, "nsteps < 0");
  if (nsteps == 0) return 0;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!graph_candidate(plan, prm, nsteps, logs, stream)) return enqueue_steps_dumps(plan, st, prm, nsteps, step0, flags, logs, stream);
  const GraphKey key = make_graph_key(st, prm, nsteps, flags);
  if (plan->have_graph && memcmp(&key, &plan->graph_key, sizeof(key)) == 0) {
    SW_CUDA(cudaGraphLaunch(plan->graph_exec, stream));
    count_launch((unsigned)plan->graph_launches);
    return 0;
  }
  if (!(plan->have_seen && memcmp(&key, &plan->seen_key, sizeof(key)) == 0)) {  // first sighting: plain launches
    plan->seen_key = key; plan->have_seen = true;
    return enqueue_steps(plan, st, prm, nsteps, step0, flags, logs, stream);
  }
  // second sighting: capture the same launches on the plan's own stream (the caller's may be the legacy stream, which
  // cannot be captured), instantiate, and launch the graph in the caller's stream
  if (!plan->cap_stream) SW_CUDA(cudaStreamCreateWithFlags(&plan->cap_stream, cudaStreamNonBlocking));
  SW_CUDA(cudaStreamBeginCapture(plan->cap_stream, cudaStreamCaptureModeThreadLocal));
  const unsigned long long launches0 = swalbe_launch_count();
  const int rc = enqueue_steps(plan, st, prm, nsteps, step0, flags, logs, plan->cap_stream);
  const int captured = (int)(swalbe_launch_count() - launches0);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(plan->cap_stream, &graph);
  if (rc != 0 || ce != cudaSuccess || !graph) {  // capture failed: fall back to plain launches, never to silence
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    plan->have_seen = false;
    if (rc != 0) return rc;
    return enqueue_steps(plan, st, prm, nsteps, step0, flags, logs, stream);
  }
  if (plan->graph_exec) { cudaGraphExecDestroy(plan->graph_exec); plan->graph_exec = nullptr; plan->have_graph = false; }
  const cudaError_t ie = cudaGraphInstantiate(&plan->graph_exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess) {
    plan->graph_exec = nullptr;
    return set_error(SWALBE_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
  }
  plan->graph_key = key; plan->have_graph = true; plan->graph_nsteps = nsteps; plan->graph_launches = captured;
  SW_CUDA(cudaGraphLaunch(plan->graph_exec, stream));
  return 0;
}

}  // extern "C"

static int enqueue_steps(swalbe_plan *plan, const swalbe_state *st, const swalbe_params *prm, int nsteps,
                         unsigned long long step0, int flags, const swalbe_loop_logs *logs, cudaStream_t stream) {
  const int Lx = plan->Lx, Ly = plan->Ly;
  const size_t N = (size_t)Lx * Ly;
#define NEED(f) if (!st->f) return set_error(SWALBE_ERR_ARG, "swalbe_time_loop: state." #f " is NULL")
  NEED(fout); NEED(ftemp); NEED(feq); NEED(height); NEED(velx); NEED(vely); NEED(vsq); NEED(pressure);
  NEED(Fx); NEED(Fy); NEED(slipx); NEED(slipy); NEED(hgradpx); NEED(hgradpy);
#undef NEED
  const bool tau1 = prm->tau == 1.0;
  const bool thermal = prm->use_thermal != 0;
  if (thermal && (!st->kbtx || !st->kbty)) return set_error(SWALBE_ERR_ARG, "thermal loop needs state.kbtx/kbty");
  const bool lazy = (flags & SWALBE_LOOP_LAZY_POPULATIONS) != 0;
  const bool skip_aux = (flags & SWALBE_LOOP_SKIP_AUX) != 0;
  if (lazy && !tau1) return set_error(SWALBE_ERR_ARG, "SWALBE_LOOP_LAZY_POPULATIONS requires tau == 1");

  FusedArgs a = {};
  if (int e = fill_consts(a, *prm)) return e;
  KernelKey key_full = make_key(*prm, a.pc.pmode, false);
  KernelKey key_mid = make_key(*prm, a.pc.pmode, true);  // lean kernel for the steps before the last
  if (logs && (logs->hmin || logs->wetted)) key_mid.opts = true;
  // tau != 1: every step after the first derives h and u from the populations the previous step streamed (FM kernels:
  // 72 B read + 72 B written per lattice update); the first one reads the caller's planes unless the caller vouches
  // for them (an initial condition written into `height` is NOT the zeroth moment of ftemp: src/simulate.jl:349-356).
  const bool fm = !tau1 && env_int("SWALBE_FM", 1) != 0;
  const bool fm_first = fm && (flags & SWALBE_LOOP_MOMENTS_CONSISTENT) != 0;
  KernelKey key_first = key_mid, key_first_full = key_full;  // step 0 when it has to read the moment planes
  if (fm) {
    key_full.fm = key_mid.fm = true;
    if (!key_mid.opts && !key_mid.thermal && a.pc.pmode != PM_GENERIC && !env_int("SWALBE_NO_LEAN", 0)) key_mid.lean_pm = a.pc.pmode;
    // measured on B200 (8192^2, tau = 0.9): the L2 prefetch changes nothing up to 3 rows ahead and costs 5-13 % beyond
    // (profiles/r02_fm_sweep.txt), so it is off; what the second read of a row competes with in L2 is the stream of
    // new populations, hence the residency hints
    a.fm_prefetch = std::max(0, std::min(16, env_int("SWALBE_FM_PREFETCH", 0)));
    // streaming stores: +1.3 %; evict_last / evict_first on the two reads cut the DRAM traffic from 160 to 146 B per
    // update but expose the DRAM latency of the first read (26 % of the stall samples on its first use): 3 % slower
    a.fm_hints = env_int("SWALBE_FM_HINTS", 1) & 7;
  }
  // bulk-copy (TMA unit) row prefetch: needs 16-byte aligned row segments, i.e. even Lx and 16-B aligned planes
  auto aligned16 = [](const void *p) { return ((uintptr_t)p & 15u) == 0; };
  // Measured on B200: +3 % where the step is HBM-bound (8192^2: 43.2 vs 41.9 GLUPS), -2..3 % where it is latency- or
  // compute-bound (1024^2, moments-only steps) -> used for large lattices whose populations are written every step.
  key_mid.bulk = key_mid.lean_pm > 0 && !key_mid.opts && !key_mid.thermal && !lazy && bulk_eligible(Lx, N) && aligned16(st->height) &&
                 aligned16(st->velx) && aligned16(st->vely) && aligned16(plan->scratch);
  key_mid.lazy = lazy;
  LaunchGeom *g_full = nullptr, *g_mid = nullptr, *g_first = nullptr, *g_first_full = nullptr;
  if (int e = plan_geometry(plan, key_full, &g_full)) return e;
  if (int e = plan_geometry(plan, key_mid, &g_mid)) return e;
  if (int e = plan_geometry(plan, key_first, &g_first)) return e;
  if (int e = plan_geometry(plan, key_first_full, &g_first_full)) return e;
  a.Lx = Lx; a.Ly = Ly; a.jbeg = 0; a.jend = Ly;
  a.wrap_y = 1; a.jglobal0 = 0; a.Ly_global = Ly;
  a.fstride_in = a.fstride_out = a.fstride_out2 = N;
  a.ct_field = prm->cospi_theta_field;

  const bool log_mm = logs && logs->hmin && logs->hmax;
  const bool log_wet = logs && logs->wetted;
  if (logs && ((logs->hmin == nullptr) != (logs->hmax == nullptr)))
    return set_error(SWALBE_ERR_ARG, "logs.hmin and logs.hmax must both be set or both NULL");
  if (log_mm || log_wet) {
    k_init_logs<<<(nsteps + 255) / 256, 256, 0, stream>>>(log_mm ? logs->hmin : nullptr, log_mm ? logs->hmax : nullptr,
                                                            log_wet ? logs->wetted : nullptr, nsteps);
    SW_LAUNCH_CHECK();
    a.hthresh = logs->hthresh;
  }

  // lattices up to SWALBE_TILE_MAX sites run their lean steps through the tile kernel (tile.cu): measured on B200 it wins
  // where the marching kernel is latency-bound
  const size_t tile_max = (size_t)std::max(0, env_int("SWALBE_TILE_MAX", 512 * 512));

  // moment ping-pong: the caller's planes (A) and the plan's scratch (B); arrange for the LAST step to land in A
  double *A[3] = {st->height, st->velx, st->vely};
  double *B[3] = {plan->scratch, plan->scratch + N, plan->scratch + 2 * N};

  // Lattices that fit the shared memory of one thread-block cluster (<= ~128^2; the README example is 100^2) run all
  // their lean steps inside ONE persistent kernel launch, in place on the caller's planes (cluster.cuh); the step that
  // materialises the reference's intermediate fields, if any, follows through the ordinary path.
  int s_begin = 0;
  {
    a.h_in = A[0]; a.ux_in = A[1]; a.uy_in = A[2];
    a.log_min = log_mm ? logs->hmin : nullptr; a.log_wet = log_wet ? logs->wetted : nullptr;  // (eligibility only)
    int R = 0;
    size_t smem_bytes = 0;
    const int ncl = skip_aux ? nsteps : nsteps - 1;
    const int C = ncl >= 1 ? cluster_plan(key_mid, a, &R, &smem_bytes) : 0;
    if (C > 0) {
      ClusterArgs ca = {};
      ca.a = a;
      ca.a.h_out = A[0]; ca.a.ux_out = A[1]; ca.a.uy_out = A[2];
      ca.a.f_out = st->fout;
      ca.a.f_out2 = ncl == nsteps ? st->ftemp : nullptr;  // fout == ftemp on return (src/collide.jl:103)
      ca.nsteps = ncl; ca.lazy = lazy ? 1 : 0; ca.rows_max = R;
      if (log_mm || log_wet) {
        const size_t need = (size_t)ncl * C * 3;
        if (need > plan->log_part_doubles) {  // (one-time growth; loops with logs are never graph-captured)
          if (plan->log_part) SW_CUDA(cudaFree(plan->log_part));
          plan->log_part = nullptr; plan->log_part_doubles = 0;
          SW_CUDA(cudaMalloc((void **)&plan->log_part, need * sizeof(double)));
          plan->log_part_doubles = need;
        }
        ca.log_part = plan->log_part;
      }
      if (int e = launch_cluster(ca, key_mid, C, smem_bytes, stream)) return e;
      if (ca.log_part)
        if (int e = launch_cluster_logs(ca.log_part, C, ncl, log_mm ? logs->hmin : nullptr, log_mm ? logs->hmax : nullptr,
                                        log_wet ? logs->wetted : nullptr, stream)) return e;
      s_begin = ncl;
    }
  }
  const int nrem = nsteps - s_begin;
  bool src_is_A = true;
  // FM: the moment planes are read by step 0 at most and written by the last step only -- no ping-pong, and no copy
  // unless the same launch does both
  const bool fm_pingpong_free = fm && (fm_first || nsteps > 1);
  if (fm_pingpong_free ? false : (nrem & 1)) {
    for (int q = 0; q < 3; ++q) SW_CUDA(cudaMemcpyAsync(B[q], A[q], sizeof(double) * N, cudaMemcpyDeviceToDevice, stream));
    src_is_A = false;
  }
  // population ping-pong (tau != 1): the reference reads ftemp; streamed result alternates fout/ftemp
  bool fsrc_is_ftemp = true;

  for (int s = s_begin; s < nsteps; ++s) {
    const bool last = s == nsteps - 1;
    double **src = src_is_A ? A : B, **dst = src_is_A ? B : A;
    a.h_in = src[0]; a.ux_in = src[1]; a.uy_in = src[2];
    a.h_out = dst[0]; a.ux_out = dst[1]; a.uy_out = dst[2];
    const bool step_fm = fm && (s > 0 || fm_first);
    if (fm_pingpong_free) {
      a.h_in = A[0]; a.ux_in = A[1]; a.uy_in = A[2];  // (read by a non-FM step 0 only)
      a.h_out = last ? A[0] : nullptr; a.ux_out = last ? A[1] : nullptr; a.uy_out = last ? A[2] : nullptr;
    }
    if (tau1) {
      a.f_in = nullptr;
      a.f_out = (!lazy || last) ? st->fout : nullptr;
      a.f_out2 = last ? st->ftemp : nullptr;  // fout == ftemp on return (src/collide.jl:103)
    } else {
      a.f_in = fsrc_is_ftemp ? st->ftemp : st->fout;
      a.f_out = fsrc_is_ftemp ? st->fout : st->ftemp;
      a.f_out2 = nullptr;
    }
    if (last && !skip_aux) {  // materialise every intermediate field the reference's state would hold
      a.pressure = st->pressure; a.hgx = st->hgradpx; a.hgy = st->hgradpy; a.slipx = st->slipx; a.slipy = st->slipy;
      a.Fx = st->Fx; a.Fy = st->Fy; a.feq = st->feq; a.vsq = st->vsq;
      a.kbtx = thermal ? st->kbtx : nullptr; a.kbty = thermal ? st->kbty : nullptr;
    }
    a.step = step0 + (unsigned long long)s;
    a.log_min = log_mm ? logs->hmin + s : nullptr;
    a.log_max = log_mm ? logs->hmax + s : nullptr;
    a.log_wet = log_wet ? logs->wetted + s : nullptr;
    const bool use_full = last && !skip_aux;
    const KernelKey &key = (fm && !step_fm) ? (use_full ? key_first_full : key_first) : (use_full ? key_full : key_mid);
    const LaunchGeom &g = (fm && !step_fm) ? (use_full ? *g_first_full : *g_first) : (use_full ? *g_full : *g_mid);
    a.rows_per_cta = g.rows_per_cta; a.W = g.W;
    if (!use_full && N <= tile_max && tile_eligible(key_mid, a)) {  // latency-bound lattices: three-phase tile kernel
      if (int e = launch_tile(a, key_mid, stream)) return e;
    } else if (int e = launch_fused(g, a, key, stream)) return e;
    src_is_A = !src_is_A;
    fsrc_is_ftemp = !fsrc_is_ftemp;
  }
  if (!tau1) {  // make fout == ftemp: the newest populations are in the array written last
    double *newest = fsrc_is_ftemp ? st->ftemp : st->fout;  // (flag already flipped: source of the NEXT step)
    double *other = fsrc_is_ftemp ? st->fout : st->ftemp;
    SW_CUDA(cudaMemcpyAsync(other, newest, sizeof(double) * 9 * N, cudaMemcpyDeviceToDevice, stream));
  }
  return 0;
}

// ---- mass log: sum(height) BEFORE selected steps (logs->hsum) -----------------------------------------------------------

struct MassLog {
  double *out = nullptr;  // device-side address of logs->hsum
  int first = 0, every = 0, nsteps = 0;
  bool on() const { return out != nullptr; }
  bool wants(int i) const { return on() && i >= first && i < nsteps && (i - first) % every == 0; }
  double *slot(int i) const { return out + (i - first) / every; }
  int next_after(int i) const {  // smallest logged state index > i, or nsteps
    if (!on()) return nsteps;
    int n = i < first ? first : first + ((i - first) / every + 1) * every;
    return n < nsteps ? n : nsteps;
  }
};

// planes: states whose rows may be summed concurrently (sweeps of the host loop that span more than one logged step)
static int mass_log_setup(swalbe_plan *plan, const swalbe_loop_logs *logs, int nsteps, MassLog *m, int planes = 1) {
  *m = MassLog();
  m->nsteps = nsteps;
  if (!logs || !logs->hsum) return 0;
  if (logs->hsum_every < 1 || logs->hsum_first < 0) return set_error(SWALBE_ERR_ARG, "logs.hsum needs hsum_every >= 1 and hsum_first >= 0");
  cudaPointerAttributes at;
  SW_CUDA(cudaPointerGetAttributes(&at, logs->hsum));
  if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) m->out = logs->hsum;
  else if (at.type == cudaMemoryTypeHost && at.devicePointer) m->out = (double *)at.devicePointer;
  else return set_error(SWALBE_ERR_ARG, "logs.hsum must be device memory or page-locked (mapped) host memory");
  m->first = logs->hsum_first; m->every = logs->hsum_every;
  if (plan->rowsum_planes < planes) {
    if (plan->rowsum) SW_CUDA(cudaFree(plan->rowsum));  // (stream-ordered with every launch that used it: cudaFree synchronises)
    plan->rowsum = nullptr; plan->rowsum_planes = 0;
    SW_CUDA(cudaMalloc((void **)&plan->rowsum, sizeof(double) * (size_t)plan->Ly * planes));
    plan->rowsum_planes = planes;
  }
  return 0;
}
static int mass_rows(swalbe_plan *plan, const double *h, int jbeg, int jend, cudaStream_t stream, int plane = 0) {
  if (jend <= jbeg) return 0;
  k_rowsum<<<jend - jbeg, 256, 0, stream>>>(h, plan->Lx, jbeg, plan->rowsum + (size_t)plane * plan->Ly);
  SW_LAUNCH_CHECK();
  return 0;
}
static int mass_final(swalbe_plan *plan, double *slot, cudaStream_t stream, int plane = 0) {
  k_rowsum_final<<<1, 1024, 0, stream>>>(plan->rowsum + (size_t)plane * plan->Ly, plan->Ly, slot);
  SW_LAUNCH_CHECK();
  return 0;
}

// the plain loop cut at the logged steps (every piece lands in the caller's planes, where the rows are summed)
static int enqueue_steps_dumps(swalbe_plan *plan, const swalbe_state *st, const swalbe_params *prm, int nsteps,
                               unsigned long long step0, int flags, const swalbe_loop_logs *logs, cudaStream_t stream) {
  MassLog ml;
  if (int e = mass_log_setup(plan, logs, nsteps, &ml)) return e;
  if (!ml.on()) return enqueue_steps(plan, st, prm, nsteps, step0, flags, logs, stream);
  if (!st->height) return set_error(SWALBE_ERR_ARG, "swalbe_time_loop: state.height is NULL");
  for (int s = 0; s < nsteps;) {
    if (ml.wants(s)) {
      if (int e = mass_rows(plan, st->height, 0, plan->Ly, stream)) return e;
      if (int e = mass_final(plan, ml.slot(s), stream)) return e;
    }
    const int nxt = ml.next_after(s);
    swalbe_loop_logs sub = *logs;
    sub.hsum = nullptr;
    if (sub.hmin) sub.hmin += s;
    if (sub.hmax) sub.hmax += s;
    if (sub.wetted) sub.wetted += s;
    const int f = flags | (nxt < nsteps ? SWALBE_LOOP_SKIP_AUX : 0) | (s > 0 ? SWALBE_LOOP_MOMENTS_CONSISTENT : 0);
    if (int e = enqueue_steps(plan, st, prm, nxt - s, step0 + (unsigned long long)s, f, &sub, stream)) return e;
    s = nxt;
  }
  return 0;
}

// ---- time loop from / to host memory (sweep.h) --------------------------------------------------------------------------

static int host_streams(swalbe_plan *plan) {
  if (plan->have_host_streams) return 0;
  SW_CUDA(cudaStreamCreateWithFlags(&plan->s_h2d, cudaStreamNonBlocking));
  SW_CUDA(cudaStreamCreateWithFlags(&plan->s_d2h, cudaStreamNonBlocking));
  SW_CUDA(cudaStreamCreateWithFlags(&plan->s_aux, cudaStreamNonBlocking));
  SW_CUDA(cudaEventCreateWithFlags(&plan->ev_join, cudaEventDisableTiming));
  SW_CUDA(cudaEventCreateWithFlags(&plan->ev_phase, cudaEventDisableTiming));
  for (int q = 0; q < 2; ++q)
    for (cudaEvent_t &ev : plan->ev_wave[q]) SW_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  SW_CUDA(cudaEventCreateWithFlags(&plan->ev_user, cudaEventDisableTiming));
  SW_CUDA(cudaEventCreateWithFlags(&plan->ev_dn, cudaEventDisableTiming));
  SW_CUDA(cudaEventCreateWithFlags(&plan->ev_done, cudaEventDisableTiming));
  for (cudaEvent_t &ev : plan->ev_up) SW_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  plan->have_host_streams = true;
  return 0;
}

static SweepConfig host_loop_config(int Lx, int Ly, int nsteps, bool has_in, bool has_out) {
  if (!env_int("SWALBE_HOST_STREAM", 1)) return SweepConfig{0, 0, 0, 0};
  return sweep_configure(Lx, Ly, nsteps, has_in, has_out, env_int("SWALBE_BAND_ROWS", 0), env_int("SWALBE_HOST_KMAX", 0),
                         (long long)env_int("SWALBE_HOST_MIN_SITES", 0));
}

static int enqueue_steps_host(swalbe_plan *plan, const swalbe_state *st, const swalbe_params *prm, int nsteps,
                              unsigned long long step0, int flags, const swalbe_loop_logs *logs, const double *hin,
                              double *hout, cudaStream_t stream) {
  const int Lx = plan->Lx, Ly = plan->Ly;
  const size_t N = (size_t)Lx * Ly;
  if (!st->height) return set_error(SWALBE_ERR_ARG, "swalbe_time_loop_host: state.height is NULL");
  const bool tau1 = prm->tau == 1.0;
  SweepConfig cfg = host_loop_config(Lx, Ly, nsteps, hin != nullptr, hout != nullptr);
  if (!tau1) cfg.nbands = 0;  // (the sweeps are wired for the kernels that read no populations)
  if (cfg.nbands == 0) {      // small lattices, tau != 1: the copies simply bracket the loop on the caller's stream
    if (hin) SW_CUDA(cudaMemcpyAsync(st->height, hin, N * sizeof(double), cudaMemcpyHostToDevice, stream));
    if (nsteps > 0)
      if (int e = enqueue_steps_dumps(plan, st, prm, nsteps, step0, flags, logs, stream)) return e;
    if (hout) SW_CUDA(cudaMemcpyAsync(hout, st->height, N * sizeof(double), cudaMemcpyDeviceToHost, stream));
    return 0;
  }
#define NEED(f) if (!st->f) return set_error(SWALBE_ERR_ARG, "swalbe_time_loop_host: state." #f " is NULL")
  NEED(fout); NEED(ftemp); NEED(feq); NEED(velx); NEED(vely); NEED(vsq); NEED(pressure);
  NEED(Fx); NEED(Fy); NEED(slipx); NEED(slipy); NEED(hgradpx); NEED(hgradpy);
#undef NEED
  const bool thermal = prm->use_thermal != 0;
  if (thermal && (!st->kbtx || !st->kbty)) return set_error(SWALBE_ERR_ARG, "thermal loop needs state.kbtx/kbty");
  const bool lazy = (flags & SWALBE_LOOP_LAZY_POPULATIONS) != 0;
  const bool skip_aux = (flags & SWALBE_LOOP_SKIP_AUX) != 0;
  if (logs && ((logs->hmin == nullptr) != (logs->hmax == nullptr)))
    return set_error(SWALBE_ERR_ARG, "logs.hmin and logs.hmax must both be set or both NULL");
  if (int e = host_streams(plan)) return e;

  FusedArgs a = {};
  if (int e = fill_consts(a, *prm)) return e;
  const KernelKey key_full = make_key(*prm, a.pc.pmode, false);
  KernelKey key_mid = make_key(*prm, a.pc.pmode, true);
  if (logs && (logs->hmin || logs->wetted)) key_mid.opts = true;
  auto aligned16 = [](const void *p) { return ((uintptr_t)p & 15u) == 0; };
  key_mid.bulk = key_mid.lean_pm > 0 && !key_mid.opts && !key_mid.thermal && !lazy && bulk_eligible(Lx, N) && aligned16(st->height) &&
                 aligned16(st->velx) && aligned16(st->vely) && aligned16(plan->scratch);
  key_mid.lazy = lazy;
  const int kphase = std::max(cfg.k_up, cfg.k_dn);
  const int band_rows = Ly / cfg.nbands, seam_rows = std::max(3, 3 * kphase / 2 + 1);
  // [whole lattice | band | seam strip] x [lean step | the step that ends the call]
  LaunchGeom *g_mid[3] = {nullptr, nullptr, nullptr}, *g_full[3] = {nullptr, nullptr, nullptr};
  const int geom_rows[3] = {Ly, band_rows, seam_rows};
  for (int q = 0; q < 3; ++q) {
    if (int e = plan_geometry(plan, key_mid, &g_mid[q], geom_rows[q])) return e;
    if (int e = plan_geometry(plan, key_full, &g_full[q], geom_rows[q])) return e;
  }
  a.Lx = Lx; a.Ly = Ly;
  a.wrap_y = 1; a.jglobal0 = 0; a.Ly_global = Ly;
  a.fstride_in = a.fstride_out = a.fstride_out2 = N;
  a.ct_field = prm->cospi_theta_field;
  const bool log_mm = logs && logs->hmin && logs->hmax;
  const bool log_wet = logs && logs->wetted;
  // mass log: the rows of a logged state are summed right behind the launch (or upload) that produces them
  MassLog ml;
  const int mass_planes = (logs && logs->hsum && logs->hsum_every > 0) ? kphase / logs->hsum_every + 2 : 1;
  if (int e = mass_log_setup(plan, logs, nsteps, &ml, mass_planes)) return e;
  std::vector<int> mass_rows_done(mass_planes, 0), mass_pending(mass_planes, -1);
  bool aux_busy = false;   // work on the second compute stream that the caller's stream has not waited for yet
  // The rows of a state are summed on the stream of the launch (or upload wait) that makes them final -- before the next
  // launches of the same band stage overwrite them.  The fold over the row sums needs all of them: it is issued on the
  // caller's stream, at once when no piece can still be in flight on the second stream, else at the next join.
  auto mass_piece = [&](int state, const double *h, int jbeg, int jend, cudaStream_t on) -> int {
    const int pl = ((state - ml.first) / ml.every) % mass_planes;
    if (int e = mass_rows(plan, h, jbeg, jend, on, pl)) return e;
    if (on != stream && jend > jbeg) aux_busy = true;
    mass_rows_done[pl] += jend - jbeg;
    if (mass_rows_done[pl] == Ly) {
      mass_rows_done[pl] = 0;
      if (on == stream && !aux_busy) return mass_final(plan, ml.slot(state), stream, pl);
      mass_pending[pl] = state;
    }
    return 0;
  };
  int up_beg[64] = {0}, up_end[64] = {0};

  // moment ping-pong as in the plain loop: the last step lands in the caller's planes (A), so step s reads A when
  // (nsteps - s) is even; the upload goes straight into the planes step 0 reads
  double *A[3] = {st->height, st->velx, st->vely};
  double *B[3] = {plan->scratch, plan->scratch + N, plan->scratch + 2 * N};
  const bool src0_is_A = (nsteps % 2) == 0;
  double **src0 = src0_is_A ? A : B;

  // SWALBE_HOST_TRACE=1 (diagnostics; blocks the host at the end of the call): device timeline of the copies and sweeps
  const bool trace = env_int("SWALBE_HOST_TRACE", 0) != 0, nocopy = env_int("SWALBE_HOST_NOCOPY", 0) != 0;
  struct Mark { cudaEvent_t ev; char what[48]; };
  std::vector<Mark> marks;
  auto mark = [&](cudaStream_t s_, const char *fmt, int x, int y) {
    if (!trace) return;
    Mark m;
    cudaEventCreate(&m.ev);
    snprintf(m.what, sizeof(m.what), fmt, x, y);
    cudaEventRecord(m.ev, s_);
    marks.push_back(m);
  };
  // order the copy streams after whatever the caller has queued on the state (the upload overwrites a plane of it)
  mark(stream, "start", 0, 0);
  SW_CUDA(cudaEventRecord(plan->ev_user, stream));
  SW_CUDA(cudaStreamWaitEvent(plan->s_h2d, plan->ev_user, 0));
  SW_CUDA(cudaStreamWaitEvent(plan->s_d2h, plan->ev_user, 0));
  const std::vector<SweepOp> ops = sweep_schedule(cfg, Ly, nsteps, hin != nullptr, hout != nullptr);
  for (const SweepOp &op : ops)  // every upload is queued up front: the copy engine never waits for the launch loop
    if (op.kind == SWEEP_UPLOAD) {
      const size_t off = (size_t)op.jbeg * Lx, cnt = (size_t)(op.jend - op.jbeg) * Lx;
      if (!nocopy) SW_CUDA(cudaMemcpyAsync(src0[0] + off, hin + off, cnt * sizeof(double), cudaMemcpyHostToDevice, plan->s_h2d));
      SW_CUDA(cudaEventRecord(plan->ev_up[op.band], plan->s_h2d));
      up_beg[op.band] = op.jbeg; up_end[op.band] = op.jend;
      mark(plan->s_h2d, "upload band %d done", op.band, 0);
    }
  if (!src0_is_A) {
    if (!hin) SW_CUDA(cudaMemcpyAsync(B[0], A[0], sizeof(double) * N, cudaMemcpyDeviceToDevice, stream));
    for (int q = 1; q < 3; ++q) SW_CUDA(cudaMemcpyAsync(B[q], A[q], sizeof(double) * N, cudaMemcpyDeviceToDevice, stream));
  }
  if (log_mm || log_wet) {
    k_init_logs<<<(nsteps + 255) / 256, 256, 0, stream>>>(log_mm ? logs->hmin : nullptr, log_mm ? logs->hmax : nullptr,
                                                            log_wet ? logs->wetted : nullptr, nsteps);
    SW_LAUNCH_CHECK();
    a.hthresh = logs->hthresh;
  }
  if (ml.wants(0) && !hin)
    if (int e = mass_piece(0, src0[0], 0, Ly, stream)) return e;
  // Two compute streams inside a sweep (sweep.h): even band stages on the caller's stream, odd ones on the plan's second
  // stream, launch (b, k) gated by one event from (b-1, k-1); everything else -- seam strips, whole-lattice steps, the
  // end of the call -- runs on the caller's stream after the two have joined.  SWALBE_HOST_STREAMS=1: one stream.
  const bool two_streams = env_int("SWALBE_HOST_STREAMS", 2) >= 2;
  cudaStream_t lanes[2] = {stream, two_streams ? plan->s_aux : stream};
  int sweep_s0 = 0;        // first step of the sweep in progress
  auto join = [&]() -> int {
    if (aux_busy) {
      SW_CUDA(cudaEventRecord(plan->ev_join, plan->s_aux));
      SW_CUDA(cudaStreamWaitEvent(stream, plan->ev_join, 0));
      aux_busy = false;
    }
    for (int pl = 0; pl < mass_planes; ++pl)
      if (mass_pending[pl] >= 0) {
        if (int e = mass_final(plan, ml.slot(mass_pending[pl]), stream, pl)) return e;
        mass_pending[pl] = -1;
      }
    return 0;
  };
  for (size_t qi = 0; qi < ops.size(); ++qi) {
    const SweepOp &op = ops[qi];
    if (op.kind == SWEEP_UPLOAD) continue;
    cudaStream_t on = op.stage >= 0 ? lanes[op.stage & 1] : stream;
    if (op.stage < 0)
      if (int e = join()) return e;
    if (op.kind == SWEEP_DOWNLOAD) {
      const size_t off = (size_t)op.jbeg * Lx, cnt = (size_t)(op.jend - op.jbeg) * Lx;
      mark(on, "rows [%d, %d) final", op.jbeg, op.jend);
      SW_CUDA(cudaEventRecord(plan->ev_dn, on));
      SW_CUDA(cudaStreamWaitEvent(plan->s_d2h, plan->ev_dn, 0));
      if (!nocopy) SW_CUDA(cudaMemcpyAsync(hout + off, A[0] + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, plan->s_d2h));
      mark(plan->s_d2h, "download %d done", op.band, 0);
      continue;
    }
    const int kk = op.stage >= 0 ? op.step - sweep_s0 + 1 : 0;  // k-th step of its sweep (set below for stage 0, k = 1)
    if (op.stage == 0 && (qi == 0 || ops[qi - 1].stage != 0 || ops[qi - 1].kind != SWEEP_STEP)) {
      // a sweep starts: the second stream must see everything the caller's stream has done so far
      sweep_s0 = op.step;
      if (two_streams) {
        SW_CUDA(cudaEventRecord(plan->ev_phase, stream));
        SW_CUDA(cudaStreamWaitEvent(plan->s_aux, plan->ev_phase, 0));
      }
    }
    const int k = op.stage >= 0 ? op.step - sweep_s0 + 1 : kk;
    if (op.band >= 0) {
      SW_CUDA(cudaStreamWaitEvent(on, plan->ev_up[op.band], 0));
      if (ml.wants(0))  // (on the stage's own stream, ahead of the launches that will overwrite the band)
        if (int e = mass_piece(0, src0[0], up_beg[op.band], up_end[op.band], on)) return e;
    }
    if (two_streams && op.stage >= 1 && k >= 2) SW_CUDA(cudaStreamWaitEvent(on, plan->ev_wave[(op.stage - 1) & 1][k - 1], 0));
    if (op.jend <= op.jbeg) continue;
    const int s = op.step;
    const bool last = s == nsteps - 1;
    const bool reads_A = ((nsteps - s) % 2) == 0;
    double **src = reads_A ? A : B, **dst = reads_A ? B : A;
    FusedArgs b = a;
    b.h_in = src[0]; b.ux_in = src[1]; b.uy_in = src[2];
    b.h_out = dst[0]; b.ux_out = dst[1]; b.uy_out = dst[2];
    b.f_in = nullptr;
    b.f_out = (!lazy || last) ? st->fout : nullptr;
    b.f_out2 = last ? st->ftemp : nullptr;  // fout == ftemp on return (src/collide.jl:103)
    const bool use_full = last && !skip_aux;
    if (use_full) {
      b.pressure = st->pressure; b.hgx = st->hgradpx; b.hgy = st->hgradpy; b.slipx = st->slipx; b.slipy = st->slipy;
      b.Fx = st->Fx; b.Fy = st->Fy; b.feq = st->feq; b.vsq = st->vsq;
      b.kbtx = thermal ? st->kbtx : nullptr; b.kbty = thermal ? st->kbty : nullptr;
    }
    b.step = step0 + (unsigned long long)s;
    b.log_min = log_mm ? logs->hmin + s : nullptr;
    b.log_max = log_mm ? logs->hmax + s : nullptr;
    b.log_wet = log_wet ? logs->wetted + s : nullptr;
    const int gq = op.seam ? 2 : (op.jend - op.jbeg == Ly ? 0 : 1);
    const LaunchGeom &g = use_full ? *g_full[gq] : *g_mid[gq];
    b.rows_per_cta = g.rows_per_cta; b.W = g.W;
    b.jbeg = op.jbeg; b.jend = op.jend;
    if (int e = launch_fused(g, b, use_full ? key_full : key_mid, on)) return e;
    if (two_streams && op.stage >= 0) {
      SW_CUDA(cudaEventRecord(plan->ev_wave[op.stage & 1][k], on));
      if (on != stream) aux_busy = true;
    }
    if (ml.wants(s + 1))
      if (int e = mass_piece(s + 1, dst[0], op.jbeg, op.jend, on)) return e;
    if (trace && (op.seam == 0) && (s == cfg.k_up - 1 || s == nsteps - 1 || op.jend - op.jbeg == Ly))
      mark(on, "step %d rows from %d done", s, op.jbeg);
  }
  if (int e = join()) return e;
  if (trace) {
    mark(stream, "compute done", 0, 0);
    cudaStreamSynchronize(stream); cudaStreamSynchronize(plan->s_h2d); cudaStreamSynchronize(plan->s_d2h);
    cudaStreamSynchronize(plan->s_aux);
    fprintf(stderr, "[swalbe] host loop %d x %d, %d steps: %d bands, sweeps of %d / %d steps%s\n", Lx, Ly, nsteps, cfg.nbands,
            cfg.k_up, cfg.k_dn, cfg.single ? " (single)" : "");
    for (size_t q = 0; q < marks.size(); ++q) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, marks[0].ev, marks[q].ev);
      fprintf(stderr, "[swalbe]   %8.3f ms  %s\n", ms, marks[q].what);
      if (q) cudaEventDestroy(marks[q].ev);
    }
    cudaEventDestroy(marks[0].ev);
  }
  if (hout) {  // the caller's stream is done when the last band has left
    SW_CUDA(cudaEventRecord(plan->ev_done, plan->s_d2h));
    SW_CUDA(cudaStreamWaitEvent(stream, plan->ev_done, 0));
  }
  return 0;
}

extern "C" {

int swalbe_time_loop_host(swalbe_plan *plan, const swalbe_state *st, const swalbe_params *prm, int nsteps,
                          unsigned long long step0, int flags, const swalbe_loop_logs *logs, const double *height_in_host,
                          double *height_out_host, void *stream_) {
  if (!plan || !st || !prm) return set_error(SWALBE_ERR_ARG, "swalbe_time_loop_host: NULL plan/state/params");
  if (nsteps < 0) return set_error(SWALBE_ERR_ARG, "nsteps < 0");
  return enqueue_steps_host(plan, st, prm, nsteps, step0, flags, logs, height_in_host, height_out_host, (cudaStream_t)stream_);
}

int swalbe_selftest_host_loop_schedule(int Lx, int Ly, int nsteps, int has_in, int has_out, int band_rows, int kmax,
                                       int min_sites, int *ops7, int max_ops, int *nops) {
  if (!nops) return set_error(SWALBE_ERR_ARG, "nops is NULL");
  if (int e = check_extent(Lx, Ly)) return e;
  const SweepConfig cfg = sweep_configure(Lx, Ly, nsteps, has_in != 0, has_out != 0, band_rows, kmax, min_sites);
  const std::vector<SweepOp> ops = sweep_schedule(cfg, Ly, nsteps, has_in != 0, has_out != 0);
  *nops = (int)ops.size();
  if (!ops7) return 0;
  if ((int)ops.size() > max_ops) return set_error(SWALBE_ERR_ARG, "schedule has %d operations, room for %d", (int)ops.size(), max_ops);
  for (size_t q = 0; q < ops.size(); ++q) {
    const int v[7] = {ops[q].kind, ops[q].step, ops[q].jbeg, ops[q].jend, ops[q].band, ops[q].seam, ops[q].stage};
    memcpy(ops7 + 7 * q, v, sizeof(v));
  }
  return 0;
}

}  // extern "C"
