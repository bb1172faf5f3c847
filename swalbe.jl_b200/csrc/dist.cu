// Multi-GPU slab runtime: row-slab decomposition along j, one process per GPU, periodic halo rows exchanged
// with NCCL send/recv on a high-priority communication stream while the interior rows are updated.
// The reference has no multi-GPU path -- this is new.
//
// Slab layout.  Every moment plane (h, ux, uy) of a rank is a (Ly_loc + 2*GH) x Lx array with GH = 3 ghost
// rows below row 0 and above row Ly_loc-1: one fused step has dependency radius 3 in h (p <- h, ∇p <- p,
// streaming <- f*), 1 in u.  For tau != 1 the population planes carry one ghost row on either side.
// A halo "row" is a contiguous run of Lx doubles, so every message is one contiguous chunk.
//
// Per step s (three streams; src/dst are the ping-pong sets, dst(s) == src(s-1)):
//   edge stream  : edge strips [0,GH) and [Ly_loc-GH, Ly_loc)      after halo(s-1) and interior(s-1)
//   comm stream  : NCCL group {send both edge strips, recv both ghost strips} of dst   after edges(s)
//   main stream  : interior rows [GH, Ly_loc-GH)                    after edges(s-1)  (needs no ghost row)
// so the interior kernels run back to back while the thin edge kernels and the exchange of step s overlap
// the interior update of the same step.
// With nranks == 1 the exchange degenerates to two device-to-device copies (self-neighbour), which also lets
// the ghost-row kernel path be tested on a single GPU.
//
// Peer-memory halos (tau == 1, default when the ranks can map each other's memory; SWALBE_DIST_P2P=0: NCCL).  All six
// moment planes of a rank and three 64-bit flags live in ONE allocation whose CUDA IPC handle the ranks all-gather at
// create time.  Per step, instead of the NCCL group, ONE kernel on the communication stream stores the freshly computed
// edge rows straight into the two neighbours' ghost rows over NVLink, fences, and publishes the step's sequence number
// in their flags; the next step's edge kernels are preceded by a one-warp kernel that waits for both of its own flags.
// No receive side, no rendezvous, no proxy thread.  Why no "ready to receive" handshake is needed: a neighbour can only
// compute the edge strips of step s+1 -- and push them into the ghost rows my step-s edge kernels read -- after it has
// seen my push of step s, which my stream issues after those kernels.  A wait that does not come true within 20 s sets a
// sticky error flag and gives up (reported by swalbe_dist_last_loop_ms) instead of hanging the device.
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "fused.cuh"
#include "launch.h"
#include "sweep.h"

namespace swalbe {

constexpr int GH = 3;

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl() {
  if (g_nccl.handle) return 0;
  // RTLD_NOLOAD first: inside a torch process this returns torch's already-loaded libnccl.so.2
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return set_error(SWALBE_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
#define SW_SYM(name)                                                                      \
  *(void **)(&g_nccl.name) = dlsym(h, "nccl" #name);                                      \
  if (!g_nccl.name) return set_error(SWALBE_ERR_NCCL, "libnccl.so.2 lacks symbol nccl" #name)
  SW_SYM(GetUniqueId); SW_SYM(CommInitRank); SW_SYM(CommDestroy); SW_SYM(Send); SW_SYM(Recv);
  SW_SYM(GroupStart); SW_SYM(GroupEnd); SW_SYM(AllGather); SW_SYM(GetErrorString);
#undef SW_SYM
  g_nccl.handle = h;
  return 0;
}

#define SW_NCCL(expr)                                                                                       \
  do {                                                                                                      \
    ncclResult_t _r = (expr);                                                                               \
    if (_r != ncclSuccess)                                                                                  \
      return set_error(SWALBE_ERR_NCCL, "%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(_r), __FILE__, __LINE__); \
  } while (0)

}  // namespace swalbe

using namespace swalbe;

struct swalbe_dist {
  int rank, nranks, Lx, Ly_global, Ly_loc, j_begin;
  swalbe_params prm;
  bool tau1, thermal;
  size_t mplane;          // elements of one moment plane incl. ghosts
  size_t fplane;          // elements of one population plane incl. ghosts
  int gh_f;
  double *arena;          // the six moment planes + the halo flags in one allocation (one IPC handle)
  double *m[2][3];        // ping-pong sets of (h, ux, uy)
  // peer-memory halos
  bool p2p;               // halo rows are stored into the neighbours' ghost rows by k_halo_push
  bool ghost_via_p2p;     // the ghost rows of the current set were (are being) filled by the neighbours' pushes
  void *peer_base[2];     // mapped arena of the rank below [0] / above [1] (the same mapping when they are one rank)
  unsigned long long *flags;        // own: [0] pushes received from below, [1] from above, [2] sticky time-out flag
  unsigned long long seq;           // pushes issued so far (every rank issues one per step)
  unsigned int *push_count;         // last-block detection of k_halo_push
  // swalbe_dist_time_loop_host: copy streams and events, created on first use
  cudaStream_t s_h2d, s_d2h;
  cudaEvent_t ev_up[64], ev_dn, ev_done, ev_seam;
  bool have_host_streams;
  double *f[2];           // population sets (tau == 1: only f[0], no ghosts)
  double *ct;             // cospi(theta) slab with GH ghost rows (NULL: scalar theta)
  double *ct_spare;       // a slab released by swalbe_dist_set_theta(NULL), kept for the next field (no free in the loop)
  bool looped;            // ev_t1 has been recorded (a time loop has run)
  double *ct_alt;         // second buffer of the same shape: target of swalbe_dist_shift_theta (allocated on first use)
  int cur;                // index of the set holding the current moments
  int fcur;               // index of the set holding the current populations (tau != 1)
  ncclComm_t comm;
  cudaStream_t s_comp, s_comm, s_edge;
  cudaEvent_t ev_edges, ev_halo, ev_int, ev_t0, ev_t1, ev_user;
  LaunchGeom g_int, g_edge;
  KernelKey key, key_edge;  // interior / edge-strip kernel flavours
  FusedArgs base;
  float last_ms;
};

static int env_flag(const char *name, int dflt) {
  const char *v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}
static void close_peer_memory(swalbe_dist *d) {
  if (d->peer_base[1] && d->peer_base[1] != d->peer_base[0]) cudaIpcCloseMemHandle(d->peer_base[1]);
  if (d->peer_base[0]) cudaIpcCloseMemHandle(d->peer_base[0]);
  d->peer_base[0] = d->peer_base[1] = nullptr;
  d->p2p = false;
  cudaGetLastError();
}

// dst[j, i] = src[j - sy, i - sx] for the owned rows j in [0, Ly_loc) of a ghosted slab plane (|sy| <= GH: the source rows
// come from the ghost rows, which the last exchange made current); x is periodic inside the slab
__global__ void k_shift_slab(double *__restrict__ dst, const double *__restrict__ src, int sx, int sy, int Lx, int Ly_loc) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)Lx * (size_t)Ly_loc) return;
  const int j = (int)(idx / (size_t)Lx), i = (int)(idx - (size_t)j * Lx);
  int is = i - sx;
  is += is < 0 ? Lx : 0;
  dst[(size_t)(j + GH) * Lx + i] = src[(size_t)(j + GH - sy) * Lx + is];
}

struct PushArgs {
  const double *src[6];  // own rows: [q] top GH rows of plane q (go up), [3 + q] bottom GH rows (go down)
  double *dst[6];        // [q] ghost rows below row 0 of the rank above, [3 + q] ghost rows above the slab of the rank below
  size_t n;              // doubles per part (GH * Lx)
  unsigned long long *flag_up, *flag_down;  // the rank above counts pushes "from below", the rank below "from above"
  unsigned long long seq;
  unsigned int *count;
};

// Stores the six edge-row blocks into the neighbours' memory; the last CTA to finish publishes the sequence number.
__global__ void __launch_bounds__(256) k_halo_push(const PushArgs a) {
  const int part = blockIdx.y;
  const double *__restrict__ src = a.src[part];
  double *__restrict__ dst = a.dst[part];
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < a.n; i += (size_t)gridDim.x * 256) dst[i] = src[i];
  __threadfence_system();  // this thread's rows are visible system-wide before the CTA reports in
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int total = gridDim.x * gridDim.y;
    if (atomicAdd(a.count, 1u) == total - 1u) {
      *a.count = 0u;
      __threadfence_system();
      *(volatile unsigned long long *)a.flag_up = a.seq;
      *(volatile unsigned long long *)a.flag_down = a.seq;
      __threadfence_system();
    }
  }
}

// One warp: wait until both neighbours have published push number `seq` (or the sticky time-out flag is set).
__global__ void k_halo_wait(unsigned long long *flags, unsigned long long seq) {
  if (threadIdx.x != 0) return;
  volatile unsigned long long *f = flags;
  if (f[2]) return;
  unsigned long long t0 = 0, t1 = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (f[0] < seq || f[1] < seq) {
    __nanosleep(200);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 20000000000ull) { f[2] = 1ull; break; }  // 20 s: the neighbour is gone; do not hang the device
  }
  __threadfence_system();
}

static int exchange_rows(swalbe_dist *d, double *plane, int gh, size_t rows_total) {
  // plane has rows [-gh, Ly_loc+gh) stored at physical rows [0, Ly_loc+2gh); send the gh top/bottom owned rows,
  // receive the gh ghost rows on both sides.  up = rank+1 owns larger j, down = rank-1.
  (void)rows_total;
  const size_t n = (size_t)gh * d->Lx;
  double *ghost_lo = plane;                                       // rows [-gh, 0)
  double *own_lo = plane + (size_t)gh * d->Lx;                    // rows [0, gh)
  double *own_hi = plane + (size_t)d->Ly_loc * d->Lx;             // rows [Ly_loc-gh, Ly_loc)
  double *ghost_hi = plane + (size_t)(d->Ly_loc + gh) * d->Lx;    // rows [Ly_loc, Ly_loc+gh)
  if (d->nranks == 1) {
    SW_CUDA(cudaMemcpyAsync(ghost_lo, own_hi, n * sizeof(double), cudaMemcpyDeviceToDevice, d->s_comm));
    SW_CUDA(cudaMemcpyAsync(ghost_hi, own_lo, n * sizeof(double), cudaMemcpyDeviceToDevice, d->s_comm));
    return 0;
  }
  const int up = (d->rank + 1) % d->nranks, down = (d->rank + d->nranks - 1) % d->nranks;
  SW_NCCL(g_nccl.Send(own_hi, n, ncclDouble, up, d->comm, d->s_comm));
  SW_NCCL(g_nccl.Send(own_lo, n, ncclDouble, down, d->comm, d->s_comm));
  SW_NCCL(g_nccl.Recv(ghost_lo, n, ncclDouble, down, d->comm, d->s_comm));
  SW_NCCL(g_nccl.Recv(ghost_hi, n, ncclDouble, up, d->comm, d->s_comm));
  return 0;
}

// halo exchange of moment set `set` (and population set `fset` when tau != 1) on the comm stream
static int exchange_halos(swalbe_dist *d, int set, int fset) {
  if (d->nranks > 1) SW_NCCL(g_nccl.GroupStart());
  for (int q = 0; q < 3; ++q)
    if (int e = exchange_rows(d, d->m[set][q], GH, d->Ly_loc + 2 * GH)) return e;
  if (!d->tau1)
    for (int k = 0; k < 9; ++k)
      if (int e = exchange_rows(d, d->f[fset] + k * d->fplane, 1, d->Ly_loc + 2)) return e;
  if (d->nranks > 1) SW_NCCL(g_nccl.GroupEnd());
  return 0;
}

// the peer-memory exchange of moment set `set`: one kernel, no receive side
static int push_halos(swalbe_dist *d, int set) {
  PushArgs a;
  const size_t row0 = (size_t)GH * d->Lx;                    // first owned row inside a ghosted plane
  const size_t top = (size_t)d->Ly_loc * d->Lx;              // rows [Ly_loc - GH, Ly_loc)
  const size_t ghost_hi = (size_t)(d->Ly_loc + GH) * d->Lx;  // rows [Ly_loc, Ly_loc + GH)
  double *down_arena = (double *)d->peer_base[0], *up_arena = (double *)d->peer_base[1];
  for (int q = 0; q < 3; ++q) {
    const size_t plane = (size_t)(set * 3 + q) * d->mplane;
    a.src[q] = d->m[set][q] + top;      a.dst[q] = up_arena + plane;                 // my top rows -> ghost rows below its row 0
    a.src[3 + q] = d->m[set][q] + row0; a.dst[3 + q] = down_arena + plane + ghost_hi;  // my bottom rows -> ghost rows above its slab
  }
  a.n = (size_t)GH * d->Lx;
  a.flag_up = (unsigned long long *)(up_arena + 6 * d->mplane) + 0;      // "received from below"
  a.flag_down = (unsigned long long *)(down_arena + 6 * d->mplane) + 1;  // "received from above"
  a.seq = ++d->seq;
  a.count = d->push_count;
  const unsigned nb = (unsigned)std::min<size_t>(8, (a.n + 2047) / 2048);
  k_halo_push<<<dim3(nb, 6), 256, 0, d->s_comm>>>(a);
  SW_LAUNCH_CHECK();
  d->ghost_via_p2p = true;
  return 0;
}

extern "C" {

int swalbe_dist_unique_id(void *id128) {
  if (!id128) return set_error(SWALBE_ERR_ARG, "id buffer is NULL");
  if (int e = load_nccl()) return e;
  static_assert(sizeof(ncclUniqueId) == SWALBE_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  SW_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int swalbe_dist_destroy(swalbe_dist *d);

static int dist_allocate(swalbe_dist *d, const void *id128, int rank, int nranks, int Lx, int Ly_loc,
                         const swalbe_params *prm);

int swalbe_dist_create(swalbe_dist **out, const void *id128, int rank, int nranks, int Lx, int Ly_global,
                       const swalbe_params *prm) {
  if (!out || !prm) return set_error(SWALBE_ERR_ARG, "swalbe_dist_create: NULL argument");
  if (int e = check_extent(Lx, Ly_global)) return e;
  if (nranks < 1 || rank < 0 || rank >= nranks) return set_error(SWALBE_ERR_ARG, "bad rank %d / nranks %d", rank, nranks);
  if (Ly_global % nranks) return set_error(SWALBE_ERR_EXTENT, "Ly=%d is not divisible by nranks=%d", Ly_global, nranks);
  const int Ly_loc = Ly_global / nranks;
  if (Ly_loc < 2 * GH) return set_error(SWALBE_ERR_EXTENT, "slab of %d rows is thinner than 2x the halo depth %d", Ly_loc, GH);
  if (nranks > 1 && !id128) return set_error(SWALBE_ERR_ARG, "nranks > 1 needs a NCCL unique id");
  swalbe_dist *d = new swalbe_dist();
  memset(d, 0, sizeof(*d));
  d->rank = rank; d->nranks = nranks; d->Lx = Lx; d->Ly_global = Ly_global; d->Ly_loc = Ly_loc; d->j_begin = rank * Ly_loc;
  d->prm = *prm;
  d->tau1 = prm->tau == 1.0; d->thermal = prm->use_thermal != 0;
  if (prm->cospi_theta_field) { delete d; return set_error(SWALBE_ERR_ARG, "pass theta fields to the slab runtime with swalbe_dist_set_theta"); }
  d->base = FusedArgs{};
  if (int e = fill_consts(d->base, *prm)) { delete d; return e; }
  if (int e = dist_allocate(d, id128, rank, nranks, Lx, Ly_loc, prm)) {
    swalbe_dist_destroy(d);  // releases whatever was created before the failure
    return e;
  }
  d->cur = 0; d->fcur = 0;
  *out = d;
  return 0;
}

// Map the neighbours' arenas: all-gather the IPC handles over the communicator that exists anyway, open the two that
// matter, then all-gather whether that worked -- every rank must take the same transport, or one would wait for flags
// that nobody writes.  Any failure (no peer access, IPC not permitted in this container) leaves the NCCL path in charge.
static int setup_peer_memory(swalbe_dist *d) {
  d->p2p = false;
  if (!d->tau1 || !env_flag("SWALBE_DIST_P2P", 1)) return 0;
  const int n = d->nranks;
  const int up = (d->rank + 1) % n, down = (d->rank + n - 1) % n;
  struct DevBuf {  // (freed on every return path)
    unsigned char *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
  } buf;
  SW_CUDA(cudaMalloc((void **)&buf.p, 64 * (size_t)(n + 1)));
  unsigned char *const dev = buf.p;
  std::vector<unsigned char> host(64 * (size_t)n);
  unsigned char mine[64] = {0};
  cudaIpcMemHandle_t h;
  static_assert(sizeof(cudaIpcMemHandle_t) <= 64, "IPC handle size");
  bool ok = cudaIpcGetMemHandle(&h, d->arena) == cudaSuccess;
  cudaGetLastError();
  memcpy(mine, &h, sizeof(h));
  SW_CUDA(cudaMemcpyAsync(dev, mine, 64, cudaMemcpyHostToDevice, d->s_comm));
  SW_NCCL(g_nccl.AllGather(dev, dev + 64, 64, ncclChar, d->comm, d->s_comm));
  SW_CUDA(cudaMemcpyAsync(host.data(), dev + 64, 64 * (size_t)n, cudaMemcpyDeviceToHost, d->s_comm));
  SW_CUDA(cudaStreamSynchronize(d->s_comm));
  const int peers[2] = {down, up};
  for (int q = 0; q < 2 && ok; ++q) {
    if (q == 1 && up == down) { d->peer_base[1] = d->peer_base[0]; break; }
    cudaIpcMemHandle_t ph;
    memcpy(&ph, host.data() + 64 * (size_t)peers[q], sizeof(ph));
    ok = cudaIpcOpenMemHandle(&d->peer_base[q], ph, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
    if (!ok) { cudaGetLastError(); d->peer_base[q] = nullptr; }
  }
  memset(mine, 0, 64);
  mine[0] = ok ? 1 : 0;
  SW_CUDA(cudaMemcpyAsync(dev, mine, 64, cudaMemcpyHostToDevice, d->s_comm));
  SW_NCCL(g_nccl.AllGather(dev, dev + 64, 64, ncclChar, d->comm, d->s_comm));
  SW_CUDA(cudaMemcpyAsync(host.data(), dev + 64, 64 * (size_t)n, cudaMemcpyDeviceToHost, d->s_comm));
  SW_CUDA(cudaStreamSynchronize(d->s_comm));
  bool all = true;
  for (int r = 0; r < n; ++r) all = all && host[64 * (size_t)r] == 1;
  if (!all) {
    close_peer_memory(d);
    if (env_flag("SWALBE_DEBUG", 0)) fprintf(stderr, "[swalbe] rank %d: peer memory unavailable, halos go through NCCL\n", d->rank);
    return 0;
  }
  d->p2p = true;
  return 0;
}

static int dist_allocate(swalbe_dist *d, const void *id128, int rank, int nranks, int Lx, int Ly_loc,
                         const swalbe_params *prm) {
  d->mplane = (size_t)(Ly_loc + 2 * GH) * Lx;
  d->gh_f = d->tau1 ? 0 : 1;
  d->fplane = (size_t)(Ly_loc + 2 * d->gh_f) * Lx;
  const size_t arena_bytes = 6 * d->mplane * sizeof(double) + 64;
  SW_CUDA(cudaMalloc((void **)&d->arena, arena_bytes));
  SW_CUDA(cudaMemset(d->arena, 0, arena_bytes));
  for (int s = 0; s < 2; ++s)
    for (int q = 0; q < 3; ++q) d->m[s][q] = d->arena + (size_t)(s * 3 + q) * d->mplane;
  d->flags = (unsigned long long *)(d->arena + 6 * d->mplane);
  d->push_count = (unsigned int *)(d->flags + 4);
  for (int s = 0; s < (d->tau1 ? 1 : 2); ++s) {
    SW_CUDA(cudaMalloc((void **)&d->f[s], 9 * d->fplane * sizeof(double)));
    SW_CUDA(cudaMemset(d->f[s], 0, 9 * d->fplane * sizeof(double)));
  }
  int lo = 0, hi = 0;
  SW_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  SW_CUDA(cudaStreamCreateWithPriority(&d->s_comp, cudaStreamNonBlocking, lo));
  SW_CUDA(cudaStreamCreateWithPriority(&d->s_comm, cudaStreamNonBlocking, hi));
  SW_CUDA(cudaStreamCreateWithPriority(&d->s_edge, cudaStreamNonBlocking, hi));
  SW_CUDA(cudaEventCreateWithFlags(&d->ev_int, cudaEventDisableTiming));
  SW_CUDA(cudaEventCreateWithFlags(&d->ev_edges, cudaEventDisableTiming));
  SW_CUDA(cudaEventCreateWithFlags(&d->ev_halo, cudaEventDisableTiming));
  SW_CUDA(cudaEventCreateWithFlags(&d->ev_user, cudaEventDisableTiming));
  SW_CUDA(cudaEventCreate(&d->ev_t0));
  SW_CUDA(cudaEventCreate(&d->ev_t1));
  if (nranks > 1) {
    if (int e = load_nccl()) return e;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    SW_NCCL(g_nccl.CommInitRank(&d->comm, nranks, id, rank));
    if (int e = setup_peer_memory(d)) return e;
  }
  d->key = make_key(*prm, d->base.pc.pmode, true);
  d->key_edge = d->key;
  // interior kernel: bulk-copy row prefetch where it pays (slab planes are cudaMalloc'ed, ghost offset = GH*Lx*8 bytes)
  d->key.bulk = d->key.lean_pm > 0 && !d->key.opts && !d->key.thermal && bulk_eligible(Lx, (size_t)Lx * Ly_loc);
  if (int e = choose_geometry(Lx, Ly_loc - 2 * GH > 0 ? Ly_loc - 2 * GH : Ly_loc, d->key, &d->g_int)) return e;
  if (int e = choose_geometry(Lx, GH, d->key_edge, &d->g_edge)) return e;
  return 0;
}

int swalbe_dist_destroy(swalbe_dist *d) {
  if (!d) return 0;
  // the only blocking entry point besides create: the handle's own streams are drained before its memory goes away
  if (d->s_comp) cudaStreamSynchronize(d->s_comp);
  if (d->s_edge) cudaStreamSynchronize(d->s_edge);
  if (d->s_comm) cudaStreamSynchronize(d->s_comm);
  if (d->have_host_streams) {
    cudaStreamSynchronize(d->s_h2d); cudaStreamSynchronize(d->s_d2h);
    cudaStreamDestroy(d->s_h2d); cudaStreamDestroy(d->s_d2h);
    for (cudaEvent_t ev : d->ev_up) cudaEventDestroy(ev);
    cudaEventDestroy(d->ev_dn); cudaEventDestroy(d->ev_done); cudaEventDestroy(d->ev_seam);
  }
  close_peer_memory(d);
  if (d->comm) g_nccl.CommDestroy(d->comm);
  cudaFree(d->arena);
  for (int s = 0; s < 2; ++s) cudaFree(d->f[s]);
  cudaFree(d->ct);
  cudaFree(d->ct_alt);
  cudaFree(d->ct_spare);
  if (d->s_comp) cudaStreamDestroy(d->s_comp);
  if (d->s_comm) cudaStreamDestroy(d->s_comm);
  if (d->s_edge) cudaStreamDestroy(d->s_edge);
  cudaEvent_t evs[] = {d->ev_int, d->ev_edges, d->ev_halo, d->ev_user, d->ev_t0, d->ev_t1};
  for (cudaEvent_t ev : evs)
    if (ev) cudaEventDestroy(ev);
  delete d;
  return 0;
}

int swalbe_dist_local_rows(const swalbe_dist *d, int *j_begin, int *j_count) {
  if (!d) return set_error(SWALBE_ERR_ARG, "dist is NULL");
  if (j_begin) *j_begin = d->j_begin;
  if (j_count) *j_count = d->Ly_loc;
  return 0;
}

int swalbe_dist_set_state(swalbe_dist *d, const double *height, const double *velx, const double *vely,
                          const double *ftemp, void *stream_) {
  if (!d || !height || !velx || !vely) return set_error(SWALBE_ERR_ARG, "swalbe_dist_set_state: NULL argument");
  if (!d->tau1 && !ftemp) return set_error(SWALBE_ERR_ARG, "tau != 1 needs the ftemp populations of the slab");
  cudaStream_t user = (cudaStream_t)stream_;
  const size_t n = (size_t)d->Ly_loc * d->Lx;
  const double *src[3] = {height, velx, vely};
  d->cur = 0; d->fcur = 0;
  d->ghost_via_p2p = false;  // the exchange below is a NCCL group whose completion is a local event
  for (int q = 0; q < 3; ++q)
    SW_CUDA(cudaMemcpyAsync(d->m[0][q] + (size_t)GH * d->Lx, src[q], n * sizeof(double), cudaMemcpyDeviceToDevice, user));
  if (ftemp)
    for (int k = 0; k < 9; ++k)
      SW_CUDA(cudaMemcpyAsync(d->f[0] + k * d->fplane + (size_t)d->gh_f * d->Lx, ftemp + k * n, n * sizeof(double),
                              cudaMemcpyDeviceToDevice, user));
  SW_CUDA(cudaEventRecord(d->ev_user, user));
  SW_CUDA(cudaStreamWaitEvent(d->s_comm, d->ev_user, 0));
  if (int e = exchange_halos(d, 0, 0)) return e;
  SW_CUDA(cudaEventRecord(d->ev_halo, d->s_comm));
  SW_CUDA(cudaStreamWaitEvent(d->s_comp, d->ev_halo, 0));
  SW_CUDA(cudaStreamWaitEvent(d->s_edge, d->ev_halo, 0));
  SW_CUDA(cudaStreamWaitEvent(user, d->ev_halo, 0));
  // "previous step" events of the first step: everything is ready once the halo exchange above is done
  SW_CUDA(cudaEventRecord(d->ev_int, d->s_comp));
  SW_CUDA(cudaEventRecord(d->ev_edges, d->s_edge));
  return 0;
}

static void slab_args(const swalbe_dist *d, int src, int dst, unsigned long long step, FusedArgs &a) {
  a = d->base;
  a.Lx = d->Lx; a.Ly = d->Ly_loc; a.wrap_y = 0;  // ghost rows: pointers are handed over at logical row 0
  a.jglobal0 = d->j_begin; a.Ly_global = d->Ly_global;
  const size_t mo = (size_t)GH * d->Lx;
  a.h_in = d->m[src][0] + mo; a.ux_in = d->m[src][1] + mo; a.uy_in = d->m[src][2] + mo;
  a.h_out = d->m[dst][0] + mo; a.ux_out = d->m[dst][1] + mo; a.uy_out = d->m[dst][2] + mo;
  a.f_in = nullptr; a.f_out = d->f[0];
  a.fstride_in = a.fstride_out = a.fstride_out2 = d->fplane;
  a.ct_field = d->ct ? d->ct + mo : nullptr;
  a.step = step;
}

// nsteps steps on the handle's own streams (edge strips, exchange and interior overlapped); the caller has ordered the
// streams after its input and orders its output after ev_halo / ev_edges / ev_int
static int dist_steps(swalbe_dist *d, int nsteps, unsigned long long step0) {
  const int Ly = d->Ly_loc;
  for (int s = 0; s < nsteps; ++s) {
    const int src = d->cur, dst = d->cur ^ 1;
    FusedArgs a = d->base;
    a.Lx = d->Lx; a.Ly = Ly; a.wrap_y = 0;  // ghost rows: pointers are handed over at logical row 0
    a.jglobal0 = d->j_begin; a.Ly_global = d->Ly_global;
    const size_t mo = (size_t)GH * d->Lx, fo = (size_t)d->gh_f * d->Lx;
    a.h_in = d->m[src][0] + mo; a.ux_in = d->m[src][1] + mo; a.uy_in = d->m[src][2] + mo;
    a.h_out = d->m[dst][0] + mo; a.ux_out = d->m[dst][1] + mo; a.uy_out = d->m[dst][2] + mo;
    int fdst = 0;
    if (d->tau1) { a.f_in = nullptr; a.f_out = d->f[0]; }
    else { fdst = d->fcur ^ 1; a.f_in = d->f[d->fcur] + fo; a.f_out = d->f[fdst] + fo; }
    a.fstride_in = a.fstride_out = a.fstride_out2 = d->fplane;
    a.ct_field = d->ct ? d->ct + mo : nullptr;
    a.step = step0 + (unsigned long long)s;
    // edge strips: need the ghost rows of `src` (previous exchange) and the rows the previous interior kernel wrote
    SW_CUDA(cudaStreamWaitEvent(d->s_edge, d->ev_halo, 0));
    if (d->p2p && d->ghost_via_p2p) {  // ... which the neighbours' push number `seq` filled: wait for both flags
      k_halo_wait<<<1, 32, 0, d->s_edge>>>(d->flags, d->seq);
      SW_LAUNCH_CHECK();
    }
    SW_CUDA(cudaStreamWaitEvent(d->s_edge, d->ev_int, 0));
    // interior rows: need the edge rows of `src` written by the previous step's edge kernels, no ghost row
    SW_CUDA(cudaStreamWaitEvent(d->s_comp, d->ev_edges, 0));
    a.W = d->g_edge.W; a.rows_per_cta = d->g_edge.rows_per_cta;
    a.jbeg = 0; a.jend = GH;
    if (int e = launch_fused(d->g_edge, a, d->key_edge, d->s_edge)) return e;
    a.jbeg = Ly - GH; a.jend = Ly;
    if (int e = launch_fused(d->g_edge, a, d->key_edge, d->s_edge)) return e;
    SW_CUDA(cudaEventRecord(d->ev_edges, d->s_edge));
    // halo exchange of the freshly written edge rows of `dst`
    SW_CUDA(cudaStreamWaitEvent(d->s_comm, d->ev_edges, 0));
    if (d->p2p) {
      if (int e = push_halos(d, dst)) return e;
    } else if (int e = exchange_halos(d, dst, fdst)) return e;
    SW_CUDA(cudaEventRecord(d->ev_halo, d->s_comm));
    if (Ly > 2 * GH) {
      a.W = d->g_int.W; a.rows_per_cta = d->g_int.rows_per_cta;
      a.jbeg = GH; a.jend = Ly - GH;
      if (int e = launch_fused(d->g_int, a, d->key, d->s_comp)) return e;
    }
    SW_CUDA(cudaEventRecord(d->ev_int, d->s_comp));
    d->cur = dst;
    if (!d->tau1) d->fcur = fdst;
  }
  return 0;
}

int swalbe_dist_time_loop(swalbe_dist *d, int nsteps, unsigned long long step0, void *stream_) {
  if (!d) return set_error(SWALBE_ERR_ARG, "dist is NULL");
  if (nsteps <= 0) return 0;
  cudaStream_t user = (cudaStream_t)stream_;
  SW_CUDA(cudaEventRecord(d->ev_user, user));
  SW_CUDA(cudaStreamWaitEvent(d->s_comp, d->ev_user, 0));
  SW_CUDA(cudaStreamWaitEvent(d->s_edge, d->ev_user, 0));
  SW_CUDA(cudaEventRecord(d->ev_t0, d->s_comp));
  if (int e = dist_steps(d, nsteps, step0)) return e;
  SW_CUDA(cudaStreamWaitEvent(d->s_comp, d->ev_halo, 0));
  SW_CUDA(cudaStreamWaitEvent(d->s_comp, d->ev_edges, 0));
  SW_CUDA(cudaEventRecord(d->ev_t1, d->s_comp));
  SW_CUDA(cudaStreamWaitEvent(user, d->ev_t1, 0));
  d->looped = true;
  return 0;
}

// ---- the slab's time loop from / to host memory --------------------------------------------------------------------
// Same sweeps as swalbe_time_loop_host (sweep.h) on the rows of one slab: the band stages never touch a ghost row (band 0
// shrinks away from the lower edge by 3 rows per step, the last band from the upper edge), and what is a periodic seam
// on one GPU is here the slab boundary: after the last band, step k of the strips [0, 3k) and [Ly_loc - 3k, Ly_loc) is
// computed from the neighbours' rows of step k-1, exchanged after every strip pair like in the ordinary loop.
static int dist_host_streams(swalbe_dist *d) {
  if (d->have_host_streams) return 0;
  SW_CUDA(cudaStreamCreateWithFlags(&d->s_h2d, cudaStreamNonBlocking));
  SW_CUDA(cudaStreamCreateWithFlags(&d->s_d2h, cudaStreamNonBlocking));
  for (cudaEvent_t &ev : d->ev_up) SW_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  SW_CUDA(cudaEventCreateWithFlags(&d->ev_dn, cudaEventDisableTiming));
  SW_CUDA(cudaEventCreateWithFlags(&d->ev_done, cudaEventDisableTiming));
  SW_CUDA(cudaEventCreateWithFlags(&d->ev_seam, cudaEventDisableTiming));
  d->have_host_streams = true;
  return 0;
}

// exchange of the ghost rows of moment set `set` after the work queued on s_comp so far; `rendezvous`: through NCCL even
// when peer memory is on (the first exchange of a call: a neighbour may still be inside its previous loop)
static int seam_exchange(swalbe_dist *d, int set, bool rendezvous) {
  SW_CUDA(cudaEventRecord(d->ev_seam, d->s_comp));
  SW_CUDA(cudaStreamWaitEvent(d->s_comm, d->ev_seam, 0));
  if (d->p2p && !rendezvous) {
    if (int e = push_halos(d, set)) return e;
  } else {
    if (int e = exchange_halos(d, set, d->fcur)) return e;
    d->ghost_via_p2p = false;
  }
  SW_CUDA(cudaEventRecord(d->ev_halo, d->s_comm));
  return 0;
}
static int seam_wait(swalbe_dist *d) {  // the ghost rows of the last exchange are there (s_comp)
  SW_CUDA(cudaStreamWaitEvent(d->s_comp, d->ev_halo, 0));
  if (d->p2p && d->ghost_via_p2p) {
    k_halo_wait<<<1, 32, 0, d->s_comp>>>(d->flags, d->seq);
    SW_LAUNCH_CHECK();
  }
  return 0;
}

int swalbe_dist_time_loop_host(swalbe_dist *d, int nsteps, unsigned long long step0, const double *h_in_host,
                               const double *velx, const double *vely, double *h_out_host, void *stream_) {
  if (!d) return set_error(SWALBE_ERR_ARG, "dist is NULL");
  if (nsteps < 0) return set_error(SWALBE_ERR_ARG, "nsteps < 0");
  if (!d->tau1) return set_error(SWALBE_ERR_ARG, "swalbe_dist_time_loop_host: tau == 1 only (the populations of a tau != 1 state "
                                                  "do not travel with the height; use set_state / time_loop / get_state)");
  if (int e = dist_host_streams(d)) return e;
  cudaStream_t user = (cudaStream_t)stream_;
  const int Lx = d->Lx, Ly = d->Ly_loc;
  const size_t mo = (size_t)GH * Lx, nown = (size_t)Ly * Lx;
  SW_CUDA(cudaEventRecord(d->ev_user, user));
  for (cudaStream_t st : {d->s_comp, d->s_edge, d->s_comm, d->s_h2d, d->s_d2h}) SW_CUDA(cudaStreamWaitEvent(st, d->ev_user, 0));
  SW_CUDA(cudaEventRecord(d->ev_t0, d->s_comp));
  const bool hin = h_in_host != nullptr, hout = h_out_host != nullptr;
  auto envi = [](const char *n) { return env_flag(n, 0); };
  SweepConfig cfg = {0, 0, 0, 0};
  if (env_flag("SWALBE_HOST_STREAM", 1))
    cfg = sweep_configure(Lx, Ly, nsteps, hin, hout, envi("SWALBE_BAND_ROWS"), envi("SWALBE_HOST_KMAX"), envi("SWALBE_HOST_MIN_SITES"));
  const std::vector<SweepOp> ops = sweep_schedule(cfg, Ly, nsteps, hin, hout);
  if (hin) {  // the state is replaced: height from the host (band by band), velocities from the caller's slabs or zero
    d->cur = 0; d->fcur = 0;
    const double *vel[2] = {velx, vely};
    for (int q = 0; q < 2; ++q) {
      if (vel[q]) SW_CUDA(cudaMemcpyAsync(d->m[0][q + 1] + mo, vel[q], nown * sizeof(double), cudaMemcpyDeviceToDevice, d->s_comp));
      else SW_CUDA(cudaMemsetAsync(d->m[0][q + 1], 0, d->mplane * sizeof(double), d->s_comp));
    }
    for (const SweepOp &op : ops)
      if (op.kind == SWEEP_UPLOAD) {
        const size_t off = (size_t)op.jbeg * Lx, cnt = (size_t)(op.jend - op.jbeg) * Lx;
        SW_CUDA(cudaMemcpyAsync(d->m[0][0] + mo + off, h_in_host + off, cnt * sizeof(double), cudaMemcpyHostToDevice, d->s_h2d));
        SW_CUDA(cudaEventRecord(d->ev_up[op.band], d->s_h2d));
      }
  }
  const int src0 = d->cur;
  bool ghosts_current = !hin;  // (the runtime keeps the ghost rows of its current set current between calls)
  LaunchGeom g_band = d->g_int, g_seam = d->g_edge;
  if (cfg.nbands > 0) {
    if (int e = choose_geometry(Lx, Ly / cfg.nbands, d->key, &g_band)) return e;
    if (int e = choose_geometry(Lx, std::max(GH, 3 * std::max(cfg.k_up, cfg.k_dn) / 2 + 1), d->key_edge, &g_seam)) return e;
  }
  // the streams of the ordinary loop hang on three events; keep them meaningful around every piece issued here
  auto publish = [&]() -> int {
    SW_CUDA(cudaEventRecord(d->ev_int, d->s_comp));
    SW_CUDA(cudaEventRecord(d->ev_edges, d->s_comp));
    return 0;
  };
  int seam_launches = 0;
  for (size_t qi = 0; qi < ops.size(); ++qi) {
    const SweepOp &op = ops[qi];
    if (op.kind == SWEEP_UPLOAD) continue;
    if (op.kind == SWEEP_DOWNLOAD) {
      const size_t off = (size_t)op.jbeg * Lx, cnt = (size_t)(op.jend - op.jbeg) * Lx;
      const int fin = src0 ^ (nsteps & 1);
      if (nsteps == 0 && hin)
        for (const SweepOp &u : ops)
          if (u.kind == SWEEP_UPLOAD) SW_CUDA(cudaStreamWaitEvent(d->s_comp, d->ev_up[u.band], 0));
      SW_CUDA(cudaEventRecord(d->ev_dn, d->s_comp));
      SW_CUDA(cudaStreamWaitEvent(d->s_d2h, d->ev_dn, 0));
      SW_CUDA(cudaMemcpyAsync(h_out_host + off, d->m[fin][0] + mo + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, d->s_d2h));
      continue;
    }
    if (op.band >= 0) SW_CUDA(cudaStreamWaitEvent(d->s_comp, d->ev_up[op.band], 0));
    const int s = op.step;
    if (op.stage < 0 && !op.seam) {  // a run of whole-slab steps: the ordinary loop (edge strips, exchange, interior)
      int n = 1;
      while (qi + n < ops.size() && ops[qi + n].kind == SWEEP_STEP && ops[qi + n].stage < 0 && !ops[qi + n].seam) ++n;
      if (!ghosts_current) {
        if (int e = seam_exchange(d, src0 ^ (s & 1), true)) return e;
        ghosts_current = true;
      }
      if (int e = publish()) return e;
      d->cur = src0 ^ (s & 1);
      if (int e = dist_steps(d, n, step0 + (unsigned long long)s)) return e;
      SW_CUDA(cudaStreamWaitEvent(d->s_comp, d->ev_edges, 0));  // the pieces below are issued on s_comp alone
      qi += (size_t)n - 1;
      continue;
    }
    const int src = src0 ^ (s & 1), dst = src ^ 1;
    FusedArgs a;
    slab_args(d, src, dst, step0 + (unsigned long long)s, a);
    if (op.seam) {
      if (!ghosts_current) {  // first strip of an upload sweep: the input state's ghost rows
        if (int e = seam_exchange(d, src, true)) return e;
        ghosts_current = true;
      }
      if ((seam_launches & 1) == 0)
        if (int e = seam_wait(d)) return e;
      a.W = g_seam.W; a.rows_per_cta = g_seam.rows_per_cta; a.jbeg = op.jbeg; a.jend = op.jend;
      if (int e = launch_fused(g_seam, a, d->key_edge, d->s_comp)) return e;
      if ((++seam_launches & 1) == 0)  // both strips of this step are queued: their rows travel
        if (int e = seam_exchange(d, dst, false)) return e;
      continue;
    }
    if (op.jend <= op.jbeg) continue;
    a.W = g_band.W; a.rows_per_cta = g_band.rows_per_cta; a.jbeg = op.jbeg; a.jend = op.jend;
    if (int e = launch_fused(g_band, a, d->key, d->s_comp)) return e;
  }
  d->cur = src0 ^ (nsteps & 1);
  if (!ghosts_current) {  // (nsteps == 0: the uploaded state still owes its ghost rows)
    for (const SweepOp &op : ops)
      if (op.kind == SWEEP_UPLOAD) SW_CUDA(cudaStreamWaitEvent(d->s_comp, d->ev_up[op.band], 0));
    if (int e = seam_exchange(d, d->cur, true)) return e;
  }
  SW_CUDA(cudaStreamWaitEvent(d->s_comp, d->ev_halo, 0));
  if (int e = publish()) return e;
  SW_CUDA(cudaEventRecord(d->ev_t1, d->s_comp));
  SW_CUDA(cudaStreamWaitEvent(user, d->ev_t1, 0));
  if (hout) {
    SW_CUDA(cudaEventRecord(d->ev_done, d->s_d2h));
    SW_CUDA(cudaStreamWaitEvent(user, d->ev_done, 0));
  }
  d->looped = true;
  return 0;
}

int swalbe_dist_get_state(swalbe_dist *d, double *height, double *velx, double *vely, double *fout, void *stream_) {
  if (!d) return set_error(SWALBE_ERR_ARG, "dist is NULL");
  cudaStream_t user = (cudaStream_t)stream_;
  const size_t n = (size_t)d->Ly_loc * d->Lx;
  double *dstp[3] = {height, velx, vely};
  for (int q = 0; q < 3; ++q)
    if (dstp[q])
      SW_CUDA(cudaMemcpyAsync(dstp[q], d->m[d->cur][q] + (size_t)GH * d->Lx, n * sizeof(double), cudaMemcpyDeviceToDevice, user));
  if (fout) {
    const double *fs = d->tau1 ? d->f[0] : d->f[d->fcur];
    for (int k = 0; k < 9; ++k)
      SW_CUDA(cudaMemcpyAsync(fout + k * n, fs + k * d->fplane + (size_t)d->gh_f * d->Lx, n * sizeof(double),
                              cudaMemcpyDeviceToDevice, user));
  }
  return 0;
}

int swalbe_dist_set_theta(swalbe_dist *d, const double *ct_slab, void *stream_) {
  if (!d) return set_error(SWALBE_ERR_ARG, "dist is NULL");
  cudaStream_t user = (cudaStream_t)stream_;
  // No host synchronisation: launches already enqueued keep the kernel flavour and pointers they were given; the copy
  // below is ordered after the last time loop by its end event (the loop's kernels read the slab this call overwrites).
  if (d->looped) SW_CUDA(cudaStreamWaitEvent(user, d->ev_t1, 0));
  if (!ct_slab) {
    if (d->ct) d->ct_spare = d->ct;  // (never both set: a spare slab is taken back before a new one is allocated)
    d->ct = nullptr;
    d->prm.cospi_theta_field = nullptr;
  } else {
    if (!d->ct && d->ct_spare) { d->ct = d->ct_spare; d->ct_spare = nullptr; }
    if (!d->ct) SW_CUDA(cudaMalloc((void **)&d->ct, d->mplane * sizeof(double)));
    const size_t n = (size_t)d->Ly_loc * d->Lx;
    SW_CUDA(cudaMemcpyAsync(d->ct + (size_t)GH * d->Lx, ct_slab, n * sizeof(double), cudaMemcpyDeviceToDevice, user));
    SW_CUDA(cudaEventRecord(d->ev_user, user));
    SW_CUDA(cudaStreamWaitEvent(d->s_comm, d->ev_user, 0));
    if (d->nranks > 1) SW_NCCL(g_nccl.GroupStart());
    if (int e = exchange_rows(d, d->ct, GH, d->Ly_loc + 2 * GH)) return e;
    if (d->nranks > 1) SW_NCCL(g_nccl.GroupEnd());
    SW_CUDA(cudaEventRecord(d->ev_halo, d->s_comm));
    SW_CUDA(cudaStreamWaitEvent(d->s_comp, d->ev_halo, 0));
    SW_CUDA(cudaStreamWaitEvent(d->s_edge, d->ev_halo, 0));
    SW_CUDA(cudaStreamWaitEvent(user, d->ev_halo, 0));
    d->prm.cospi_theta_field = d->ct;  // only its NULL-ness matters for the kernel flavour
  }
  d->key = make_key(d->prm, d->base.pc.pmode, true);
  d->key_edge = d->key;
  const int Ly_loc = d->Ly_loc;
  d->key.bulk = d->key.lean_pm > 0 && !d->key.opts && !d->key.thermal && bulk_eligible(d->Lx, (size_t)d->Lx * Ly_loc);
  if (int e = choose_geometry(d->Lx, Ly_loc - 2 * GH > 0 ? Ly_loc - 2 * GH : Ly_loc, d->key, &d->g_int)) return e;
  if (int e = choose_geometry(d->Lx, GH, d->key_edge, &d->g_edge)) return e;
  return 0;
}

int swalbe_dist_shift_theta(swalbe_dist *d, int sx, int sy, void *stream_) {
  if (!d) return set_error(SWALBE_ERR_ARG, "dist is NULL");
  if (!d->ct) return set_error(SWALBE_ERR_ARG, "swalbe_dist_shift_theta: no theta field is set (swalbe_dist_set_theta)");
  if (sy < -GH || sy > GH) return set_error(SWALBE_ERR_ARG, "swalbe_dist_shift_theta: |sy| = %d exceeds the ghost depth %d", sy, GH);
  cudaStream_t user = (cudaStream_t)stream_;
  if (!d->ct_alt) {
    SW_CUDA(cudaMalloc((void **)&d->ct_alt, d->mplane * sizeof(double)));
    SW_CUDA(cudaMemset(d->ct_alt, 0, d->mplane * sizeof(double)));
  }
  sx %= d->Lx;
  if (sx < 0) sx += d->Lx;
  // ordered on the caller's stream: after the previous time loop (its end event was awaited on this stream) and after
  // the ghost exchange of the last set/shift
  const size_t n = (size_t)d->Lx * d->Ly_loc;
  k_shift_slab<<<(unsigned)((n + 255) / 256), 256, 0, user>>>(d->ct_alt, d->ct, sx, sy, d->Lx, d->Ly_loc);
  SW_LAUNCH_CHECK();
  double *t = d->ct; d->ct = d->ct_alt; d->ct_alt = t;
  d->prm.cospi_theta_field = d->ct;
  SW_CUDA(cudaEventRecord(d->ev_user, user));
  SW_CUDA(cudaStreamWaitEvent(d->s_comm, d->ev_user, 0));
  if (d->nranks > 1) SW_NCCL(g_nccl.GroupStart());
  if (int e = exchange_rows(d, d->ct, GH, d->Ly_loc + 2 * GH)) return e;
  if (d->nranks > 1) SW_NCCL(g_nccl.GroupEnd());
  SW_CUDA(cudaEventRecord(d->ev_halo, d->s_comm));
  SW_CUDA(cudaStreamWaitEvent(d->s_comp, d->ev_halo, 0));
  SW_CUDA(cudaStreamWaitEvent(d->s_edge, d->ev_halo, 0));
  SW_CUDA(cudaStreamWaitEvent(user, d->ev_halo, 0));
  return 0;
}

int swalbe_dist_height_stats(swalbe_dist *d, double *out4, double thresh, void *stream_) {
  if (!d || !out4) return set_error(SWALBE_ERR_ARG, "NULL argument");
  // the owned rows of a ghosted plane are one contiguous Lx * Ly_loc block
  return swalbe_field_stats(out4, d->m[d->cur][0] + (size_t)GH * d->Lx, thresh, d->Lx, d->Ly_loc, stream_);
}

int swalbe_dist_uses_peer_memory(const swalbe_dist *d, int *yes) {
  if (!d || !yes) return set_error(SWALBE_ERR_ARG, "NULL argument");
  *yes = d->p2p ? 1 : 0;
  return 0;
}

int swalbe_dist_last_loop_ms(swalbe_dist *d, float *ms) {
  if (!d || !ms) return set_error(SWALBE_ERR_ARG, "NULL argument");
  SW_CUDA(cudaEventSynchronize(d->ev_t1));
  if (d->p2p) {
    unsigned long long err = 0;
    SW_CUDA(cudaMemcpy(&err, d->flags + 2, sizeof(err), cudaMemcpyDeviceToHost));
    if (err) return set_error(SWALBE_ERR_NCCL, "peer-memory halo exchange: a neighbour's rows did not arrive within 20 s");
  }
  SW_CUDA(cudaEventElapsedTime(ms, d->ev_t0, d->ev_t1));
  d->last_ms = *ms;
  return 0;
}

}  // extern "C"
