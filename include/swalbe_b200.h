/* libswalbe_b200 -- C ABI of the B200-native (sm_100a) thin-film D2Q9 lattice-Boltzmann step.
 *
 * This header is the drop-in boundary for ONE path of Swalbe.jl: the 2-D time step
 *   filmpressure! -> h∇p! -> slippage! -> force sum -> equilibrium! -> BGKandStream! -> moments!
 * (src/simulate.jl:15-22 of the reference).  Swalbe itself is pure Julia and has no FFI; each entry
 * point below is what a `ccall((:sym, "libswalbe_b200"), Cint, (...), ...)` method on Swalbe's GPU
 * state types (CuState / CuState_thermal, src/initialize.jl:214-256) would bind.  The reference
 * function every symbol replaces is cited as file:line into the reference tree; INTEGRATION.md shows
 * the Julia-side glue.
 *
 * Conventions
 *  - All field pointers are DEVICE pointers to Float64 arrays in Julia's native column-major layout,
 *    un-padded: A[i,j,k] (0-based) at i + Lx*(j + Ly*k); populations are nine contiguous Lx*Ly planes
 *    (SoA), plane k = population k with the D2Q9 velocity set c0=(0,0) c1=(1,0) c2=(0,1) c3=(-1,0)
 *    c4=(0,-1) c5=(1,1) c6=(-1,1) c7=(-1,-1) c8=(1,-1)   (src/collide.jl:92-100, :270-282).
 *  - Zero-copy: the library never frees or retains caller pointers beyond a call.
 *  - Every call is asynchronous on the caller's `stream` (a cudaStream_t passed as void*; NULL = the
 *    legacy default stream), so it is ordered with the caller's own device work.  No hidden
 *    cudaDeviceSynchronize.  The only host-blocking entry points are the ones that own memory:
 *    swalbe_plan_create / swalbe_dist_create (cudaMalloc, NCCL communicator set-up, IPC mapping of the
 *    neighbours' memory), their destroy counterparts (swalbe_dist_destroy waits for the handle's own
 *    streams before freeing), swalbe_dist_last_loop_ms (waits for the loop's end event, by definition)
 *    and, once per handle, the FIRST call that needs library-owned resources created lazily: the copy
 *    streams of swalbe_time_loop_host / swalbe_dist_time_loop_host and the row-sum buffer of the mass log
 *    (swalbe_loop_logs.hsum).
 *  - Every function returns 0 on success or a swalbe_status code; swalbe_last_error() gives the
 *    message (thread-local).  Nothing throws across the ABI.
 *  - Arithmetic is IEEE-754 double with NO fused multiply-add contraction and the reference's exact
 *    evaluation order, so results equal the Julia CPU path operation for
 *    operation.  cospi(theta) is evaluated by the caller and passed in as data.
 *  - There is no CPU fallback: without a CUDA device every compute entry point returns
 *    SWALBE_ERR_CUDA.
 */
#ifndef SWALBE_B200_H
#define SWALBE_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWALBE_B200_VERSION 100 /* 0.1.0 */

typedef enum swalbe_status {
  SWALBE_OK = 0,
  SWALBE_ERR_DOMAIN = 1, /* unsupported (n,m) for the array-form pressure == Julia DomainError, src/pressure.jl:101-107 */
  SWALBE_ERR_EXTENT = 2, /* Lx or Ly < 1, or too large */
  SWALBE_ERR_CUDA = 3,   /* CUDA runtime error (message holds cudaGetErrorString) */
  SWALBE_ERR_NCCL = 4,   /* NCCL error / libnccl.so.2 not loadable */
  SWALBE_ERR_ARG = 5     /* NULL pointer or inconsistent arguments */
} swalbe_status;

int swalbe_version(void);
const char *swalbe_last_error(void);
/* number of kernels launched by this library in this process (for bench.py's gpu_launches claim) */
unsigned long long swalbe_launch_count(void);

/* pressure_variant */
#define SWALBE_PRESSURE_POWER_BROAD 0 /* state form, src/pressure.jl:119-155 (any n,m) */
#define SWALBE_PRESSURE_FAST 1        /* array form, src/pressure.jl:72-115 ((9,3) or (3,2) only) */
/* slip_variant */
#define SWALBE_SLIP_STANDARD 0 /* slippage!          src/forcing.jl:42-46  */
#define SWALBE_SLIP_HCRIT 1    /* slippage2!         src/forcing.jl:85-99  */
#define SWALBE_SLIP_RING_RIV 2 /* slippage_ring_riv! src/forcing.jl:107-111 */

/* ---------------------------------------------------------------------------------------------
 * Per-operator entry points (array forms; the state forms of the reference only unpack fields).
 * ------------------------------------------------------------------------------------------- */

/* equilibrium!(feq, height, velx, vely, vsq, g)              src/equilibrium.jl:63-116 */
int swalbe_equilibrium_d2q9(double *feq, const double *height, const double *velx, const double *vely, double *vsq,
                            double g, int Lx, int Ly, void *stream);

/* BGKandStream!(fout, feq, ftemp, Fx, Fy, tau)               src/collide.jl:70-105
 * On return fout == ftemp == streamed post-collision populations (src/collide.jl:103). */
int swalbe_bgk_stream_d2q9(double *fout, const double *feq, double *ftemp, const double *Fx, const double *Fy,
                           double tau, int Lx, int Ly, void *stream);

/* moments!(height, velx, vely, fout)                         src/moments.jl:43-52 */
int swalbe_moments_d2q9(double *height, double *velx, double *vely, const double *fout, int Lx, int Ly, void *stream);

/* filmpressure!(output, f, dgrad, gamma, theta, n, m, hmin, hcrit)   src/pressure.jl:72-115 (variant FAST)
 * filmpressure!(state, sys; theta, gamma, n, m, hmin, hcrit)         src/pressure.jl:119-155 (variant POWER_BROAD)
 * cospi_theta_field: NULL -> the scalar cospi_theta is used; else an Lx*Ly device field of cospi.(theta).
 * dgrad is the reference's 8-plane scratch; it is accepted for signature parity and never touched. */
int swalbe_filmpressure(double *pressure, const double *height, double *dgrad, double gamma, double cospi_theta,
                        const double *cospi_theta_field, int n, int m, double hmin, double hcrit, int pressure_variant,
                        int Lx, int Ly, void *stream);

/* h∇p!(state)                                                src/forcing.jl:168-187
 * == ∇f!(outx, outy, f, dgrad, a) with f = pressure, a = height      src/differences.jl:189-206 */
int swalbe_hgradp(double *hgradpx, double *hgradpy, const double *pressure, const double *height, int Lx, int Ly,
                  void *stream);

/* ∇f!(outx, outy, f [, a])                                   src/differences.jl:153-187 (a == NULL: 3-arg form) */
int swalbe_grad9(double *outx, double *outy, const double *f, const double *a, int Lx, int Ly, void *stream);

/* ∇²f!(output, f, gamma)                                     src/differences.jl:57-75 */
int swalbe_lap9(double *output, const double *f, double gamma, int Lx, int Ly, void *stream);

/* slippage! / slippage2! / slippage_ring_riv!                src/forcing.jl:42-46, 85-99, 107-111 */
int swalbe_slippage(double *slipx, double *slipy, const double *height, const double *velx, const double *vely,
                    double delta, double mu, double hcrit, int slip_variant, int Lx, int Ly, void *stream);

/* the inline force sum of every driver ("update!")           src/simulate.jl:18-19
 * kbtx/kbty NULL -> F = -h∇p - slip; else F = -h∇p - slip - kbt     scripts/Rivulet_stability.jl:123-124 */
int swalbe_force_sum(double *Fx, double *Fy, const double *hgradpx, const double *hgradpy, const double *slipx,
                     const double *slipy, const double *kbtx, const double *kbty, int Lx, int Ly, void *stream);

/* thermal!(kbtx, kbty, height, kbt, mu, delta)               src/forcing.jl:297-311
 * The N(0,1) draws come from a counter-based Philox4x32-10 keyed on (seed, step, cell index) so the field is
 * independent of any domain decomposition (Julia's randn! stream is not reproducible outside Julia). */
int swalbe_thermal(double *kbtx, double *kbty, const double *height, double kbt, double mu, double delta,
                   unsigned long long seed, unsigned long long step, int Lx, int Ly, void *stream);

/* inclination!(alpha, state; t, tstart, tsmooth)             src/forcing.jl:363-368
 * factor = 0.5 + 0.5*tanh((t - tstart)/tsmooth), evaluated by the caller. F += h*alpha*factor. */
int swalbe_inclination(double *Fx, double *Fy, const double *height, double alpha_x, double alpha_y, double factor,
                       int Lx, int Ly, void *stream);

/* diagnostics used inside the drivers' loops, computed on the device, result written to out[] (device):
 * out[0] = min(f), out[1] = max(f), out[2] = sum(f) (fixed-order two-pass sum), out[3] = count(f > thresh).
 * sum(state.height) src/simulate.jl:8-14; maximum-minimum :56; wetted! src/measures.jl:13-17. */
int swalbe_field_stats(double *out4, const double *f, double thresh, int Lx, int Ly, void *stream);

/* out[c] = cospi(theta[c]) (CUDA libdevice cospi, <= 1 ulp): what a host without Base.cospi on the device uses to turn a
 * contact-angle field into the cospi_theta_field argument above (Julia hosts broadcast cospi.(theta) themselves). */
int swalbe_cospi_field(double *out, const double *theta, size_t count, void *stream);

/* ---------------------------------------------------------------------------------------------
 * On-device initial conditions and substrate motion.  The reference builds these on one host
 * thread and uploads them; here every rank fills its own row slab: rows [j_begin, j_begin + Ly_local) of the global
 * lattice, `height` pointing at the slab's first row (single GPU: j_begin = 0, Ly_local = Ly).  Coordinates are the
 * reference's 1-based (i, j).  cospi_theta is cospi(θ) evaluated by the caller, as everywhere in this ABI.
 * sin/cos/asin are CUDA's: equal to Julia's openlibm within a few ulp, not bit for bit.
 * ------------------------------------------------------------------------------------------- */

/* singledroplet(height, radius, θ, center)                    src/initialvalues.jl:203-224
 * spherical cap (cos(asin(r/radius)) - cospi θ)*radius inside r <= radius, `precursor` (0.05 upstream) elsewhere and
 * wherever the cap is negative. */
int swalbe_ic_singledroplet(double *height, double radius, double cospi_theta, double cx, double cy, double precursor,
                            int Lx, int Ly_local, int j_begin, void *stream);

/* torus(lx, ly, r1, R2, θ, center, hmin; noise)               src/initialvalues.jl:144-168
 * noise != 0 adds noise*N(0,1) from the counter-based generator keyed on (seed, global cell). */
int swalbe_ic_torus(double *height, double r1, double R2, double cospi_theta, double cx, double cy, double hmin,
                    double noise, unsigned long long seed, int Lx, int Ly_local, int j_begin, void *stream);

/* rivulet(Lx, Ly, radius, θ, orientation, center, hmin; noise) src/initialvalues.jl:69-104
 * orientation 0 = :y (profile varies with i, ridge along j), 1 = :x. */
int swalbe_ic_rivulet(double *height, double radius, double cospi_theta, int orientation, double center, double hmin,
                      double noise, unsigned long long seed, int Lx, int Ly_local, int j_begin, void *stream);

/* the initial condition of run_rayleightaylor                 src/simulate.jl:350-353
 * h0*(1 + eps*sin(2π kx i/(Lx-1))*sin(2π ky j/(Ly-1))); Ly is the GLOBAL extent (the divisor), Ly_local the slab. */
int swalbe_ic_sinewave2d(double *height, double h0, double eps, double kx, double ky, int Lx, int Ly, int Ly_local,
                         int j_begin, void *stream);

/* randinterface!(height, h0, eps)                              src/initialvalues.jl:23-33
 * h0*(1 + eps*N(0,1)); Julia's unseeded randn! stream is not reproducible, the normals come from Philox4x32-10 keyed
 * on (seed, global cell) -- identical for every decomposition, compared with the reference statistically. */
int swalbe_ic_randinterface(double *height, double h0, double eps, unsigned long long seed, int Lx, int Ly_local,
                            int j_begin, void *stream);

/* circshift!(dst, src, (sx, sy)): dst[i,j] = src[i-sx, j-sy] periodic -- the body of move_substrate!
 * (scripts/Moving_wettability_structs.jl:139-152: circshift!(θ, input, (1,1)); input .= θ).  dst must not alias src. */
int swalbe_circshift(double *dst, const double *src, int sx, int sy, int Lx, int Ly, void *stream);

/* self-test (diagnostics, not part of the reference's API): runs the library's exact-division helper (shared
 * reciprocal, zero-numerator fast path) against the compiler's IEEE `/` on n pseudo-random operand triples of every
 * class (all exponents, +-0, Inf, NaN, denormals, near-equal operands) and writes the number of bitwise mismatches to
 * *mismatches (device). Must be 0. */
int swalbe_selftest_division(unsigned long long n, unsigned long long seed, unsigned long long *mismatches, void *stream);

/* self-test: one block of the library's Philox4x32-10 (the generator behind swalbe_thermal, the thermal time loop and
 * the noisy initial conditions) for counter ctr[4] and key key[2] (host arrays); the four output words go to out4
 * (device).  Checked against the published known-answer vectors of the Random123 distribution. */
int swalbe_selftest_philox(const unsigned int ctr[4], const unsigned int key[2], unsigned int *out4, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Fused time loop (what time_loop / run_* call when the device string is "GPU").
 * ------------------------------------------------------------------------------------------- */

/* State / CuState / CuState_thermal  src/initialize.jl:149-168, 214-256 (x/y split fields as in the reference) */
typedef struct swalbe_state {
  double *fout, *ftemp, *feq;                       /* Lx*Ly*9 */
  double *height, *velx, *vely, *vsq, *pressure;    /* Lx*Ly   */
  double *Fx, *Fy, *slipx, *slipy, *hgradpx, *hgradpy;
  double *dgrad;                                    /* Lx*Ly*8 scratch of the reference; unused, may be NULL */
  double *kbtx, *kbty;                              /* thermal states only; may be NULL */
} swalbe_state;

/* Taumucs  src/initialize.jl:43-60 (+ the per-call keyword overrides of filmpressure!) */
typedef struct swalbe_params {
  double tau, mu, delta, kbt, gamma, hmin, hcrit, g;
  int n, m;
  double cospi_theta;              /* cospi(theta) */
  const double *cospi_theta_field; /* device Lx*Ly field of cospi.(theta), or NULL */
  int pressure_variant;            /* SWALBE_PRESSURE_* */
  int slip_variant;                /* SWALBE_SLIP_* */
  int use_inclination;             /* adds h*alpha*factor to F between force sum and equilibrium! (src/simulate.jl:89) */
  double incl_ax, incl_ay, incl_factor;
  int use_thermal;                 /* thermal! before the force sum, F = -h∇p - slip - kbt */
  unsigned long long seed;         /* Philox key for the thermal noise */
} swalbe_params;

/* loop flags */
#define SWALBE_LOOP_DEFAULT 0
/* tau == 1 only: populations are written on the LAST step of the call only ("moments-only" steps, reported
 * separately from the 144-B/LU accounting).  Field values on return are identical to the default mode. */
#define SWALBE_LOOP_LAZY_POPULATIONS 1
/* Do not materialise feq/vsq/pressure/h∇p/slip/F[/kbt] on the last step of the call (they keep whatever they held).
 * height/velx/vely and fout == ftemp are always current on return.  For drivers that call the loop in chunks (mass
 * print every tdump steps, moving substrates) and only need the intermediate fields at the very end. */
#define SWALBE_LOOP_SKIP_AUX 2
/* tau != 1 only: the caller promises that height/velx/vely ARE the moments of ftemp (src/moments.jl:47-50) -- true on
 * return of every swalbe_time_loop call, false after an initial condition was written into height.  The first step of
 * the call can then derive h and u from the populations like every later step does (144 B per lattice update instead
 * of 192).  Drivers set it on every chunk after their first. */
#define SWALBE_LOOP_MOMENTS_CONSISTENT 4
/* per-step device logs (see swalbe_time_loop) */
typedef struct swalbe_loop_logs {
  double *hmin, *hmax;        /* device, nsteps each: min/max of height BEFORE each step (src/simulate.jl:56); NULL = off */
  unsigned long long *wetted; /* device, nsteps: count(height > hthresh) in the callback slot (src/simulate.jl:89); NULL = off */
  double hthresh;             /* 0.055 in wetted!  src/measures.jl:13 */
  /* mass log: `mass = sum(state.height)` of the drivers' dump steps (src/simulate.jl:8-14) without leaving the loop.
   * hsum[m] = sum of the height BEFORE step hsum_first + m*hsum_every of the call (0-based), for every such step below
   * nsteps.  Device memory or page-locked host memory: each value is written with a system-scope store as soon as it is
   * known, so a host thread may poll a slot it pre-set to NaN and print progress while the loop runs.  The sum is formed
   * row by row in a fixed order (it does not depend on how the loop is cut into launches).  NULL = off. */
  double *hsum;
  int hsum_first, hsum_every;
} swalbe_loop_logs;

typedef struct swalbe_plan swalbe_plan; /* opaque: library-owned scratch (3 moment planes) + launch geometry */

int swalbe_plan_create(swalbe_plan **plan, int Lx, int Ly);
int swalbe_plan_destroy(swalbe_plan *plan);

/* nsteps iterations of the loop body of time_loop (src/simulate.jl:15-22), fused into one kernel per step.
 * `step0` is the index of the first step (only used as the thermal-noise counter).  On return (stream-ordered)
 * EVERY field of `state` holds exactly what the reference's state holds after the same nsteps: height/velx/vely,
 * fout == ftemp == streamed populations, and feq/vsq/pressure/h∇p/slip/F[/kbt] of the last step.  dgrad is scratch
 * in the reference and is left untouched here. */
int swalbe_time_loop(swalbe_plan *plan, const swalbe_state *state, const swalbe_params *params, int nsteps,
                     unsigned long long step0, int flags, const swalbe_loop_logs *logs, void *stream);

/* The same loop for a job whose initial height lives in HOST memory and / or whose final height goes back to it -- the
 * pattern of every shipped GPU script: `state.height .= CUDA.adapt(CuArray, h)` ... loop ... `Array(state.height)`
 * (scripts/Moving_wettability_structs.jl:39,60-71, src/simulate.jl:349-357).  On return (stream-ordered) the state and
 * height_out_host hold bit for bit what
 *     cudaMemcpyAsync(state->height, height_in_host, H2D); swalbe_time_loop(...); cudaMemcpyAsync(height_out_host, state->height, D2H)
 * leaves, but on large lattices (tau == 1) the copies travel as row bands on the plan's own copy streams while the
 * first steps run behind the upload front and the last steps ahead of the download (csrc/sweep.h): the PCIe time of
 * the two planes disappears behind ~10 steps of compute each.  Either host pointer may be NULL (no copy in that
 * direction; both NULL == swalbe_time_loop).  Host buffers: Lx*Ly doubles in the layout of state->height, page-locked
 * (pageable memory works but serialises).  velx / vely are inputs of the first step as in swalbe_time_loop.
 * The plan holds the copy streams; the call is asynchronous and returns with the caller's stream waiting on the last
 * band's download. */
int swalbe_time_loop_host(swalbe_plan *plan, const swalbe_state *state, const swalbe_params *params, int nsteps,
                          unsigned long long step0, int flags, const swalbe_loop_logs *logs, const double *height_in_host,
                          double *height_out_host, void *stream);

/* self-test (host only, needs no device): the operation list of swalbe_time_loop_host for an Lx x Ly lattice -- seven ints
 * per operation {kind (0 upload rows, 1 step launch, 2 download rows), step, jbeg, jend, band, seam, stage}, in issue
 * order -- so that the schedule can be replayed and checked on the CPU.  Launches with stage >= 0 belong to band stage
 * `stage` of a sweep: even and odd stages run on two streams, launch (stage, k-th step of the sweep) ordered after
 * (stage, k-1) and (stage-1, k-1) only.  band_rows / kmax / min_sites <= 0: the defaults.  ops7 == NULL: only *nops. */
int swalbe_selftest_host_loop_schedule(int Lx, int Ly, int nsteps, int has_in, int has_out, int band_rows, int kmax,
                                       int min_sites, int *ops7, int max_ops, int *nops);

/* ---------------------------------------------------------------------------------------------
 * The 1-D (D1Q3) family: State_1D / SysConst_1D, src/initialize.jl:587-598.  The reference runs it
 * on the CPU only (no device string in its 1-D allocator or drivers), so these entry points have no upstream GPU
 * counterpart; they take DEVICE vectors of length L (populations: three contiguous length-L columns, column k =
 * population k with c0 = 0, c1 = +1, c2 = -1, src/collide.jl:194-196, :303-309).
 * ------------------------------------------------------------------------------------------- */

/* equilibrium!(feq, height, velocity, gravity)               src/equilibrium.jl:169-181 */
int swalbe_equilibrium_d1q3(double *feq, const double *height, const double *vel, double g, int L, void *stream);
/* BGKandStream!(fout, feq, ftemp, F::Vector, tau)            src/collide.jl:179-201 (fout == ftemp on return) */
int swalbe_bgk_stream_d1q3(double *fout, const double *feq, double *ftemp, const double *F, double tau, int L, void *stream);
/* moments!(height::Vector, vel, fout)                        src/moments.jl:54-62 */
int swalbe_moments_d1q3(double *height, double *vel, const double *fout, int L, void *stream);
/* filmpressure!(output::Vector, f, dgrad, gamma, theta, n, m, hmin, hcrit)   src/pressure.jl:196-227 (variant FAST)
 * filmpressure!(state::LBM_state_1D, sys; ...)                              src/pressure.jl:230-256 (POWER_BROAD)
 * dgrad (L x 2 in the reference) is accepted for signature parity and never touched. */
int swalbe_filmpressure_1d(double *pressure, const double *height, double *dgrad, double gamma, double cospi_theta,
                           const double *cospi_theta_field, int n, int m, double hmin, double hcrit, int pressure_variant,
                           int L, void *stream);
/* ∇f!(output::Vector, f, dgrad, a) | ∇f!(output, f::Vector, dgrad)   src/differences.jl:208-230 (a == NULL: no multiplier);
 * with f = pressure, a = height this is h∇p!(state::LBM_state_1D)    src/forcing.jl:189-198 */
int swalbe_grad_1d(double *output, const double *f, const double *a, int L, void *stream);
/* ∇²f!(output, f::Vector, dgrad)                             src/differences.jl:77-85 */
int swalbe_lap_1d(double *output, const double *f, int L, void *stream);
/* slippage!(slip, height, vel, delta, mu)                    src/forcing.jl:68-71 */
int swalbe_slippage_1d(double *slip, const double *height, const double *vel, double delta, double mu, int L, void *stream);
/* state.F .= -state.h∇p .- state.slip                        src/simulate.jl:110 */
int swalbe_force_sum_1d(double *F, const double *hgradp, const double *slip, int L, void *stream);

/* force sum of the loops that carry a third term: F = -h∇p - slip - extra
 * (run_gamma: extra = ∇γ, src/simulate.jl:544; thermal 1-D loops: extra = kbt) */
int swalbe_force_sum3_1d(double *F, const double *hgradp, const double *slip, const double *extra, int L, void *stream);
/* thermal!(fluc, height, kbt, mu, delta)                     src/forcing.jl:322-333 (State_thermal_1D: :335-336)
 * counter-based normals keyed on (seed, step, site), like swalbe_thermal */
int swalbe_thermal_1d(double *fluc, const double *height, double kbt, double mu, double delta, unsigned long long seed,
                      unsigned long long step, int L, void *stream);
/* inclination!(alpha::Float64, state::State_1D; t, tstart, tsmooth)   src/forcing.jl:379-389
 * factor = 0.5 + 0.5*tanh((t - tstart)/tsmooth), evaluated by the caller.  F += h*alpha*factor */
int swalbe_inclination_1d(double *F, const double *height, double alpha, double factor, int L, void *stream);
/* ∇γ!(state)        src/forcing.jl:423-432   (height == NULL):  -3/2 * (γ[i-1] - γ[i+1]) / 2
 * ∇γ!(state, sys)   src/forcing.jl:434-447   (height != NULL):  (2h² + 6δh + 3δ²)/(6h) * h/2 * (γ[i-1] - γ[i+1]) / 2 */
int swalbe_gradgamma_1d(double *dgamma, const double *gamma, const double *height, double delta, int L, void *stream);
/* filmpressure!(state::State_gamma_1D, sys; θ, n, m, hmin, hcrit, γ)   src/pressure.jl:284-315: power_broad, the surface
 *   tension a scalar (gamma_field == NULL) or a per-site field (run_gamma passes γ = gamma::Vector, src/simulate.jl:541);
 *   ftemp != NULL: the disjoining and the Laplace contribution are also stored in columns 1 and 2 of ftemp (L*3), as the
 *   reference does (:307-312).
 * filmpressure!(output::Vector, f, dgrad, rho, γ, θ, n, m, hmin, hcrit; Gamma)   src/pressure.jl:318-338 (active matter):
 *   rho != NULL, tension γ + Gamma*rho.
 * filmpressure!(state::Expanded_1D, sys; ...)   src/pressure.jl:258-282: gamma_field == rho == ftemp == NULL. */
int swalbe_filmpressure_gamma_1d(double *pressure, const double *height, double gamma, const double *gamma_field,
                                 const double *rho, double Gamma, double cospi_theta, const double *cospi_theta_field, int n,
                                 int m, double hmin, double hcrit, double *ftemp, int L, void *stream);
/* BGKandStream!(state::StateWithBound_1D, sys::SysConstWithBound_1D)   src/collide.jl:214-249: bounce-back walls.
 * border0 / border1: the masks sys.border[1] / sys.border[2] written by obslist! (src/obstacle.jl:41-53) as device
 * vectors; fbound: L*3, columns 1 and 2 receive the held-back populations.  fout == ftemp on return. */
int swalbe_bgk_stream_bound_d1q3(double *fout, const double *feq, double *ftemp, double *fbound, const double *F,
                                 const double *border0, const double *border1, double tau, int L, void *stream);
/* update_rho!(rho, rho_int, height, dgrad, differentials; D, M)   src/forcing.jl:399-417
 * differentials: L*4 (lap rho, grad rho, lap h, grad h are left there as in the reference) */
int swalbe_update_rho_1d(double *rho, double *rho_int, const double *height, double *differentials, double D, double M, int L,
                         void *stream);

/* State_1D  src/initialize.jl:587-598 and the fields its expanded kinds add (:304-341); NULL where the kind has none */
typedef struct swalbe_state_1d {
  double *fout, *ftemp, *feq;                     /* L*3 */
  double *height, *vel, *pressure, *F, *slip, *hgradp; /* L */
  double *dgrad;                                  /* L*2 scratch of the reference; unused, may be NULL */
  double *gamma, *dgamma;                         /* L: State_gamma_1D / StateWithBound_1D γ, ∇γ */
  double *kbt;                                    /* L: State_thermal_1D */
  double *fbound;                                 /* L*3: StateWithBound_1D */
} swalbe_state_1d;

/* flags of swalbe_time_loop_1d on top of SWALBE_LOOP_SKIP_AUX: the loop body of run_gamma (src/simulate.jl:541-547) */
#define SWALBE_LOOP_GAMMA_FIELD 8  /* filmpressure!(state, sys, γ = state.gamma): the tension is the per-site field */
#define SWALBE_LOOP_MARANGONI 16   /* F = -h∇p - slip - state.dgamma */

/* nsteps iterations of time_loop(sys::SysConst_1D, state::State_1D[, theta | Δh])   src/simulate.jl:98-157.
 * params: the Taumucs fields, pressure_variant / cospi_theta[_field] and use_inclination / incl_ax / incl_factor (the
 * callback slot of time_loop(sys, state, inclination!, α), src/simulate.jl:159-179) of swalbe_params (slip variant and
 * thermal fields are ignored); flags: SWALBE_LOOP_SKIP_AUX, SWALBE_LOOP_GAMMA_FIELD, SWALBE_LOOP_MARANGONI; logs:
 * hmin / hmax per step (wetted and hsum are ignored).
 * Lattices that fit the shared memory of one CTA (L <= ~4800 at tau == 1) run all steps but the materialising one
 * inside a single persistent launch.  On return every field of the state holds what the reference's holds. */
int swalbe_time_loop_1d(const swalbe_state_1d *state, const swalbe_params *params, int L, int nsteps, int flags,
                        const swalbe_loop_logs *logs, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU: row-slab decomposition along j (Ly), one process per GPU; halo rows as peer-memory stores over NVLink
 * (one kernel per step; see swalbe_dist_uses_peer_memory) or, where that is not available, over NCCL send/recv.
 * The reference has no multi-GPU path; this is new.  Rank r owns global rows
 * [r*Ly/nranks, (r+1)*Ly/nranks) of every plane; Ly % nranks must be 0.
 * ------------------------------------------------------------------------------------------- */
typedef struct swalbe_dist swalbe_dist; /* opaque: slabs with ghost rows, NCCL communicator, streams, events */

#define SWALBE_NCCL_UNIQUE_ID_BYTES 128
/* rank 0 calls this and ships the 128 bytes to the other ranks by any means (MPI / torch.distributed / file) */
int swalbe_dist_unique_id(void *id128);
int swalbe_dist_create(swalbe_dist **dist, const void *id128, int rank, int nranks, int Lx, int Ly_global,
                       const swalbe_params *params);
int swalbe_dist_destroy(swalbe_dist *dist);
int swalbe_dist_local_rows(const swalbe_dist *dist, int *j_begin, int *j_count);
/* upload this rank's rows of height/velx/vely (+ the nine ftemp planes when tau != 1; ftemp may be NULL at tau == 1)
 * from DEVICE arrays holding only the local slab (Lx * j_count each), then exchange halos */
int swalbe_dist_set_state(swalbe_dist *dist, const double *height, const double *velx, const double *vely,
                          const double *ftemp, void *stream);
/* contact-angle field: this rank's rows of cospi.(theta) (Lx * j_count, device); NULL switches back to the scalar of
 * `params`.  Ghost rows are exchanged once here; call again after moving the substrate. */
int swalbe_dist_set_theta(swalbe_dist *dist, const double *cospi_theta_slab, void *stream);

/* move_substrate! on the slab runtime (scripts/Moving_wettability_structs.jl:139-152): the contact-angle field of the
 * whole lattice is shifted periodically by (sx, sy), theta[i,j] <- theta[i-sx, j-sy].  Rows that cross a slab boundary
 * come out of the ghost rows, so |sy| <= 3; the ghost rows are exchanged again afterwards.  Collective: every rank
 * calls it with the same shift. */
int swalbe_dist_shift_theta(swalbe_dist *d, int sx, int sy, void *stream);
/* min / max / sum / count(h > thresh) of this rank's rows of the height field -> out4[4] (device); combine across
 * ranks on the host (min, max, +, +).  sum(state.height) src/simulate.jl:8-14, wetted! src/measures.jl:13-17 */
int swalbe_dist_height_stats(swalbe_dist *dist, double *out4, double thresh, void *stream);
/* nsteps fused steps with halo exchange overlapped with the interior update */
int swalbe_dist_time_loop(swalbe_dist *dist, int nsteps, unsigned long long step0, void *stream);
/* The slab's time loop from / to HOST memory (tau == 1): the counterpart of swalbe_time_loop_host on a rank's rows.
 * height_in_host != NULL: the state is replaced -- this rank's rows of the height come from page-locked host memory
 * (Lx * j_count doubles), the velocities from the device slabs velx / vely (NULL: zero) -- and the ghost rows are
 * exchanged; height_out_host != NULL: the rows of the final height go back to host memory.  On large slabs the planes
 * travel in row bands while the first steps run behind the upload and the last ones ahead of the download (csrc/sweep.h;
 * the strips next to the slab boundaries are stepped last, with one halo exchange per step).  Bit for bit what
 * set_state + time_loop + get_state give.  Collective: every rank calls it with the same nsteps. */
int swalbe_dist_time_loop_host(swalbe_dist *dist, int nsteps, unsigned long long step0, const double *height_in_host,
                               const double *velx, const double *vely, double *height_out_host, void *stream);
/* copy this rank's slab rows of height/velx/vely and (optional, may be NULL) the nine population planes out */
int swalbe_dist_get_state(swalbe_dist *dist, double *height, double *velx, double *vely, double *fout, void *stream);
/* *yes = 1 when the halo rows of the time loop travel as stores into the neighbours' memory (CUDA IPC mapping over
 * NVLink, one kernel per step, tau == 1), 0 when they go through NCCL send/recv (SWALBE_DIST_P2P=0, tau != 1, or the
 * ranks cannot map each other's memory) */
int swalbe_dist_uses_peer_memory(const swalbe_dist *dist, int *yes);
/* device time (ms) spent in the last swalbe_dist_time_loop call, measured with CUDA events on its streams */
int swalbe_dist_last_loop_ms(swalbe_dist *dist, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* SWALBE_B200_H */
