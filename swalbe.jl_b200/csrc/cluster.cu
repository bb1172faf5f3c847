// Launcher of the persistent cluster kernel (cluster.cuh): eligibility, cluster size, shared-memory opt-in.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "cluster.cuh"
#include "launch.h"

namespace swalbe {
namespace {

constexpr int CNT = 512;  // threads per CTA: 100^2 on 16 CTAs is ~1.6 site updates per thread and phase
typedef void (*cluster_fn)(const ClusterArgs);

template <bool GZ>
cluster_fn pick_cluster(int pm) {
  switch (pm) {
    case PM_BROAD_93: return k_cluster_steps<CNT, PM_BROAD_93, GZ>;
    case PM_BROAD_32: return k_cluster_steps<CNT, PM_BROAD_32, GZ>;
    case PM_FAST_93: return k_cluster_steps<CNT, PM_FAST_93, GZ>;
    case PM_FAST_32: return k_cluster_steps<CNT, PM_FAST_32, GZ>;
    default: return nullptr;
  }
}

int env_i(const char *name, int dflt) {
  const char *s = getenv(name);
  return s && *s ? atoi(s) : dflt;
}

}  // namespace

// Cluster size for an Lx x Ly lattice, 0 = the lattice does not fit (or the kernel is switched off).  The largest
// cluster whose slabs are at least 4 rows tall wins: fewer sites per thread, and the halo recomputation (pressure on
// rows+4, collisions on rows+2) is latency the other CTAs' threads would otherwise idle through.
int cluster_plan(const KernelKey &key, const FusedArgs &a, int *rows_max, size_t *smem_bytes) {
  if (!env_i("SWALBE_CLUSTER", 1)) return 0;
  // measured on B200 (profiles/r02_probes_call4.txt, us per step against tile kernel + graph replay): 32^2 3.0 vs 4.5,
  // 64^2 3.7 vs 4.7, 100^2 5.9 vs 4.9 without logs; with per-step logs (no graph replay possible) 100^2 8.6 vs 10.6,
  // 128^2 a tie -> the default bounds below (SWALBE_CLUSTER_MAX / SWALBE_CLUSTER_MAX_LOGS sites)
  const bool logs = a.log_min != nullptr || a.log_wet != nullptr;
  const size_t max_sites = (size_t)std::max(0, logs ? env_i("SWALBE_CLUSTER_MAX_LOGS", 128 * 128) : env_i("SWALBE_CLUSTER_MAX", 80 * 80));
  if ((size_t)a.Lx * a.Ly > max_sites) return 0;
  if (!(key.tau1 && key.lean_pm > 0 && !key.thermal && a.wrap_y == 1 && a.jbeg == 0 && a.jend == a.Ly && a.ct_field == nullptr &&
        a.sc.variant == SWALBE_SLIP_STANDARD && !a.use_incl))
    return 0;
  int dev = 0, max_optin = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  // (keys of loops with per-step logs do not carry the g == 0 specialisation; the constants do)
  cluster_fn fn = a.ec.g == 0.0 ? pick_cluster<true>(key.lean_pm) : pick_cluster<false>(key.lean_pm);
  if (!fn) return 0;
  const int forced = env_i("SWALBE_CLUSTER_SIZE", 0);
  for (int C : {16, 8, 4, 2, 1}) {
    if (forced ? C != forced : (C > 1 && a.Ly / C < 4)) continue;
    if (a.Ly / C < 3) continue;  // a halo (3 rows) must come from the immediate neighbour alone
    if ((size_t)a.Lx * ((a.Ly + C - 1) / C + 4) >= 65536 || a.Lx > 1024) continue;  // (range of the kernel's division-free row index)
    const int R = (a.Ly + C - 1) / C;
    const size_t bytes = cluster_smem_doubles(a.Lx, R) * sizeof(double);
    if (bytes + 2048 > (size_t)max_optin) continue;
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) { cudaGetLastError(); continue; }
    if (C > 8 && cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); continue; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C); cfg.blockDim = dim3(CNT); cfg.dynamicSmemBytes = bytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, fn, &cfg) != cudaSuccess || nclusters < 1) { cudaGetLastError(); continue; }
    *rows_max = R; *smem_bytes = bytes;
    return C;
  }
  return 0;
}

int launch_cluster_logs(const double *part, int C, int nsteps, double *log_min, double *log_max, unsigned long long *log_wet,
                        cudaStream_t stream) {
  k_cluster_logs<<<(nsteps + 127) / 128, 128, 0, stream>>>(part, C, nsteps, log_min, log_max, log_wet);
  SW_LAUNCH_CHECK();
  return 0;
}

int launch_cluster(const ClusterArgs &ca, const KernelKey &key, int C, size_t smem_bytes, cudaStream_t stream) {
  cluster_fn fn = ca.a.ec.g == 0.0 ? pick_cluster<true>(key.lean_pm) : pick_cluster<false>(key.lean_pm);
  if (!fn) return set_error(SWALBE_ERR_ARG, "no cluster kernel for pressure mode %d", key.lean_pm);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C); cfg.blockDim = dim3(CNT); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  SW_CUDA(cudaLaunchKernelEx(&cfg, fn, ca));
  SW_LAUNCH_CHECK();
  if (env_i("SWALBE_DEBUG", 0))
    fprintf(stderr, "[swalbe] cluster kernel: %d x %d lattice, %d steps, cluster of %d CTAs x %d threads, %zu B smem each\n",
            ca.a.Lx, ca.a.Ly, ca.nsteps, C, CNT, smem_bytes);
  return 0;
}

}  // namespace swalbe
