// Launch geometry of the fused step kernel, shared by fused.cu (single GPU) and dist.cu (slabs).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace swalbe {

struct FusedArgs;

struct LaunchGeom {
  int nt;             // threads per CTA (kernel variant)
  int variant;        // index into the variant table
  int W;              // output columns per CTA
  int nstrips;        // CTAs along x
  int rows_per_cta;   // rows marched by one CTA
  int nchunks;        // CTAs along y for the full row range
  int blocks_per_sm;  // occupancy of the chosen variant
};

// which kernel instantiation: lean_pm > 0 selects the lean tau==1 kernel compiled for that pressure mode
struct KernelKey {
  bool tau1, thermal;
  int lean_pm;
  bool bulk;  // lean non-thermal kernel whose row prefetch uses cp.async.bulk (needs even Lx >= NT and 16-B aligned planes)
  bool gz;    // gravity == 0: lean kernels with the (+-0)*h terms of the equilibrium folded away
  bool lazy;  // populations are not written by this launch (geometry only: lower HBM floor)
  bool opts;  // lean kernel that takes theta field / slip variant / inclination / logs at run time
  bool fm;    // tau != 1: height / velocity derived from the streamed populations instead of read from their planes
  bool ns;    // strict lean kernels: neighbour-warp hand-shake instead of the per-row CTA barrier
};

int choose_geometry(int Lx, int nrows, const KernelKey &key, LaunchGeom *g);
int launch_fused(const LaunchGeom &g, const FusedArgs &a, const KernelKey &key, cudaStream_t stream);

}  // namespace swalbe

struct swalbe_params;
namespace swalbe {
int fill_consts(FusedArgs &a, const swalbe_params &p);
KernelKey make_key(const swalbe_params &p, int pmode, bool want_lean);
bool bulk_eligible(int Lx, size_t ncells);
// persistent multi-step kernel for lattices that fit one thread-block cluster (cluster.cu)
struct ClusterArgs;
int cluster_plan(const KernelKey &key, const FusedArgs &a, int *rows_max, size_t *smem_bytes);
int launch_cluster(const ClusterArgs &ca, const KernelKey &key, int C, size_t smem_bytes, cudaStream_t stream);
int launch_cluster_logs(const double *part, int C, int nsteps, double *log_min, double *log_max, unsigned long long *log_wet,
                        cudaStream_t stream);
// small-lattice tile flavour of the strict lean step (tile.cu)
bool tile_eligible(const KernelKey &key, const FusedArgs &a);
int launch_tile(const FusedArgs &a, const KernelKey &key, cudaStream_t stream);
}
