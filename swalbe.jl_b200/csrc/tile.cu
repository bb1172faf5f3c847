// Launcher of the small-lattice tile kernel (tile.cuh): instantiations per pressure mode / gravity, eligibility.
#include "tile.cuh"
#include "launch.h"

namespace swalbe {
namespace {

typedef void (*tile_fn)(FusedArgs);
template <bool GZ>
tile_fn pick_tile(int pm) {
  switch (pm) {
    case PM_BROAD_93: return k_tile_step<PM_BROAD_93, GZ>;
    case PM_BROAD_32: return k_tile_step<PM_BROAD_32, GZ>;
    case PM_FAST_93: return k_tile_step<PM_FAST_93, GZ>;
    case PM_FAST_32: return k_tile_step<PM_FAST_32, GZ>;
    default: return nullptr;
  }
}

}  // namespace

bool tile_eligible(const KernelKey &k, const FusedArgs &a) {
  return k.lean_pm > 0 && k.tau1 && !k.thermal && !k.opts && a.wrap_y == 1 && a.jbeg == 0 && a.jend == a.Ly &&
         a.ct_field == nullptr && a.log_min == nullptr && a.log_wet == nullptr && a.pressure == nullptr;
}

int launch_tile(const FusedArgs &a, const KernelKey &key, cudaStream_t stream) {
  tile_fn fn = key.gz ? pick_tile<true>(key.lean_pm) : pick_tile<false>(key.lean_pm);
  if (!fn) return set_error(SWALBE_ERR_ARG, "no tile kernel for pressure mode %d", key.lean_pm);
  dim3 grid((a.Lx + TX - 1) / TX, (a.Ly + TY - 1) / TY);
  if (grid.y > 65535u) return set_error(SWALBE_ERR_EXTENT, "lattice too tall for the tile kernel");
  fn<<<grid, TT, 0, stream>>>(a);
  SW_LAUNCH_CHECK();
  return 0;
}

}  // namespace swalbe
