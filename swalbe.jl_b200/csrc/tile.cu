// Launcher of the small-lattice tile kernel (tile.cuh): instantiations per pressure mode / gravity, eligibility.
#include <stdlib.h>

#include "tile.cuh"
#include "launch.h"

namespace swalbe {
namespace {

typedef void (*tile_fn)(FusedArgs);
template <bool GZ, bool TF>
tile_fn pick_tile(int pm) {
  switch (pm) {
    case PM_BROAD_93: return k_tile_step<PM_BROAD_93, GZ, TF>;
    case PM_BROAD_32: return k_tile_step<PM_BROAD_32, GZ, TF>;
    case PM_FAST_93: return k_tile_step<PM_FAST_93, GZ, TF>;
    case PM_FAST_32: return k_tile_step<PM_FAST_32, GZ, TF>;
    default: return nullptr;
  }
}

}  // namespace

// A contact-angle field and nothing else (standard slip, no inclination, no logs: the moving-wettability scripts, mostly
// 512^2) takes the TF instantiations.  Measured on B200 (profiles/r02_probes_call2.txt, 98-step loops, graph replay):
// 256^2 10.9 -> 9.6 us/step, 384^2 14.9 -> 9.2, 512^2 18.2 -> 15.2 against the run-time-option marching kernel, equal
// at 128^2; bit parity on the CPU emulation and on the GPU (tests).  SWALBE_TILE_THETA=0 switches it off.
static bool tile_theta_enabled() {
  const char *s = getenv("SWALBE_TILE_THETA");
  return !(s && *s) || atoi(s) != 0;
}

bool tile_eligible(const KernelKey &k, const FusedArgs &a) {
  if (!(k.lean_pm > 0 && k.tau1 && !k.thermal && a.wrap_y == 1 && a.jbeg == 0 && a.jend == a.Ly &&
        a.log_min == nullptr && a.log_wet == nullptr && a.pressure == nullptr))
    return false;
  if (a.ct_field == nullptr) return !k.opts;
  return a.sc.variant == SWALBE_SLIP_STANDARD && !a.use_incl && tile_theta_enabled();
}

int launch_tile(const FusedArgs &a, const KernelKey &key, cudaStream_t stream) {
  const bool gz = a.ec.g == 0.0;  // (OPTS keys do not carry the g == 0 specialisation; the constants do)
  tile_fn fn = a.ct_field != nullptr ? (gz ? pick_tile<true, true>(key.lean_pm) : pick_tile<false, true>(key.lean_pm))
                                     : (key.gz ? pick_tile<true, false>(key.lean_pm) : pick_tile<false, false>(key.lean_pm));
  if (!fn) return set_error(SWALBE_ERR_ARG, "no tile kernel for pressure mode %d", key.lean_pm);
  dim3 grid((a.Lx + TX - 1) / TX, (a.Ly + TY - 1) / TY);
  if (grid.y > 65535u) return set_error(SWALBE_ERR_EXTENT, "lattice too tall for the tile kernel");
  fn<<<grid, TT, 0, stream>>>(a);
  SW_LAUNCH_CHECK();
  return 0;
}

}  // namespace swalbe
