// Launch geometry of the fused step kernel, shared by fused.cu (single GPU) and dist.cu (slabs).
#pragma once
#include <cuda_runtime.h>

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

int choose_geometry(int Lx, int nrows, bool tau1, bool thermal, LaunchGeom *g);
int launch_fused(const LaunchGeom &g, const FusedArgs &a, bool tau1, bool thermal, cudaStream_t stream);

}  // namespace swalbe

struct swalbe_params;
namespace swalbe {
int fill_consts(FusedArgs &a, const swalbe_params &p);
}
