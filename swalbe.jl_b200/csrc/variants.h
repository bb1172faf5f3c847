// Table of fused-kernel instantiations, one translation unit per CTA width so that they compile in parallel.
#pragma once
#include "fused.cuh"

namespace swalbe {

typedef void (*fused_fn)(const FusedArgs);

struct Variant {
  int nt;
  fused_fn full[2][2];     // [tau1][thermal]        PM = -1: every option at run time
  fused_fn lean[2][5][2];  // [thermal][pmode][g==0] tau == 1 only, pmode in PM_BROAD_93..PM_FAST_32 (index 0 unused)
  fused_fn opts[2][5];     // [thermal][pmode]       lean + theta field / slip variant / inclination / logs at run time
  fused_fn bulk[5][2];     // [pmode][g==0]          lean, non-thermal, rows prefetched with cp.async.bulk (TMA unit)
  fused_fn fm_full[2];     // [thermal]              tau != 1, moments derived from the streamed populations (FM), run-time options
  fused_fn fm_lean[5][2];  // [pmode][g==0]          tau != 1, FM, strict lean
  fused_fn lean_ns[2][5][2];  // [thermal][pmode][g==0]  strict lean, neighbour-warp hand-shake instead of the CTA barrier (NS)
  fused_fn bulk_ns[5][2];     // [pmode][g==0]           NS + per-warp bulk-copy rows
};

extern const Variant g_variant_128, g_variant_160, g_variant_192, g_variant_224, g_variant_256;

// MB1: minimum CTAs/SM requested for the tau == 1 kernels, MB0: for the general-tau kernels (18 more registers)
#define SW_LEAN_ROW(NT, MB1, TH, GZ, OPTS)                                                                    \
  k_fused_step<NT, MB1, true, TH, PM_BROAD_93, false, GZ, OPTS>, k_fused_step<NT, MB1, true, TH, PM_BROAD_32, false, GZ, OPTS>, \
      k_fused_step<NT, MB1, true, TH, PM_FAST_93, false, GZ, OPTS>, k_fused_step<NT, MB1, true, TH, PM_FAST_32, false, GZ, OPTS>
#define SW_BULK_ROW(NT, MB1, GZ)                                                                              \
  k_fused_step<NT, MB1, true, false, PM_BROAD_93, true, GZ, false>, k_fused_step<NT, MB1, true, false, PM_BROAD_32, true, GZ, false>, \
      k_fused_step<NT, MB1, true, false, PM_FAST_93, true, GZ, false>, k_fused_step<NT, MB1, true, false, PM_FAST_32, true, GZ, false>

#define SW_FM_ROW(NT, MB0, GZ)                                                                                \
  k_fused_step<NT, MB0, false, false, PM_BROAD_93, false, GZ, false, true>, k_fused_step<NT, MB0, false, false, PM_BROAD_32, false, GZ, false, true>, \
      k_fused_step<NT, MB0, false, false, PM_FAST_93, false, GZ, false, true>, k_fused_step<NT, MB0, false, false, PM_FAST_32, false, GZ, false, true>

#define SW_NS_ROW(NT, MB1, TH, BULK, GZ)                                                                     \
  k_fused_step<NT, MB1, true, TH, PM_BROAD_93, BULK, GZ, false, false, true>, k_fused_step<NT, MB1, true, TH, PM_BROAD_32, BULK, GZ, false, false, true>, \
      k_fused_step<NT, MB1, true, TH, PM_FAST_93, BULK, GZ, false, false, true>, k_fused_step<NT, MB1, true, TH, PM_FAST_32, BULK, GZ, false, false, true>

#define SW_DEFINE_VARIANT(NT, MB1, MB0)                                                                       \
  namespace {                                                                                                 \
  const fused_fn lean_##NT[2][2][4] = {{{SW_LEAN_ROW(NT, MB1, false, false, false)}, {SW_LEAN_ROW(NT, MB1, false, true, false)}}, \
                                       {{SW_LEAN_ROW(NT, MB1, true, false, false)}, {SW_LEAN_ROW(NT, MB1, true, true, false)}}};  \
  const fused_fn opts_##NT[2][4] = {{SW_LEAN_ROW(NT, MB1, false, false, true)}, {SW_LEAN_ROW(NT, MB1, true, false, true)}}; \
  const fused_fn bulk_##NT[2][4] = {{SW_BULK_ROW(NT, MB1, false)}, {SW_BULK_ROW(NT, MB1, true)}};              \
  const fused_fn fm_##NT[2][4] = {{SW_FM_ROW(NT, MB0, false)}, {SW_FM_ROW(NT, MB0, true)}};                   \
  const fused_fn lns_##NT[2][2][4] = {{{SW_NS_ROW(NT, MB1, false, false, false)}, {SW_NS_ROW(NT, MB1, false, false, true)}}, \
                                      {{SW_NS_ROW(NT, MB1, true, false, false)}, {SW_NS_ROW(NT, MB1, true, false, true)}}};  \
  const fused_fn bns_##NT[2][4] = {{SW_NS_ROW(NT, MB1, false, true, false)}, {SW_NS_ROW(NT, MB1, false, true, true)}};       \
  Variant make_##NT() {                                                                                       \
    Variant v = {};                                                                                           \
    v.nt = NT;                                                                                                \
    v.full[0][0] = k_fused_step<NT, MB0, false, false, -1, false, false, true>;                               \
    v.full[0][1] = k_fused_step<NT, MB0, false, true, -1, false, false, true>;                                \
    v.full[1][0] = k_fused_step<NT, MB1, true, false, -1, false, false, true>;                                \
    v.full[1][1] = k_fused_step<NT, MB1, true, true, -1, false, false, true>;                                 \
    for (int th = 0; th < 2; ++th)                                                                            \
      for (int gz = 0; gz < 2; ++gz)                                                                          \
        for (int pm = 1; pm <= 4; ++pm) v.lean[th][pm][gz] = lean_##NT[th][gz][pm - 1];                       \
    for (int th = 0; th < 2; ++th)                                                                            \
      for (int pm = 1; pm <= 4; ++pm) v.opts[th][pm] = opts_##NT[th][pm - 1];                                 \
    for (int gz = 0; gz < 2; ++gz)                                                                            \
      for (int pm = 1; pm <= 4; ++pm) v.bulk[pm][gz] = bulk_##NT[gz][pm - 1];                                 \
    v.fm_full[0] = k_fused_step<NT, MB0, false, false, -1, false, false, true, true>;                         \
    v.fm_full[1] = k_fused_step<NT, MB0, false, true, -1, false, false, true, true>;                          \
    for (int gz = 0; gz < 2; ++gz)                                                                            \
      for (int pm = 1; pm <= 4; ++pm) v.fm_lean[pm][gz] = fm_##NT[gz][pm - 1];                                \
    for (int th = 0; th < 2; ++th)                                                                            \
      for (int gz = 0; gz < 2; ++gz)                                                                          \
        for (int pm = 1; pm <= 4; ++pm) v.lean_ns[th][pm][gz] = lns_##NT[th][gz][pm - 1];                     \
    for (int gz = 0; gz < 2; ++gz)                                                                            \
      for (int pm = 1; pm <= 4; ++pm) v.bulk_ns[pm][gz] = bns_##NT[gz][pm - 1];                               \
    return v;                                                                                                 \
  }                                                                                                           \
  }                                                                                                           \
  const Variant g_variant_##NT = make_##NT();

}  // namespace swalbe
