// Skewed sweep of a time loop whose initial height arrives from, and whose final height leaves to, HOST memory
// (swalbe_time_loop_host).  Pure host C++, no CUDA: the schedule is a list of operations that fused.cu issues and that
// tests/test_simt_emulation.py replays on the CPU through the emulated kernels.
//
// One step has dependency radius 3 rows (h -> p -> grad p -> pull), so step k of a row needs step k-1 of the rows within
// +-3.  The lattice is cut into row bands [Y_b, Y_b+1); once bands 0..b are on the device, step k can be computed on
// the rows [Y_b - 3k, Y_b+1 - 3k): the band, lagging 3 rows more with every step.  Issued band after band, the first K
// steps of the loop run behind the upload front instead of after the whole copy; run the same way, the last K steps
// finish band after band and every band leaves for the host while the next one is still being computed.  The lattice
// is periodic in y, so the sweep cannot close on itself: the rows within 3k of the seam between the last band and
// the first, [Ly - 3k, Ly) + [0, 3k), are computed after the last band ("seam" launches, a few rows each).
// Two moment buffers suffice (the ping-pong of the plain loop): what step k+1 overwrites in a band is state k-1 of rows
// that every later launch reads only beyond -- the skew equals the dependency radius (checked on the CPU emulation
// with the copies replayed as early and as late as the stream order allows).
#pragma once
#include <vector>

namespace swalbe {

enum SweepKind { SWEEP_UPLOAD = 0, SWEEP_STEP = 1, SWEEP_DOWNLOAD = 2 };

struct SweepOp {
  int kind;        // SweepKind
  int step;        // SWEEP_STEP: index of the step inside the call, 0-based
  int jbeg, jend;  // rows [jbeg, jend)
  int band;        // SWEEP_UPLOAD: band index; SWEEP_STEP: band whose upload the launch has to wait for (-1: none);
                   // SWEEP_DOWNLOAD: running index of the download
  int seam;        // SWEEP_STEP: 1 = a seam launch (a few rows)
  int stage;       // SWEEP_STEP / SWEEP_DOWNLOAD inside a sweep: the band stage b it belongs to; -1: whole lattice, seam
};
// What a launch depends on (besides its band's upload): launch (stage b, k-th step of the sweep) reads state k-1 of the rows
// [Y_b - 3(k+1), Y_b+1 - 3(k-1)), produced by (b, k-1) and (b-1, k-1), and overwrites state k-2 of its own rows, last read by
// those same two launches.  Nothing else: the stages of a sweep may run on two streams, even stages on one, odd stages on
// the other, with one event from (b-1, k-1) to (b, k) -- the tail of one launch fills with the head of the next.
// Seam launches and whole-lattice steps need everything before them.

struct SweepConfig {
  int nbands;   // 0: not worth it / not possible -> plain copy, loop, copy
  int k_up;     // steps that run behind the upload front
  int k_dn;     // steps that run ahead of the download
  int single;   // 1: the whole loop is one sweep that uploads and downloads (nsteps <= kmax)
};

inline int sweep_band_begin(int b, int nbands, int Ly) { return (int)((long long)b * Ly / nbands); }

// band_rows_req / kmax_req / min_sites_req <= 0: defaults (bands of ~8 Mi sites = 64 MiB of height; 12 steps;
// lattices of at least 2^21 sites)
inline SweepConfig sweep_configure(int Lx, int Ly, int nsteps, bool has_in, bool has_out, int band_rows_req, int kmax_req,
                                   long long min_sites_req) {
  SweepConfig c = {0, 0, 0, 0};
  const long long min_sites = min_sites_req > 0 ? min_sites_req : (1ll << 21);
  if ((!has_in && !has_out) || nsteps < 1 || (long long)Lx * Ly < min_sites) return c;
  int band = band_rows_req > 0 ? band_rows_req : (int)((8ll << 20) / Lx);
  if (band < 64 && band_rows_req <= 0) band = 64;
  int nb = (Ly + band / 2) / band;
  if (nb < 2) nb = 2;
  if (nb > 64) nb = 64;
  const int minband = Ly / nb;  // bands are [b*Ly/nb, (b+1)*Ly/nb): never shorter than this
  int kmax = kmax_req > 0 ? (kmax_req < 32 ? kmax_req : 32) : 12;
  if (kmax > (minband - 2) / 6) kmax = (minband - 2) / 6;  // band 0 shrinks by 3 rows at both ends with every step
  if (kmax < 1) return c;
  c.nbands = nb;
  if (nsteps <= kmax && has_in && has_out) { c.single = 1; c.k_up = nsteps; return c; }
  if (has_in && has_out && nsteps < 2 * kmax) { c.k_up = nsteps / 2; c.k_dn = nsteps - c.k_up; return c; }
  c.k_up = has_in ? (nsteps < kmax ? nsteps : kmax) : 0;
  c.k_dn = has_out ? (nsteps - c.k_up < kmax ? nsteps - c.k_up : kmax) : 0;
  return c;
}

// one sweep over steps [s0, s0 + K)
inline void sweep_phase(std::vector<SweepOp> &ops, int Ly, int nbands, int s0, int K, bool upload, bool download, int *ndown) {
  for (int b = 0; b < nbands; ++b) {
    const int y0 = sweep_band_begin(b, nbands, Ly), y1 = sweep_band_begin(b + 1, nbands, Ly);
    if (upload) ops.push_back({SWEEP_UPLOAD, 0, y0, y1, b, 0, -1});
    int lo = 0, hi = 0;
    for (int k = 1; k <= K; ++k) {
      lo = b == 0 ? 3 * k : y0 - 3 * k;
      hi = y1 - 3 * k;
      ops.push_back({SWEEP_STEP, s0 + k - 1, lo, hi, upload && k == 1 ? b : -1, 0, b});
    }
    if (download) ops.push_back({SWEEP_DOWNLOAD, 0, lo, hi, (*ndown)++, 0, b});
  }
  for (int k = 1; k <= K; ++k) {
    ops.push_back({SWEEP_STEP, s0 + k - 1, Ly - 3 * k, Ly, -1, 1, -1});
    ops.push_back({SWEEP_STEP, s0 + k - 1, 0, 3 * k, -1, 1, -1});
  }
  if (download) {
    ops.push_back({SWEEP_DOWNLOAD, 0, Ly - 3 * K, Ly, (*ndown)++, 0, -1});
    ops.push_back({SWEEP_DOWNLOAD, 0, 0, 3 * K, (*ndown)++, 0, -1});
  }
}

// the whole call: [sweep behind the upload] [plain whole-lattice steps] [sweep ahead of the download]
inline std::vector<SweepOp> sweep_schedule(const SweepConfig &c, int Ly, int nsteps, bool has_in, bool has_out) {
  std::vector<SweepOp> ops;
  int ndown = 0;
  if (c.nbands == 0) {
    if (has_in) ops.push_back({SWEEP_UPLOAD, 0, 0, Ly, 0, 0, -1});
    for (int s = 0; s < nsteps; ++s) ops.push_back({SWEEP_STEP, s, 0, Ly, s == 0 && has_in ? 0 : -1, 0, -1});
    if (has_out) ops.push_back({SWEEP_DOWNLOAD, 0, 0, Ly, ndown++, 0, -1});
    return ops;
  }
  if (c.single) {
    sweep_phase(ops, Ly, c.nbands, 0, c.k_up, true, true, &ndown);
    return ops;
  }
  if (c.k_up > 0) sweep_phase(ops, Ly, c.nbands, 0, c.k_up, true, false, &ndown);
  else if (has_in) ops.push_back({SWEEP_UPLOAD, 0, 0, Ly, 0, 0, -1});
  for (int s = c.k_up; s < nsteps - c.k_dn; ++s)
    ops.push_back({SWEEP_STEP, s, 0, Ly, s == 0 && has_in ? 0 : -1, 0, -1});
  if (c.k_dn > 0) sweep_phase(ops, Ly, c.nbands, nsteps - c.k_dn, c.k_dn, false, true, &ndown);
  else if (has_out) ops.push_back({SWEEP_DOWNLOAD, 0, 0, Ly, ndown++, 0, -1});
  return ops;
}

}  // namespace swalbe
