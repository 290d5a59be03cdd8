// TEST INFRASTRUCTURE ONLY.  Runs the per-ray optics of attosecondraytracing_b200/csrc/art_device.cuh
// (the very code the trace kernel inlines) on the host, one ray at a time, so that the numerics can be
// checked against the golden fixtures in the GPU-less build container.  Not linked into
// libart_b200.so, not importable from the package.
#include <cstdint>
#include <string>
#include <vector>

#include "../../attosecondraytracing_b200/csrc/art_device.cuh"
#include "../../attosecondraytracing_b200/csrc/art_lowering.h"

using namespace art;

static std::string g_msg;
extern "C" const char* hc_last_error() { return g_msg.c_str(); }

// out_* arrays: n_elements x n, row k = bundle after element k (alive flags in out_alive)
extern "C" int hc_trace(const ArtElementDesc* els, int n_el, const ArtZernikeDesc* defs, int n_def,
                        const ArtGridMapDesc* gmaps, int n_maps, long long n,
                        const double* px, const double* py, const double* pz, const double* ux, const double* uy,
                        const double* uz, unsigned flags, double* opx, double* opy, double* opz, double* oux,
                        double* ouy, double* ouz, double* opath, double* oinc, uint8_t* oalive) {
  std::vector<ElemDev> E(n_el);
  for (int k = 0; k < n_el; ++k) {
    std::string why = lower_element(els[k], E[k]);
    if (!why.empty()) {
      g_msg = why;
      return -1;
    }
  }
  for (int k = 0; k + 1 < n_el; ++k) link_elements(E[k], E[k + 1]);
  std::vector<double> ztab;
  std::vector<int> zoff;
  for (int i = 0; i < n_def; ++i) {
    std::vector<double> t;
    std::string why = build_zernike_table(defs[i], t);
    if (!why.empty()) {
      g_msg = why;
      return -1;
    }
    zoff.push_back((int)ztab.size());
    ztab.insert(ztab.end(), t.begin(), t.end());
  }
  std::vector<MapDev> maps(n_maps);
  for (int i = 0; i < n_maps; ++i) {
    std::string why = lower_gridmap(gmaps[i], maps[i]);  // host pointers here
    if (!why.empty()) {
      g_msg = why;
      return -1;
    }
  }
  const bool ign = (flags & ART_TRACE_IGNORE_DEFECTS) != 0;
  for (long long i = 0; i < n; ++i) {
    Ray r;
    r.px = px[i]; r.py = py[i]; r.pz = pz[i];
    r.ux = ux[i]; r.uy = uy[i]; r.uz = uz[i];
    r.path = 0.0;
    r.inc = ART_NAN;
    r.alive = true;
    // the trace kernel's element loop: inner elements hand the ray over in the next element's frame and
    // the lab-frame bundle of the history is recovered with frame_to_lab
    to_element_frame(E[0], r);
    for (int k = 0; k < n_el; ++k) {
      const bool inner = k + 1 < n_el;
      if (r.alive) apply_element<true, true, SURFS_ANY, double>(E[k], r, ztab.data(), zoff.data(), ign, true, maps.data(), inner);
      Ray w = r;
      if (inner) frame_to_lab(E[k + 1], r, w);
      const long long o = (long long)k * n + i;
      oalive[o] = w.alive;
      opx[o] = w.px; opy[o] = w.py; opz[o] = w.pz;
      oux[o] = w.ux; ouy[o] = w.uy; ouz[o] = w.uz;
      opath[o] = w.path; oinc[o] = w.inc;
    }
  }
  return 0;
}
