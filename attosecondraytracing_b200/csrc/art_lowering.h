// art_lowering.h -- host-side lowering of the C-ABI scene description (ArtElementDesc,
// ArtZernikeDesc) into the packed tables the kernels read.  Host arithmetic only.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "art_device.cuh"

namespace art {

// Pre-digest one element: rotation from (normal, majoraxis) with the reference's branching
// (ART/ModuleProcessing.py:289-294, ART/ModuleGeometry.py:333-343), squared radii, half sizes.
// Returns an empty string on success, else what is wrong.
inline std::string lower_element(const ArtElementDesc& d, ElemDev& e) {
  std::memset(&e, 0, sizeof(e));
  for (int i = 0; i < 3; ++i) {
    if (!std::isfinite(d.position[i]) || !std::isfinite(d.normal[i]) || !std::isfinite(d.majoraxis[i]) ||
        !std::isfinite(d.centre[i]))
      return "non-finite pose";
    e.pos[i] = d.position[i];
    e.ctr[i] = d.centre[i];
  }
  const double nn = d.normal[0] * d.normal[0] + d.normal[1] * d.normal[1] + d.normal[2] * d.normal[2];
  const double mm = d.majoraxis[0] * d.majoraxis[0] + d.majoraxis[1] * d.majoraxis[1] + d.majoraxis[2] * d.majoraxis[2];
  if (!(nn > 0.0) || !(mm > 0.0)) return "zero normal or majoraxis";
  element_rotation(d.normal, d.majoraxis, e.rot);
  e.surface = d.surface;
  e.support = d.support;
  e.n_defects = d.n_defects;
  e.first_defect = d.first_defect;
  const double* s = d.surface_params;
  switch (d.surface) {
    case ART_SURF_PLANE:
    case ART_SURF_MASK:
      break;
    case ART_SURF_SPHERICAL:   // ART/ModuleMirror.py:163-167: c = p.p - R^2
    case ART_SURF_CYLINDRICAL: // :831-833
      if (!(s[0] > 0.0)) return "radius must be > 0";
      e.sp[0] = s[0];
      e.sp[1] = s[0] * s[0];
      break;
    case ART_SURF_PARABOLIC:  // :276 p = feff (1 + cos(offaxis)); x^2 + y^2 = 2 p z
      if (!std::isfinite(s[0])) return "parabola parameter p must be finite";
      e.sp[0] = s[0];
      e.soff[0] = d.centre[0];  // support is tested about the optic centre (:344)
      e.soff[1] = d.centre[1];
      break;
    case ART_SURF_TOROIDAL:  // :391-395
      if (!(s[0] > 0.0) || !(s[1] > 0.0)) return "toroid radii must be > 0";
      e.sp[0] = s[0];
      e.sp[1] = s[1];
      e.sp[2] = s[1] * s[1];
      e.sp[3] = s[0] * s[0] - s[1] * s[1];
      e.sp[4] = s[0] * s[0];
      break;
    case ART_SURF_ELLIPSOIDAL:  // :565-569, :667-669
      if (!(s[0] > 0.0) || !(s[1] > 0.0)) return "ellipsoid semi-axes must be > 0";
      e.sp[0] = s[0];
      e.sp[1] = s[1];
      e.sp[2] = 1.0 / (s[0] * s[0]);
      e.sp[3] = 1.0 / (s[1] * s[1]);
      e.soff[0] = d.centre[0];  // :678
      e.soff[1] = d.centre[1];
      break;
    default:
      return "unknown surface kind";
  }
  const double* a = d.support_params;
  switch (d.support) {
    case ART_SUPP_ROUND:  // IncludeDisk: x^2 + y^2 <= R^2, ART/ModuleGeometry.py:259-268
      e.ap[0] = a[0] * a[0];
      break;
    case ART_SUPP_ROUND_HOLE:
      e.ap[0] = a[0] * a[0];
      e.ap[1] = a[1] * a[1];
      e.ap[2] = a[2];
      e.ap[3] = a[3];
      break;
    case ART_SUPP_RECT:  // IncludeRectangle: |x| <= |X/2|, ART/ModuleGeometry.py:249-255
      e.ap[0] = std::fabs(a[0] / 2);
      e.ap[1] = std::fabs(a[1] / 2);
      break;
    case ART_SUPP_RECT_HOLE:
      e.ap[0] = std::fabs(a[0] / 2);
      e.ap[1] = std::fabs(a[1] / 2);
      e.ap[2] = a[2] * a[2];
      e.ap[3] = a[3];
      e.ap[4] = a[4];
      break;
    case ART_SUPP_RECT_RECT_HOLE:
      e.ap[0] = std::fabs(a[0] / 2);
      e.ap[1] = std::fabs(a[1] / 2);
      e.ap[2] = std::fabs(a[2] / 2);
      e.ap[3] = std::fabs(a[3] / 2);
      e.ap[4] = a[4];
      e.ap[5] = a[5];
      break;
    default:
      return "unknown support kind";
  }
  if (d.n_defects < 0 || d.first_defect < 0 || d.n_gridmaps < 0 || d.first_gridmap < 0) return "negative defect index";
  if ((d.n_defects > 0 || d.n_gridmaps > 0) && (d.surface == ART_SURF_MASK)) return "a mask cannot carry defects";
  e.n_maps = d.n_gridmaps;
  e.first_map = d.first_gridmap;
  return std::string();
}

// Hand-over map between consecutive elements: a point h / direction o in element `e`'s frame arrive in
// `next`'s frame as  M h + b  /  M o  with  M = R_next R^T,  b = R_next (pos - pos_next - R^T ctr) + ctr_next
// -- the composition of ART/ModuleProcessing.py:306-309 (element -> lab) and :289-295 of the next loop trip
// (lab -> next element), evaluated once per element pair in extended precision.
inline void link_elements(ElemDev& e, const ElemDev& next) {
  long double t[3];
  for (int i = 0; i < 3; ++i) {
    long double rc = 0;  // (R^T ctr)[i]
    for (int j = 0; j < 3; ++j) rc += (long double)e.rot[3 * j + i] * e.ctr[j];
    t[i] = (long double)e.pos[i] - next.pos[i] - rc;
  }
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) {
      long double m = 0;
      for (int q = 0; q < 3; ++q) m += (long double)next.rot[3 * i + q] * e.rot[3 * j + q];
      e.nrot[3 * i + j] = (double)m;
    }
    long double b = next.ctr[i];
    for (int q = 0; q < 3; ++q) b += (long double)next.rot[3 * i + q] * t[q];
    e.noff[i] = (double)b;
  }
}

inline std::string lower_gridmap(const ArtGridMapDesc& g, MapDev& m) {
  if (g.nx < 2 || g.ny < 2) return "a grid map needs at least 2 x 2 points";
  if (!(g.x1 > g.x0) || !(g.y1 > g.y0)) return "grid extents must be increasing";
  if (!g.h || !g.dx || !g.dy) return "grid map arrays are NULL";
  m.h = g.h; m.dx = g.dx; m.dy = g.dy;
  m.x0 = g.x0; m.y0 = g.y0;
  m.sx = (g.nx - 1) / (g.x1 - g.x0);
  m.sy = (g.ny - 1) / (g.y1 - g.y0);
  m.nx = g.nx; m.ny = g.ny;
  return std::string();
}

// Zernike table of one defect in the layout zernike_eval (art_device.cuh) walks:
//   [0] radius, [1] max order N (the reference: max n over the keys, at least 2,
//   ART/ModuleDefects.py:151-154 + ART/recursive_zernike_generator.py:37-38), then for l = 0..N and
//   k = 0..(N-l)/2 the record {alpha, beta, gamma, c_cos, c_sin, pad}:
//   Q_k^l(s) = (alpha s + beta) Q_{k-1}^l(s) - gamma Q_{k-2}^l(s), Q_0 = 1, where
//   rho^l Q_k^l(rho^2) is the radial Zernike polynomial R_{l+2k}^l(rho) = (-1)^k rho^l P_k^{(l,0)}(1 - 2 rho^2);
//   the three-term recurrence is Jacobi's.  The reference key (n, m) is the polynomial
//   R_n^{|2m-n|} times cos(l theta) for 2m >= n and sin(l theta) for 2m < n.
inline std::string build_zernike_table(const ArtZernikeDesc& z, std::vector<double>& out) {
  if (!(z.radius > 0.0) || !std::isfinite(z.radius)) return "zernike radius must be > 0";
  if (z.n_coefficients < 0) return "negative coefficient count";
  int N = 2;
  std::map<std::pair<int, int>, double> coef;
  for (int i = 0; i < z.n_coefficients; ++i) {
    const int n = z.n[i], m = z.m[i];
    if (n < 0 || m < 0 || m > n) return "zernike key (n, m) needs 0 <= m <= n";
    if (n > 64) return "zernike order above 64 is not supported";
    coef[std::make_pair(n, m)] = z.c[i];  // dict semantics: a repeated key overwrites
    if (n > N) N = n;
  }
  out.clear();
  out.push_back(z.radius);
  out.push_back((double)N);
  auto get = [&](int n, int m) {
    auto it = coef.find(std::make_pair(n, m));
    return it == coef.end() ? 0.0 : it->second;
  };
  for (int l = 0; l <= N; ++l) {
    const int K = (N - l) / 2;
    for (int k = 0; k <= K; ++k) {
      const int n = l + 2 * k;
      long double alpha = 0, beta = 0, gamma = 0;
      if (k == 1) {  // Q_1 = (l+2) s - (l+1)
        alpha = l + 2;
        beta = -(long double)(l + 1);
      } else if (k >= 2) {
        const long double c = 2.0L * k + l;
        const long double D = 2.0L * k * (k + l) * (c - 2);
        alpha = 2 * (c - 1) * c * (c - 2) / D;
        beta = -(c - 1) * (c * (c - 2) + (long double)l * l) / D;
        gamma = 2 * (long double)(k + l - 1) * (k - 1) * c / D;
      }
      out.push_back((double)alpha);
      out.push_back((double)beta);
      out.push_back((double)gamma);
      out.push_back(get(n, (n + l) / 2));
      out.push_back(l > 0 ? get(n, (n - l) / 2) : 0.0);
      out.push_back(0.0);  // pad: six doubles per record = three aligned 16-byte loads (zernike_record)
    }
  }
  return std::string();
}

}  // namespace art
