// art_device.cuh -- device-side data layout and per-ray optics of libart_b200.
//
// Everything here works on ONE ray held in registers, in FP64.  The reference semantics each
// function reproduces are cited as file:line of the reference repository (ART v0.93).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/art_b200.h"

// Per-ray functions are __host__ __device__ so that tests/hostcheck can run the very same code on
// the CPU while debugging numerics (test infrastructure only -- libart_b200.so exports no host path).
#define ART_HD __host__ __device__ __forceinline__
#ifdef __CUDA_ARCH__
#define ART_NAN CUDART_NAN
#else
#define ART_NAN __builtin_nan("")
#endif

namespace art {

// ---------------------------------------------------------------------------------------------
// packed element table entry (one per variant x element), read at warp-uniform addresses
// ---------------------------------------------------------------------------------------------
struct __align__(16) ElemDev {
  double rot[9];   // lab -> element rotation, row-major (R2 R1 of ART/ModuleProcessing.py:289-294)
  double pos[3];   // OpticalElement.position
  double ctr[3];   // Type.get_centre()
  double sp[6];    // surface parameters, pre-digested (see lower_element)
  double ap[6];    // support parameters, pre-digested
  double soff[2];  // (x,y) subtracted before the support test (parabola / ellipsoid: centre)
  int32_t surface;
  int32_t support;
  int32_t n_defects;
  int32_t first_defect;
};

struct BundleDev {
  double *px, *py, *pz, *ux, *uy, *uz, *path, *inc, *inten;
  uint8_t* alive;
};

// A ray in registers.
struct Ray {
  double px, py, pz, ux, uy, uz, path, inc;
  bool alive;
};

ART_HD double sq(double x) { return x * x; }

// ---------------------------------------------------------------------------------------------
// Branch-free FP64 division / square root / reciprocal square root for the device: an MUFU seed
// (rcp.approx / rsqrt.approx, ~2^-23) refined by two Newton steps in FMAs and a final residual
// correction -> results within 1 ulp for normal-range operands, with none of the slow-path calls,
// predicate fix-ups and register shuffles of the IEEE-exact library sequences.  Operands here
// are lengths, direction cosines and their products (1e-300 < |x| < 1e300).  A zero divisor gives NaN
// instead of +-inf; every caller treats both as "no hit" (comparisons with NaN are false).
// On the host (tests/hostcheck) the plain operators are used.
// ---------------------------------------------------------------------------------------------
ART_HD double fdiv(double a, double b) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  r = fma(r, fma(-b, r, 1.0), r);
  r = fma(r, fma(-b, r, 1.0), r);
  const double q = a * r;
  return fma(fma(-b, q, a), r, q);
#else
  return a / b;
#endif
}
ART_HD double frsqrt(double x) {
#ifdef __CUDA_ARCH__
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double h = 0.5 * x;
  y = y * fma(-h * y, y, 1.5);
  y = y * fma(-h * y, y, 1.5);
  return y;
#else
  return 1.0 / sqrt(x);
#endif
}
ART_HD double fsqrt(double x) {
#ifdef __CUDA_ARCH__
  const double y = frsqrt(x);
  double s = x * y;
  s = fma(fma(-s, s, x), 0.5 * y, s);
  return x == 0.0 ? 0.0 : s;
#else
  return sqrt(x);
#endif
}

// ---------------------------------------------------------------------------------------------
// RotationPoint as a matrix, ART/ModuleGeometry.py:333-343 with AngleBetweenTwoVectors :40-44
// (Kahan) and RotationAroundAxis :321-329 (quaternion rotation == Rodrigues).  Shared by the
// host lowering and the device autoplace kernel so both branch identically.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline double kahan_angle(const double* U, const double* V) {
  double u = sqrt(U[0] * U[0] + U[1] * U[1] + U[2] * U[2]);
  double v = sqrt(V[0] * V[0] + V[1] * V[1] + V[2] * V[2]);
  double d0 = U[0] * v - V[0] * u, d1 = U[1] * v - V[1] * u, d2 = U[2] * v - V[2] * u;
  double s0 = U[0] * v + V[0] * u, s1 = U[1] * v + V[1] * u, s2 = U[2] * v + V[2] * u;
  return 2.0 * atan2(sqrt(d0 * d0 + d1 * d1 + d2 * d2), sqrt(s0 * s0 + s1 * s1 + s2 * s2));
}

__host__ __device__ inline void rotation_from_to(const double* a1, const double* a2, double* M) {
  const double PI = 3.141592653589793;
  double ang = kahan_angle(a1, a2);
  for (int i = 0; i < 9; ++i) M[i] = 0.0;
  if (fabs(ang) < 1e-10) {
    M[0] = M[4] = M[8] = 1.0;
    return;
  }
  if (fabs(ang - PI) < 1e-10) {  // the reference returns -Point here (ModuleGeometry.py:338-339)
    M[0] = M[4] = M[8] = -1.0;
    return;
  }
  double kx = a1[1] * a2[2] - a1[2] * a2[1];
  double ky = a1[2] * a2[0] - a1[0] * a2[2];
  double kz = a1[0] * a2[1] - a1[1] * a2[0];
  double kn = sqrt(kx * kx + ky * ky + kz * kz);
  kx /= kn; ky /= kn; kz /= kn;
  double c = cos(ang), s = sin(ang), oc = 1.0 - c;
  M[0] = c + oc * kx * kx;      M[1] = oc * kx * ky - s * kz; M[2] = oc * kx * kz + s * ky;
  M[3] = oc * ky * kx + s * kz; M[4] = c + oc * ky * ky;      M[5] = oc * ky * kz - s * kx;
  M[6] = oc * kz * kx - s * ky; M[7] = oc * kz * ky + s * kx; M[8] = c + oc * kz * kz;
}

// R = R2 R1: R1 = normal -> ez, mPrime = R1 majoraxis, R2 = mPrime -> ex (ModuleProcessing.py:290-294)
__host__ __device__ inline void element_rotation(const double* normal, const double* major, double* R) {
  const double ez[3] = {0, 0, 1}, ex[3] = {1, 0, 0};
  double R1[9], R2[9], mp[3];
  rotation_from_to(normal, ez, R1);
  for (int i = 0; i < 3; ++i) mp[i] = R1[3 * i] * major[0] + R1[3 * i + 1] * major[1] + R1[3 * i + 2] * major[2];
  rotation_from_to(mp, ex, R2);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      R[3 * i + j] = R2[3 * i] * R1[j] + R2[3 * i + 1] * R1[3 + j] + R2[3 * i + 2] * R1[6 + j];
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// supports: `_IncludeSupport`, ART/ModuleSupport.py:68,151,228,322,431 (all comparisons inclusive)
//   ROUND            ap = {R^2}
//   ROUND_HOLE       ap = {R^2, Rh^2, cx, cy}
//   RECT             ap = {|X/2|, |Y/2|}
//   RECT_HOLE        ap = {|X/2|, |Y/2|, Rh^2, cx, cy}
//   RECT_RECT_HOLE   ap = {|X/2|, |Y/2|, |hX/2|, |hY/2|, cx, cy}
// NaN coordinates compare false, as in numpy.
// ---------------------------------------------------------------------------------------------
ART_HD bool in_support(const ElemDev& E, double x, double y) {
  x -= E.soff[0];
  y -= E.soff[1];
  switch (E.support) {
    case ART_SUPP_ROUND:
      return x * x + y * y <= E.ap[0];
    case ART_SUPP_ROUND_HOLE: {
      double hx = x - E.ap[2], hy = y - E.ap[3];
      return (x * x + y * y <= E.ap[0]) && !(hx * hx + hy * hy <= E.ap[1]);
    }
    case ART_SUPP_RECT:
      return fabs(x) <= E.ap[0] && fabs(y) <= E.ap[1];
    case ART_SUPP_RECT_HOLE: {
      double hx = x - E.ap[3], hy = y - E.ap[4];
      return (fabs(x) <= E.ap[0] && fabs(y) <= E.ap[1]) && !(hx * hx + hy * hy <= E.ap[2]);
    }
    default: {  // ART_SUPP_RECT_RECT_HOLE
      double hx = x - E.ap[4], hy = y - E.ap[5];
      return (fabs(x) <= E.ap[0] && fabs(y) <= E.ap[1]) && !(fabs(hx) <= E.ap[2] && fabs(hy) <= E.ap[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// real roots of a t^2 + b t + c as np.roots would give them (ART/ModuleGeometry.py:80-91), in the
// cancellation-free form: q = -(b + sgn(b) sqrt(D))/2, t1 = q/a, t2 = c/q.
// a == 0 degrades to the single root -c/b (t1 becomes inf and fails every later test), D < 0 to none.
// ---------------------------------------------------------------------------------------------
ART_HD void solve_quadratic(double a, double b, double c, double& t1, double& t2) {
  double w = 4.0 * a * c;
  double e = fma(-4.0 * a, c, w);  // rounding error of w
  double f = fma(b, b, -w);
  double disc = f + e;
  if (!(disc >= 0.0)) {
    t1 = t2 = ART_NAN;
    return;
  }
  double q = -0.5 * (b + copysign(fsqrt(disc), b));
  t1 = fdiv(q, a);
  t2 = fdiv(c, q);
}

// Candidate rule shared by the quadrics: t > 1e-12 (KeepPositiveSolution, ModuleGeometry.py:110-120),
// surface-side test, support test; one candidate -> it, two -> the nearer one
// (_IntersectionRayMirror ART/ModuleMirror.py:27-38, ClosestPoint ModuleGeometry.py:138-147).
template <bool SIDE_Z_NEG>
ART_HD double pick_candidate(const ElemDev& E, const Ray& r, double t1, double t2, double zlim) {
  bool c1 = t1 > 1e-12, c2 = t2 > 1e-12;
  {
    double x = fma(t1, r.ux, r.px), y = fma(t1, r.uy, r.py), z = fma(t1, r.uz, r.pz);
    c1 = c1 && (!SIDE_Z_NEG || z < zlim) && in_support(E, x, y);
  }
  {
    double x = fma(t2, r.ux, r.px), y = fma(t2, r.uy, r.py), z = fma(t2, r.uz, r.pz);
    c2 = c2 && (!SIDE_Z_NEG || z < zlim) && in_support(E, x, y);
  }
  if (c1 && c2) return t1 < t2 ? t1 : t2;
  if (c1) return t1;
  if (c2) return t2;
  return ART_NAN;
}

// ---------------------------------------------------------------------------------------------
// Toroid, ART/ModuleMirror.py:443-478: (sqrt(x^2+z^2) - R)^2 + y^2 = r^2.
//
// The reference solves the expanded quartic with np.roots and keeps roots with t > 1e-12,
// z < -R and (x,y) on the support.  B200 path: every such point lies on the OUTER sheet
// (z < -R => rho > R), which is part of the boundary of the convex solid {dist(., disk of radius
// R in y=0) <= r}.  Along the ray,  F(t) = max(rho-R,0)^2 + y^2 - r^2  is therefore CONVEX with at
// most two zeros ta <= tb, and Newton's iteration on a convex function converges monotonically
// from outside the root interval.  tb is found from the right (start: hit with the tangent plane
// z = -(R+r), else a provably-right start), ta from t = 0 when the origin lies outside the solid.
// F keeps full relative accuracy near the surface (no s^2 - 4R^2 rho^2 cancellation).
// ---------------------------------------------------------------------------------------------
struct TorEval {
  double F, dF;
};
ART_HD TorEval tor_eval(const Ray& r, double t, double R, double r2) {
  double x = fma(t, r.ux, r.px), y = fma(t, r.uy, r.py), z = fma(t, r.uz, r.pz);
  double s = fma(x, x, z * z);
  double inv = frsqrt(s);
  double rho = s * inv;
  double q = rho - R;
  TorEval e;
  if (q > 0.0) {
    e.F = fma(q, q, fma(y, y, -r2));
    e.dF = 2.0 * fma(q * inv, fma(x, r.ux, z * r.uz), y * r.uy);
  } else {
    e.F = fma(y, y, -r2);
    e.dF = 2.0 * y * r.uy;
  }
  return e;
}

// 1/d to ~2^-46: MUFU.RCP64H seed (2^-23) + one Newton step.  Only used for Newton CORRECTIONS of the
// root search, which are self-correcting; the converged root does not depend on the step's last bits.
ART_HD double fast_rcp(double d) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  return fma(r, fma(-d, r, 1.0), r);
#else
  return 1.0 / d;
#endif
}

// Newton stops once the NEXT correction would be below rounding level: F is convex with
// F'' <= 2 |u|^2 = 2, so the step after dt is at most dt^2 / |F'|; at the noise floor of F
// (~ eps r^2) dt itself is ~1e-13 mm and the test holds as well.
// largest zero of F, Newton from the right of it (e = evaluation at t); NaN if the line misses the solid
ART_HD double tor_root_right(const Ray& r, double t, TorEval e, double R, double r2, double scale) {
  for (int it = 0; it < 64; ++it) {
    if (!(e.dF > 0.0)) return ART_NAN;  // walked past the minimum of F without meeting a zero
    const double dt = e.F * fast_rcp(e.dF);
    t -= dt;
    if (dt * dt <= 2e-16 * (fabs(t) + scale) * e.dF) return t;
    e = tor_eval(r, t, R, r2);
  }
  return ART_NAN;
}
// smallest zero of F, Newton from the left of it (start t = 0 with F > 0, dF < 0; e = evaluation at 0)
ART_HD double tor_root_left(const Ray& r, TorEval e, double R, double r2, double scale) {
  double t = 0.0;
  for (int it = 0; it < 64; ++it) {
    if (!(e.dF < 0.0)) return ART_NAN;
    const double dt = e.F * fast_rcp(e.dF);
    t -= dt;
    if (dt * dt <= 2e-16 * (fabs(t) + scale) * -e.dF) return t;
    e = tor_eval(r, t, R, r2);
  }
  return ART_NAN;
}

ART_HD double intersect_toroid(const ElemDev& E, const Ray& r) {
  const double R = E.sp[0], rr = E.sp[1], r2 = E.sp[2];
  // start for the right root: the tangent plane z = -(R+r) lies outside the solid
  double t0 = fdiv(-(R + rr) - r.pz, r.uz);
  TorEval e0 = tor_eval(r, t0, R, r2);
  if (!(t0 > 0.0 && e0.F >= 0.0 && e0.dF > 0.0 && t0 < 1e300)) {
    // beyond closest approach to the axis point by more than R + r the solid is behind us
    double tc = -(r.px * r.ux + r.py * r.uy + r.pz * r.uz);
    t0 = tc + 1.0009765625 * (R + rr);
    e0 = tor_eval(r, t0, R, r2);
  }
  double tb = tor_root_right(r, t0, e0, R, r2, rr);
  double ta = ART_NAN;
  TorEval o = tor_eval(r, 0.0, R, r2);
  if (o.F > 0.0) {
    if (!(o.dF < 0.0)) return ART_NAN;  // moving away from the solid: no forward root
    ta = tor_root_left(r, o, R, r2, rr);
  }
  return pick_candidate<true>(E, r, ta, tb, -R);
}

// ---------------------------------------------------------------------------------------------
// Zernike defect, ART/ModuleDefects.py:149-177.  The reference evaluates Andersen's Cartesian
// recurrences for ALL (n,m) up to max_order with Python lists.  Device path: the same polynomials
// written as  Z = Q_k^l(s) * {C_l, S_l}(x,y),  s = x^2+y^2,  C_l + i S_l = (x + i y)^l,
// Q_k^l(s) = R_{l+2k}^l(rho)/rho^l = (-1)^k P_k^{(l,0)}(1-2s)  (Jacobi), advanced in k by a
// three-term recurrence -> O(1) registers per ray for any order.  Table (host-built, smem):
//   zt[0] = radius R (= Support._CircumCirc()), zt[1] = max order N, then for l = 0..N,
//   k = 0..(N-l)/2: {alpha, beta, gamma, c_cos, c_sin} with
//   Q_k = (alpha s + beta) Q_{k-1} - gamma Q_{k-2};  c_cos / c_sin = coefficients of the reference
//   keys (n, (n+l)/2) / (n, (n-l)/2), n = l + 2k.
// Returns value and Cartesian gradient (already divided by R where the reference does).
// ---------------------------------------------------------------------------------------------
template <bool WANT_VALUE, bool WANT_GRAD>
ART_HD void zernike_eval(const double* __restrict__ zt, double X, double Y, double& val,
                                             double& gx, double& gy) {
  const double Rz = zt[0];
  const int N = (int)zt[1];
  const double iR = fdiv(1.0, Rz);
  const double x = X * iR, y = Y * iR;
  const double s = fma(x, x, y * y);
  const double* rec = zt + 2;
  double Cl = 1.0, Sl = 0.0, Cm = 0.0, Sm = 0.0;  // (x+iy)^l and (x+iy)^(l-1)
  double v = 0.0, dx = 0.0, dy = 0.0;
  for (int l = 0; l <= N; ++l) {
    const int K = (N - l) >> 1;
    double Q = 1.0, Qp = 0.0, dQ = 0.0, dQp = 0.0;
    double A = rec[3], B = rec[4], dA = 0.0, dB = 0.0;
    rec += 5;
    for (int k = 1; k <= K; ++k) {
      const double al = rec[0], be = rec[1], ga = rec[2], cc = rec[3], cs = rec[4];
      rec += 5;
      const double lin = fma(al, s, be);
      const double Qn = fma(lin, Q, -ga * Qp);
      if (WANT_GRAD) {
        const double dQn = fma(al, Q, fma(lin, dQ, -ga * dQp));
        dQp = dQ;
        dQ = dQn;
        dA = fma(cc, dQn, dA);
        dB = fma(cs, dQn, dB);
      }
      Qp = Q;
      Q = Qn;
      A = fma(cc, Qn, A);
      B = fma(cs, Qn, B);
    }
    if (WANT_VALUE) v = fma(A, Cl, fma(B, Sl, v));
    if (WANT_GRAD) {
      const double rad = fma(dA, Cl, dB * Sl);      // sum c dQ/ds * angular part
      const double fl = (double)l;
      dx = fma(2.0 * x, rad, fma(fl, fma(A, Cm, B * Sm), dx));
      dy = fma(2.0 * y, rad, fma(fl, fma(B, Cm, -A * Sm), dy));
    }
    Cm = Cl;
    Sm = Sl;
    const double Cn = fma(x, Cl, -y * Sl);
    Sl = fma(x, Sl, y * Cl);
    Cl = Cn;
  }
  val = v;
  gx = dx * iR;  // ModuleDefects.py:163-164
  gy = dy * iR;
}

// ---------------------------------------------------------------------------------------------
// surface normal, `get_normal` of each mirror class (unit vector)
// ---------------------------------------------------------------------------------------------
ART_HD void surface_normal(const ElemDev& E, double x, double y, double z, double& nx,
                                               double& ny, double& nz) {
  double gx, gy, gz;
  switch (E.surface) {
    case ART_SURF_SPHERICAL:  // ART/ModuleMirror.py:180-183
      gx = -x; gy = -y; gz = -z;
      break;
    case ART_SURF_PARABOLIC:  // :349-355
      gx = -x; gy = -y; gz = E.sp[0];
      break;
    case ART_SURF_TOROIDAL: {  // :480-498 (common factor 4 dropped)
      const double S = fma(x, x, fma(y, y, z * z));
      const double a = S + E.sp[3];        // + (R^2 - r^2)
      const double b = a - 2.0 * E.sp[4];  // - 2 R^2
      gx = -x * b; gy = -y * a; gz = -z * b;
      break;
    }
    case ART_SURF_ELLIPSOIDAL:  // :685-693
      gx = -x * E.sp[2]; gy = -y * E.sp[3]; gz = -z * E.sp[3];
      break;
    case ART_SURF_CYLINDRICAL:  // :846-849
      gx = 0.0; gy = -y; gz = -z;
      break;
    default:  // plane, mask: :84-87
      nx = 0.0; ny = 0.0; nz = 1.0;
      return;
  }
  const double inv = frsqrt(fma(gx, gx, fma(gy, gy, gz * gz)));
  nx = gx * inv; ny = gy * inv; nz = gz * inv;
}

// atan2(y, x) for y >= 0 (result in [0, pi]), branch-free: one division for the reduced argument
// z (|z| <= tan(pi/8)) and the Maclaurin series of atan to z^41 (truncation < 2e-18 z).
ART_HD double fatan2_ypos(double y, double x) {
  const double ax = fabs(x);
  const bool swap = y > ax;
  const double lo = swap ? ax : y, hi = swap ? y : ax;          // lo/hi in [0, 1]
  const bool big = lo > 0.41421356237309503 * hi;                // beyond tan(pi/8): rotate by pi/4
  const double z = fdiv(big ? lo - hi : lo, big ? lo + hi : hi);
  const double w = z * z;
  double p = -1.0 / 41.0;
  p = fma(p, w, 1.0 / 39.0);  p = fma(p, w, -1.0 / 37.0); p = fma(p, w, 1.0 / 35.0);  p = fma(p, w, -1.0 / 33.0);
  p = fma(p, w, 1.0 / 31.0);  p = fma(p, w, -1.0 / 29.0); p = fma(p, w, 1.0 / 27.0);  p = fma(p, w, -1.0 / 25.0);
  p = fma(p, w, 1.0 / 23.0);  p = fma(p, w, -1.0 / 21.0); p = fma(p, w, 1.0 / 19.0);  p = fma(p, w, -1.0 / 17.0);
  p = fma(p, w, 1.0 / 15.0);  p = fma(p, w, -1.0 / 13.0); p = fma(p, w, 1.0 / 11.0);  p = fma(p, w, -1.0 / 9.0);
  p = fma(p, w, 1.0 / 7.0);   p = fma(p, w, -1.0 / 5.0);  p = fma(p, w, 1.0 / 3.0);
  double a = fma(-z * w, p, z);                                  // atan(z)
  a += big ? 0.78539816339744831 : 0.0;
  a = swap ? 1.5707963267948966 - a : a;
  return x < 0.0 ? 3.1415926535897931 - a : a;
}

// Angle between two UNIT vectors a, b.  The reference uses Kahan's 2 atan2(|a-b|, |a+b|)
// (ART/ModuleGeometry.py:40-44); atan2(|a x b|, a.b) is the same angle, equally well conditioned over
// [0, pi], and needs one square root instead of two.
ART_HD double unit_angle(double ax, double ay, double az, double bx, double by, double bz) {
  const double cx = fma(ay, bz, -az * by), cy = fma(az, bx, -ax * bz), cz = fma(ax, by, -ay * bx);
  return fatan2_ypos(fsqrt(fma(cx, cx, fma(cy, cy, cz * cz))), fma(ax, bx, fma(ay, by, az * bz)));
}

// ---------------------------------------------------------------------------------------------
// one element acting on one ray: ART/ModuleProcessing.py:284-309 (frame in, optic, frame out)
// ---------------------------------------------------------------------------------------------
template <bool WANT_INC, bool HAS_DEF = true>
ART_HD void apply_element(const ElemDev& E, Ray& r, const double* __restrict__ ztab,
                                              const int* __restrict__ zoff, bool ignore_defects,
                                              bool inc_here = true) {
  // lab -> element frame (:289-295): p_e = R (p - pos) + centre, u_e = R u
  Ray e;
  {
    const double dx = r.px - E.pos[0], dy = r.py - E.pos[1], dz = r.pz - E.pos[2];
    e.px = fma(E.rot[0], dx, fma(E.rot[1], dy, fma(E.rot[2], dz, E.ctr[0])));
    e.py = fma(E.rot[3], dx, fma(E.rot[4], dy, fma(E.rot[5], dz, E.ctr[1])));
    e.pz = fma(E.rot[6], dx, fma(E.rot[7], dy, fma(E.rot[8], dz, E.ctr[2])));
    e.ux = fma(E.rot[0], r.ux, fma(E.rot[1], r.uy, E.rot[2] * r.uz));
    e.uy = fma(E.rot[3], r.ux, fma(E.rot[4], r.uy, E.rot[5] * r.uz));
    e.uz = fma(E.rot[6], r.ux, fma(E.rot[7], r.uy, E.rot[8] * r.uz));
  }
  double t;
  switch (E.surface) {
    case ART_SURF_PLANE: {  // ART/ModuleMirror.py:73-82: t > 0 (no epsilon) and on the support
      t = fdiv(-e.pz, e.uz);
      const double x = fma(t, e.ux, e.px), y = fma(t, e.uy, e.py);
      if (!(t > 0.0 && in_support(E, x, y))) t = ART_NAN;
      break;
    }
    case ART_SURF_MASK: {  // ART/ModuleMask.py:51-61: passes iff t > 0 and NOT on the support
      t = fdiv(-e.pz, e.uz);
      const double x = fma(t, e.ux, e.px), y = fma(t, e.uy, e.py);
      if (!(t > 0.0 && !in_support(E, x, y))) t = ART_NAN;
      break;
    }
    case ART_SURF_SPHERICAL: {  // :163-178
      const double a = fma(e.ux, e.ux, fma(e.uy, e.uy, e.uz * e.uz));
      const double b = 2.0 * fma(e.ux, e.px, fma(e.uy, e.py, e.uz * e.pz));
      const double c = fma(e.px, e.px, fma(e.py, e.py, fma(e.pz, e.pz, -E.sp[1])));
      double t1, t2;
      solve_quadratic(a, b, c, t1, t2);
      t = pick_candidate<true>(E, e, t1, t2, 0.0);
      break;
    }
    case ART_SURF_PARABOLIC: {  // :325-347 (no z test)
      const double p = E.sp[0];
      const double a = fma(e.ux, e.ux, e.uy * e.uy);
      const double b = 2.0 * fma(e.ux, e.px, fma(e.uy, e.py, -p * e.uz));
      const double c = fma(e.px, e.px, fma(e.py, e.py, -2.0 * p * e.pz));
      double t1, t2;
      solve_quadratic(a, b, c, t1, t2);
      t = pick_candidate<false>(E, e, t1, t2, 0.0);
      break;
    }
    case ART_SURF_TOROIDAL:
      t = intersect_toroid(E, e);
      break;
    case ART_SURF_ELLIPSOIDAL: {  // :662-683, sp[2] = 1/a^2, sp[3] = 1/b^2
      const double ia = E.sp[2], ib = E.sp[3];
      const double a = fma(fma(e.uy, e.uy, e.uz * e.uz), ib, e.ux * e.ux * ia);
      const double b = 2.0 * fma(fma(e.uy, e.py, e.uz * e.pz), ib, e.ux * e.px * ia);
      const double c = fma(fma(e.py, e.py, e.pz * e.pz), ib, fma(e.px * e.px, ia, -1.0));
      double t1, t2;
      solve_quadratic(a, b, c, t1, t2);
      t = pick_candidate<true>(E, e, t1, t2, 0.0);
      break;
    }
    default: {  // ART_SURF_CYLINDRICAL :824-844
      const double a = fma(e.uy, e.uy, e.uz * e.uz);
      const double b = 2.0 * fma(e.uy, e.py, e.uz * e.pz);
      const double c = fma(e.py, e.py, fma(e.pz, e.pz, -E.sp[1]));
      double t1, t2;
      solve_quadratic(a, b, c, t1, t2);
      t = pick_candidate<true>(E, e, t1, t2, 0.0);
      break;
    }
  }
  if (!(t == t)) {  // miss: the reference drops the ray (ModuleMirror.py:932, ModuleMask.py:132)
    r.alive = false;
    return;
  }
  double hx = fma(t, e.ux, e.px), hy = fma(t, e.uy, e.py), hz = fma(t, e.uz, e.pz);
  double ox = e.ux, oy = e.uy, oz = e.uz;  // outgoing direction, element frame
  if (E.surface == ART_SURF_MASK) {
    // _TransmitMaskRay, ART/ModuleMask.py:93-108: direction unchanged, incidence vs ez
    if (WANT_INC && inc_here) r.inc = unit_angle(e.ux, e.uy, e.uz, 0.0, 0.0, 1.0);
  } else {
    double nx, ny, nz;
    surface_normal(E, hx, hy, hz, nx, ny, nz);
    if (HAS_DEF && E.n_defects > 0) {
      // DeformedMirror._get_intersection, ART/ModuleMirror.py:969-980:
      //   h = sum offsets(P - C); alpha = angle(-u, n_base(P)); P -= u h / cos(alpha)
      double h = 0.0;
      for (int d = 0; d < E.n_defects; ++d) {
        double v, g0, g1;
        zernike_eval<true, false>(ztab + zoff[E.first_defect + d], hx - E.ctr[0], hy - E.ctr[1], v, g0, g1);
        h += v;
      }
      const double cosa = -(nx * e.ux + ny * e.uy + nz * e.uz);
      const double sh = fdiv(h, cosa);
      t -= sh;
      hx = fma(-sh, e.ux, hx); hy = fma(-sh, e.uy, hy); hz = fma(-sh, e.uz, hz);
      surface_normal(E, hx, hy, hz, nx, ny, nz);  // the reflection uses get_normal(shifted point)
      if (!ignore_defects) {
        // DeformedMirror.get_normal + normal_add, ART/ModuleMirror.py:952-961, ModuleGeometry.py:394-407:
        // slopes add; the result is (-gx, -gy, 1) normalised
        const double inz = fdiv(1.0, nz);
        double sx = -nx * inz, sy = -ny * inz;
        for (int d = 0; d < E.n_defects; ++d) {
          double v, g0, g1;
          zernike_eval<false, true>(ztab + zoff[E.first_defect + d], hx - E.ctr[0], hy - E.ctr[1], v, g0, g1);
          sx += g0;
          sy += g1;
        }
        const double inv = frsqrt(fma(sx, sx, fma(sy, sy, 1.0)));
        nx = -sx * inv; ny = -sy * inv; nz = inv;
      }
    }
    // _ReflectionMirrorRay, ART/ModuleMirror.py:878-906: u' = u - 2 (n.u) n, incidence = angle(-u, n)
    const double d = fma(nx, e.ux, fma(ny, e.uy, nz * e.uz));
    ox = fma(-2.0 * d, nx, e.ux); oy = fma(-2.0 * d, ny, e.uy); oz = fma(-2.0 * d, nz, e.uz);
    // Ray.vector setter renormalises (ART/ModuleOpticalRay.py:85-90); one Newton step is exact here
    const double sc = fma(-0.5, fma(ox, ox, fma(oy, oy, oz * oz)), 1.5);
    ox *= sc; oy *= sc; oz *= sc;
    if (WANT_INC && inc_here) r.inc = unit_angle(-e.ux, -e.uy, -e.uz, nx, ny, nz);
  }
  r.path += fabs(t);  // |P - A| with |u| = 1 (ModuleMirror.py:904, ModuleMask.py:100)
  // element -> lab frame (:306-309): p = R^T (p_e - centre) + pos, u = R^T u_e
  {
    const double dx = hx - E.ctr[0], dy = hy - E.ctr[1], dz = hz - E.ctr[2];
    r.px = fma(E.rot[0], dx, fma(E.rot[3], dy, fma(E.rot[6], dz, E.pos[0])));
    r.py = fma(E.rot[1], dx, fma(E.rot[4], dy, fma(E.rot[7], dz, E.pos[1])));
    r.pz = fma(E.rot[2], dx, fma(E.rot[5], dy, fma(E.rot[8], dz, E.pos[2])));
    r.ux = fma(E.rot[0], ox, fma(E.rot[3], oy, E.rot[6] * oz));
    r.uy = fma(E.rot[1], ox, fma(E.rot[4], oy, E.rot[7] * oz));
    r.uz = fma(E.rot[2], ox, fma(E.rot[5], oy, E.rot[8] * oz));
  }
}
#endif  // __CUDACC__

}  // namespace art
