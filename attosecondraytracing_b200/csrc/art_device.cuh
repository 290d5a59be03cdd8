// art_device.cuh -- device-side data layout, branch-free FP64 primitives and the pose -> rotation
// arithmetic of libart_b200.  The per-ray optics are in art_optics.cuh (included at the end).
// The reference semantics each function reproduces are cited as file:line of the reference
// repository (ART v0.93).
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
  double nrot[9];  // this element's frame -> the NEXT element's frame: p' = nrot h + noff, u' = nrot u
  double noff[3];  // (link_elements, art_lowering.h; unused for the last element of a chain)
  int32_t surface;
  int32_t support;
  int32_t n_defects;
  int32_t first_defect;
  int32_t n_maps;      // gridded defects
  int32_t first_map;
};

// one gridded defect (ArtGridMapDesc, pre-digested): cell index = (x - x0) * sx
struct MapDev {
  const double *h, *dx, *dy;
  double x0, sx, y0, sy;
  int32_t nx, ny;
};

struct BundleDev {
  double *px, *py, *pz, *ux, *uy, *uz, *path, *inc, *inten;
  uint8_t* alive;
};

ART_HD double sq(double x) { return x * x; }

// ---------------------------------------------------------------------------------------------
// Branch-free FP64 division / square root / reciprocal square root for the device: an MUFU seed
// (rcp.approx / rsqrt.approx, ~2^-23) refined by Newton steps in FMAs and a final residual
// correction -> results within 1 ulp for normal-range operands, with none of the slow-path calls,
// predicate fix-ups and register shuffles of the IEEE-exact library sequences.  Quotient and square
// root need ONE Newton step on the seed (2^-46): the residual correction squares that error again
// (emulated with exact rational FMAs over 2e4 random operands: both stay correctly rounded).  Operands here
// are lengths, direction cosines and their products (1e-300 < |x| < 1e300).  A zero divisor gives NaN
// instead of +-inf; every caller treats both as "no hit" (comparisons with NaN are false).
// On the host (tests/hostcheck) the plain operators are used.
// ---------------------------------------------------------------------------------------------
ART_HD double fdiv(double a, double b) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
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
  // x + 1e-300 is x itself for every x >= 1e-284 and keeps the seed finite at x = 0 (0 * rsqrt(0) would be
  // NaN); an addition instead of a compare and two selects.  Negative x -> NaN, as sqrt().
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x + 1e-300));
  y = y * fma(-0.5 * x * y, y, 1.5);
  const double s = x * y;
  return fma(fma(-s, s, x), 0.5 * y, s);
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

}  // namespace art

#ifdef __CUDACC__
#include "art_optics.cuh"
#endif
