// art_optics.cuh -- the per-ray optics of libart_b200, written once for a "lane pack" type T:
//   T = double : one ray per thread (also what tests/hostcheck runs on the host)
//   T = D2     : two rays per thread advanced in lock-step.  Every per-ray decision is a select
//                on a lane mask, loops run until both lanes are done (finished lanes are frozen,
//                so a ray's result never depends on its partner), and the two independent
//                dependency chains give the FP64 pipe twice the instruction-level parallelism.
// Only warp-uniform quantities (surface / support kind, Zernike order) steer real branches.
// The reference semantics each function reproduces are cited as file:line of the reference
// repository (ART v0.93).
#pragma once

namespace art {

// ---------------------------------------------------------------------------------------------
// lane packs
// ---------------------------------------------------------------------------------------------
struct D2 {
  double a, b;
};
struct B2 {
  bool a, b;
};
template <class T>
struct MaskOf {
  typedef bool type;
};
template <>
struct MaskOf<D2> {
  typedef B2 type;
};

ART_HD D2 operator+(D2 x, D2 y) { return {x.a + y.a, x.b + y.b}; }
ART_HD D2 operator-(D2 x, D2 y) { return {x.a - y.a, x.b - y.b}; }
ART_HD D2 operator*(D2 x, D2 y) { return {x.a * y.a, x.b * y.b}; }
ART_HD D2 operator+(D2 x, double y) { return {x.a + y, x.b + y}; }
ART_HD D2 operator-(D2 x, double y) { return {x.a - y, x.b - y}; }
ART_HD D2 operator*(D2 x, double y) { return {x.a * y, x.b * y}; }
ART_HD D2 operator+(double x, D2 y) { return {x + y.a, x + y.b}; }
ART_HD D2 operator-(double x, D2 y) { return {x - y.a, x - y.b}; }
ART_HD D2 operator*(double x, D2 y) { return {x * y.a, x * y.b}; }
ART_HD D2 operator-(D2 x) { return {-x.a, -x.b}; }
ART_HD B2 operator<(D2 x, D2 y) { return {x.a < y.a, x.b < y.b}; }
ART_HD B2 operator>(D2 x, D2 y) { return {x.a > y.a, x.b > y.b}; }
ART_HD B2 operator<=(D2 x, D2 y) { return {x.a <= y.a, x.b <= y.b}; }
ART_HD B2 operator>=(D2 x, D2 y) { return {x.a >= y.a, x.b >= y.b}; }
ART_HD B2 operator==(D2 x, D2 y) { return {x.a == y.a, x.b == y.b}; }
ART_HD B2 operator<(D2 x, double y) { return {x.a < y, x.b < y}; }
ART_HD B2 operator>(D2 x, double y) { return {x.a > y, x.b > y}; }
ART_HD B2 operator<=(D2 x, double y) { return {x.a <= y, x.b <= y}; }
ART_HD B2 operator>=(D2 x, double y) { return {x.a >= y, x.b >= y}; }
ART_HD B2 operator&(B2 x, B2 y) { return {x.a && y.a, x.b && y.b}; }
ART_HD B2 operator|(B2 x, B2 y) { return {x.a || y.a, x.b || y.b}; }
ART_HD B2 operator!(B2 x) { return {!x.a, !x.b}; }

// the same vocabulary for one lane
ART_HD double mfma(double x, double y, double z) { return fma(x, y, z); }
ART_HD D2 mfma(D2 x, D2 y, D2 z) { return {fma(x.a, y.a, z.a), fma(x.b, y.b, z.b)}; }
ART_HD D2 mfma(double x, D2 y, D2 z) { return {fma(x, y.a, z.a), fma(x, y.b, z.b)}; }
ART_HD D2 mfma(D2 x, double y, D2 z) { return {fma(x.a, y, z.a), fma(x.b, y, z.b)}; }
ART_HD D2 mfma(D2 x, D2 y, double z) { return {fma(x.a, y.a, z), fma(x.b, y.b, z)}; }
ART_HD D2 mfma(double x, D2 y, double z) { return {fma(x, y.a, z), fma(x, y.b, z)}; }
ART_HD D2 mfma(D2 x, double y, double z) { return {fma(x.a, y, z), fma(x.b, y, z)}; }
ART_HD double sel(bool m, double x, double y) { return m ? x : y; }
ART_HD D2 sel(B2 m, D2 x, D2 y) { return {m.a ? x.a : y.a, m.b ? x.b : y.b}; }
ART_HD D2 sel(B2 m, D2 x, double y) { return {m.a ? x.a : y, m.b ? x.b : y}; }
ART_HD D2 sel(B2 m, double x, D2 y) { return {m.a ? x : y.a, m.b ? x : y.b}; }
ART_HD D2 sel(B2 m, double x, double y) { return {m.a ? x : y, m.b ? x : y}; }
ART_HD bool any(bool m) { return m; }
ART_HD bool any(B2 m) { return m.a || m.b; }
ART_HD bool mand(bool x, bool y) { return x && y; }
ART_HD B2 mand(B2 x, B2 y) { return x & y; }
ART_HD B2 mand(B2 x, bool y) { return {x.a && y, x.b && y}; }
ART_HD bool mor(bool x, bool y) { return x || y; }
ART_HD B2 mor(B2 x, B2 y) { return x | y; }
ART_HD bool mnot(bool x) { return !x; }
ART_HD B2 mnot(B2 x) { return !x; }
ART_HD double mabs(double x) { return fabs(x); }
ART_HD D2 mabs(D2 x) { return {fabs(x.a), fabs(x.b)}; }
ART_HD double mcopysign(double x, double y) { return copysign(x, y); }
ART_HD D2 mcopysign(D2 x, D2 y) { return {copysign(x.a, y.a), copysign(x.b, y.b)}; }
ART_HD D2 fdiv(D2 x, D2 y) { return {fdiv(x.a, y.a), fdiv(x.b, y.b)}; }
ART_HD D2 fdiv(double x, D2 y) { return {fdiv(x, y.a), fdiv(x, y.b)}; }
ART_HD D2 frsqrt(D2 x) { return {frsqrt(x.a), frsqrt(x.b)}; }
ART_HD D2 fsqrt(D2 x) { return {fsqrt(x.a), fsqrt(x.b)}; }
template <class T>
ART_HD T splat(double x);
template <>
ART_HD double splat<double>(double x) { return x; }
template <>
ART_HD D2 splat<D2>(double x) { return {x, x}; }
template <class T>
ART_HD typename MaskOf<T>::type splat_mask(bool x);
template <>
ART_HD bool splat_mask<double>(bool x) { return x; }
template <>
ART_HD B2 splat_mask<D2>(bool x) { return {x, x}; }

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
ART_HD D2 fast_rcp(D2 d) { return {fast_rcp(d.a), fast_rcp(d.b)}; }

// A ray (or a pair of rays) in registers.
template <class T>
struct RayT {
  T px, py, pz, ux, uy, uz, path, inc;
  typename MaskOf<T>::type alive;
};
typedef RayT<double> Ray;

ART_HD RayT<D2> pack_rays(const Ray& p, const Ray& q) {
  RayT<D2> r;
  r.px = {p.px, q.px}; r.py = {p.py, q.py}; r.pz = {p.pz, q.pz};
  r.ux = {p.ux, q.ux}; r.uy = {p.uy, q.uy}; r.uz = {p.uz, q.uz};
  r.path = {p.path, q.path}; r.inc = {p.inc, q.inc};
  r.alive = {p.alive, q.alive};
  return r;
}
ART_HD void unpack_rays(const RayT<D2>& r, Ray& p, Ray& q) {
  p.px = r.px.a; p.py = r.py.a; p.pz = r.pz.a; p.ux = r.ux.a; p.uy = r.uy.a; p.uz = r.uz.a;
  p.path = r.path.a; p.inc = r.inc.a; p.alive = r.alive.a;
  q.px = r.px.b; q.py = r.py.b; q.pz = r.pz.b; q.ux = r.ux.b; q.uy = r.uy.b; q.uz = r.uz.b;
  q.path = r.path.b; q.inc = r.inc.b; q.alive = r.alive.b;
}

// ---------------------------------------------------------------------------------------------
// supports: `_IncludeSupport`, ART/ModuleSupport.py:68,151,228,322,431 (all comparisons inclusive)
//   ROUND            ap = {R^2}
//   ROUND_HOLE       ap = {R^2, Rh^2, cx, cy}
//   RECT             ap = {|X/2|, |Y/2|}
//   RECT_HOLE        ap = {|X/2|, |Y/2|, Rh^2, cx, cy}
//   RECT_RECT_HOLE   ap = {|X/2|, |Y/2|, |hX/2|, |hY/2|, cx, cy}
// NaN coordinates compare false, as in numpy.
// ---------------------------------------------------------------------------------------------
#ifndef ART_SKIP_SOFF
#define ART_SKIP_SOFF 1
#endif
// OFFSET = false: the caller knows soff = 0 (every surface but the parabola and the ellipsoid, art_lowering.h)
template <bool OFFSET = true, class T>
ART_HD typename MaskOf<T>::type in_support(const ElemDev& E, T x, T y) {
  if (OFFSET || !ART_SKIP_SOFF) {
    x = x - E.soff[0];
    y = y - E.soff[1];
  }
  // Two binary, warp-uniform decisions (outline round / rectangular, hole or not) instead of a five-way switch: the
  // compiler turns a switch -- and an if-chain over the kind -- into a jump table that costs seven instructions per call.
  typedef typename MaskOf<T>::type M;
  const int kind = E.support;
  M in;
  if (kind >= ART_SUPP_RECT) in = mand(mabs(x) <= E.ap[0], mabs(y) <= E.ap[1]);
  else in = mfma(x, x, y * y) <= E.ap[0];
  if ((kind & 1) == 0 && kind != ART_SUPP_RECT_RECT_HOLE) return in;   // SupportRound, SupportRectangle
  if (kind == ART_SUPP_RECT_RECT_HOLE) {
    const T hx = x - E.ap[4], hy = y - E.ap[5];
    return mand(in, mnot(mand(mabs(hx) <= E.ap[2], mabs(hy) <= E.ap[3])));
  }
  // round hole: ap = {.., Rh^2, cx, cy} from index 1 (SupportRoundHole) or 2 (SupportRectangleHole)
  const double* h = E.ap + (kind == ART_SUPP_ROUND_HOLE ? 1 : 2);
  const T hx = x - h[1], hy = y - h[2];
  return mand(in, mnot(mfma(hx, hx, hy * hy) <= h[0]));
}

// ---------------------------------------------------------------------------------------------
// real roots of a t^2 + b t + c as np.roots would give them (ART/ModuleGeometry.py:80-91), in the
// cancellation-free form: q = -(b + sgn(b) sqrt(D))/2, t1 = q/a, t2 = c/q.
// a == 0 degrades to the single root -c/b (t1 is not finite and fails every later test), D < 0 to
// none (both NaN).
// ---------------------------------------------------------------------------------------------
template <class T>
ART_HD void solve_quadratic(T a, T b, T c, T& t1, T& t2) {
  const T w = 4.0 * a * c;
  const T e = mfma(-4.0 * a, c, w);  // rounding error of w
  const T f = mfma(b, b, -w);
  const T disc = f + e;
  // disc < 0: the square root is NaN and so are q, t1 and t2 -- no masks needed
  const T q = -0.5 * (b + mcopysign(fsqrt(disc), b));
  t1 = fdiv(q, a);
  t2 = fdiv(c, q);
}

// Candidate rule shared by the curved mirrors: t > 1e-12 (KeepPositiveSolution,
// ModuleGeometry.py:110-120), surface-side test, support test; one candidate -> it, two -> the
// nearer one (_IntersectionRayMirror ART/ModuleMirror.py:27-38, ClosestPoint ModuleGeometry.py:138-147).
// A root that is not positive for either lane (the far root of a ray that starts inside a sphere, ...)
// skips its hit-point evaluation altogether.
template <bool SIDE_Z_NEG, class T>
ART_HD T pick_candidate(const ElemDev& E, const RayT<T>& r, T t1, T t2, double zlim, typename MaskOf<T>::type& valid) {
  typedef typename MaskOf<T>::type M;
  M c1 = t1 > 1e-12, c2 = t2 > 1e-12;
  if (any(c1)) {
    const T x = mfma(t1, r.ux, r.px), y = mfma(t1, r.uy, r.py), z = mfma(t1, r.uz, r.pz);
    if (SIDE_Z_NEG) c1 = mand(c1, z < zlim);
    c1 = mand(c1, in_support(E, x, y));
  }
  if (any(c2)) {
    const T x = mfma(t2, r.ux, r.px), y = mfma(t2, r.uy, r.py), z = mfma(t2, r.uz, r.pz);
    if (SIDE_Z_NEG) c2 = mand(c2, z < zlim);
    c2 = mand(c2, in_support(E, x, y));
  }
  // t2 wins when it is a candidate and either t1 is none or t2 is nearer
  const M use2 = mand(c2, mor(mnot(c1), t2 < t1));
  valid = mor(use2, c1);   // the hit distance of a lane without a candidate is meaningless (the caller masks it)
  return sel(use2, t2, t1);
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
template <class T>
struct TorEval {
  T F, dF;
};
template <class T>
ART_HD TorEval<T> tor_eval(const RayT<T>& r, T t, double R, double r2) {
  const T x = mfma(t, r.ux, r.px), y = mfma(t, r.uy, r.py), z = mfma(t, r.uz, r.pz);
  const T s = mfma(x, x, z * z);
  const T inv = frsqrt(s);
  const T rho = s * inv;
  const T q = rho - R;
  // max(q, 0) as (q + |q|) / 2 (exact): a select would be matched to fmax and its NaN-quieting
  // sequence costs 8 instructions per lane
  const T qq = 0.5 * (q + mabs(q));
  TorEval<T> e;
  e.F = mfma(qq, qq, mfma(y, y, -r2));
  e.dF = 2.0 * mfma(qq * inv, mfma(x, r.ux, z * r.uz), y * r.uy);
  return e;
}

// Newton from outside the root interval of the convex F.  DIR = +1: largest zero, approached from
// the right (needs F' > 0 on the way); DIR = -1: smallest zero from the left (F' < 0).
// A lane stops once the NEXT correction would be below rounding level: F'' <= 2 |u|^2 = 2, so the
// step after dt is at most dt^2 / |F'|; at the noise floor of F (~ eps r^2) dt itself is ~1e-13 mm
// and the test holds as well.  Lanes that are done are frozen; a lane that walks past the minimum of
// F without meeting a zero (the line misses the solid) or needs more than 64 steps yields NaN.
// PRE leading steps are taken without the convergence test (the start is known to need them: from
// the tangent plane a ray needs three evaluations); only their slope signs are accumulated.  They are
// taken by every lane alike, so a ray's result still never depends on its partner.
#ifndef ART_PEEL_NEWTON
#define ART_PEEL_NEWTON 1   // first checked step outside the loop: the common case converges on it
#endif
template <int DIR, class T>
ART_HD void tor_newton_step(const TorEval<T>& e, T& t, typename MaskOf<T>::type& done, typename MaskOf<T>::type& todo,
                            double scale) {
  typedef typename MaskOf<T>::type M;
  const T slope = DIR > 0 ? e.dF : -e.dF;
  const M good = slope > 0.0;
  const T dt = e.F * fast_rcp(e.dF);
  const T tn = t - dt;
  const M conv = mand(good, dt * dt <= 2e-16 * (mabs(tn) + scale) * slope);
  t = sel(todo, tn, t);   // every lane still iterating moves; one that converges with this step keeps tn for good
  done = mor(done, mand(todo, conv));
  todo = mand(todo, mand(good, mnot(conv)));
}
// Returns t; `done` = the lanes whose iteration converged (the value of any other lane is meaningless).
template <int DIR, int PRE, class T>
ART_HD T tor_newton(const RayT<T>& r, T t, TorEval<T> e, double R, double r2, double scale,
                    typename MaskOf<T>::type active, typename MaskOf<T>::type& done) {
  typedef typename MaskOf<T>::type M;
  done = splat_mask<T>(false);
  M todo = active;
#pragma unroll
  for (int p = 0; p < PRE; ++p) {
    todo = mand(todo, (DIR > 0 ? e.dF : -e.dF) > 0.0);
    t = t - e.F * fast_rcp(e.dF);
    e = tor_eval(r, t, R, r2);
  }
#if ART_PEEL_NEWTON
  tor_newton_step<DIR>(e, t, done, todo, scale);
  for (int it = 1; it < 64 && any(todo); ++it) {
    e = tor_eval(r, t, R, r2);
    tor_newton_step<DIR>(e, t, done, todo, scale);
  }
#else
  for (int it = 0; it < 64 && any(todo); ++it) {
    tor_newton_step<DIR>(e, t, done, todo, scale);
    if (any(todo)) e = tor_eval(r, t, R, r2);
  }
#endif
  return t;
}

#ifndef ART_NEWTON_PRE
#define ART_NEWTON_PRE 2
#endif
#ifndef ART_CHEAP_INSIDE
#define ART_CHEAP_INSIDE 1
#endif
#ifndef ART_QUARTIC_PRE
#define ART_QUARTIC_PRE 1
#endif
#ifndef ART_GRADIENT_REFLECT
#define ART_GRADIENT_REFLECT 1
#endif
#ifndef ART_T0_RCP
#define ART_T0_RCP 1
#endif
#ifndef ART_QUARTIC_STEPS
#define ART_QUARTIC_STEPS 2   // leading unchecked Newton steps taken on the quartic form
#endif
// The expanded quartic of the reference (ART/ModuleMirror.py:450-465) as P(t) = A^2 - 4 R^2 s with
// s = x^2 + z^2, A = s + y^2 + R^2 - r^2:  P = F * (A + 2 R rho), the second factor positive and nearly
// constant, so Newton on P takes (almost) the steps of Newton on F -- without the reciprocal square root
// (15 FP64 instructions and a dependency depth of 5 instead of 22 + MUFU and 15).  Its cancellation noise
// (~1e-10 mm in t) is irrelevant for the LEADING steps; the final, checked steps use F.
template <class T>
ART_HD TorEval<T> tor_eval_quartic(const RayT<T>& r, T t, double c, double R2) {
  const T x = mfma(t, r.ux, r.px), y = mfma(t, r.uy, r.py), z = mfma(t, r.uz, r.pz);
  const T s = mfma(x, x, z * z);
  const T hs = mfma(x, r.ux, z * r.uz);  // s' / 2
  const T A = mfma(y, y, s) + c;
  const T hA = mfma(y, r.uy, hs);        // A' / 2
  TorEval<T> e;
  e.F = mfma(A, A, -((4.0 * R2) * s));
  e.dF = 4.0 * mfma(A, hA, -((2.0 * R2) * hs));
  return e;
}

template <class T>
ART_HD T intersect_toroid(const ElemDev& E, const RayT<T>& r, typename MaskOf<T>::type act,
                          typename MaskOf<T>::type& valid) {
  typedef typename MaskOf<T>::type M;
  const double R = E.sp[0], rr = E.sp[1], r2 = E.sp[2];
  // start for the right root: the tangent plane z = -(R+r) lies outside the solid
#if ART_T0_RCP
  T t0 = (-(R + rr) - r.pz) * fast_rcp(r.uz);  // a START value: 2^-46 is plenty (uz = 0 -> NaN -> the fallback start)
#else
  T t0 = fdiv(-(R + rr) - r.pz, r.uz);
#endif
#if ART_QUARTIC_PRE && ART_NEWTON_PRE == 2
  TorEval<T> e0 = tor_eval_quartic(r, t0, E.sp[3], E.sp[4]);
#else
  TorEval<T> e0 = tor_eval(r, t0, R, r2);
#endif
  const M fine = mand(mand(t0 > 0.0, e0.F >= 0.0), mand(e0.dF > 0.0, t0 < 1e300));
  if (any(mand(act, mnot(fine)))) {
    // beyond closest approach to the axis point by more than R + r the solid is behind us
    const T tc = -mfma(r.px, r.ux, mfma(r.py, r.uy, r.pz * r.uz));
    t0 = sel(fine, t0, tc + 1.0009765625 * (R + rr));
#if ART_QUARTIC_PRE && ART_NEWTON_PRE == 2
    const TorEval<T> e1 = tor_eval_quartic(r, t0, E.sp[3], E.sp[4]);
#else
    const TorEval<T> e1 = tor_eval(r, t0, R, r2);
#endif
    e0.F = sel(fine, e0.F, e1.F);
    e0.dF = sel(fine, e0.dF, e1.dF);
  }
#if ART_QUARTIC_PRE && ART_NEWTON_PRE == 2
  // ART_QUARTIC_STEPS leading Newton steps on the quartic, then the checked iteration on F
  M ok = mand(act, e0.dF > 0.0);
  t0 = t0 - e0.F * fast_rcp(e0.dF);
#pragma unroll
  for (int q = 1; q < ART_QUARTIC_STEPS; ++q) {
    e0 = tor_eval_quartic(r, t0, E.sp[3], E.sp[4]);
    ok = mand(ok, e0.dF > 0.0);
    t0 = t0 - e0.F * fast_rcp(e0.dF);
  }
  M vb;
  const T tb = tor_newton<+1, 0>(r, t0, tor_eval(r, t0, R, r2), R, r2, rr, ok, vb);
#else
  M vb;
  const T tb = tor_newton<+1, ART_NEWTON_PRE>(r, t0, e0, R, r2, rr, act, vb);
#endif
  // the far root as a candidate: t > 1e-12, z < -R, on the support (pick_candidate's rule for one root)
  M cb = mand(vb, tb > 1e-12);
  {
    const T x = mfma(tb, r.ux, r.px), y = mfma(tb, r.uy, r.py), z = mfma(tb, r.uz, r.pz);
    cb = mand(mand(cb, z < -R), in_support<false>(E, x, y));
  }
  // A second, nearer root exists only when the origin lies OUTSIDE the solid (F(0) > 0).  Sign of F(0)
  // without the square root: with A = x^2 + y^2 + z^2 + R^2 - r^2 the origin is inside iff y^2 <= r^2 and
  // (rho <= R or A <= 2 R rho); a relative margin of 1e-9 (1e7 times the rounding error of either side)
  // decides "clearly inside", anything closer to the surface takes the exact evaluation below.
  M maybe_out = act;
#if ART_CHEAP_INSIDE
  {
    const T s = mfma(r.px, r.px, r.pz * r.pz), yy = r.py * r.py;
    const T A = s + yy + E.sp[3];
    const M in_y = yy < 0.999999999 * r2;
    const M in_rho = mor(mor(s <= E.sp[4], A <= 0.0), A * A < (3.999999996 * E.sp[4]) * s);
    maybe_out = mand(act, mnot(mand(in_y, in_rho)));
  }
#endif
  if (!any(maybe_out)) {
    valid = cb;
    return tb;
  }
  const TorEval<T> o = tor_eval(r, splat<T>(0.0), R, r2);
  const M outside = mand(maybe_out, o.F > 0.0);    // origin outside the solid: a second, nearer root may exist
  const M away = mand(outside, mnot(o.dF < 0.0));  // ... but not if we move away from the solid
  const M need = mand(outside, mnot(away));
  T ta = splat<T>(0.0);
  M va = splat_mask<T>(false);
  if (any(need)) ta = tor_newton<-1, 0>(r, splat<T>(0.0), o, R, r2, rr, need, va);
  M ca = mand(mand(need, va), ta > 1e-12);
  {
    const T x = mfma(ta, r.ux, r.px), y = mfma(ta, r.uy, r.py), z = mfma(ta, r.uz, r.pz);
    ca = mand(mand(ca, z < -R), in_support<false>(E, x, y));
  }
  // one candidate -> it, two -> the nearer (ta <= tb)
  valid = mand(mor(ca, cb), mnot(away));
  return sel(ca, ta, tb);
}

// ---------------------------------------------------------------------------------------------
// Zernike defect, ART/ModuleDefects.py:149-177.  The reference evaluates Andersen's Cartesian
// recurrences for ALL (n,m) up to max_order with Python lists.  Device path: the same polynomials
// written as  Z = Q_k^l(s) * {C_l, S_l}(x,y),  s = x^2+y^2,  C_l + i S_l = (x + i y)^l,
// Q_k^l(s) = R_{l+2k}^l(rho)/rho^l = (-1)^k P_k^{(l,0)}(1-2s)  (Jacobi), advanced in k by a
// three-term recurrence -> O(1) registers per ray for any order.  Table (host-built, smem):
//   zt[0] = radius R (= Support._CircumCirc()), zt[1] = max order N, then for l = 0..N,
//   k = 0..(N-l)/2: {alpha, beta, gamma, c_cos, c_sin, pad} with
//   Q_k = (alpha s + beta) Q_{k-1} - gamma Q_{k-2};  c_cos / c_sin = coefficients of the reference
//   keys (n, (n+l)/2) / (n, (n-l)/2), n = l + 2k.
// Returns value and Cartesian gradient (already divided by R where the reference does).
// ---------------------------------------------------------------------------------------------
// One table record per (l, k): {alpha, beta, gamma, c_cos, c_sin, pad} -- six doubles, so that every record is
// three aligned 16-byte loads from the shared-memory table.
struct ZRec {
  double al, be, ga, cc, cs;
};
ART_HD ZRec zernike_record(const double* __restrict__ rec) {
#ifdef __CUDA_ARCH__
  const double2* q = reinterpret_cast<const double2*>(rec);
  const double2 a = q[0], b = q[1], c = q[2];
  return {a.x, a.y, b.x, b.y, c.x};
#else
  return {rec[0], rec[1], rec[2], rec[3], rec[4]};
#endif
}
constexpr int ZREC = 6;  // doubles per record
// Unrolling of the radial recurrence.  Measured on cfg4 (10^7 rays, order 20, profiles/r02_summary.md): 1: 0.445,
// 2: 0.422, 4: 0.418 ms.  One flat loop over all records with a block-start mark per record (no nested loops) was
// tried as well and is slower (0.512 ms: a uniform but unpredictable branch per record, and spills).
#ifndef ART_ZUNROLL
#define ART_ZUNROLL 4
#endif
constexpr int ZUNROLL = ART_ZUNROLL;

template <bool WANT_VALUE, bool WANT_GRAD, class T>
ART_HD void zernike_eval(const double* __restrict__ zt, T X, T Y, T& val, T& gx, T& gy) {
  const double Rz = zt[0];
  const int N = (int)zt[1];
  const double iR = fdiv(1.0, Rz);
  const T x = X * iR, y = Y * iR;
  const T s = mfma(x, x, y * y);
  const double* rec = zt + 2;
  T Cl = splat<T>(1.0), Sl = splat<T>(0.0), Cm = splat<T>(0.0), Sm = splat<T>(0.0);  // (x+iy)^l, (x+iy)^(l-1)
  T v = splat<T>(0.0), dx = splat<T>(0.0), dy = splat<T>(0.0);
  for (int l = 0; l <= N; ++l) {
    const int K = (N - l) >> 1;
    T Q = splat<T>(1.0), Qp = splat<T>(0.0), dQ = splat<T>(0.0), dQp = splat<T>(0.0);
    const ZRec r0 = zernike_record(rec);
    T A = splat<T>(r0.cc), B = splat<T>(r0.cs), dA = splat<T>(0.0), dB = splat<T>(0.0);
    rec += ZREC;
#pragma unroll(ZUNROLL)
    for (int k = 1; k <= K; ++k) {
      const ZRec z = zernike_record(rec);
      rec += ZREC;
      const T lin = mfma(z.al, s, z.be);
      const T Qn = mfma(lin, Q, -z.ga * Qp);
      if (WANT_GRAD) {
        const T dQn = mfma(z.al, Q, mfma(lin, dQ, -z.ga * dQp));
        dQp = dQ;
        dQ = dQn;
        dA = mfma(z.cc, dQn, dA);
        dB = mfma(z.cs, dQn, dB);
      }
      Qp = Q;
      Q = Qn;
      A = mfma(z.cc, Qn, A);
      B = mfma(z.cs, Qn, B);
    }
    if (WANT_VALUE) v = mfma(A, Cl, mfma(B, Sl, v));
    if (WANT_GRAD) {
      const T rad = mfma(dA, Cl, dB * Sl);  // sum c dQ/ds * angular part
      const double fl = (double)l;
      dx = mfma(2.0 * x, rad, mfma(fl, mfma(A, Cm, B * Sm), dx));
      dy = mfma(2.0 * y, rad, mfma(fl, mfma(B, Cm, -(A * Sm)), dy));
    }
    Cm = Cl;
    Sm = Sl;
    const T Cn = mfma(x, Cl, -(y * Sl));
    Sl = mfma(x, Sl, y * Cl);
    Cl = Cn;
  }
  val = v;
  gx = dx * iR;  // ModuleDefects.py:163-164
  gy = dy * iR;
}

// ---------------------------------------------------------------------------------------------
// Gridded defect (MeasuredMap / Fourrier, ART/ModuleDefects.py:34-146): bilinear interpolation of the
// height map (WHICH = 0) or of the two slope maps (WHICH = 1) at (X, Y) relative to the optic centre,
// as scipy's RegularGridInterpolator(method="linear") does on the reference's grids.
// ---------------------------------------------------------------------------------------------
template <int WHICH>
ART_HD void gridmap_eval(const MapDev& M, double X, double Y, double& a, double& b) {
  const double fx = (X - M.x0) * M.sx, fy = (Y - M.y0) * M.sy;
  int ix = (int)floor(fx), iy = (int)floor(fy);
  ix = ix < 0 ? 0 : (ix > M.nx - 2 ? M.nx - 2 : ix);
  iy = iy < 0 ? 0 : (iy > M.ny - 2 ? M.ny - 2 : iy);
  const double tx = fx - ix, ty = fy - iy;
  const double w00 = (1.0 - tx) * (1.0 - ty), w10 = tx * (1.0 - ty), w01 = (1.0 - tx) * ty, w11 = tx * ty;
  const long long o = (long long)ix * M.ny + iy;
  if (WHICH == 0) {
    a = w00 * M.h[o] + w10 * M.h[o + M.ny] + w01 * M.h[o + 1] + w11 * M.h[o + M.ny + 1];
    b = 0.0;
  } else {
    a = w00 * M.dx[o] + w10 * M.dx[o + M.ny] + w01 * M.dx[o + 1] + w11 * M.dx[o + M.ny + 1];
    b = w00 * M.dy[o] + w10 * M.dy[o + M.ny] + w01 * M.dy[o + 1] + w11 * M.dy[o + M.ny + 1];
  }
}
template <int WHICH>
ART_HD void gridmap_eval(const MapDev& M, D2 X, D2 Y, D2& a, D2& b) {
  gridmap_eval<WHICH>(M, X.a, Y.a, a.a, b.a);
  gridmap_eval<WHICH>(M, X.b, Y.b, a.b, b.b);
}

// ---------------------------------------------------------------------------------------------
// surface normal, `get_normal` of each mirror class: surface_gradient gives its direction (any length),
// surface_normal the unit vector
// ---------------------------------------------------------------------------------------------
// SURFS (0 any / 1 toroid class / 2 quadric class, the enum further down): the surface kinds a kernel instantiation can
// meet.  Written as nested two-way decisions on the warp-uniform kind: a five-way switch becomes a jump table.
template <int SURFS = 0, class T>
ART_HD void surface_gradient(const ElemDev& E, T x, T y, T z, T& gx, T& gy, T& gz) {
  const int kind = E.surface;
  if (SURFS != 2 && kind == ART_SURF_TOROIDAL) {  // ART/ModuleMirror.py:480-498 (common factor 4 dropped)
    const T S = mfma(x, x, mfma(y, y, z * z));
    const T a = S + E.sp[3];        // + (R^2 - r^2)
    const T b = a - 2.0 * E.sp[4];  // - 2 R^2
    gx = -(x * b); gy = -(y * a); gz = -(z * b);
    return;
  }
  if (kind == ART_SURF_PLANE || kind == ART_SURF_MASK || SURFS == 1) {  // :84-87
    gx = splat<T>(0.0); gy = splat<T>(0.0); gz = splat<T>(1.0);
    return;
  }
  if (kind == ART_SURF_SPHERICAL || kind == ART_SURF_PARABOLIC) {
    gx = -x; gy = -y;
    if (kind == ART_SURF_SPHERICAL) gz = -z;     // :180-183
    else gz = splat<T>(E.sp[0]);                 // :349-355
    return;
  }
  if (kind == ART_SURF_ELLIPSOIDAL) {  // :685-693
    gx = -(x * E.sp[2]); gy = -(y * E.sp[3]); gz = -(z * E.sp[3]);
    return;
  }
  gx = splat<T>(0.0); gy = -y; gz = -z;  // ART_SURF_CYLINDRICAL :846-849
}
template <class T>
ART_HD void surface_normal(const ElemDev& E, T x, T y, T z, T& nx, T& ny, T& nz) {
  T gx, gy, gz;
  surface_gradient(E, x, y, z, gx, gy, gz);
  if (E.surface == ART_SURF_PLANE || E.surface == ART_SURF_MASK) {
    nx = gx; ny = gy; nz = gz;
    return;
  }
  const T inv = frsqrt(mfma(gx, gx, mfma(gy, gy, gz * gz)));
  nx = gx * inv; ny = gy * inv; nz = gz * inv;
}

// atan2(y, x) for y >= 0 (result in [0, pi]), branch-free: one division for the reduced argument
// z (|z| <= tan(pi/8)) and atan(z) = z - z w P(w), w = z^2, P of degree 11 interpolating (1 - atan(z)/z)/w
// at the Chebyshev nodes of [0, tan^2(pi/8)] (relative error of atan below 3.2e-18 with the coefficients
// rounded to double; fitted with mpmath at 60 digits).  Estrin evaluation: dependency depth 5 instead of
// 12 -- this kernel is bound by dependent-FP64 latency, not by the FMA count.  The coefficients are
// constant-bank operands of the FMAs on the device (no UMOV pairs to materialise 64-bit immediates).
#define ART_ATAN_COEFFS                                                                                     \
  {0.3333333333333333, -0.19999999999999804, 0.14285714285659828, -0.11111111105155447,                     \
   0.09090908753500877, -0.07692296375032143, 0.06666424885738255, -0.05878928997834775,                    \
   0.05230454270650244, -0.04551593220626549, 0.034570561981427744, -0.016285756855221028}
#ifdef __CUDACC__
__constant__ double c_atan[12] = ART_ATAN_COEFFS;
#endif
template <class T>
ART_HD T atan_poly(T w) {
#ifdef __CUDA_ARCH__
  const double* c = c_atan;
#else
  const double c[12] = ART_ATAN_COEFFS;
#endif
  const T w2 = w * w, w4 = w2 * w2, w8 = w4 * w4;
  const T p01 = mfma(w, c[1], c[0]), p23 = mfma(w, c[3], c[2]), p45 = mfma(w, c[5], c[4]);
  const T p67 = mfma(w, c[7], c[6]), p89 = mfma(w, c[9], c[8]), pab = mfma(w, c[11], c[10]);
  const T q0 = mfma(p23, w2, p01), q1 = mfma(p67, w2, p45), q2 = mfma(pab, w2, p89);
  return mfma(q2, w8, mfma(q1, w4, q0));
}
template <class T>
ART_HD T fatan2_ypos(T y, T x) {
  typedef typename MaskOf<T>::type M;
  const T ax = mabs(x);
  const M swap = y > ax;
  const T lo = sel(swap, ax, y), hi = sel(swap, y, ax);      // lo/hi in [0, 1]
  const M big = lo > 0.41421356237309503 * hi;               // beyond tan(pi/8): rotate by pi/4
  const T z = fdiv(sel(big, lo - hi, lo), sel(big, lo + hi, hi));
  const T w = z * z;
  const T p = atan_poly(w);
  T a = mfma(-(z * w), p, z);                                // atan(z)
  a = a + sel(big, 0.78539816339744831, 0.0);
  a = sel(swap, 1.5707963267948966 - a, a);
  return sel(x < 0.0, 3.1415926535897931 - a, a);
}

// Angle between two UNIT vectors a, b.  The reference uses Kahan's 2 atan2(|a-b|, |a+b|)
// (ART/ModuleGeometry.py:40-44); atan2(|a x b|, a.b) is the same angle, equally well conditioned over
// [0, pi], and needs one square root instead of two.
template <class T>
ART_HD T unit_angle(T ax, T ay, T az, T bx, T by, T bz) {
  const T cx = mfma(ay, bz, -(az * by)), cy = mfma(az, bx, -(ax * bz)), cz = mfma(ax, by, -(ay * bx));
  return fatan2_ypos(fsqrt(mfma(cx, cx, mfma(cy, cy, cz * cz))), mfma(ax, bx, mfma(ay, by, az * bz)));
}

// surface classes a kernel instantiation is compiled for (the chain says which it needs)
enum { SURFS_ANY = 0, SURFS_TOROID = 1, SURFS_QUADRIC = 2 };

// ---------------------------------------------------------------------------------------------
// lab -> element frame, ART/ModuleProcessing.py:289-295: p_e = R (p - pos) + centre, u_e = R u, in place.
// eorg: the origin shared by all rays (point source), already in this element's frame.
// ---------------------------------------------------------------------------------------------
template <class T>
ART_HD void to_element_frame(const ElemDev& E, RayT<T>& r, const double* __restrict__ eorg = nullptr) {
  T px, py, pz;
  if (eorg) {
    px = splat<T>(eorg[0]); py = splat<T>(eorg[1]); pz = splat<T>(eorg[2]);
  } else {
    const T dx = r.px - E.pos[0], dy = r.py - E.pos[1], dz = r.pz - E.pos[2];
    px = mfma(E.rot[0], dx, mfma(E.rot[1], dy, mfma(E.rot[2], dz, E.ctr[0])));
    py = mfma(E.rot[3], dx, mfma(E.rot[4], dy, mfma(E.rot[5], dz, E.ctr[1])));
    pz = mfma(E.rot[6], dx, mfma(E.rot[7], dy, mfma(E.rot[8], dz, E.ctr[2])));
  }
  const T ux = mfma(E.rot[0], r.ux, mfma(E.rot[1], r.uy, E.rot[2] * r.uz));
  const T uy = mfma(E.rot[3], r.ux, mfma(E.rot[4], r.uy, E.rot[5] * r.uz));
  const T uz = mfma(E.rot[6], r.ux, mfma(E.rot[7], r.uy, E.rot[8] * r.uz));
  r.px = px; r.py = py; r.pz = pz;
  r.ux = ux; r.uy = uy; r.uz = uz;
}

// ---------------------------------------------------------------------------------------------
// one element acting on one ray (pair) that is ALREADY in the element's frame: the optic of
// ART/ModuleProcessing.py:298-301 and the frame change that follows, in place.
//
// The reference takes every ray back to the lab frame after each element and into the next element's frame
// right afterwards (:306-309 then :289-295 of the next loop trip).  Between two elements the chain here hands
// the ray over DIRECTLY: p' = M h + b, u' = M o with M = R_next R^T and b = R_next (pos - pos_next - R^T ctr)
// + ctr_next composed once per element pair on the host (link_elements, art_lowering.h) -- one affine map
// instead of two, 18 FMAs less per ray and element boundary.
//   out_next  hand the ray over in the NEXT element's frame (E.nrot / E.noff); otherwise r leaves in the lab frame
// The lab-frame bundle after an inner element (the per-element history the reference API returns) is obtained
// from the handed-over ray with frame_to_lab(next element).
// ---------------------------------------------------------------------------------------------
template <bool WANT_INC, bool HAS_DEF, int SURFS, class T>
ART_HD void apply_element(const ElemDev& E, RayT<T>& r, const double* __restrict__ ztab,
                          const int* __restrict__ zoff, bool ignore_defects, bool inc_here,
                          const MapDev* __restrict__ maps, bool out_next) {
  typedef typename MaskOf<T>::type M;
  const M act = r.alive;
  RayT<T>& e = r;   // the ray in this element's frame; r.p / r.u are overwritten only once h and o are known
  T t;
  M valid;
  const int surf = E.surface;
  if (surf == ART_SURF_PLANE || surf == ART_SURF_MASK) {
    // ART/ModuleMirror.py:73-82: t > 0 (no epsilon) and on the support;
    // ART/ModuleMask.py:51-61: passes iff t > 0 and NOT on the support
    t = fdiv(-e.pz, e.uz);
    const T x = mfma(t, e.ux, e.px), y = mfma(t, e.uy, e.py);
    M ok = in_support<false>(E, x, y);
    if (surf == ART_SURF_MASK) ok = mnot(ok);
    valid = mand(t > 0.0, ok);
  } else if (SURFS != SURFS_QUADRIC && surf == ART_SURF_TOROIDAL) {
    t = intersect_toroid(E, e, act, valid);
  } else if (SURFS != SURFS_TOROID) {
    T a, b, c;
    bool side = true;
    if (surf == ART_SURF_SPHERICAL) {  // :163-178
      a = mfma(e.ux, e.ux, mfma(e.uy, e.uy, e.uz * e.uz));
      b = 2.0 * mfma(e.ux, e.px, mfma(e.uy, e.py, e.uz * e.pz));
      c = mfma(e.px, e.px, mfma(e.py, e.py, mfma(e.pz, e.pz, -E.sp[1])));
    } else if (surf == ART_SURF_PARABOLIC) {  // :325-347 (no z test)
      const double p = E.sp[0];
      a = mfma(e.ux, e.ux, e.uy * e.uy);
      b = 2.0 * mfma(e.ux, e.px, mfma(e.uy, e.py, -p * e.uz));
      c = mfma(e.px, e.px, mfma(e.py, e.py, -2.0 * p * e.pz));
      side = false;
    } else if (surf == ART_SURF_ELLIPSOIDAL) {  // :662-683, sp[2] = 1/a^2, sp[3] = 1/b^2
      const double ia = E.sp[2], ib = E.sp[3];
      a = mfma(mfma(e.uy, e.uy, e.uz * e.uz), ib, e.ux * e.ux * ia);
      b = 2.0 * mfma(mfma(e.uy, e.py, e.uz * e.pz), ib, e.ux * e.px * ia);
      c = mfma(mfma(e.py, e.py, e.pz * e.pz), ib, mfma(e.px * e.px, ia, -1.0));
    } else {  // ART_SURF_CYLINDRICAL :824-844
      a = mfma(e.uy, e.uy, e.uz * e.uz);
      b = 2.0 * mfma(e.uy, e.py, e.uz * e.pz);
      c = mfma(e.py, e.py, mfma(e.pz, e.pz, -E.sp[1]));
    }
    T t1, t2;
    solve_quadratic(a, b, c, t1, t2);
    t = side ? pick_candidate<true>(E, e, t1, t2, 0.0, valid) : pick_candidate<false>(E, e, t1, t2, 0.0, valid);
  } else {
    t = splat<T>(ART_NAN);
    valid = splat_mask<T>(false);
  }
  // miss: the reference drops the ray (ModuleMirror.py:932, ModuleMask.py:132).  Validity travels as a lane mask, not
  // as a NaN in t: no selects to poison t and no FP64 compare to read the poison back.
  const M hit = mand(act, valid);
  if (!any(hit)) {
    r.alive = hit;
    return;
  }
  T hx = mfma(t, e.ux, e.px), hy = mfma(t, e.uy, e.py), hz = mfma(t, e.uz, e.pz);
  T inc = r.inc;
  if (surf == ART_SURF_MASK) {
    // _TransmitMaskRay, ART/ModuleMask.py:93-108: direction unchanged (r.u stays), incidence vs ez
    if (WANT_INC && inc_here) inc = unit_angle(e.ux, e.uy, e.uz, splat<T>(0.0), splat<T>(0.0), splat<T>(1.0));
  } else {
    const bool deformed = HAS_DEF && (E.n_defects > 0 || E.n_maps > 0);
    if (ART_GRADIENT_REFLECT && !deformed) {
      // _ReflectionMirrorRay, ART/ModuleMirror.py:878-906 with the normal left unnormalised: u' = u - 2 (g.u)/(g.g) g
      // and angle(-u, g) are both independent of |g| -- no reciprocal square root, no renormalisation (the
      // reflection keeps |u| to rounding; Ray.vector's renormalisation, ART/ModuleOpticalRay.py:85-90, moves the
      // last bit only)
      T gx, gy, gz;
      surface_gradient<SURFS>(E, hx, hy, hz, gx, gy, gz);
      const T d = mfma(gx, e.ux, mfma(gy, e.uy, gz * e.uz));
      if (WANT_INC && inc_here) inc = unit_angle(-e.ux, -e.uy, -e.uz, gx, gy, gz);
      const T k = fdiv(-2.0 * d, mfma(gx, gx, mfma(gy, gy, gz * gz)));
      r.ux = mfma(k, gx, e.ux); r.uy = mfma(k, gy, e.uy); r.uz = mfma(k, gz, e.uz);
    } else {
      T nx, ny, nz;
      surface_normal(E, hx, hy, hz, nx, ny, nz);
      if (HAS_DEF && (E.n_defects > 0 || E.n_maps > 0)) {
        // DeformedMirror._get_intersection, ART/ModuleMirror.py:969-980:
        //   h = sum offsets(P - C); alpha = angle(-u, n_base(P)); P -= u h / cos(alpha)
        T h = splat<T>(0.0);
        for (int d = 0; d < E.n_defects; ++d) {
          T v, g0, g1;
          zernike_eval<true, false>(ztab + zoff[E.first_defect + d], hx - E.ctr[0], hy - E.ctr[1], v, g0, g1);
          h = h + v;
        }
        for (int d = 0; d < E.n_maps; ++d) {
          T v, unused;
          gridmap_eval<0>(maps[E.first_map + d], hx - E.ctr[0], hy - E.ctr[1], v, unused);
          h = h + v;
        }
        const T cosa = -mfma(nx, e.ux, mfma(ny, e.uy, nz * e.uz));
        const T sh = fdiv(h, cosa);
        t = t - sh;
        hx = mfma(-sh, e.ux, hx); hy = mfma(-sh, e.uy, hy); hz = mfma(-sh, e.uz, hz);
        surface_normal(E, hx, hy, hz, nx, ny, nz);  // the reflection uses get_normal(shifted point)
        if (!ignore_defects) {
          // DeformedMirror.get_normal + normal_add, ART/ModuleMirror.py:952-961, ModuleGeometry.py:394-407:
          // slopes add; the result is (-gx, -gy, 1) normalised
          const T inz = fdiv(1.0, nz);
          T sx = -(nx * inz), sy = -(ny * inz);
          for (int d = 0; d < E.n_defects; ++d) {
            T v, g0, g1;
            zernike_eval<false, true>(ztab + zoff[E.first_defect + d], hx - E.ctr[0], hy - E.ctr[1], v, g0, g1);
            sx = sx + g0;
            sy = sy + g1;
          }
          for (int d = 0; d < E.n_maps; ++d) {
            // MeasuredMap / Fourrier.get_normal returns (+dX, +dY, 1)/norm (ART/ModuleDefects.py:52-58,119-129),
            // so normal_add's slope -n_x/n_z is MINUS the interpolated derivative
            T g0, g1;
            gridmap_eval<1>(maps[E.first_map + d], hx - E.ctr[0], hy - E.ctr[1], g0, g1);
            sx = sx - g0;
            sy = sy - g1;
          }
          const T inv = frsqrt(mfma(sx, sx, mfma(sy, sy, 1.0)));
          nx = -(sx * inv); ny = -(sy * inv); nz = inv;
        }
      }
      // _ReflectionMirrorRay, ART/ModuleMirror.py:878-906: u' = u - 2 (n.u) n, incidence = angle(-u, n)
      const T d = mfma(nx, e.ux, mfma(ny, e.uy, nz * e.uz));
      if (WANT_INC && inc_here) inc = unit_angle(-e.ux, -e.uy, -e.uz, nx, ny, nz);
      T ox = mfma(-2.0 * d, nx, e.ux), oy = mfma(-2.0 * d, ny, e.uy), oz = mfma(-2.0 * d, nz, e.uz);
      // Ray.vector setter renormalises (ART/ModuleOpticalRay.py:85-90); one Newton step is exact here
      const T sc = mfma(-0.5, mfma(ox, ox, mfma(oy, oy, oz * oz)), 1.5);
      r.ux = ox * sc; r.uy = oy * sc; r.uz = oz * sc;   // the outgoing direction, still in this element's frame
    }
  }
  r.path = r.path + mabs(t);  // |P - A| with |u| = 1 (ModuleMirror.py:904, ModuleMask.py:100)
  r.inc = inc;
  r.alive = hit;
  // Leave the element: lab frame (:306-309: p = R^T (p_e - centre) + pos, u = R^T u_e) after the last element,
  // else straight into the next element's frame (p' = M h + b, u' = M o).  Written unconditionally: a lane that
  // missed is dead from here on and its columns are never read or stored again.
  const T ox = r.ux, oy = r.uy, oz = r.uz;
  if (!out_next) {
    const T dx = hx - E.ctr[0], dy = hy - E.ctr[1], dz = hz - E.ctr[2];
    r.px = mfma(E.rot[0], dx, mfma(E.rot[3], dy, mfma(E.rot[6], dz, E.pos[0])));
    r.py = mfma(E.rot[1], dx, mfma(E.rot[4], dy, mfma(E.rot[7], dz, E.pos[1])));
    r.pz = mfma(E.rot[2], dx, mfma(E.rot[5], dy, mfma(E.rot[8], dz, E.pos[2])));
    r.ux = mfma(E.rot[0], ox, mfma(E.rot[3], oy, E.rot[6] * oz));
    r.uy = mfma(E.rot[1], ox, mfma(E.rot[4], oy, E.rot[7] * oz));
    r.uz = mfma(E.rot[2], ox, mfma(E.rot[5], oy, E.rot[8] * oz));
  } else {
    r.px = mfma(E.nrot[0], hx, mfma(E.nrot[1], hy, mfma(E.nrot[2], hz, E.noff[0])));
    r.py = mfma(E.nrot[3], hx, mfma(E.nrot[4], hy, mfma(E.nrot[5], hz, E.noff[1])));
    r.pz = mfma(E.nrot[6], hx, mfma(E.nrot[7], hy, mfma(E.nrot[8], hz, E.noff[2])));
    r.ux = mfma(E.nrot[0], ox, mfma(E.nrot[1], oy, E.nrot[2] * oz));
    r.uy = mfma(E.nrot[3], ox, mfma(E.nrot[4], oy, E.nrot[5] * oz));
    r.uz = mfma(E.nrot[6], ox, mfma(E.nrot[7], oy, E.nrot[8] * oz));
  }
}

// A ray that sits in element E's frame (handed over by the previous element) expressed in the lab frame:
// ART/ModuleProcessing.py:306-309 -- the bundle the reference stores after the previous element.
template <class T>
ART_HD void frame_to_lab(const ElemDev& E, const RayT<T>& in, RayT<T>& out) {
  const T dx = in.px - E.ctr[0], dy = in.py - E.ctr[1], dz = in.pz - E.ctr[2];
  out.px = mfma(E.rot[0], dx, mfma(E.rot[3], dy, mfma(E.rot[6], dz, E.pos[0])));
  out.py = mfma(E.rot[1], dx, mfma(E.rot[4], dy, mfma(E.rot[7], dz, E.pos[1])));
  out.pz = mfma(E.rot[2], dx, mfma(E.rot[5], dy, mfma(E.rot[8], dz, E.pos[2])));
  out.ux = mfma(E.rot[0], in.ux, mfma(E.rot[3], in.uy, E.rot[6] * in.uz));
  out.uy = mfma(E.rot[1], in.ux, mfma(E.rot[4], in.uy, E.rot[7] * in.uz));
  out.uz = mfma(E.rot[2], in.ux, mfma(E.rot[5], in.uy, E.rot[8] * in.uz));
  out.path = in.path;
  out.inc = in.inc;
  out.alive = in.alive;
}

}  // namespace art
