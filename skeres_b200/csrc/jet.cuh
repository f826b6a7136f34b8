// jet.cuh — forward-mode dual numbers held in registers, and the registered device cost functors.
//
// Mirrors spire `Jet` as used by the reference's autodiff bridge
// (core/.../AutodiffCostFunction.scala:95-130: seed Jet(x, k) per scalar parameter, run the functor
// once, read .real -> residuals and .infinitesimal -> row-major per-block Jacobians), with ordering
// on real parts only (core/.../package.scala:27).
//
// Functors:
//   snavely_reprojection   examples/.../SimpleBundleAdjuster.scala:79-119  (2; 9, 3)
//   angle_axis_rotate_point core/.../Rotation.scala:449-522
//   exponential_residual   examples/.../CurveFitting.scala:92-98           (1; 1, 1)
//   test functors          core/src/test/.../AutodiffCostFuntionSpec.scala (golden vectors)
//
// Everything is __host__ __device__ so the arithmetic can be unit-checked on the build host
// (tests/hostcheck); the product only ever calls it from kernels.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define SK_HD __host__ __device__ __forceinline__
#else
#define SK_HD inline
#endif

namespace sk {

template <int N>
struct Jet {
  double a;
  double v[N];
  SK_HD Jet() {}
  SK_HD Jet(double real) : a(real) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = 0.0;
  }
  SK_HD Jet(double real, int k) : a(real) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = (i == k) ? 1.0 : 0.0;
  }
};

#define SK_JET_LOOP _Pragma("unroll") for (int i = 0; i < N; ++i)

template <int N> SK_HD Jet<N> operator+(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a + y.a; SK_JET_LOOP r.v[i] = x.v[i] + y.v[i]; return r; }
template <int N> SK_HD Jet<N> operator-(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a - y.a; SK_JET_LOOP r.v[i] = x.v[i] - y.v[i]; return r; }
template <int N> SK_HD Jet<N> operator-(const Jet<N>& x) { Jet<N> r; r.a = -x.a; SK_JET_LOOP r.v[i] = -x.v[i]; return r; }
template <int N> SK_HD Jet<N> operator*(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a * y.a; SK_JET_LOOP r.v[i] = y.a * x.v[i] + x.a * y.v[i]; return r; }
template <int N> SK_HD Jet<N> operator/(const Jet<N>& x, const Jet<N>& y) {
  Jet<N> r; const double binv = 1.0 / y.a; const double q = x.a * binv; r.a = q;
  SK_JET_LOOP r.v[i] = binv * (x.v[i] - q * y.v[i]);
  return r;
}
template <int N> SK_HD Jet<N> operator+(const Jet<N>& x, double s) { Jet<N> r = x; r.a = x.a + s; return r; }
template <int N> SK_HD Jet<N> operator+(double s, const Jet<N>& x) { Jet<N> r = x; r.a = s + x.a; return r; }
template <int N> SK_HD Jet<N> operator-(const Jet<N>& x, double s) { Jet<N> r = x; r.a = x.a - s; return r; }
template <int N> SK_HD Jet<N> operator-(double s, const Jet<N>& x) { Jet<N> r; r.a = s - x.a; SK_JET_LOOP r.v[i] = -x.v[i]; return r; }
template <int N> SK_HD Jet<N> operator*(const Jet<N>& x, double s) { Jet<N> r; r.a = x.a * s; SK_JET_LOOP r.v[i] = x.v[i] * s; return r; }
template <int N> SK_HD Jet<N> operator*(double s, const Jet<N>& x) { return x * s; }
template <int N> SK_HD Jet<N> operator/(double s, const Jet<N>& y) {
  Jet<N> r; const double binv = 1.0 / y.a; const double q = s * binv; r.a = q;
  SK_JET_LOOP r.v[i] = binv * (-(q * y.v[i]));
  return r;
}
template <int N> SK_HD bool operator>(const Jet<N>& x, double y) { return x.a > y; }   // package.scala:27

template <int N> SK_HD Jet<N> jsqrt(const Jet<N>& x) { Jet<N> r; const double sa = sqrt(x.a); const double h = 1.0 / (2.0 * sa); r.a = sa; SK_JET_LOOP r.v[i] = x.v[i] * h; return r; }
template <int N> SK_HD void jsincos(const Jet<N>& x, Jet<N>* s, Jet<N>* c) {
  double sa, ca;
#ifdef __CUDA_ARCH__
  sincos(x.a, &sa, &ca);
#else
  sa = sin(x.a); ca = cos(x.a);
#endif
  s->a = sa; c->a = ca;
  SK_JET_LOOP { s->v[i] = ca * x.v[i]; c->v[i] = -sa * x.v[i]; }
}
template <int N> SK_HD Jet<N> jexp(const Jet<N>& x) { Jet<N> r; const double e = exp(x.a); r.a = e; SK_JET_LOOP r.v[i] = e * x.v[i]; return r; }
template <int N> SK_HD Jet<N> jlog(const Jet<N>& x) { Jet<N> r; const double inv = 1.0 / x.a; r.a = log(x.a); SK_JET_LOOP r.v[i] = x.v[i] * inv; return r; }
// The names a functor written for T = double uses (spire's Trig / NRoot instances on Jet, package.scala): found by argument-dependent
// lookup when T = Jet<N>, so that ONE functor body serves both instantiations -- also for functors compiled at run time.
using ::sqrt; using ::exp; using ::log; using ::sin; using ::cos;      // keep the double versions visible next to the overloads
template <int N> SK_HD Jet<N> sqrt(const Jet<N>& x) { return jsqrt(x); }
template <int N> SK_HD Jet<N> exp(const Jet<N>& x) { return jexp(x); }
template <int N> SK_HD Jet<N> log(const Jet<N>& x) { return jlog(x); }
template <int N> SK_HD Jet<N> sin(const Jet<N>& x) { Jet<N> s, c; jsincos(x, &s, &c); return s; }
template <int N> SK_HD Jet<N> cos(const Jet<N>& x) { Jet<N> s, c; jsincos(x, &s, &c); return c; }
template <int N> SK_HD bool operator<(const Jet<N>& x, double y) { return x.a < y; }
template <int N> SK_HD bool operator>(const Jet<N>& x, const Jet<N>& y) { return x.a > y.a; }
template <int N> SK_HD bool operator<(const Jet<N>& x, const Jet<N>& y) { return x.a < y.a; }
SK_HD double jsqrt(double x) { return sqrt(x); }
SK_HD void jsincos(double x, double* s, double* c) {
#ifdef __CUDA_ARCH__
  sincos(x, s, c);
#else
  *s = sin(x); *c = cos(x);
#endif
}
SK_HD double jexp(double x) { return exp(x); }

#define SK_EPSILON_DOUBLE 2.220446049250313e-16 /* Math.ulp(1.0), package.scala:15 */

// Rotation.scala:449-522, statement order preserved.
template <class T>
SK_HD void angle_axis_rotate_point(const T* aa, const T* pt, T* result) {
  const T theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (theta2 > SK_EPSILON_DOUBLE) {
    const T theta = jsqrt(theta2);
    T sinTheta, cosTheta;
    jsincos(theta, &sinTheta, &cosTheta);
    const T thetaInverse = 1.0 / theta;
    const T w0 = aa[0] * thetaInverse, w1 = aa[1] * thetaInverse, w2 = aa[2] * thetaInverse;
    const T c0 = w1 * pt[2] - w2 * pt[1];
    const T c1 = w2 * pt[0] - w0 * pt[2];
    const T c2 = w0 * pt[1] - w1 * pt[0];
    const T tmp = (w0 * pt[0] + w1 * pt[1] + w2 * pt[2]) * (1.0 - cosTheta);
    result[0] = pt[0] * cosTheta + c0 * sinTheta + w0 * tmp;
    result[1] = pt[1] * cosTheta + c1 * sinTheta + w1 * tmp;
    result[2] = pt[2] * cosTheta + c2 * sinTheta + w2 * tmp;
  } else {
    result[0] = pt[0] + (aa[1] * pt[2] - aa[2] * pt[1]);
    result[1] = pt[1] + (aa[2] * pt[0] - aa[0] * pt[2]);
    result[2] = pt[2] + (aa[0] * pt[1] - aa[1] * pt[0]);
  }
}

// Projection half of SimpleBundleAdjuster.scala:99-117 given p = R X + t.
template <class T>
SK_HD void snavely_project(const T* p, const T& focal, const T& l1, const T& l2, double ox, double oy, T* res) {
  const T xp = (-p[0]) / p[2];
  const T yp = (-p[1]) / p[2];
  const T r2 = xp * xp + yp * yp;
  const T distortion = 1.0 + r2 * (l1 + l2 * r2);
  const T fd = focal * distortion;
  res[0] = fd * xp - ox;
  res[1] = fd * yp - oy;
}

// Residual only (T = double): the `jacobians.isNull` branch, AutodiffCostFunction.scala:80-93.
SK_HD void snavely_residual(const double* cam, const double* pt, double ox, double oy, double* res) {
  double p[3];
  angle_axis_rotate_point(cam, pt, p);
  p[0] += cam[3]; p[1] += cam[4]; p[2] += cam[5];
  snavely_project(p, cam[6], cam[7], cam[8], ox, oy, res);
}

// Residual + Jacobian.  The 12-wide dual of the reference (JetDim(9 + 3)) is evaluated in two
// 6-wide stages that exploit the seed structure: the rotation sees only (angle-axis, point); the
// translation enters additively; the projection sees only (p, focal, l1, l2).  The two stages are
// chained exactly (dres/daa = dres/dp * dp/daa, ...), so the result equals the 12-wide evaluation
// up to rounding while using half the registers and ~1/3 of the FP64 work.
//   F: 2 x 9 row-major (d res / d camera), E: 2 x 3 row-major (d res / d point)
SK_HD void snavely_residual_jacobian(const double* cam, const double* pt, double ox, double oy,
                                     double* res, double* F, double* E) {
  typedef Jet<6> J6;
  J6 q[3];
  {
    const J6 aa[3] = {J6(cam[0], 0), J6(cam[1], 1), J6(cam[2], 2)};
    const J6 X[3] = {J6(pt[0], 3), J6(pt[1], 4), J6(pt[2], 5)};
    angle_axis_rotate_point(aa, X, q);
  }
  J6 r[2];
  {
    const J6 p[3] = {J6(q[0].a + cam[3], 0), J6(q[1].a + cam[4], 1), J6(q[2].a + cam[5], 2)};
    snavely_project(p, J6(cam[6], 3), J6(cam[7], 4), J6(cam[8], 5), ox, oy, r);
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    res[k] = r[k].a;
    const double d0 = r[k].v[0], d1 = r[k].v[1], d2 = r[k].v[2];   // d res_k / d p
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      F[k * 9 + c] = d0 * q[0].v[c] + d1 * q[1].v[c] + d2 * q[2].v[c];          // angle-axis
      E[k * 3 + c] = d0 * q[0].v[3 + c] + d1 * q[1].v[3 + c] + d2 * q[2].v[3 + c];  // point
    }
    F[k * 9 + 3] = d0; F[k * 9 + 4] = d1; F[k * 9 + 5] = d2;                   // translation
    F[k * 9 + 6] = r[k].v[3]; F[k * 9 + 7] = r[k].v[4]; F[k * 9 + 8] = r[k].v[5];  // focal, l1, l2
  }
}

// ---- generic registered functors (full-width duals), T = double or Jet<N> --------------------
// params: pointer to the concatenated parameter blocks as T (block i starts at off[i]).
template <class T> SK_HD bool functor_snavely(const double* c, const T* cam, const T* pt, T* res) {
  T p[3];
  angle_axis_rotate_point(cam, pt, p);
  p[0] = p[0] + cam[3]; p[1] = p[1] + cam[4]; p[2] = p[2] + cam[5];
  snavely_project(p, cam[6], cam[7], cam[8], c[0], c[1], res);
  return true;
}
template <class T> SK_HD bool functor_exponential(const double* c, const T* m, const T* cc, T* res) {
  res[0] = c[1] - jexp(m[0] * c[0] + cc[0]);      // CurveFitting.scala:96
  return true;
}
template <class T> SK_HD bool functor_hello_world(const T* x, T* res) {
  res[0] = 10.0 - x[0];                            // HelloWorld.scala:13
  return true;
}
// Powell.scala:13-51; block 0 / block 1 of each functor are a[0] / b[0]
template <class T> SK_HD bool functor_powell_f1(const T* a, const T* b, T* res) { res[0] = a[0] + 10.0 * b[0]; return true; }          // :18
template <class T> SK_HD bool functor_powell_f2(const T* a, const T* b, T* res) { res[0] = sqrt(5.0) * a[0] - b[0]; return true; }    // :28 (sic)
template <class T> SK_HD bool functor_powell_f2a(const T* a, const T* b, T* res) { res[0] = sqrt(5.0) * (a[0] - b[0]); return true; }   // PowellAnalytic.scala:36
template <class T> SK_HD bool functor_powell_f3(const T* a, const T* b, T* res) { const T d = a[0] - 2.0 * b[0]; res[0] = d * d; return true; }   // :38-39
template <class T> SK_HD bool functor_powell_f4(const T* a, const T* b, T* res) { const T d = a[0] - b[0]; res[0] = (sqrt(10.0) * d) * d; return true; }   // :48-49
template <class T> SK_HD bool functor_bilinear_scalar(const double* c, const T* x, const T* y, T* z) {
  z[0] = x[0] * y[0] + x[1] * y[1] - c[0];        // AutodiffCostFuntionSpec.scala:23
  return true;
}
template <class T> SK_HD bool functor_bilinear_vector3(const double* c, const T* x, const T* y, T* z) {
  const double a = c[0];                           // AutodiffCostFuntionSpec.scala:64-67
  z[0] = x[0] * y[0] + x[1] * y[1] - a;
  z[1] = x[0] * y[0] - x[1] * y[1] + a;
  z[2] = x[0] * x[1] + y[0] * y[1] + 10.0 * a;
  return true;
}
template <class T> SK_HD bool functor_sum10(const T* p, T* z) {
  T s = p[0];                                      // AutodiffCostFuntionSpec.scala:117
#pragma unroll
  for (int i = 1; i < 10; ++i) s = s + p[i];
  z[0] = s;
  return true;
}

struct FunctorInfo { int id, nres, nblk, sizes[SK_MAX_PARAMETER_BLOCKS], nconsts, ntot; };

SK_HD bool functor_info(int id, FunctorInfo* f) {
  f->id = id;
  for (int i = 0; i < SK_MAX_PARAMETER_BLOCKS; ++i) f->sizes[i] = 0;
  switch (id) {
    case SK_FUNCTOR_SNAVELY_REPROJECTION_ERROR: f->nres = 2; f->nblk = 2; f->sizes[0] = 9; f->sizes[1] = 3; f->nconsts = 2; f->ntot = 12; return true;
    case SK_FUNCTOR_EXPONENTIAL_RESIDUAL: f->nres = 1; f->nblk = 2; f->sizes[0] = 1; f->sizes[1] = 1; f->nconsts = 2; f->ntot = 2; return true;
    case SK_FUNCTOR_HELLO_WORLD: f->nres = 1; f->nblk = 1; f->sizes[0] = 1; f->nconsts = 0; f->ntot = 1; return true;
    case SK_FUNCTOR_POWELL_F1: case SK_FUNCTOR_POWELL_F2: case SK_FUNCTOR_POWELL_F3: case SK_FUNCTOR_POWELL_F4:
    case SK_FUNCTOR_POWELL_ANALYTIC_F2:
      f->nres = 1; f->nblk = 2; f->sizes[0] = 1; f->sizes[1] = 1; f->nconsts = 0; f->ntot = 2; return true;
    case SK_FUNCTOR_TEST_BILINEAR_SCALAR: f->nres = 1; f->nblk = 2; f->sizes[0] = 2; f->sizes[1] = 2; f->nconsts = 1; f->ntot = 4; return true;
    case SK_FUNCTOR_TEST_BILINEAR_VECTOR3: f->nres = 3; f->nblk = 2; f->sizes[0] = 2; f->sizes[1] = 2; f->nconsts = 1; f->ntot = 4; return true;
    case SK_FUNCTOR_TEST_SUM10: f->nres = 1; f->nblk = 10; for (int i = 0; i < 10; ++i) f->sizes[i] = 1; f->nconsts = 0; f->ntot = 10; return true;
  }
  return false;
}

#define SK_MAX_RESIDUALS 3
#define SK_MAX_TOTAL_PARAMS 12

// AutoDiffCostFunction.evaluate for one residual block, generic functor.
//   x: the concatenated parameter values (ntot doubles); res: nres outputs;
//   jac: nullptr (residual only, :80) or nres x ntot row-major over the CONCATENATED parameters —
//   the caller slices per-block, row-major nres x N_i blocks out of it (:121-127).
SK_HD bool evaluate_functor(int id, const double* c, const double* x, double* res, double* jac) {
  switch (id) {
    case SK_FUNCTOR_SNAVELY_REPROJECTION_ERROR: {
      if (!jac) { snavely_residual(x, x + 9, c[0], c[1], res); return true; }
      double F[18], E[6];
      snavely_residual_jacobian(x, x + 9, c[0], c[1], res, F, E);
      for (int k = 0; k < 2; ++k) { for (int j = 0; j < 9; ++j) jac[k * 12 + j] = F[k * 9 + j]; for (int j = 0; j < 3; ++j) jac[k * 12 + 9 + j] = E[k * 3 + j]; }
      return true;
    }
    case SK_FUNCTOR_EXPONENTIAL_RESIDUAL: {
      if (!jac) return functor_exponential(c, x, x + 1, res);
      Jet<2> jx[2] = {Jet<2>(x[0], 0), Jet<2>(x[1], 1)}, jr[1];
      if (!functor_exponential(c, jx, jx + 1, jr)) return false;
      res[0] = jr[0].a; jac[0] = jr[0].v[0]; jac[1] = jr[0].v[1];
      return true;
    }
    case SK_FUNCTOR_HELLO_WORLD: {
      if (!jac) return functor_hello_world(x, res);
      Jet<1> jx[1] = {Jet<1>(x[0], 0)}, jr[1];
      if (!functor_hello_world(jx, jr)) return false;
      res[0] = jr[0].a; jac[0] = jr[0].v[0];
      return true;
    }
    case SK_FUNCTOR_POWELL_F1: case SK_FUNCTOR_POWELL_F2: case SK_FUNCTOR_POWELL_F3: case SK_FUNCTOR_POWELL_F4:
    case SK_FUNCTOR_POWELL_ANALYTIC_F2: {
      if (!jac) {
        switch (id) {
          case SK_FUNCTOR_POWELL_F1: return functor_powell_f1(x, x + 1, res);
          case SK_FUNCTOR_POWELL_F2: return functor_powell_f2(x, x + 1, res);
          case SK_FUNCTOR_POWELL_F3: return functor_powell_f3(x, x + 1, res);
          case SK_FUNCTOR_POWELL_ANALYTIC_F2: return functor_powell_f2a(x, x + 1, res);
          default: return functor_powell_f4(x, x + 1, res);
        }
      }
      Jet<2> jx[2] = {Jet<2>(x[0], 0), Jet<2>(x[1], 1)}, jr[1];
      bool ok;
      switch (id) {
        case SK_FUNCTOR_POWELL_F1: ok = functor_powell_f1(jx, jx + 1, jr); break;
        case SK_FUNCTOR_POWELL_F2: ok = functor_powell_f2(jx, jx + 1, jr); break;
        case SK_FUNCTOR_POWELL_F3: ok = functor_powell_f3(jx, jx + 1, jr); break;
        case SK_FUNCTOR_POWELL_ANALYTIC_F2: ok = functor_powell_f2a(jx, jx + 1, jr); break;
        default: ok = functor_powell_f4(jx, jx + 1, jr); break;
      }
      if (!ok) return false;
      res[0] = jr[0].a; jac[0] = jr[0].v[0]; jac[1] = jr[0].v[1];
      return true;
    }
    case SK_FUNCTOR_TEST_BILINEAR_SCALAR: {
      if (!jac) return functor_bilinear_scalar(c, x, x + 2, res);
      Jet<4> jx[4], jr[1];
      for (int i = 0; i < 4; ++i) jx[i] = Jet<4>(x[i], i);
      if (!functor_bilinear_scalar(c, jx, jx + 2, jr)) return false;
      res[0] = jr[0].a; for (int i = 0; i < 4; ++i) jac[i] = jr[0].v[i];
      return true;
    }
    case SK_FUNCTOR_TEST_BILINEAR_VECTOR3: {
      if (!jac) return functor_bilinear_vector3(c, x, x + 2, res);
      Jet<4> jx[4], jr[3];
      for (int i = 0; i < 4; ++i) jx[i] = Jet<4>(x[i], i);
      if (!functor_bilinear_vector3(c, jx, jx + 2, jr)) return false;
      for (int k = 0; k < 3; ++k) { res[k] = jr[k].a; for (int i = 0; i < 4; ++i) jac[k * 4 + i] = jr[k].v[i]; }
      return true;
    }
    case SK_FUNCTOR_TEST_SUM10: {
      if (!jac) return functor_sum10(x, res);
      Jet<10> jx[10], jr[1];
      for (int i = 0; i < 10; ++i) jx[i] = Jet<10>(x[i], i);
      if (!functor_sum10(jx, jr)) return false;
      res[0] = jr[0].a; for (int i = 0; i < 10; ++i) jac[i] = jr[0].v[i];
      return true;
    }
  }
  return false;
}

// ---- LossFunction::Evaluate (ceres/loss_function.cc) and Corrector (corrector.cc) -------------
struct LossSpec { int type; double a; double b = 0.0; };   // b: second parameter (TolerantLoss only)

SK_HD void loss_evaluate(const LossSpec& l, double s, double* rho) {
  const double kMin = 2.2250738585072014e-308;
  if (l.type == SK_LOSS_HUBER) {
    const double b = l.a * l.a;
    if (s > b) { const double r = sqrt(s); rho[0] = 2.0 * l.a * r - b; rho[1] = fmax(kMin, l.a / r); rho[2] = -rho[1] / (2.0 * s); }
    else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
  } else if (l.type == SK_LOSS_CAUCHY) {
    const double b = l.a * l.a, c = 1.0 / b;
    const double sum = 1.0 + s * c, inv = 1.0 / sum;
    rho[0] = b * log(sum); rho[1] = fmax(kMin, inv); rho[2] = -c * (inv * inv);
  } else if (l.type == SK_LOSS_TOLERANT) {                  // ceres/loss_function.cc TolerantLoss::Evaluate
    const double c = l.b * log(1.0 + exp(-l.a / l.b));
    const double x = (s - l.a) / l.b;
    if (x > 36.7) { rho[0] = s - l.a - c; rho[1] = 1.0; rho[2] = 0.0; }     // e^x would swamp the 1 (kLog2Pow53)
    else {
      const double ex = exp(x);
      rho[0] = l.b * log(1.0 + ex) - c;
      rho[1] = fmax(kMin, ex / (1.0 + ex));
      rho[2] = 0.5 / (l.b * (1.0 + cosh(x)));
    }
  } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
}

struct Corrector {
  double sqrt_rho1, residual_scaling, alpha_sq_norm;
  SK_HD Corrector(double sq_norm, const double* rho) {
    sqrt_rho1 = sqrt(rho[1]);
    if (sq_norm == 0.0 || rho[2] <= 0.0) { residual_scaling = sqrt_rho1; alpha_sq_norm = 0.0; return; }
    const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1];
    const double alpha = 1.0 - ((D > 0.0) ? sqrt(D) : 0.0);
    residual_scaling = sqrt_rho1 / (1 - alpha);
    alpha_sq_norm = alpha / sq_norm;
  }
  // J: nrow x ncol row-major with leading dimension ld.
  SK_HD void correct_jacobian(int nrow, int ncol, int ld, const double* r, double* J) const {
    if (alpha_sq_norm == 0.0) { if (sqrt_rho1 != 1.0) for (int q = 0; q < nrow; ++q) for (int c = 0; c < ncol; ++c) J[q * ld + c] *= sqrt_rho1; return; }
    for (int c = 0; c < ncol; ++c) {
      double rtj = 0.0;
      for (int q = 0; q < nrow; ++q) rtj += J[q * ld + c] * r[q];
      for (int q = 0; q < nrow; ++q) J[q * ld + c] = sqrt_rho1 * (J[q * ld + c] - alpha_sq_norm * r[q] * rtj);
    }
  }
  SK_HD void correct_residuals(int n, double* r) const { for (int i = 0; i < n; ++i) r[i] *= residual_scaling; }
};

// ---- small SPD inverses (InvertPSDMatrix: llt().solve(Identity)) ------------------------------
// 3x3 symmetric, input upper triangle m = (m00, m01, m02, m11, m12, m22); output same packing.
SK_HD bool invert_spd3(const double* m, double* inv) {
  // Cholesky A = L L^T
  if (!(m[0] > 0.0)) return false;
  const double l00 = sqrt(m[0]);
  const double l10 = m[1] / l00, l20 = m[2] / l00;
  const double d1 = m[3] - l10 * l10;
  if (!(d1 > 0.0)) return false;
  const double l11 = sqrt(d1);
  const double l21 = (m[4] - l20 * l10) / l11;
  const double d2 = m[5] - l20 * l20 - l21 * l21;
  if (!(d2 > 0.0)) return false;
  const double l22 = sqrt(d2);
  // inverse of L (lower): Li
  const double i00 = 1.0 / l00, i11 = 1.0 / l11, i22 = 1.0 / l22;
  const double i10 = -l10 * i00 * i11;
  const double i21 = -l21 * i11 * i22;
  const double i20 = -(l20 * i00 + l21 * i10) * i22;
  // A^-1 = Li^T Li
  inv[0] = i00 * i00 + i10 * i10 + i20 * i20;
  inv[1] = i10 * i11 + i20 * i21;
  inv[2] = i20 * i22;
  inv[3] = i11 * i11 + i21 * i21;
  inv[4] = i21 * i22;
  inv[5] = i22 * i22;
  return true;
}

// n x n SPD (n <= 9), row-major full storage; in place inverse via Cholesky. Returns false if not PD.
template <int NMAX>
SK_HD bool invert_spd(double* A, int n) {
  double L[NMAX * NMAX];
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int k = 0; k < j; ++k) d -= L[j * NMAX + k] * L[j * NMAX + k];
    if (!(d > 0.0)) return false;
    d = sqrt(d);
    L[j * NMAX + j] = d;
    const double inv = 1.0 / d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= L[i * NMAX + k] * L[j * NMAX + k];
      L[i * NMAX + j] = s * inv;
    }
  }
  for (int c = 0; c < n; ++c) {
    double y[NMAX];
    for (int i = 0; i < n; ++i) {          // L y = e_c
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = 0; k < i; ++k) s -= L[i * NMAX + k] * y[k];
      y[i] = s / L[i * NMAX + i];
    }
    for (int i = n - 1; i >= 0; --i) {     // L^T x = y
      double s = y[i];
      for (int k = i + 1; k < n; ++k) s -= L[k * NMAX + i] * y[k];
      y[i] = s / L[i * NMAX + i];
    }
    for (int i = 0; i < n; ++i) A[i * n + c] = y[i];
  }
  return true;
}

}  // namespace sk
