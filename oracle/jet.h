// oracle/jet.h — TEST INFRASTRUCTURE ONLY (see oracle/README.md). Not linked into libskeres.so.
//
// CPU restatement of the reference's per-residual arithmetic:
//   * forward-mode dual numbers with the semantics of spire 0.11.0 `Jet` (build.sbt:3), the type
//     AutodiffCostFunction.scala:102 seeds with Jet(x, k);  ordering compares real parts only
//     (package.scala:27);
//   * Rotation.angleAxisRotatePoint            core/.../Rotation.scala:449-522
//   * Rotation.angleAxisToRotationMatrix        core/.../Rotation.scala:211-255 (test helper only)
//   * SnavelyReprojectionError.apply            examples/.../SimpleBundleAdjuster.scala:79-119
//   * ExponentialResidual.apply                 examples/.../CurveFitting.scala:92-98
//   * the three functors of                     core/src/test/.../AutodiffCostFuntionSpec.scala
// Operation order follows the Scala sources statement by statement so that the oracle is a
// restatement, not a re-derivation.
#pragma once
#include <cmath>

namespace oracle {

template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0.0) { for (int i = 0; i < N; ++i) v[i] = 0.0; }
  Jet(double real) : a(real) { for (int i = 0; i < N; ++i) v[i] = 0.0; }  // implicit scalar lift
  Jet(double real, int k) : a(real) { for (int i = 0; i < N; ++i) v[i] = 0.0; v[k] = 1.0; }
};

template <int N> inline Jet<N> operator+(const Jet<N>& x, const Jet<N>& y) {
  Jet<N> r; r.a = x.a + y.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] + y.v[i]; return r; }
template <int N> inline Jet<N> operator-(const Jet<N>& x, const Jet<N>& y) {
  Jet<N> r; r.a = x.a - y.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] - y.v[i]; return r; }
template <int N> inline Jet<N> operator-(const Jet<N>& x) {
  Jet<N> r; r.a = -x.a; for (int i = 0; i < N; ++i) r.v[i] = -x.v[i]; return r; }
// spire Jet.*: Jet(real * b.real, (b.real *: infinitesimal) + (real *: b.infinitesimal))
template <int N> inline Jet<N> operator*(const Jet<N>& x, const Jet<N>& y) {
  Jet<N> r; r.a = x.a * y.a; for (int i = 0; i < N; ++i) r.v[i] = y.a * x.v[i] + x.a * y.v[i]; return r; }
// spire Jet./: br_inv = 1/b.real; ar_div_br = real*br_inv; (infinitesimal - ar_div_br*b.inf)*br_inv
template <int N> inline Jet<N> operator/(const Jet<N>& x, const Jet<N>& y) {
  Jet<N> r; const double binv = 1.0 / y.a; const double q = x.a * binv; r.a = q;
  for (int i = 0; i < N; ++i) r.v[i] = binv * (x.v[i] - q * y.v[i]); return r; }
template <int N> inline Jet<N> operator+(const Jet<N>& x, double s) { Jet<N> r = x; r.a = x.a + s; return r; }
template <int N> inline Jet<N> operator+(double s, const Jet<N>& x) { Jet<N> r = x; r.a = s + x.a; return r; }
template <int N> inline Jet<N> operator-(const Jet<N>& x, double s) { Jet<N> r = x; r.a = x.a - s; return r; }
template <int N> inline Jet<N> operator-(double s, const Jet<N>& x) { Jet<N> r; r.a = s - x.a; for (int i = 0; i < N; ++i) r.v[i] = -x.v[i]; return r; }
template <int N> inline Jet<N> operator*(const Jet<N>& x, double s) { Jet<N> r; r.a = x.a * s; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * s; return r; }
template <int N> inline Jet<N> operator*(double s, const Jet<N>& x) { return x * s; }
template <int N> inline Jet<N> operator/(const Jet<N>& x, double s) { Jet<N> r; const double si = 1.0 / s; r.a = x.a * si; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * si; return r; }
template <int N> inline Jet<N> operator/(double s, const Jet<N>& y) { return Jet<N>(s) / y; }
template <int N> inline bool operator>(const Jet<N>& x, const Jet<N>& y) { return x.a > y.a; }  // package.scala:27
template <int N> inline bool operator>(const Jet<N>& x, double y) { return x.a > y; }

template <int N> inline Jet<N> sqrt(const Jet<N>& x) {
  Jet<N> r; const double sa = std::sqrt(x.a); const double h = 1.0 / (2.0 * sa); r.a = sa;
  for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * h; return r; }
template <int N> inline Jet<N> cos(const Jet<N>& x) {
  Jet<N> r; r.a = std::cos(x.a); const double d = -std::sin(x.a);
  for (int i = 0; i < N; ++i) r.v[i] = d * x.v[i]; return r; }
template <int N> inline Jet<N> sin(const Jet<N>& x) {
  Jet<N> r; r.a = std::sin(x.a); const double d = std::cos(x.a);
  for (int i = 0; i < N; ++i) r.v[i] = d * x.v[i]; return r; }
template <int N> inline Jet<N> exp(const Jet<N>& x) {
  Jet<N> r; const double e = std::exp(x.a); r.a = e;
  for (int i = 0; i < N; ++i) r.v[i] = e * x.v[i]; return r; }
inline double sqrt(double x) { return std::sqrt(x); }
inline double cos(double x) { return std::cos(x); }
inline double sin(double x) { return std::sin(x); }
inline double exp(double x) { return std::exp(x); }

inline double real_part(double x) { return x; }
template <int N> inline double real_part(const Jet<N>& x) { return x.a; }

constexpr double kEpsilonDouble = 2.220446049250313e-16;  // Math.ulp(1.0), package.scala:15

// Rotation.scala:449-522
template <class T>
inline void angleAxisRotatePoint(const T* angleAxis, const T* pt, T* result) {
  // dotProduct (Rotation.scala:445-446) = sum of products accumulated from zero, in index order.
  const T theta2 = angleAxis[0] * angleAxis[0] + angleAxis[1] * angleAxis[1] + angleAxis[2] * angleAxis[2];
  if (theta2 > kEpsilonDouble) {                                  // :456
    const T theta = sqrt(theta2);                                 // :467
    const T cosTheta = cos(theta);
    const T sinTheta = sin(theta);
    const T thetaInverse = 1.0 / theta;                           // :470
    const T w[3] = {angleAxis[0] * thetaInverse, angleAxis[1] * thetaInverse, angleAxis[2] * thetaInverse};
    const T wCrossPt[3] = {w[1] * pt[2] - w[2] * pt[1],           // :480-484
                           w[2] * pt[0] - w[0] * pt[2],
                           w[0] * pt[1] - w[1] * pt[0]};
    const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (1.0 - cosTheta);  // :485
    result[0] = pt[0] * cosTheta + wCrossPt[0] * sinTheta + w[0] * tmp;             // :486-490
    result[1] = pt[1] * cosTheta + wCrossPt[1] * sinTheta + w[1] * tmp;
    result[2] = pt[2] * cosTheta + wCrossPt[2] * sinTheta + w[2] * tmp;
  } else {                                                        // :492-520, R*pt = pt + w x pt
    const T wCrossPt[3] = {angleAxis[1] * pt[2] - angleAxis[2] * pt[1],
                           angleAxis[2] * pt[0] - angleAxis[0] * pt[2],
                           angleAxis[0] * pt[1] - angleAxis[1] * pt[0]};
    result[0] = pt[0] + wCrossPt[0];
    result[1] = pt[1] + wCrossPt[1];
    result[2] = pt[2] + wCrossPt[2];
  }
}

// Rotation.scala:211-255, column-major 3x3 (ColumnMajorMatrixAdapter3x3, :206-209): R(i,j)=R[i+3j].
inline void angleAxisToRotationMatrix(const double* aa, double* R) {
  const double theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (theta2 > kEpsilonDouble) {
    const double theta = std::sqrt(theta2);
    const double wx = aa[0] / theta, wy = aa[1] / theta, wz = aa[2] / theta;
    const double c = std::cos(theta), s = std::sin(theta);
    R[0 + 3 * 0] = c + wx * wx * (1.0 - c);
    R[1 + 3 * 0] = wz * s + wx * wy * (1.0 - c);
    R[2 + 3 * 0] = -wy * s + wx * wz * (1.0 - c);
    R[0 + 3 * 1] = wx * wy * (1.0 - c) - wz * s;
    R[1 + 3 * 1] = c + wy * wy * (1.0 - c);
    R[2 + 3 * 1] = wx * s + wy * wz * (1.0 - c);
    R[0 + 3 * 2] = wy * s + wx * wz * (1.0 - c);
    R[1 + 3 * 2] = -wx * s + wy * wz * (1.0 - c);
    R[2 + 3 * 2] = c + wz * wz * (1.0 - c);
  } else {
    R[0] = 1.0; R[1] = aa[2]; R[2] = -aa[1];
    R[3] = -aa[2]; R[4] = 1.0; R[5] = aa[0];
    R[6] = aa[1]; R[7] = -aa[0]; R[8] = 1.0;
  }
}

// SimpleBundleAdjuster.scala:79-119. camera = params[0] (9), point = params[1] (3).
template <class T>
inline bool snavelyReprojectionError(const double* consts, T const* const* params, T* residuals) {
  const T* camera = params[0];
  const T* point = params[1];
  T p[3];
  angleAxisRotatePoint(camera, point, p);          // :91-92
  p[0] = p[0] + camera[3];                         // :95-97
  p[1] = p[1] + camera[4];
  p[2] = p[2] + camera[5];
  const T xp = (-p[0]) / p[2];                     // :102-103
  const T yp = (-p[1]) / p[2];
  const T l1 = camera[7];                          // :106-109
  const T l2 = camera[8];
  const T r2 = xp * xp + yp * yp;
  const T distortion = 1.0 + r2 * (l1 + l2 * r2);
  const T focal = camera[6];                       // :112-114
  const T predictedX = focal * distortion * xp;
  const T predictedY = focal * distortion * yp;
  residuals[0] = predictedX - consts[0];           // :117
  residuals[1] = predictedY - consts[1];
  return true;
}

// NOT in the reference: a camera model of the bundle-adjustment shape (2; 9, 3) the GPU library has no built-in for -- the
// division model of radial distortion; fails for a point behind the camera.  The GPU tests hand the library the same body as
// CUDA source (tests/user_functor_sources.py: DIVISION_MODEL, sk_functor_register_source) and check its tile kernels
// against this one.
template <class T>
inline bool divisionModelReprojectionError(const double* consts, T const* const* params, T* residuals) {
  const T* camera = params[0];
  const T* point = params[1];
  T p[3];
  angleAxisRotatePoint(camera, point, p);
  p[0] = p[0] + camera[3];
  p[1] = p[1] + camera[4];
  p[2] = p[2] + camera[5];
  if (!((-p[2]) > 0.0)) return false;
  const T xp = (-p[0]) / p[2];
  const T yp = (-p[1]) / p[2];
  const T r2 = xp * xp + yp * yp;
  const T distortion = 1.0 / (1.0 + r2 * (camera[7] + camera[8] * r2));
  residuals[0] = camera[6] * distortion * xp - consts[0];
  residuals[1] = camera[6] * distortion * yp - consts[1];
  return true;
}

// CurveFitting.scala:92-98. consts = (x, y); params = m (1), c (1).
template <class T>
inline bool exponentialResidual(const double* consts, T const* const* params, T* residuals) {
  const T* m = params[0];
  const T* c = params[1];
  residuals[0] = consts[1] - exp(m[0] * consts[0] + c[0]);
  return true;
}

// AutodiffCostFuntionSpec.scala:14-26
template <class T>
inline bool testBilinearScalar(const double* consts, T const* const* p, T* z) {
  const T* x = p[0]; const T* y = p[1];
  z[0] = x[0] * y[0] + x[1] * y[1] - consts[0];
  return true;
}
// AutodiffCostFuntionSpec.scala:55-69
template <class T>
inline bool testBilinearVector3(const double* consts, T const* const* p, T* z) {
  const T* x = p[0]; const T* y = p[1]; const double a = consts[0];
  z[0] = x[0] * y[0] + x[1] * y[1] - a;
  z[1] = x[0] * y[0] - x[1] * y[1] + a;
  z[2] = x[0] * x[1] + y[0] * y[1] + 10 * a;
  return true;
}
// HelloWorld.scala:11-14
template <class T>
inline bool helloWorld(const double*, T const* const* x, T* res) { res[0] = 10.0 - x[0][0]; return true; }
// Powell.scala:13-51 (F2 as the code computes it, :28: sqrt(5) * x3 - x4)
template <class T>
inline bool powellF1(const double*, T const* const* x, T* res) { res[0] = x[0][0] + 10.0 * x[1][0]; return true; }
template <class T>
inline bool powellF2(const double*, T const* const* x, T* res) { res[0] = std::sqrt(5.0) * x[0][0] - x[1][0]; return true; }
template <class T>   // PowellAnalytic.scala:36
inline bool powellF2a(const double*, T const* const* x, T* res) { res[0] = std::sqrt(5.0) * (x[0][0] - x[1][0]); return true; }
template <class T>
inline bool powellF3(const double*, T const* const* x, T* res) { const T d = x[0][0] - 2.0 * x[1][0]; res[0] = d * d; return true; }
template <class T>
inline bool powellF4(const double*, T const* const* x, T* res) { const T d = x[0][0] - x[1][0]; res[0] = (std::sqrt(10.0) * d) * d; return true; }

// AutodiffCostFuntionSpec.scala:110-119
template <class T>
inline bool testSum10(const double*, T const* const* p, T* z) {
  T s = p[0][0];
  for (int i = 1; i < 10; ++i) s = s + p[i][0];
  z[0] = s;
  return true;
}

}  // namespace oracle
