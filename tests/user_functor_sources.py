"""CUDA sources of functors handed to sk_functor_register_source by the tests: the three functors of the reference's
AutodiffCostFuntionSpec.scala and CurveFitting.scala's ExponentialResidual, written the way a user of the reference would port
the body of `apply[T](x: Array[T]*)`."""

BILINEAR_SCALAR = """
// AutodiffCostFuntionSpec.scala:14-26   case class BinaryScalarCost(a: Double) extends AutoDiffCostFunctor(1, 2, 2)
template <class T> __device__ bool BinaryScalarCost(const double* consts, T const* const* p, T* z) {
  const T* x = p[0]; const T* y = p[1];
  z[0] = x[0] * y[0] + x[1] * y[1] - consts[0];
  return true;
}
"""

BILINEAR_VECTOR3 = """
// AutodiffCostFuntionSpec.scala:55-69   case class BinaryVectorCost(a: Double) extends AutoDiffCostFunctor(3, 2, 2)
template <class T> __device__ bool BinaryVectorCost(const double* consts, T const* const* p, T* z) {
  const double a = consts[0];
  const T* x = p[0]; const T* y = p[1];
  z[0] = x[0] * y[0] + x[1] * y[1] - a;
  z[1] = x[0] * y[0] - x[1] * y[1] + a;
  z[2] = x[0] * x[1] + y[0] * y[1] + 10.0 * a;
  return true;
}
"""

SUM10 = """
// AutodiffCostFuntionSpec.scala:110-119   object TenParamsCost extends AutoDiffCostFunctor(1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1)
template <class T> __device__ bool TenParamsCost(const double* consts, T const* const* p, T* z) {
  T s = p[0][0];
  for (int i = 1; i < 10; ++i) s = s + p[i][0];
  z[0] = s;
  return true;
}
"""

EXPONENTIAL = """
// CurveFitting.scala:92-98   class ExponentialResidual(x: Double, y: Double) extends AutoDiffCostFunctor(1, 1, 1)
template <class T> __device__ bool UserExponentialResidual(const double* consts, T const* const* p, T* residual) {
  const double x = consts[0], y = consts[1];
  const T m = p[0][0], c = p[1][0];
  residual[0] = y - exp(m * x + c);
  return true;
}
"""

# name, source, kNumResiduals, N, num_consts -- in the order of tests/golden/autodiff_spec_vectors.json
SPEC = [("BinaryScalarCost", BILINEAR_SCALAR, 1, [2, 2], 1), ("BinaryVectorCost", BILINEAR_VECTOR3, 3, [2, 2], 1),
        ("TenParamsCost", SUM10, 1, [1] * 10, 0)]

SNAVELY = """
// SimpleBundleAdjuster.scala:79-119   class SnavelyReprojectionError(observedX, observedY) extends AutoDiffCostFunctor(2, 9, 3)
// Rotation.angleAxisRotatePoint (core/.../Rotation.scala:449-522) is part of the library on both sides: sk::angle_axis_rotate_point.
template <class T> __device__ bool UserSnavelyReprojectionError(const double* consts, T const* const* params, T* residuals) {
  const double observedX = consts[0], observedY = consts[1];
  const T* camera = params[0];
  const T* point = params[1];
  T p[3];
  sk::angle_axis_rotate_point(camera, point, p);   // camera(0, 1, 2) are the angle-axis rotation
  p[0] = p[0] + camera[3];                          // camera(3, 4, 5) are the translation
  p[1] = p[1] + camera[4];
  p[2] = p[2] + camera[5];
  const T xp = (-p[0]) / p[2];                      // Bundler's camera looks down the negative z axis
  const T yp = (-p[1]) / p[2];
  const T l1 = camera[7], l2 = camera[8];
  const T r2 = xp * xp + yp * yp;
  const T distortion = 1.0 + r2 * (l1 + l2 * r2);
  const T focal = camera[6];
  residuals[0] = focal * distortion * xp - observedX;
  residuals[1] = focal * distortion * yp - observedY;
  return true;
}
"""

DIVISION_MODEL = """
// A camera model the library has never seen: the same (2; 9, 3) shape with the DIVISION model of radial distortion,
// distortion = 1 / (1 + l1 r^2 + l2 r^4), and a functor that reports failure for a point behind the camera.
template <class T> __device__ bool DivisionModelReprojectionError(const double* consts, T const* const* params, T* residuals) {
  const T* camera = params[0];
  const T* point = params[1];
  T p[3];
  sk::angle_axis_rotate_point(camera, point, p);
  p[0] = p[0] + camera[3]; p[1] = p[1] + camera[4]; p[2] = p[2] + camera[5];
  if (!((-p[2]) > 0.0)) return false;
  const T xp = (-p[0]) / p[2];
  const T yp = (-p[1]) / p[2];
  const T r2 = xp * xp + yp * yp;
  const T distortion = 1.0 / (1.0 + r2 * (camera[7] + camera[8] * r2));
  residuals[0] = camera[6] * distortion * xp - consts[0];
  residuals[1] = camera[6] * distortion * yp - consts[1];
  return true;
}
"""

# bundle-adjustment shaped functors: compiled into the tile evaluation kernel of the Schur solvers as well
BA_SHAPED = [("UserSnavelyReprojectionError", SNAVELY, 2, [9, 3], 2), ("DivisionModelReprojectionError", DIVISION_MODEL, 2, [9, 3], 2)]
