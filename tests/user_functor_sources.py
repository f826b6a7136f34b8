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
