// Facade.scala -- the reference's public names over Native (UNTESTED, see Native.scala).  Callers of
// org.somelightprojections.skeres keep writing `new Problem`, `problem.addResidualBlock(cost, loss, x.toPointer)`,
// `functor.toAutoDiffCostFunction`, `new Solver.Options`, `ceres.solve(options, problem, summary)`.
package com.google.ceres

/** DoubleArray (ceres.i:95-96): device-resident; freed explicitly, not by a finaliser. */
final class DoubleArray(val n: Long) extends AutoCloseable {
  val handle: Long = Native.doubleArrayCreate(n)
  def get(i: Int): Double = Native.doubleArrayGet(handle, i)                    // RichDoubleArray.get
  def set(i: Int, v: Double): Unit = Native.doubleArraySet(handle, i, v)       // RichDoubleArray.set
  def toPointer: DoublePointer = DoublePointer(this, 0)                        // RichDoubleArray.toPointer
  def slice(offset: Long): DoublePointer = DoublePointer(this, offset)         // RichDoubleArray.slice (:52)
  def copyFrom(a: Array[Double]): Unit = Native.doubleArrayUpload(handle, 0, a)          // ONE crossing (:60-66)
  def toArray(length: Int): Array[Double] = { val a = new Array[Double](length); Native.doubleArrayDownload(handle, 0, a); a }
  override def close(): Unit = Native.doubleArrayDestroy(handle)
}

/** package.scala:11 `type DoublePointer = SWIGTYPE_p_double` becomes (array, offset): an interior pointer (ceres.i:99-107). */
final case class DoublePointer(array: DoubleArray, offset: Long) {
  def slice(k: Long): DoublePointer = DoublePointer(array, offset + k)
  def isNull: Boolean = array == null
}

final class LossFunction private[ceres] (val handle: Long)
object PredefinedLossFunctions {                                               // ceres.i:160-184
  def trivialLoss: LossFunction = new LossFunction(Native.lossTrivial())
  def huberLoss(a: Double): LossFunction = new LossFunction(Native.lossHuber(a))
  def cauchyLoss(a: Double): LossFunction = new LossFunction(Native.lossCauchy(a))
  def tolerantLoss(a: Double, b: Double): LossFunction = new LossFunction(Native.lossTolerant(a, b))
}

/** AutoDiffCostFunction of a registered device functor (AutodiffCostFunction.scala:68). */
final class CostFunction private[ceres] (val handle: Long) {
  /** AutodiffCostFunction.scala:74-78; jacobians == null <=> isNull, a null entry <=> getRow(i).isNull. */
  def evaluate(parameters: Seq[DoublePointer], residuals: DoublePointer, jacobians: Seq[DoublePointer]): Boolean =
    Native.costFunctionEvaluate(handle, parameters.map(_.array.handle).toArray, parameters.map(_.offset).toArray,
      residuals.array.handle, residuals.offset,
      if (jacobians == null) null else jacobians.map(j => if (j == null || j.isNull) 0L else j.array.handle).toArray,
      if (jacobians == null) null else jacobians.map(j => if (j == null) 0L else j.offset).toArray)
}

/** CostFunctor.scala:40-51.  `functorId` names the device functor (include/skeres.h sk_functor_id): SnavelyReprojectionError
  * = 1 with consts (observedX, observedY); ExponentialResidual = 2 with (x, y); HelloCostFunctor = 3; Powell F1..F4 = 4..7.
  * The generic `apply[T]` body stays for JVM-side tests; it is not what runs in a solve.  A functor without an id fails with
  * the library's message -- there is no up-call and no CPU fallback. */
abstract class AutoDiffCostFunctor(val kNumResiduals: Int, val N: Int*) {
  def functorId: Int
  def consts: Array[Double] = Array.empty
  def toAutoDiffCostFunction: CostFunction = new CostFunction(Native.costFunctionCreate(functorId, consts))
}

/** A functor whose body is handed to the library as CUDA source (sk_functor_register_source) and compiled once for the device:
  * the port of `def apply[T](x: Array[T]*): Array[T]` is
  * `template <class T> __device__ bool Name(const double* consts, T const* const* x, T* residuals)`.
  * {{{
  *   val Exp = SourceCostFunctor.define("ExponentialResidual", src, 1, Seq(1, 1), numConsts = 2)
  *   problem.addResidualBlock(Exp(x, y).toAutoDiffCostFunction, loss, m.toPointer, c.toPointer)
  * }}} */
object SourceCostFunctor {
  final class Defined(val functorId: Int, kNumResiduals: Int, N: Seq[Int], numConsts: Int) {
    def apply(cs: Double*): AutoDiffCostFunctor = {
      require(cs.length == numConsts, s"functor takes $numConsts constants")
      val id = functorId
      new AutoDiffCostFunctor(kNumResiduals, N: _*) { def functorId: Int = id; override def consts: Array[Double] = cs.toArray }
    }
  }
  def define(name: String, cudaSource: String, kNumResiduals: Int, N: Seq[Int], numConsts: Int = 0): Defined =
    new Defined(Native.functorRegisterSource(name, cudaSource, kNumResiduals, N.toArray, numConsts), kNumResiduals, N, numConsts)
}

/** Problem.scala:16-33. */
final class Problem extends AutoCloseable {
  val handle: Long = Native.problemCreate()
  private val keepAlive = scala.collection.mutable.ArrayBuffer.empty[AnyRef]   // Problem.scala:29-32: DO_NOT_TAKE_OWNERSHIP
  def addResidualBlock(cost: CostFunction, loss: LossFunction, x: DoublePointer*): Long = {
    keepAlive += cost; keepAlive += loss
    Native.addResidualBlock(handle, cost.handle, if (loss == null) 0L else loss.handle, x.map(_.array.handle).toArray, x.map(_.offset).toArray)
  }
  /** The loop of SimpleBundleAdjuster.scala:139-145 as one call: `offsets` is n x (number of blocks), row-major. */
  def addResidualBlocks(functorId: Int, consts: Array[Double], loss: LossFunction, array: DoubleArray, offsets: Array[Long]): Long = {
    keepAlive += loss; keepAlive += array
    Native.addResidualBlocks(handle, functorId, consts, if (loss == null) 0L else loss.handle, array.handle, offsets)
  }
  /** Problem::AddParameterBlock in bulk (rank-local multi-GPU mode: every rank declares all cameras). */
  def addParameterBlocks(array: DoubleArray, offsets: Array[Long], size: Int): Unit = Native.addParameterBlocks(handle, array.handle, offsets, size)
  override def close(): Unit = Native.problemDestroy(handle)
}

object Solver {
  final class Options extends AutoCloseable {                                  // setters used by the examples (SURVEY 8 a11)
    val handle: Long = Native.optionsCreate()
    def setLinearSolverType(t: Int): Unit = Native.optionsSetLinearSolverType(handle, t)
    def setPreconditionerType(t: Int): Unit = Native.optionsSetPreconditionerType(handle, t)
    def setMinimizerType(t: Int): Unit = Native.optionsSetMinimizerType(handle, t)
    def setMaxNumIterations(n: Int): Unit = Native.optionsSetMaxNumIterations(handle, n)
    def setMinimizerProgressToStdout(b: Boolean): Unit = Native.optionsSetMinimizerProgressToStdout(handle, b)
    def setFunctionTolerance(v: Double): Unit = Native.optionsSetFunctionTolerance(handle, v)
    def setComm(comm: Long): Unit = Native.optionsSetComm(handle, comm)
    def setResidualBlocksAreLocal(b: Boolean): Unit = Native.optionsSetResidualBlocksAreLocal(handle, b)
    override def close(): Unit = Native.optionsDestroy(handle)
  }
  final class Summary extends AutoCloseable {
    val handle: Long = Native.summaryCreate()
    def briefReport: String = Native.summaryBriefReport(handle)
    def fullReport: String = Native.summaryFullReport(handle)
    def message: String = Native.summaryMessage(handle)
    def finalCost: Double = Native.summaryFinalCost(handle)
    def terminationType: Int = Native.summaryTerminationType(handle)
    override def close(): Unit = Native.summaryDestroy(handle)
  }
}

object ceres {
  def initGoogleLogging(name: String): Unit = Native.initGoogleLogging(name)
  def solve(options: Solver.Options, problem: Problem, summary: Solver.Summary): Unit = Native.solve(options.handle, problem.handle, summary.handle)
}
