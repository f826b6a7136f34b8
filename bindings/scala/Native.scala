// Native.scala -- JVM side of bindings/jni/skeres_jni.c.  UNTESTED (no JDK / Scala in the build image); written against
// include/skeres.h.  Replaces the SWIG-generated com.google.ceres classes (ceres.i) on the hot path; the library is loaded
// under the name the reference already loads (ceres.i:213-223), libceres is no longer needed.
package com.google.ceres

object Native {
  System.loadLibrary("skeres")       // the C ABI (CUDA kernels inside)
  System.loadLibrary("skeres_jni")   // this shim

  @native def doubleArrayCreate(n: Long): Long
  @native def doubleArrayDestroy(h: Long): Unit
  @native def doubleArraySize(h: Long): Long
  @native def doubleArrayGet(h: Long, i: Long): Double
  @native def doubleArraySet(h: Long, i: Long, v: Double): Unit
  @native def doubleArrayUpload(h: Long, offset: Long, src: Array[Double]): Unit
  @native def doubleArrayDownload(h: Long, offset: Long, dst: Array[Double]): Unit

  @native def lossTrivial(): Long
  @native def lossHuber(a: Double): Long
  @native def lossCauchy(a: Double): Long
  @native def lossTolerant(a: Double, b: Double): Long
  @native def lossDestroy(h: Long): Unit

  @native def costFunctionCreate(functorId: Int, consts: Array[Double]): Long
  @native def costFunctionDestroy(h: Long): Unit
  @native def functorRegisterSource(name: String, cudaSource: String, numResiduals: Int, blockSizes: Array[Int], numConsts: Int): Int
  @native def costFunctionEvaluate(h: Long, paramArrays: Array[Long], paramOffsets: Array[Long], residualArray: Long,
                                   residualOffset: Long, jacobianArrays: Array[Long], jacobianOffsets: Array[Long]): Boolean

  @native def problemCreate(): Long
  @native def problemDestroy(h: Long): Unit
  @native def addResidualBlock(problem: Long, cost: Long, loss: Long, arrays: Array[Long], offsets: Array[Long]): Long
  @native def addResidualBlocks(problem: Long, functorId: Int, consts: Array[Double], loss: Long, array: Long,
                                offsets: Array[Long]): Long
  @native def addParameterBlocks(problem: Long, array: Long, offsets: Array[Long], size: Int): Unit
  @native def problemNumResidualBlocks(h: Long): Long
  @native def problemNumParameterBlocks(h: Long): Long

  @native def optionsCreate(): Long
  @native def optionsDestroy(h: Long): Unit
  @native def optionsSetLinearSolverType(h: Long, v: Int): Unit
  @native def optionsSetPreconditionerType(h: Long, v: Int): Unit
  @native def optionsSetMinimizerType(h: Long, v: Int): Unit
  @native def optionsSetMaxNumIterations(h: Long, v: Int): Unit
  @native def optionsSetMinimizerProgressToStdout(h: Long, v: Boolean): Unit
  @native def optionsSetFunctionTolerance(h: Long, v: Double): Unit
  @native def optionsSetComm(h: Long, comm: Long): Unit
  @native def optionsSetResidualBlocksAreLocal(h: Long, v: Boolean): Unit

  @native def summaryCreate(): Long
  @native def summaryDestroy(h: Long): Unit
  @native def summaryBriefReport(h: Long): String
  @native def summaryFullReport(h: Long): String
  @native def summaryMessage(h: Long): String
  @native def summaryFinalCost(h: Long): Double
  @native def summaryTerminationType(h: Long): Int
  @native def solve(options: Long, problem: Long, summary: Long): Unit
  @native def initGoogleLogging(name: String): Unit

  @native def setDevice(ordinal: Int): Unit
  @native def commUniqueId(): Array[Byte]
  @native def commCreate(id: Array[Byte], rank: Int, world: Int): Long
  @native def commDestroy(h: Long): Unit
}
