package com.linkedin.mlease.regression.gpu;

import java.nio.ByteBuffer;

/**
 * Thin JNI view of the mlease_world_* entry points of include/mlease_b200.h: all GPUs of the box behind one object.
 * Reference-side source (INTEGRATION.md section 2): drop it next to the jobs, build mlease_b200_jni.c against the JDK's jni.h
 * and libmlease_b200.so, and replace the body of RegressionAdmmTrain.run()'s iteration loop as INTEGRATION.md section 3 shows.
 * (This build image has no JDK: the C side is type-checked against a stub jni.h by tests/test_abi.py, the Java side is not compiled here.)
 */
public final class NativeAdmm implements AutoCloseable {
  static { System.loadLibrary("mlease_b200_jni"); }          // links libmlease_b200.so (which dlopens libnccl.so.2 when N > 1)
  private long handle;                                       // mlease_world*, read by the native side

  public NativeAdmm(int[] devices, int numBlocks, int numFeatures, float[] lambdas, float[] rhosOrNull,
                    float[] lambdaMapOrNull, int regularizer, boolean penalizeIntercept, double epsilon,
                    float rhoAdaptCoefficient, boolean aggressiveDecay, boolean binaryFeature) throws java.io.IOException {
    handle = create(devices, numBlocks, numFeatures, lambdas, rhosOrNull, lambdaMapOrNull, regularizer, penalizeIntercept,
                    epsilon, rhoAdaptCoefficient, aggressiveDecay, binaryFeature);
  }
  // direct ByteBuffers (native byte order): zero-copy host pointers for the one-time upload
  public native void addPartitionCsr(int partitionId, long nrows, ByteBuffer rowptrI64, ByteBuffer colidxI32, ByteBuffer valsF32,
                                     ByteBuffer responseI32, ByteBuffer weightF32OrNull, ByteBuffer offsetF32OrNull) throws java.io.IOException;
  public native void begin() throws java.io.IOException;                                     // mlease_world_begin
  public native void beginInitialized(double[] z0, float boostRate) throws java.io.IOException;   // initialize.boost.rate > 0 (RegressionAdmmTrain.java:236-266)
  public native int  fitPartition(int partitionId, double[] xInOut, double[] priorMean, double[] priorPrecision) throws java.io.IOException;
  public native boolean iterate(double[] maxdiffOut) throws java.io.IOException;             // one ADMM iteration on all GPUs; true = the reference would break (:493-496)
  public native int  run(int numIters) throws java.io.IOException;                           // whole loop in C when no per-iteration files are wanted
  public native void getZ(int lambdaIdx, double[] out) throws java.io.IOException;           // driver z (double), length numFeatures + 1, intercept last
  public native void getFinalModel(int lambdaIdx, float[] out) throws java.io.IOException;
  public native void getX(int partitionId, int lambdaIdx, double[] out) throws java.io.IOException;
  public native void getU(int partitionId, int lambdaIdx, float[] out) throws java.io.IOException;
  public native void getUplusx(int partitionId, int lambdaIdx, float[] out) throws java.io.IOException;
  @Override public native void close();
  private static native long create(int[] devices, int numBlocks, int numFeatures, float[] lambdas, float[] rhos, float[] lambdaMap,
                                    int regularizer, boolean penalizeIntercept, double epsilon, float rhoAdaptCoefficient,
                                    boolean aggressiveDecay, boolean binaryFeature) throws java.io.IOException;
}
