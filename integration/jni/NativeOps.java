package com.linkedin.mlease.regression.gpu;

import java.nio.ByteBuffer;

/**
 * Stateless entry points of include/mlease_b200.h behind the remaining jobs of the path (INTEGRATION.md section 1):
 * RegressionNaiveTrain's reducer, RegressionTest's scoring, RegressionTestLoglik.  Buffers are direct ByteBuffers in native byte
 * order (CSR arrays as mlease_rows_get or the Java-side ingest fills them); null = absent.
 */
public final class NativeOps {
  static { System.loadLibrary("mlease_b200_jni"); }
  private NativeOps() {}

  /** NaiveReducer.reduce for every "lambda#key" at once (jobs/RegressionNaiveTrain.java:348-415): key k owns rows
   *  [keyRowStart[k], keyRowStart[k+1]) of one CSR; outModel is [numLambdas][numKeys][numFeatures + 1], intercept last;
   *  skipped[k] = 1 where a key has fewer rows than data.size.threshold.  Throws IOException("Model fitting error!") like :400-412. */
  public static native void naiveTrain(int device, int numKeys, int numFeatures, ByteBuffer keyRowStartI64, ByteBuffer rowptrI64,
                                       ByteBuffer colidxI32, ByteBuffer valsF32, ByteBuffer responseI32, ByteBuffer weightF32OrNull,
                                       ByteBuffer offsetF32OrNull, float[] lambdas, float[] lambdaMapOrNull, float priorMean,
                                       boolean penalizeIntercept, boolean hasIntercept, int dataSizeThreshold, boolean binaryFeature,
                                       double[] outModel, int[] skipped) throws java.io.IOException;

  /** pred = float(offset + interceptTerm + x.beta) per record (models/LinearModel.java:241-257; jobs/RegressionTest.java:163);
   *  model has numFeatures + 1 entries, intercept last. */
  public static native void score(int device, int numFeatures, long nrows, ByteBuffer rowptrI64, ByteBuffer colidxI32, ByteBuffer valsF32,
                                  ByteBuffer offsetF32OrNull, double[] model, int numClickReplicates, boolean binaryFeature,
                                  float[] predOut) throws java.io.IOException;

  /** RegressionTestLoglik's mapper / combiner / reducer arithmetic with its float casts (jobs/RegressionTestLoglik.java:124-200):
   *  returns the float average log-likelihood, countOut[0] = the double count. */
  public static native float testLoglik(int device, int[] response, float[] pred, float[] weight, long combinerBlock,
                                        double[] countOut) throws java.io.IOException;
}
