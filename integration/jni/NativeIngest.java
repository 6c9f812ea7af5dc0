package com.linkedin.mlease.regression.gpu;

import java.nio.ByteBuffer;

/**
 * The job layer's block-parallel avro ingest (include/mlease_host.h, libmlease_host.so) for a Java host that does not want to
 * decode the data files itself (INTEGRATION.md section 5): RegressionPrepareOutput records (raw = false; a file or a directory)
 * or raw test records (raw = true; one file) -> CSR arrays with global feature ids in first-seen order, which is what a sequential
 * pass through LibLinearDataset.addInstanceAvro assigns (llf/LibLinearDataset.java:456-479).
 */
public final class NativeIngest implements AutoCloseable {
  static { System.loadLibrary("mlease_b200_jni"); }
  private long handle;                                       // mlease_rows*

  public NativeIngest(String path, boolean raw, boolean binaryFeature) throws java.io.IOException { handle = read(path, raw, binaryFeature); }
  /** {records, stored values, features} */
  public native long[] counts();
  /** Fills caller-allocated direct buffers (native byte order): rowptr int64[records + 1], colidx int32[nnz], vals float32[nnz],
   *  response int32[records], weight / offset float32[records]; any may be null. */
  public native void get(ByteBuffer rowptrI64, ByteBuffer colidxI32, ByteBuffer valsF32, ByteBuffer responseI32, ByteBuffer weightF32, ByteBuffer offsetF32);
  /** name, or name + U+0001 + term, of feature k (the dictionary key of the reference's datasets). */
  public native String feature(int k);
  /** partition key of record i (Integer.parseInt(key) is the AdmmMapper's partition id, jobs/RegressionAdmmTrain.java:558). */
  public native String key(long i);
  @Override public native void close();
  private static native long read(String path, boolean raw, boolean binaryFeature) throws java.io.IOException;
}
