/*
 * mlease_host.h -- C entry points of the host job layer (libmlease_host.so; ml-ease_b200/host/).
 * The reference's job classes (the com/linkedin/mlease/regression/jobs/ sources) keep their names, config keys (java
 * .properties job file, com/linkedin/mapred/JobConfig.java:78-90), avro schemas (the .avsc files under src/main/avro/) and output
 * directory layout; the arithmetic goes through include/mlease_b200.h.  CLI: `mlease_regression <job class> <config>`
 * mirrors `hadoop jar … com.linkedin.mlease.regression.jobs.Regression <config>` (jobs/Regression.java:88-98).
 */
#ifndef MLEASE_HOST_H
#define MLEASE_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* job_class: Regression | RegressionPrepare | RegressionAdmmTrain | RegressionTest | RegressionTestLoglik |
 * RegressionNaiveTrain (README aliases AdmmPrepare/AdmmTrain/AdmmTest/AdmmTestLoglik/NaiveTrain accepted).
 * Returns 0, or non-zero with the message (the reference's IOException / RuntimeException text) in mlease_job_last_error(). */
int mlease_job_run(const char* job_class, const char* config_path);
const char* mlease_job_last_error(void);

/* Deterministic branches of RegressionPrepare (jobs/RegressionPrepare.java:154-186): per-record partition keys (positives
 * replicated onto consecutive partitions mod nblocks when the key is drawn at random) and the prepared float weight. */
int mlease_prepare_keys(int64_t nrows, const int32_t* base_key, const int32_t* response, const double* weight_in, int32_t nblocks,
                        int32_t num_click_replicates, int32_t random_key_mode, int32_t* out_keys, int32_t* out_nkeys, float* out_weight);
/* PartitionIdAssigner ids (sorted Utf8 order of "<lambda>#<key>", jobs/PartitionIdAssigner.java:79-88) and NaivePartitioner
 * partitions (id % R, else abs(String.hashCode()) % R, jobs/RegressionNaiveTrain.java:269-283). keys: NUL-separated. */
int mlease_partition_ids(int32_t nkeys, const char* keys_packed, const float* lambdas, int32_t L, int32_t num_reducers, int32_t* out_ids,
                         int32_t* out_partition, int32_t* out_hash_partition);
/* Java Float.toString (model keys "1.0", "1.0#3"). */
int mlease_java_float_to_string(float f, char* buf, int32_t buflen);
/* Avro container round trip (decode every record generically, re-encode with `codec` = "null" | "deflate"). */
int mlease_avro_copy(const char* in_path, const char* out_path, const char* codec, int64_t* nrecords, int64_t* nblocks);
#ifdef __cplusplus
}
#endif
#endif
