#!/usr/bin/env python
"""Throughput of the host job layer's CPU paths (no GPU needed): RegressionPrepare, the prepared-record ingest of AdmmTrain /
NaiveTrain, the per-iteration model files, RegressionTest's output step -- block-parallel plan-walker code against the generic
(Value-tree) decoder it replaced (still the fallback and the test reference).  Prints one JSON object.

    python tools/bench_host_ingest.py [records] [features per record]        (default 100000 x 50, 5000 distinct features)
"""
import ctypes as C
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ml-ease_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import avro_util as au  # noqa: E402
import mlease_b200  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    nnz = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    D = 5000
    mlease_b200.lib()
    h = C.CDLL(os.path.join(ROOT, "ml-ease_b200", "lib", "libmlease_host.so"))
    h.mlease_job_last_error.restype = C.c_char_p
    h.mlease_rows_count.restype = C.c_int64
    h.mlease_rows_count.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    h.mlease_rows_free.argtypes = [C.c_void_p]
    rng = np.random.default_rng(0)
    tmp = tempfile.mkdtemp(prefix="mlease_ingest_")
    cols = rng.integers(0, D, size=(n, nnz)); vals = rng.normal(size=(n, nnz)).astype(np.float32); y = rng.integers(0, 2, n)
    recs = [{"offset": 0, "weight": 1, "response": int(y[i]), "pkey": int(i % 8),
             "features": [{"name": str(int(c)), "term": "", "value": float(v)} for c, v in zip(cols[i], vals[i])]} for i in range(n)]
    src = os.path.join(tmp, "in", "part-0.avro")
    au.write_avro(src, au.pig_schema_with_key(), recs, block=4000, codec="deflate")
    del recs
    out = {"records": n, "features_per_record": nnz, "input_MB": os.path.getsize(src) / 1e6, "host_threads": h.mlease_host_set_threads(0),
           "cpus": len(os.sched_getaffinity(0))}

    def best(fn, reps=3):
        ts = []
        for _ in range(reps):
            t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
        return min(ts)

    def prepare(generic):
        os.environ["MLEASE_HOST_GENERIC_INGEST"] = "1" if generic else "0"
        cfg = os.path.join(tmp, "p.job")
        open(cfg, "w").write("input.paths=%s\noutput.path=%s\nmap.key=pkey\nnum.blocks=8\n" % (os.path.join(tmp, "in"), os.path.join(tmp, "prep")))
        assert h.mlease_job_run(b"RegressionPrepare", cfg.encode()) == 0, h.mlease_job_last_error()
    for name, generic in (("generic", 1), ("fast", 0)):
        dt = best(lambda: prepare(generic), 2 if generic else 3)
        out["prepare_" + name] = {"seconds": dt, "records_per_s": n / dt, "features_per_s": n * nnz / dt}
    os.environ["MLEASE_HOST_GENERIC_INGEST"] = "0"

    def ingest(generic):
        hd = C.c_void_p()
        assert h.mlease_rows_read(os.path.join(tmp, "prep").encode(), 0, 0, generic, C.byref(hd)) == 0
        h.mlease_rows_free(hd)
    for name, generic in (("generic", 1), ("fast", 0)):
        dt = best(lambda: ingest(generic), 2 if generic else 4)
        out["ingest_prepared_" + name] = {"seconds": dt, "records_per_s": n / dt, "features_per_s": n * nnz / dt}

    Dm, M = 10000, 24    # the per-iteration `model` file of config 3: 8 partitions x 3 lambdas x (x, u+x) x 10 001 coefficients
    names = ("\0".join(str(k) for k in range(Dm)) + "\0").encode(); keys = ("\0".join("1.0#%d" % m for m in range(M)) + "\0").encode()
    co = rng.normal(size=(M, Dm + 1)).astype(np.float32)
    for name, generic in (("generic", 1), ("fast", 0)):
        dt = best(lambda: h.mlease_models_write(os.path.join(tmp, "m.avro").encode(), Dm, names, M, keys, co.ctypes.data_as(C.c_void_p), co.ctypes.data_as(C.c_void_p), generic))
        out["iteration_model_file_" + name] = {"seconds": dt, "coefficients_per_s": 2 * M * (Dm + 1) / dt}
    pred = np.zeros(n, np.float32)
    for name, generic in (("generic", 1), ("fast", 0)):
        dt = best(lambda: h.mlease_test_output_write(src.encode(), os.path.join(tmp, "t.avro").encode(), pred.ctypes.data_as(C.c_void_p), C.c_int64(n), generic), 2)
        out["test_output_" + name] = {"seconds": dt, "records_per_s": n / dt}
    out["note"] = ("generic = Value-tree decoder / encoder on one thread (deflate already on background threads, level 1); round 1's job layer "
                   "was this decoder with serial level-6 deflate")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
