"""Host-side mirror of the reference's ADMM train job for the accelerated path.

`AdmmSession` wraps the C ABI session = the body of RegressionAdmmTrain.run()
(jobs/RegressionAdmmTrain.java:278-501): partitions are uploaded once
(replacing AdmmReducer's per-iteration dataset rebuild, :677-690), `local_step` is the
reducer phase (:642-718) for all resident (partition, lambda) pairs, `consensus` is the
driver's z/u update (:362-404, :736-765).  Config keys keep the reference's names
(with '.' -> '_').
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._native import ALLREDUCE_FN, AdmmConfigC, MleaseError, StatsC, check, lib, ptr


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _keep(a, dtype):
    """numpy/torch passthrough with dtype/contiguity enforcement; returns (object_to_keep_alive)."""
    if a is None:
        return None
    if hasattr(a, "data_ptr"):  # torch tensor
        import torch
        td = {np.float32: torch.float32, np.int32: torch.int32, np.int64: torch.int64, np.float64: torch.float64}[dtype]
        if a.dtype != td or not a.is_contiguous():
            a = a.to(td).contiguous()
        return a
    return np.ascontiguousarray(a, dtype=dtype)


class Comm:
    """One rank of the library's NCCL communicator (include/mlease_b200.h "Multi-GPU" (a)): rank 0 makes the id with
    Comm.unique_id(), ships the 128 bytes to the other processes (any transport), every rank builds Comm(id, rank, nranks, device)
    and attaches it to its AdmmSession with set_comm(); session.run(iters) then runs the whole loop in C with one
    ncclAllReduce per iteration."""

    def __init__(self, unique_id: bytes, rank: int, nranks: int, device: int = 0):
        self._h = None
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        h = C.c_void_p()
        check(lib().mlease_comm_create(buf, int(rank), int(nranks), int(device), C.byref(h)))
        self._h, self.rank, self.nranks = h, int(rank), int(nranks)

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_char * 128)()
        check(lib().mlease_comm_unique_id(buf))
        return bytes(buf.raw)

    def nccl_version(self):
        v = C.c_int32(0)
        check(lib().mlease_comm_info(self._h, None, None, C.byref(v)))
        return v.value

    def close(self):
        if self._h is not None:
            lib().mlease_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class AdmmSession:
    def __init__(self, num_blocks, num_features, lambdas, rhos=None, *, device=0, stream=None, regularizer=2,
                 penalize_intercept=False, epsilon=1e-4, rho_adapt_coefficient=0.0, aggressive_liblinear_epsilon_decay=False,
                 binary_feature=False, lambda_map=None, newton_xtol=0.0, max_newton=0, hessian_policy=0):
        self._h = None
        self.num_blocks, self.num_features = int(num_blocks), int(num_features)
        self.lambdas = _f32(np.atleast_1d(lambdas))
        self.L = len(self.lambdas)
        self.Dt = self.num_features + 1
        self.device = int(device)
        self._rhos = _f32(rhos)
        self._lmap = _f32(lambda_map)
        cfg = AdmmConfigC()
        cfg.device, cfg.num_blocks, cfg.num_features, cfg.num_lambdas = self.device, self.num_blocks, self.num_features, self.L
        cfg.lambdas = self.lambdas.ctypes.data_as(C.POINTER(C.c_float))
        cfg.rhos = None if self._rhos is None else self._rhos.ctypes.data_as(C.POINTER(C.c_float))
        cfg.lambda_map = None if self._lmap is None else self._lmap.ctypes.data_as(C.POINTER(C.c_float))
        cfg.regularizer, cfg.penalize_intercept = int(regularizer), int(bool(penalize_intercept))
        cfg.aggressive_decay, cfg.binary_feature = int(bool(aggressive_liblinear_epsilon_decay)), int(bool(binary_feature))
        cfg.epsilon, cfg.rho_adapt_coefficient = float(epsilon), float(rho_adapt_coefficient)
        cfg.newton_xtol, cfg.max_newton, cfg.hessian_policy = float(newton_xtol), int(max_newton), int(hessian_policy)
        cfg.stream = stream
        h = C.c_void_p()
        check(lib().mlease_session_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self._cb = None

    def close(self):
        if self._h is not None:
            lib().mlease_session_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_comm(self, comm):
        """Attach (or detach with None) the NCCL communicator of a multi-process job; the session then holds only its own
        partitions (p % nranks == rank) and run()/iterate() all-reduce inside the library."""
        self._comm = comm
        check(lib().mlease_session_set_comm(self._h, None if comm is None else comm._h))

    # ---- data ----
    def add_partition_dense(self, partition_id, X, response, weight=None, offset=None):
        X = _keep(X, np.float32)
        n, d = X.shape
        ld = X.stride(0) if hasattr(X, "data_ptr") else d
        r, w, o = _keep(response, np.int32), _keep(weight, np.float32), _keep(offset, np.float32)
        check(lib().mlease_add_partition_dense(self._h, int(partition_id), n, ptr(X), ld, ptr(r), ptr(w), ptr(o)))

    def add_partition_csr(self, partition_id, rowptr, colidx, vals, response, weight=None, offset=None):
        rp, ci, v = _keep(rowptr, np.int64), _keep(colidx, np.int32), _keep(vals, np.float32)
        r, w, o = _keep(response, np.int32), _keep(weight, np.float32), _keep(offset, np.float32)
        check(lib().mlease_add_partition_csr(self._h, int(partition_id), len(r), ptr(rp), ptr(ci), ptr(v), ptr(r), ptr(w), ptr(o)))

    # ---- ADMM ----
    def begin(self, z0=None, boost_rate=0.0):
        """Cold start (z = {}), or -- initialize.boost.rate > 0 -- from z0 [L][D+1] with the reducers' rho scaled by boost_rate
        (jobs/RegressionAdmmTrain.java:236-266, 313-316)."""
        if z0 is None:
            check(lib().mlease_admm_begin(self._h))
        else:
            z = np.ascontiguousarray(z0, np.float64).reshape(self.L, self.Dt)
            check(lib().mlease_admm_begin_initialized(self._h, ptr(z), C.c_float(boost_rate)))

    def mean_naive_model(self, partition_ids, penalize_intercept=False):
        """z0 of initialize.boost.rate for a single-process job: per (partition, lambda) the RegressionNaiveTrain fit (prior
        variance 1/lambda, intercept variance 100000 unless penalised, prior mean 0, start 0: jobs/RegressionNaiveTrain.java:333-343,395),
        written as float and averaged by MeanLinearModelConsumer (cons/MeanLinearModelConsumer.java:44-70).  Multi-process jobs
        sum the per-rank partial means with one all-reduce of this array."""
        z0 = np.zeros((self.L, self.Dt), np.float64)
        nb = float(self.num_blocks)
        for li, lam in enumerate(self.lambdas):
            q = np.full(self.Dt, float(lam), np.float64)
            if not penalize_intercept:
                q[-1] = 1.0 / 100000.0
            for pid in partition_ids:
                x, _ = self.fit_partition(pid, np.zeros(self.Dt), np.zeros(self.Dt), q)
                z0[li] = 1.0 * z0[li] + (1.0 / nb) * x.astype(np.float32).astype(np.float64)
        return z0

    def local_step(self, exchange_dev_ptr):
        check(lib().mlease_admm_local_step(self._h, ptr(exchange_dev_ptr)))

    def consensus(self, exchange_sum_dev_ptr):
        md, stop = C.c_double(0), C.c_int32(0)
        check(lib().mlease_admm_consensus(self._h, ptr(exchange_sum_dev_ptr), C.byref(md), C.byref(stop)))
        return md.value, bool(stop.value)

    def iterate(self):
        """One iteration of a single-process job: local_step + consensus on the session's own exchange buffer."""
        md, stop = C.c_double(0), C.c_int32(0)
        check(lib().mlease_admm_iterate(self._h, C.byref(md), C.byref(stop)))
        return md.value, bool(stop.value)

    def run(self, num_iters, allreduce=None):
        """Single-process job (allreduce None) or with a Python all-reduce callable(buf_ptr, count, stream_ptr)."""
        done = C.c_int32(0)
        cb = None
        if allreduce is not None:
            def _cb(ctx, buf, count, stream):
                try:
                    allreduce(buf, count, stream)
                    return 0
                except Exception:  # pragma: no cover
                    import traceback
                    traceback.print_exc()
                    return 1
            cb = ALLREDUCE_FN(_cb)
            self._cb = cb
        check(lib().mlease_admm_run(self._h, int(num_iters), C.cast(cb, C.c_void_p) if cb else None, None, C.byref(done)))
        return done.value

    # ---- state ----
    def z(self, lambda_idx=0):
        out = np.zeros(self.Dt, np.float64)
        check(lib().mlease_get_z(self._h, lambda_idx, ptr(out)))
        return out

    def final_model(self, lambda_idx=0):
        out = np.zeros(self.Dt, np.float32)
        check(lib().mlease_get_final_model(self._h, lambda_idx, ptr(out)))
        return out

    def x(self, partition_id, lambda_idx=0):
        out = np.zeros(self.Dt, np.float64)
        check(lib().mlease_get_x(self._h, partition_id, lambda_idx, ptr(out)))
        return out

    def u(self, partition_id, lambda_idx=0):
        out = np.zeros(self.Dt, np.float32)
        check(lib().mlease_get_u(self._h, partition_id, lambda_idx, ptr(out)))
        return out

    def uplusx(self, partition_id, lambda_idx=0):
        out = np.zeros(self.Dt, np.float32)
        check(lib().mlease_get_uplusx(self._h, partition_id, lambda_idx, ptr(out)))
        return out

    def stats(self):
        s = StatsC()
        check(lib().mlease_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in StatsC._fields_}

    # ---- function-level entry points ----
    def objective(self, partition_id, w, prior_mean, prior_precision, want_grad=True, want_hessian=False, tensor=True):
        w, m, q = (np.ascontiguousarray(a, np.float64) for a in (w, prior_mean, prior_precision))
        f = C.c_double(0)
        g = np.zeros(self.Dt, np.float64) if want_grad else None
        H = np.zeros((self.Dt, self.Dt), np.float64) if want_hessian else None
        check(lib().mlease_objective(self._h, partition_id, ptr(w), ptr(m), ptr(q), C.byref(f), ptr(g), ptr(H), int(tensor)))
        return f.value, g, H

    def fit_partition(self, partition_id, init, prior_mean, prior_precision):
        x = np.array(init, np.float64, copy=True)
        m, q = (np.ascontiguousarray(a, np.float64) for a in (prior_mean, prior_precision))
        steps = C.c_int32(0)
        check(lib().mlease_fit_partition(self._h, partition_id, ptr(x), ptr(m), ptr(q), C.byref(steps)))
        return x, steps.value

    def posterior_variance(self, partition_id, w, prior_precision, full=False, want_cov=False):
        """LibLinear.train's computePosteriorVar tail (llf/LibLinear.java:315-334): diagonal (1 / hessianDiagonal) or full
        (diag of the inverse of the exact fp64 Hessian; want_cov also returns H^-1)."""
        w, q = (np.ascontiguousarray(a, np.float64) for a in (w, prior_precision))
        var = np.zeros(self.Dt, np.float64)
        cov = np.zeros((self.Dt, self.Dt), np.float64) if (full and want_cov) else None
        check(lib().mlease_posterior_variance(self._h, int(partition_id), ptr(w), ptr(q), int(bool(full)), ptr(var), ptr(cov)))
        return (var, cov) if want_cov else var

    def profile(self, enable=-1):
        """Per-kernel CUDA-event timing accumulators; enable: 1 on, 0 off, 2 on+reset, -1 read only."""
        ms = np.zeros(4, np.float64); cnt = np.zeros(4, np.int64)
        kb, eb, gf = C.c_double(0), C.c_double(0), C.c_double(0)
        check(lib().mlease_profile(self._h, int(enable), ptr(ms), ptr(cnt), C.byref(kb), C.byref(eb), C.byref(gf)))
        names = ("k1", "small", "gram", "cholesky")
        return dict(ms=dict(zip(names, ms.tolist())), launches=dict(zip(names, cnt.tolist())), k1_bytes=kb.value,
                    k1_emit_bytes=eb.value, gram_flops=gf.value)

    def time_kernel(self, partition_id, which, reps=5, emit_scaled=False):
        ms = C.c_float(0)
        check(lib().mlease_time_kernel(self._h, partition_id, {"k1": 1, "gram": 2, "cholesky": 3}[which], reps, int(emit_scaled), C.byref(ms)))
        return ms.value


class World:
    """N GPUs of THIS process behind the session calls (include/mlease_b200.h "Multi-GPU" (b)): partitions go to GPU
    partition_id % ndev, one worker thread per GPU inside the library, NCCL all-reduce per iteration."""

    def __init__(self, devices, num_blocks, num_features, lambdas, rhos=None, *, regularizer=2, penalize_intercept=False, epsilon=1e-4,
                 rho_adapt_coefficient=0.0, aggressive_liblinear_epsilon_decay=False, binary_feature=False, lambda_map=None,
                 newton_xtol=0.0, max_newton=0, hessian_policy=0):
        self._h = None
        self.lambdas = _f32(np.atleast_1d(lambdas))
        self.L, self.Dt, self.num_blocks = len(self.lambdas), int(num_features) + 1, int(num_blocks)
        self._rhos, self._lmap = _f32(rhos), _f32(lambda_map)
        cfg = AdmmConfigC()
        cfg.device, cfg.num_blocks, cfg.num_features, cfg.num_lambdas = 0, self.num_blocks, int(num_features), self.L
        cfg.lambdas = self.lambdas.ctypes.data_as(C.POINTER(C.c_float))
        cfg.rhos = None if self._rhos is None else self._rhos.ctypes.data_as(C.POINTER(C.c_float))
        cfg.lambda_map = None if self._lmap is None else self._lmap.ctypes.data_as(C.POINTER(C.c_float))
        cfg.regularizer, cfg.penalize_intercept = int(regularizer), int(bool(penalize_intercept))
        cfg.aggressive_decay, cfg.binary_feature = int(bool(aggressive_liblinear_epsilon_decay)), int(bool(binary_feature))
        cfg.epsilon, cfg.rho_adapt_coefficient = float(epsilon), float(rho_adapt_coefficient)
        cfg.newton_xtol, cfg.max_newton, cfg.hessian_policy = float(newton_xtol), int(max_newton), int(hessian_policy)
        cfg.stream = None
        devs = np.ascontiguousarray(devices, np.int32)
        h = C.c_void_p()
        check(lib().mlease_world_create(C.byref(cfg), ptr(devs), len(devs), C.byref(h)))
        self._h, self.ndev = h, len(devs)

    def close(self):
        if self._h is not None:
            lib().mlease_world_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def add_partition_dense(self, partition_id, X, response, weight=None, offset=None):
        X = _keep(X, np.float32)
        n, d = X.shape
        ld = X.stride(0) if hasattr(X, "data_ptr") else d
        r, w, o = _keep(response, np.int32), _keep(weight, np.float32), _keep(offset, np.float32)
        check(lib().mlease_world_add_partition_dense(self._h, int(partition_id), n, ptr(X), ld, ptr(r), ptr(w), ptr(o)))

    def add_partition_csr(self, partition_id, rowptr, colidx, vals, response, weight=None, offset=None):
        rp, ci, v = _keep(rowptr, np.int64), _keep(colidx, np.int32), _keep(vals, np.float32)
        r, w, o = _keep(response, np.int32), _keep(weight, np.float32), _keep(offset, np.float32)
        check(lib().mlease_world_add_partition_csr(self._h, int(partition_id), len(r), ptr(rp), ptr(ci), ptr(v), ptr(r), ptr(w), ptr(o)))

    def begin(self, z0=None, boost_rate=0.0):
        if z0 is None:
            check(lib().mlease_world_begin(self._h))
        else:
            z = np.ascontiguousarray(z0, np.float64).reshape(self.L, self.Dt)
            check(lib().mlease_world_begin_initialized(self._h, ptr(z), C.c_float(boost_rate)))

    def iterate(self):
        md, stop = C.c_double(0), C.c_int32(0)
        check(lib().mlease_world_iterate(self._h, C.byref(md), C.byref(stop)))
        return md.value, bool(stop.value)

    def run(self, num_iters):
        done = C.c_int32(0)
        check(lib().mlease_world_run(self._h, int(num_iters), C.byref(done)))
        return done.value

    def z(self, lambda_idx=0):
        out = np.zeros(self.Dt, np.float64)
        check(lib().mlease_world_get_z(self._h, lambda_idx, ptr(out)))
        return out

    def _vec(self, fn, pid, l, dtype):
        out = np.zeros(self.Dt, dtype)
        check(fn(self._h, int(pid), int(l), ptr(out)))
        return out

    def x(self, partition_id, lambda_idx=0):
        return self._vec(lib().mlease_world_get_x, partition_id, lambda_idx, np.float64)

    def u(self, partition_id, lambda_idx=0):
        return self._vec(lib().mlease_world_get_u, partition_id, lambda_idx, np.float32)

    def uplusx(self, partition_id, lambda_idx=0):
        return self._vec(lib().mlease_world_get_uplusx, partition_id, lambda_idx, np.float32)

    def stats(self):
        s = StatsC()
        check(lib().mlease_world_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in StatsC._fields_}


def score(vals, model, *, rowptr=None, colidx=None, offset=None, num_features=None, device=0, stream=None,
          num_click_replicates=1, binary_feature=False, out=None):
    """RegressionTest scoring (models/LinearModel.java:241-257; jobs/RegressionTest.java:163). Dense if colidx is None."""
    model = _keep(model, np.float64)
    Dg = int(num_features if num_features is not None else len(model) - 1)
    if colidx is None:
        vals = _keep(vals, np.float32)
        n, ld = vals.shape[0], (vals.stride(0) if hasattr(vals, "data_ptr") else vals.shape[1])
        rp = ci = None
    else:
        rp, ci, vals = _keep(rowptr, np.int64), _keep(colidx, np.int32), _keep(vals, np.float32)
        n, ld = len(rp) - 1, 0
    o = _keep(offset, np.float32)
    pred = np.zeros(n, np.float32) if out is None else out
    check(lib().mlease_score(device, stream, Dg, n, ptr(rp), ptr(ci), ptr(vals), ld, ptr(o), ptr(model), int(num_click_replicates),
                             int(binary_feature), ptr(pred)))
    return pred


def test_loglik(response, pred, weight=None, combiner_block=0, device=0, stream=None):
    """RegressionTestLoglik (jobs/RegressionTestLoglik.java:124-200) -> (float32 avg loglik, count)."""
    r, p, w = _keep(response, np.int32), _keep(pred, np.float32), _keep(weight, np.float32)
    ll, cnt = C.c_float(0), C.c_double(0)
    check(lib().mlease_test_loglik(device, stream, len(r), ptr(r), ptr(p), ptr(w), int(combiner_block), C.byref(ll), C.byref(cnt)))
    return np.float32(ll.value), cnt.value


test_loglik.__test__ = False


def naive_train_dense(X, key_rowstart, response, lam, weight=None, offset=None, lambda_map=None, prior_mean=0.0,
                      penalize_intercept=False, has_intercept=True, data_size_threshold=0, device=0, stream=None):
    """RegressionNaiveTrain reducer for K keys (jobs/RegressionNaiveTrain.java:302-415) -> (models [K,D+1], skipped[K])."""
    X = _keep(X, np.float32)
    krs = np.ascontiguousarray(key_rowstart, np.int64)
    K, D = len(krs) - 1, X.shape[1]
    ld = X.stride(0) if hasattr(X, "data_ptr") else D
    r, w, o, lm = _keep(response, np.int32), _keep(weight, np.float32), _keep(offset, np.float32), _keep(lambda_map, np.float32)
    out = np.zeros((K, D + 1), np.float64)
    skipped = np.zeros(K, np.int32)
    check(lib().mlease_naive_train_dense(device, stream, K, D, ptr(krs), ptr(X), ld, ptr(r), ptr(w), ptr(o), float(lam), ptr(lm),
                                         float(prior_mean), int(penalize_intercept), int(has_intercept), int(data_size_threshold),
                                         ptr(out), ptr(skipped)))
    return out, skipped.astype(bool)


def naive_train(vals, key_rowstart, response, lambdas, *, rowptr=None, colidx=None, num_features=None, weight=None, offset=None,
                lambda_map=None, prior_mean=0.0, penalize_intercept=False, has_intercept=True, data_size_threshold=0,
                binary_feature=False, device=0, stream=None):
    """RegressionNaiveTrain reducers for K keys x L lambdas on one upload (jobs/RegressionNaiveTrain.java:228-241, 302-415).
    CSR when rowptr/colidx are given (vals = stored values), dense otherwise (vals = X [rows, D]).
    -> (models [L, K, D+1] float64, skipped [K] bool)."""
    krs = np.ascontiguousarray(key_rowstart, np.int64)
    K = len(krs) - 1
    lam = _f32(np.atleast_1d(lambdas))
    L = len(lam)
    if colidx is None:
        vals = _keep(vals, np.float32)
        D = vals.shape[1]
        ld = vals.stride(0) if hasattr(vals, "data_ptr") else D
        rp = ci = None
    else:
        rp, ci, vals = _keep(rowptr, np.int64), _keep(colidx, np.int32), _keep(vals, np.float32)
        D, ld = int(num_features), 0
    r, w, o, lm = _keep(response, np.int32), _keep(weight, np.float32), _keep(offset, np.float32), _keep(lambda_map, np.float32)
    out = np.zeros((L, K, D + 1), np.float64)
    skipped = np.zeros(K, np.int32)
    check(lib().mlease_naive_train(device, stream, K, D, ptr(krs), ptr(rp), ptr(ci), ptr(vals), ld, ptr(r), ptr(w), ptr(o), L, ptr(lam), ptr(lm),
                                   float(prior_mean), int(penalize_intercept), int(has_intercept), int(data_size_threshold),
                                   int(bool(binary_feature)), ptr(out), ptr(skipped)))
    return out, skipped.astype(bool)
