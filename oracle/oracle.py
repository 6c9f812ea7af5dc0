"""ctypes loader for the CPU oracle (oracle/mlease_oracle.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Nothing under ml-ease_b200/ imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libmlease_oracle.so")
    src = os.path.join(_HERE, "mlease_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmlease_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_last_error.restype = C.c_char_p
    return _LIB


def _p(a, t):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(t))


def _chk(rc):
    if rc != 0:
        raise RuntimeError("oracle: " + lib().orc_last_error().decode())


class Csr:
    """Prepared records (RegressionPrepareOutput) in global-index CSR form."""

    def __init__(self, rowptr, colidx, val, response, weight=None, offset=None, n_features=None):
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        self.colidx = np.ascontiguousarray(colidx, dtype=np.int32)
        self.val = np.ascontiguousarray(val, dtype=np.float32)
        self.response = np.ascontiguousarray(response, dtype=np.int32)
        n = len(self.response)
        self.weight = np.ones(n, np.float32) if weight is None else np.ascontiguousarray(weight, dtype=np.float32)
        self.offset = np.zeros(n, np.float32) if offset is None else np.ascontiguousarray(offset, dtype=np.float32)
        self.n_features = int(n_features if n_features is not None else (self.colidx.max() + 1 if len(self.colidx) else 0))

    @property
    def nrows(self):
        return len(self.response)

    @staticmethod
    def from_dense(X, response, weight=None, offset=None):
        X = np.ascontiguousarray(X, dtype=np.float32)
        n, d = X.shape
        rowptr = np.arange(n + 1, dtype=np.int64) * d
        colidx = np.tile(np.arange(d, dtype=np.int32), n)
        return Csr(rowptr, colidx, X.reshape(-1), response, weight, offset, d)


def objective(mode, data: Csr, w, prior_mean, prior_var, vec=None, has_bias=True):
    """mode: 'fun'|'grad'|'Hv'|'hessian'|'hessian_diag' (llf/LogisticRegressionL2.java)."""
    m = {"fun": 0, "grad": 1, "Hv": 2, "hessian": 3, "hessian_diag": 4}[mode]
    Dt = data.n_features + (1 if has_bias else 0)
    w = np.ascontiguousarray(w, np.float64)
    pm = np.ascontiguousarray(prior_mean, np.float64)
    pv = np.ascontiguousarray(prior_var, np.float64)
    v = None if vec is None else np.ascontiguousarray(vec, np.float64)
    out_s = C.c_double(0)
    out_v = np.zeros(Dt * Dt if m == 3 else Dt, np.float64)
    _chk(lib().orc_objective(m, data.n_features, C.c_int64(data.nrows), _p(data.rowptr, C.c_int64), _p(data.colidx, C.c_int32),
                             _p(data.val, C.c_float), _p(data.response, C.c_int32), _p(data.weight, C.c_float),
                             _p(data.offset, C.c_float), int(has_bias), _p(w, C.c_double), _p(pm, C.c_double),
                             _p(pv, C.c_double), _p(v, C.c_double), C.byref(out_s), _p(out_v, C.c_double)))
    if m == 0:
        return out_s.value
    if m == 1:
        return out_s.value, out_v
    if m == 3:
        return out_v.reshape(Dt, Dt)
    return out_v


def liblinear_train(data: Csr, init, prior_mean, prior_var, epsilon, max_iter=10000, has_bias=True):
    param = np.array(init, np.float64, copy=True)
    pm = np.ascontiguousarray(prior_mean, np.float64)
    pv = np.ascontiguousarray(prior_var, np.float64)
    outer, cg, gn, ps = C.c_int(0), C.c_int(0), C.c_double(0), C.c_int64(0)
    _chk(lib().orc_liblinear_train(data.n_features, C.c_int64(data.nrows), _p(data.rowptr, C.c_int64), _p(data.colidx, C.c_int32),
                                   _p(data.val, C.c_float), _p(data.response, C.c_int32), _p(data.weight, C.c_float),
                                   _p(data.offset, C.c_float), int(has_bias), _p(param, C.c_double), _p(pm, C.c_double),
                                   _p(pv, C.c_double), C.c_double(epsilon), int(max_iter), C.byref(outer), C.byref(cg),
                                   C.byref(gn), C.byref(ps)))
    return param, dict(outer=outer.value, cg=cg.value, gnorm=gn.value, passes=ps.value)


def admm_run(data: Csr, part_rowstart, lambdas, rhos=None, niters=10, epsilon=1e-4, mode="exact", penalize_intercept=False,
             aggressive_decay=False, rho_adapt_coefficient=0.0, binary_feature=False, nthreads=1, initialize_boost_rate=0.0,
             init_liblinear_epsilon=0.01, regularizer=2, lambda_map=None):
    """RegressionAdmmTrain.run restated (jobs/RegressionAdmmTrain.java:130-522). Returns dict."""
    prs = np.ascontiguousarray(part_rowstart, np.int64)
    P = len(prs) - 1
    lam = np.ascontiguousarray(lambdas, np.float32)
    L = len(lam)
    rh = None if rhos is None else np.ascontiguousarray(rhos, np.float32)
    Dt = data.n_features + 1
    z_hist = np.zeros((niters, L, Dt), np.float64)
    diff_hist = np.zeros((niters, L), np.float64)
    eps_hist = np.zeros(niters, np.float32)
    x_last = np.zeros((P, L, Dt), np.float64)
    u_last = np.zeros((P, L, Dt), np.float32)
    uplusx_last = np.zeros((P, L, Dt), np.float32)
    done, passes, touter, tcg = C.c_int(0), C.c_int64(0), C.c_int64(0), C.c_int64(0)
    lmap = None if lambda_map is None else np.ascontiguousarray(lambda_map, np.float32)
    _chk(lib().orc_admm_run(P, data.n_features, _p(prs, C.c_int64), _p(data.rowptr, C.c_int64), _p(data.colidx, C.c_int32),
                            _p(data.val, C.c_float), _p(data.response, C.c_int32), _p(data.weight, C.c_float),
                            _p(data.offset, C.c_float), L, _p(lam, C.c_float), _p(rh, C.c_float), int(niters),
                            C.c_double(epsilon), 0 if mode == "faithful" else 1, int(penalize_intercept), int(aggressive_decay),
                            C.c_float(rho_adapt_coefficient), int(binary_feature), int(nthreads), _p(z_hist, C.c_double),
                            _p(diff_hist, C.c_double), _p(eps_hist, C.c_float), _p(x_last, C.c_double), _p(u_last, C.c_float),
                            _p(uplusx_last, C.c_float), C.byref(done), C.byref(passes), C.byref(touter), C.byref(tcg),
                            C.c_float(initialize_boost_rate), C.c_float(init_liblinear_epsilon), int(regularizer),
                            _p(lmap, C.c_float)))
    n = done.value
    return dict(z_hist=z_hist[:n], diff_hist=diff_hist[:n], eps_hist=eps_hist[:n], x_last=x_last, u_last=u_last,
                uplusx_last=uplusx_last, iters_done=n, passes=passes.value, tron_outer=touter.value, tron_cg=tcg.value,
                final_model=z_hist[n - 1].astype(np.float32))


def naive_train(data: Csr, key_rowstart, lam, lambda_map=None, prior_mean=0.0, penalize_intercept=False, has_intercept=True,
                liblinear_epsilon=0.001, data_size_threshold=0, mode="exact", nthreads=1):
    krs = np.ascontiguousarray(key_rowstart, np.int64)
    K = len(krs) - 1
    Dt = data.n_features + 1
    out = np.zeros((K, Dt), np.float64)
    skipped = np.zeros(K, np.int32)
    lm = None if lambda_map is None else np.ascontiguousarray(lambda_map, np.float32)
    passes = C.c_int64(0)
    _chk(lib().orc_naive_train(K, data.n_features, _p(krs, C.c_int64), _p(data.rowptr, C.c_int64), _p(data.colidx, C.c_int32),
                               _p(data.val, C.c_float), _p(data.response, C.c_int32), _p(data.weight, C.c_float),
                               _p(data.offset, C.c_float), C.c_float(lam), _p(lm, C.c_float), C.c_float(prior_mean),
                               int(penalize_intercept), int(has_intercept), C.c_float(liblinear_epsilon),
                               int(data_size_threshold), 0 if mode == "faithful" else 1, int(nthreads), _p(out, C.c_double),
                               _p(skipped, C.c_int32), C.byref(passes)))
    return out, skipped.astype(bool), passes.value


def score(data: Csr, model, num_click_replicates=1, binary_feature=False):
    m = np.ascontiguousarray(model, np.float64)
    pred = np.zeros(data.nrows, np.float32)
    _chk(lib().orc_score(data.n_features, C.c_int64(data.nrows), _p(data.rowptr, C.c_int64), _p(data.colidx, C.c_int32),
                         _p(data.val, C.c_float), _p(data.offset, C.c_float), _p(m, C.c_double), int(num_click_replicates),
                         int(binary_feature), _p(pred, C.c_float)))
    return pred


def test_loglik(response, pred, weight=None, combiner_block=0):
    r = np.ascontiguousarray(response, np.int32)
    p = np.ascontiguousarray(pred, np.float32)
    w = None if weight is None else np.ascontiguousarray(weight, np.float32)
    ll, cnt = C.c_float(0), C.c_double(0)
    _chk(lib().orc_test_loglik(C.c_int64(len(r)), _p(r, C.c_int32), _p(p, C.c_float), _p(w, C.c_float),
                               C.c_int64(combiner_block), C.byref(ll), C.byref(cnt)))
    return ll.value, cnt.value


test_loglik.__test__ = False  # not a pytest test


def sample_test_loglik(data: Csr, model, binary_feature=False):
    m = np.ascontiguousarray(model, np.float64)
    out = C.c_double(0)
    _chk(lib().orc_sample_test_loglik(data.n_features, C.c_int64(data.nrows), _p(data.rowptr, C.c_int64),
                                      _p(data.colidx, C.c_int32), _p(data.val, C.c_float), _p(data.response, C.c_int32),
                                      _p(data.weight, C.c_float), _p(data.offset, C.c_float), _p(m, C.c_double),
                                      int(binary_feature), C.byref(out)))
    return out.value


def prepare(base_key, response, weight, nblocks, num_click_replicates=1, random_key_mode=False):
    bk = np.ascontiguousarray(base_key, np.int32)
    r = np.ascontiguousarray(response, np.int32)
    w = None if weight is None else np.ascontiguousarray(weight, np.float64)
    n = len(r)
    keys = np.full((n, num_click_replicates), -1, np.int32)
    nk = np.zeros(n, np.int32)
    ow = np.zeros(n, np.float32)
    _chk(lib().orc_prepare(C.c_int64(n), _p(bk, C.c_int32), _p(r, C.c_int32), _p(w, C.c_double), int(nblocks),
                           int(num_click_replicates), int(random_key_mode), _p(keys, C.c_int32), _p(nk, C.c_int32),
                           _p(ow, C.c_float)))
    return keys, nk, ow


def partition_ids(keys, lambdas, num_reducers):
    packed = b"".join(k.encode() + b"\0" for k in keys)
    lam = np.ascontiguousarray(lambdas, np.float32)
    L, n = len(lam), len(keys)
    ids = np.zeros((L, n), np.int32)
    part = np.zeros((L, n), np.int32)
    hpart = np.zeros((L, n), np.int32)
    _chk(lib().orc_partition_ids(n, packed, _p(lam, C.c_float), L, int(num_reducers), _p(ids, C.c_int32), _p(part, C.c_int32),
                                 _p(hpart, C.c_int32)))
    return ids, part, hpart


def java_float_to_string(f):
    buf = C.create_string_buffer(64)
    _chk(lib().orc_java_float_to_string(C.c_float(f), buf, 64))
    return buf.value.decode()


def java_string_hash(s):
    lib().orc_java_string_hash.restype = C.c_int32
    return lib().orc_java_string_hash(s.encode())
