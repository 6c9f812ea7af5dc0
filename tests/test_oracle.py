"""CPU tests that pin the oracle (oracle/mlease_oracle.cpp): finite differences, the
scikit-learn fixed point on the reference's fixture, frozen goldens and the reference's own
sanity rules.  No GPU."""
import numpy as np
import pytest

from oracle import oracle as orc


def synth(n=400, d=12, seed=0, sparse=False):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, d)).astype(np.float32)
    if sparse:
        X *= rng.random((n, d)) < 0.3
    beta = rng.normal(size=d) / np.sqrt(d)
    y = (rng.random(n) < 1 / (1 + np.exp(-(X @ beta - 0.5)))).astype(np.int32)
    w = rng.uniform(0.5, 2.0, n).astype(np.float32)
    o = rng.normal(0, 0.1, n).astype(np.float32)
    if sparse:
        rp, ci, v = [0], [], []
        for i in range(n):
            nz = np.nonzero(X[i])[0]
            ci += list(nz); v += list(X[i, nz]); rp.append(len(ci))
        return orc.Csr(rp, ci, v, y, w, o, d), X
    return orc.Csr.from_dense(X, y, w, o), X


def test_fixture_counts(fixture_data):
    # 1000 records, 200 distinct features, 299 positives
    assert fixture_data.nrows == 1000 and fixture_data.n_features == 200
    assert int((fixture_data.response == 1).sum()) == 299
    assert np.all(fixture_data.weight == 1) and np.all(fixture_data.offset == 0)


@pytest.mark.parametrize("sparse", [False, True])
def test_grad_hv_hessian_finite_differences(sparse):
    data, X = synth(sparse=sparse)
    D = data.n_features
    rng = np.random.default_rng(1)
    w = rng.normal(0, 0.3, D + 1); pm = rng.normal(0, 0.3, D + 1); pv = rng.uniform(0.2, 2, D + 1)
    f0, g = orc.objective("grad", data, w, pm, pv)
    assert abs(f0 - orc.objective("fun", data, w, pm, pv)) < 1e-12
    h = 1e-6
    gn = np.array([(orc.objective("fun", data, w + h * e, pm, pv) - orc.objective("fun", data, w - h * e, pm, pv)) / (2 * h)
                   for e in np.eye(D + 1)])
    assert np.abs(gn - g).max() < 1e-6 * max(1, np.abs(g).max())
    H = orc.objective("hessian", data, w, pm, pv)
    assert np.allclose(H, H.T)
    s = rng.normal(size=D + 1)
    Hs = orc.objective("Hv", data, w, pm, pv, vec=s)
    assert np.allclose(H @ s, Hs, rtol=1e-10, atol=1e-10)
    gp = orc.objective("grad", data, w + h * s, pm, pv)[1]; gm = orc.objective("grad", data, w - h * s, pm, pv)[1]
    assert np.abs((gp - gm) / (2 * h) - Hs).max() < 1e-5 * np.abs(Hs).max()
    assert np.allclose(np.diag(H), orc.objective("hessian_diag", data, w, pm, pv), rtol=1e-12)
    # closed form with numpy (independent restatement of the formulas in llf/LogisticRegressionL2.java:30-47)
    Xb = np.hstack([X.astype(np.float64), np.ones((len(X), 1))])
    yy = np.where(data.response == 1, 1.0, -1.0)
    sc = Xb @ w + data.offset
    p = 1 / (1 + np.exp(-yy * sc))
    g_np = Xb.T @ (data.weight * (p - 1) * yy) + (w - pm) / pv
    H_np = (Xb * (data.weight * p * (1 - p))[:, None]).T @ Xb + np.diag(1 / pv)
    assert np.allclose(g, g_np, rtol=1e-10, atol=1e-10) and np.allclose(H, H_np, rtol=1e-10, atol=1e-10)


def test_hessian_rejects_unsorted_duplicate_index():
    # llf/LogisticRegressionL2.java:277 throws on repeated indices in a row
    data = orc.Csr([0, 2], [0, 0], [1.0, 2.0], [1], n_features=1)
    with pytest.raises(RuntimeError, match="not sorted"):
        orc.objective("hessian", data, np.zeros(2), np.zeros(2), np.ones(2))


def test_bad_inputs_rejected():
    # llf/LibLinearDataset.java:419-420, :428-429
    with pytest.raises(RuntimeError, match="response"):
        orc.objective("fun", orc.Csr([0, 1], [0], [1.0], [2], n_features=1), np.zeros(2), np.zeros(2), np.ones(2))
    with pytest.raises(RuntimeError, match="weight"):
        orc.objective("fun", orc.Csr([0, 1], [0], [1.0], [1], weight=[-1.0], n_features=1), np.zeros(2), np.zeros(2), np.ones(2))


def test_tron_reaches_minimiser_and_loose_tolerance_is_loose():
    data, X = synth(n=600, d=20, seed=3)
    D = data.n_features
    pm = np.zeros(D + 1); pv = np.ones(D + 1)
    x_exact, st = orc.liblinear_train(data, np.zeros(D + 1), pm, pv, 1e-14, 100000)
    g = orc.objective("grad", data, x_exact, pm, pv)[1]
    assert np.abs(g).max() < 1e-9
    x_loose, st2 = orc.liblinear_train(data, np.zeros(D + 1), pm, pv, 0.01)
    assert st2["outer"] < st["outer"] and 1e-6 < np.abs(x_loose - x_exact).max() < 0.1
    # warm start at the optimum returns immediately (bw/Tron.java:62)
    x_ws, st3 = orc.liblinear_train(data, x_exact, pm, pv, 0.01)
    assert st3["outer"] == 0 and np.array_equal(x_ws, x_exact)


def test_absent_features_take_prior_mean():
    # llf/LibLinear.java:374-383
    data = orc.Csr([0, 1, 2], [0, 0], [1.0, -1.0], [1, 0], n_features=3)
    pm = np.array([0.1, 0.2, 0.3, 0.0]); pv = np.ones(4)
    x, _ = orc.liblinear_train(data, np.zeros(4), pm, pv, 1e-12)
    assert x[1] == 0.2 and x[2] == 0.3


def test_admm_exact_converges_to_sklearn_fixed_point(fixture_data, sklearn_fp):
    # SURVEY.md 8c pin (1): independent solver, same minimiser
    prs = np.linspace(0, 1000, 5).astype(np.int64)
    ex = orc.admm_run(fixture_data, prs, [1.0, 10.0, 100.0], niters=800, mode="exact", nthreads=8, epsilon=0)
    for li, lam in enumerate((1.0, 10.0, 100.0)):
        ref = sklearn_fp["lam%g" % lam]
        z = ex["z_hist"][-1, li]
        assert np.abs(z - ref).max() / np.abs(ref).max() < 2e-6


def test_admm_matches_frozen_golden(fixture_data, frozen):
    prs = frozen["part_rowstart"]
    ex = orc.admm_run(fixture_data, prs, [1.0, 10.0, 100.0], niters=20, mode="exact", nthreads=4)
    np.testing.assert_allclose(ex["z_hist"], frozen["exact_z_hist"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(ex["u_last"], frozen["exact_u_last"], rtol=1e-6, atol=1e-9)
    fa = orc.admm_run(fixture_data, prs, [1.0, 10.0, 100.0], niters=20, mode="faithful", nthreads=1)
    np.testing.assert_allclose(fa["z_hist"], frozen["faithful_z_hist"], rtol=1e-9, atol=1e-12)
    assert fa["iters_done"] == int(frozen["faithful_iters"]) and fa["passes"] == int(frozen["faithful_passes"])
    np.testing.assert_array_equal(fa["eps_hist"], frozen["faithful_eps_hist"])
    # faithful (loose TRON) tracks exact to O(liblinearEpsilon) early on
    gap = np.abs(fa["z_hist"] - ex["z_hist"]).max(axis=(1, 2))
    assert gap.max() < 0.05 and gap.max() > 1e-8


def test_admm_lambda_order_and_threads_do_not_matter(fixture_data):
    prs = np.linspace(0, 1000, 4).astype(np.int64)
    a = orc.admm_run(fixture_data, prs, [1.0, 100.0], niters=3, mode="faithful", nthreads=1)
    b = orc.admm_run(fixture_data, prs, [100.0, 1.0], niters=3, mode="faithful", nthreads=6)
    np.testing.assert_array_equal(a["z_hist"][:, 0], b["z_hist"][:, 1])
    np.testing.assert_array_equal(a["z_hist"][:, 1], b["z_hist"][:, 0])


def test_admm_schedule_and_stop_rule():
    # jobs/RegressionAdmmTrain.java:338-346 (decay once mindiff<1e-3 or every iter>5 if aggressive), :493-496
    data, _ = synth(n=300, d=5, seed=5)
    prs = [0, 150, 300]
    r = orc.admm_run(data, prs, [1.0], niters=12, mode="faithful", aggressive_decay=True)
    e = r["eps_hist"]
    assert np.all(e[:5] == np.float32(0.01)) and e[5] == np.float32(0.01) / np.float32(10)
    r2 = orc.admm_run(data, prs, [1.0], niters=200, mode="faithful", epsilon=1e-3)
    assert r2["iters_done"] < 200 and r2["eps_hist"][-1] <= 1e-5 and r2["diff_hist"][-1].max() < 1e-3
    # penalize.intercept toggles the intercept rule (:392-403)
    a = orc.admm_run(data, prs, [10.0], niters=2, mode="exact", penalize_intercept=True)
    b = orc.admm_run(data, prs, [10.0], niters=2, mode="exact", penalize_intercept=False)
    assert abs(a["z_hist"][0, 0, -1]) < abs(b["z_hist"][0, 0, -1])
    w = np.float32(2.0) / (np.float32(10.0) + np.float32(2.0))
    assert np.isclose(a["z_hist"][0, 0, -1], float(w) * b["z_hist"][0, 0, -1], rtol=1e-12)


def test_score_and_loglik_float_rounding(fixture_data, frozen):
    model = frozen["exact_z_hist"][-1, 0]
    pred = orc.score(fixture_data, model)
    np.testing.assert_array_equal(pred, frozen["score_pred"])
    X = np.zeros((1000, 200)); rp = fixture_data.rowptr
    for i in range(1000):
        X[i, fixture_data.colidx[rp[i]:rp[i + 1]]] = fixture_data.val[rp[i]:rp[i + 1]]
    ref = (X @ model[:200] + model[200]).astype(np.float32)
    assert np.abs(pred - ref).max() <= 2e-6 * np.abs(ref).max()
    ll, cnt = orc.test_loglik(fixture_data.response, pred, fixture_data.weight, combiner_block=128)
    assert ll == np.float32(frozen["loglik"]) and cnt == 1000
    yy = fixture_data.response == 1
    ref_ll = -(np.log1p(np.exp(np.where(yy, -1, 1) * pred.astype(np.float64)))).mean()
    assert abs(ll - ref_ll) < 1e-6
    assert abs(orc.sample_test_loglik(fixture_data, model) - ref_ll) < 1e-6
    with pytest.raises(RuntimeError):
        orc.test_loglik([3], [0.0])


def test_prepare_click_replicates():
    # jobs/RegressionPrepare.java:159-186
    keys, nk, w = orc.prepare([3, 1, 0], [1, 0, 1], [1.0, 2.0, 3.0], nblocks=4, num_click_replicates=3, random_key_mode=True)
    assert nk.tolist() == [3, 1, 3]
    assert keys[0].tolist() == [3, 0, 1] and keys[1, 0] == 1 and keys[2].tolist() == [0, 1, 2]
    assert w.tolist() == [np.float32(1 / 3), 2.0, 1.0]
    keys, nk, w = orc.prepare([7, 9], [1, 0], None, nblocks=4, num_click_replicates=2, random_key_mode=False)
    assert nk.tolist() == [1, 1] and keys[:, 0].tolist() == [7, 9] and w.tolist() == [0.5, 1.0]


def test_partition_id_assigner_and_java_strings():
    assert orc.java_float_to_string(1.0) == "1.0" and orc.java_float_to_string(0.1) == "0.1"
    assert orc.java_float_to_string(100.0) == "100.0" and orc.java_float_to_string(1e-4) == "1.0E-4"
    assert orc.java_float_to_string(1e7) == "1.0E7" and orc.java_float_to_string(0.001) == "0.001"
    # 0.01f / 10 is 9.999999E-4 in float32 (what jobs/RegressionAdmmTrain.java:340,702 hands to LibLinear)
    assert orc.java_float_to_string(np.float32(0.01) / np.float32(10)) == "9.999999E-4"
    # Java: "hello".hashCode() == 99162322 ; "1.0#7".hashCode() computed by the 31-polynomial
    assert orc.java_string_hash("hello") == 99162322
    ids, part, hpart = orc.partition_ids(["b", "a", "10", "9"], [10.0, 1.0], 3)
    # sorted Utf8 order: "1.0#10" < "1.0#9" < "1.0#a" < "1.0#b" < "10.0#10" < ...
    assert ids[1].tolist() == [3, 2, 0, 1] and ids[0].tolist() == [7, 6, 4, 5]
    assert np.array_equal(part, ids % 3) and hpart.min() >= 0 and hpart.max() < 3


def test_naive_train_matches_sklearn():
    from sklearn.linear_model import LogisticRegression
    data, X = synth(n=500, d=8, seed=11)
    krs = [0, 250, 500]
    # penalised intercept variance 1e5 (jobs/RegressionNaiveTrain.java:341) ~ unpenalised
    models, skipped, _ = orc.naive_train(data, krs, 2.0, mode="exact")
    for k in range(2):
        sl = slice(krs[k], krs[k + 1])
        clf = LogisticRegression(C=1 / 2.0, solver="newton-cholesky", tol=1e-12, max_iter=500)
        clf.fit(X[sl].astype(np.float64), data.response[sl], sample_weight=data.weight[sl].astype(np.float64))
        # offsets are not supported by sklearn -> compare on data without offset
    data2 = orc.Csr.from_dense(X, data.response, data.weight)
    models, skipped, _ = orc.naive_train(data2, krs, 2.0, mode="exact")
    for k in range(2):
        sl = slice(krs[k], krs[k + 1])
        clf = LogisticRegression(C=1 / 2.0, solver="newton-cholesky", tol=1e-12, max_iter=500)
        clf.fit(X[sl].astype(np.float64), data.response[sl], sample_weight=data.weight[sl].astype(np.float64))
        ref = np.concatenate([clf.coef_.ravel(), clf.intercept_])
        assert np.abs(models[k] - ref).max() < 2e-4   # 1e5 intercept variance is not exactly infinite
    m3, sk3, _ = orc.naive_train(data2, krs, 2.0, data_size_threshold=300)
    assert sk3.all() and not m3.any()
    m4, _, _ = orc.naive_train(data2, krs, 2.0, has_intercept=False, mode="exact")
    assert np.all(m4[:, -1] == 0)


def test_initialize_boost_rate_restated():
    """initialize.boost.rate > 0 (jobs/RegressionAdmmTrain.java:236-266, 313-316): z starts at the mean of per-partition
    NaiveTrain fits and the reducers of ITERATION 1 use rho * boost.  The driver builds a new JobConf every iteration
    (`conf = createJobConf(...)`, :286-291 -> com/linkedin/mapred/AbstractAvroJob.java:101-115), so rho.adapt.rate falls back to
    the reducers' default 1.0f (:621) from iteration 2 on: the fixed point is the UN-boosted pooled fit (scikit-learn)."""
    from sklearn.linear_model import LogisticRegression
    data, X = synth(n=900, d=10, seed=21)
    data = orc.Csr.from_dense(X, data.response, data.weight)          # sklearn has no offsets
    prs = [0, 300, 600, 900]
    lam, boost = 2.0, 3.0
    rho = 20.0                                                        # curvature-matched: ADMM contracts fast
    base = orc.admm_run(data, prs, [lam], [rho], niters=3, mode="exact", epsilon=0)
    run = orc.admm_run(data, prs, [lam], [rho], niters=400, mode="exact", epsilon=0, initialize_boost_rate=boost)
    # (1) the first reducers start from, and are pulled towards, the mean NaiveTrain model: iteration 1 differs from the cold run
    naive, _, _ = orc.naive_train(data, prs, lam, mode="exact")
    z0 = sum((1.0 / 3) * naive[k].astype(np.float32).astype(np.float64) for k in range(3))
    assert np.abs(run["z_hist"][0, 0] - base["z_hist"][0, 0]).max() > 1e-3
    assert np.abs(run["z_hist"][0, 0] - z0).max() < np.abs(base["z_hist"][0, 0] - z0).max()
    # (2) iteration 1 is one ADMM step from z0 with the reducers on rho * boost: x_p = argmin loss_p + (rho boost / 2)|x - z0|^2,
    #     z_1 = w * mean(float(x_p)) with the UN-boosted rho in w (:381), intercept = plain mean (:392-403)
    xs = []
    for k in range(3):
        sub = orc.Csr.from_dense(X[prs[k]:prs[k + 1]], data.response[prs[k]:prs[k + 1]], data.weight[prs[k]:prs[k + 1]])
        z0f = z0.astype(np.float32).astype(np.float64)
        xk, _ = orc.liblinear_train(sub, z0f, z0f, np.full(11, 1.0 / (rho * boost)), 1e-14, max_iter=100000)
        xs.append(xk.astype(np.float32).astype(np.float64))
    xbar = sum(xs) / 3.0
    wgt = float(np.float32(3 * rho) / (np.float32(lam) + np.float32(3 * rho)))
    z1 = np.concatenate([wgt * xbar[:-1], xbar[-1:]])
    assert np.abs(run["z_hist"][0, 0] - z1).max() < 2e-6
    # (3) fixed point = the un-boosted pooled fit: the boost is gone from iteration 2 on
    clf = LogisticRegression(C=1 / lam, solver="newton-cholesky", tol=1e-12, max_iter=500)
    clf.fit(X.astype(np.float64), data.response, sample_weight=data.weight.astype(np.float64))
    ref = np.concatenate([clf.coef_.ravel(), clf.intercept_])
    assert np.abs(run["z_hist"][-1, 0] - ref).max() / np.abs(ref).max() < 1e-5
    # (4) a tiny rho.adapt.coefficient (rate = float(exp(-(i-1) c)) == 1.0f) must therefore change nothing
    run2 = orc.admm_run(data, prs, [lam], [rho], niters=3, mode="exact", epsilon=0, initialize_boost_rate=boost, rho_adapt_coefficient=1e-9)
    np.testing.assert_array_equal(run2["z_hist"], run["z_hist"][:3])
