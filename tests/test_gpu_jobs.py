"""GPU test of the host job layer end to end: the `Regression` chain (Prepare -> AdmmTrain -> Test -> TestLoglik)
and RegressionNaiveTrain run from a .job config on avro input, and every output file is compared with the oracle."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from oracle import oracle as orc

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import avro_util as au  # noqa: E402

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def host():
    import mlease_b200
    mlease_b200.lib()
    h = C.CDLL(os.path.join(ROOT, "ml-ease_b200", "lib", "libmlease_host.so"))
    h.mlease_job_last_error.restype = C.c_char_p
    return h


def _cfg(path, **kv):
    with open(path, "w") as f:
        for k, v in kv.items():
            f.write("%s=%s\n" % (k.replace("_", "."), v))
    return path


def _model_vec(model_list, names):
    idx = {n: i for i, n in enumerate(names)}
    v = np.zeros(len(names) + 1, np.float32)
    for f in model_list:
        if f["name"] == "(INTERCEPT)":
            v[-1] = f["value"]
        else:
            v[idx[f["name"]]] = f["value"]
    return v


def test_regression_chain_on_fixture(host, tmp_path, fixture_data, frozen):
    npz = np.load(os.path.join(GOLDEN, "sample_data.npz"))
    names = [str(n) for n in npz["feature_names"]]
    recs = au.fixture_records(npz, with_key=lambda i: i // 250)       # partitions = the frozen oracle run's row blocks
    au.write_avro(str(tmp_path / "in" / "part-0.avro"), au.pig_schema_with_key(), recs, codec="deflate", block=128)
    out = str(tmp_path / "out")
    cfg = _cfg(str(tmp_path / "r.job"), input_paths=str(tmp_path / "in"), output_base_path=out, test_path=str(tmp_path / "in"),
               map_key="pkey", num_blocks=4, num_iters=20, regularizer=2, epsilon=0, force_output_overwrite="true")
    open(cfg, "a").write("lambda=1,10,100\n")
    rc = host.mlease_job_run(b"Regression", cfg.encode())
    assert rc == 0, host.mlease_job_last_error().decode()

    # lambda-rho (jobs/RegressionAdmmTrain.java:200, defaults :174-181)
    lr = au.read_dir(out + "/lambda-rho")
    assert sorted((r["lambda"], r["rho"]) for r in lr) == [(1.0, 1.0), (10.0, 1.0), (100.0, 1.0)]
    # final-model = float(z) after 20 iterations, keyed by String.valueOf(float lambda)
    fm = {r["key"]: _model_vec(r["model"], names) for r in au.read_dir(out + "/final-model")}
    assert sorted(fm) == ["1.0", "10.0", "100.0"]
    for li, key in enumerate(("1.0", "10.0", "100.0")):
        ref = frozen["exact_z_hist"][-1, li]
        assert np.abs(fm[key] - ref).max() / np.abs(ref).max() < 1e-5
    # iteration files: iter-1/u is empty, iter-20/model holds float(x), float(u+x) per "lambda#partition"
    assert au.read_dir(out + "/iter-1/u") == []
    assert len(au.read_dir(out + "/iter-2/u")) == 12
    it20 = {r["key"]: r for r in au.read_dir(out + "/iter-20/model")}
    assert len(it20) == 12
    for p in range(4):
        for li, key in enumerate(("1.0", "10.0", "100.0")):
            x = _model_vec(it20["%s#%d" % (key, p)]["model"], names)
            ref = frozen["exact_x_last"][p, li]
            assert np.abs(x - ref).max() / np.abs(ref).max() < 1e-5
    # RegressionTest: pred = float(x.beta + offset) with the float final model; RegressionTestLoglik
    for li, (lam, key) in enumerate((("1", "1.0"), ("10", "10.0"), ("100", "100.0"))):
        preds = np.array([r["pred"] for r in au.read_dir(out + "/test/lambda-" + lam) if "pred" in r], np.float32)
        ref_pred = orc.score(fixture_data, fm[key].astype(np.float64))
        assert len(preds) == 1000 and np.abs(preds - ref_pred).max() <= 2e-6 * np.abs(ref_pred).max()
        ll = au.read_dir(out + "/test/lambda-" + lam + "/_loglik")
        ref_ll, ref_cnt = orc.test_loglik(fixture_data.response, ref_pred, fixture_data.weight, combiner_block=1000)
        assert ll[0]["key"] == "averageTestLoglik" and ll[0]["count"] == ref_cnt
        assert abs(ll[0]["testLoglik"] - ref_ll) <= 1e-6 * abs(ref_ll)
    # per-iteration sample test loglik + best model (:812-845)
    sl = {r["lambda"]: r for r in au.read_avro(out + "/sample-test-loglik/iteration-20.avro")[1]}
    assert sl["1.0"]["iter"] == 20
    ref_sl = orc.sample_test_loglik(fixture_data, frozen["exact_z_hist"][-1, 0])
    assert abs(sl["1.0"]["testLoglik"] - ref_sl) <= 2e-6 * abs(ref_sl)
    assert len(os.listdir(out + "/best-model")) == 1


def test_naive_train_job(host, tmp_path, fixture_data):
    npz = np.load(os.path.join(GOLDEN, "sample_data.npz"))
    names = [str(n) for n in npz["feature_names"]]
    recs = au.fixture_records(npz, with_key=lambda i: i // 250)
    au.write_avro(str(tmp_path / "in" / "part-0.avro"), au.pig_schema_with_key(), recs, block=500)
    out = str(tmp_path / "out")
    cfg = _cfg(str(tmp_path / "p.job"), input_paths=str(tmp_path / "in"), output_path=out + "/tmp-data", map_key="pkey", num_blocks=4)
    assert host.mlease_job_run(b"RegressionPrepare", cfg.encode()) == 0, host.mlease_job_last_error().decode()
    cfg = _cfg(str(tmp_path / "n.job"), output_base_path=out, num_blocks=4, heavy_per_item_train="true", remove_tmp_dir="false")
    open(cfg, "a").write("lambda=10,1\n")
    assert host.mlease_job_run(b"NaiveTrain", cfg.encode()) == 0, host.mlease_job_last_error().decode()
    models = {r["key"]: _model_vec(r["model"], names) for r in au.read_dir(out + "/models")}
    assert len(models) == 8
    prs = [0, 250, 500, 750, 1000]
    for lam, key in ((1.0, "1.0"), (10.0, "10.0")):
        ref, _, _ = orc.naive_train(fixture_data, prs, lam, mode="exact", nthreads=4)
        for k in range(4):
            m = models["%s#%d" % (key, k)]
            assert np.abs(m - ref[k]).max() / np.abs(ref[k]).max() < 1e-5
        mean = {r["key"]: _model_vec(r["model"], names) for r in au.read_dir(out + "/final-model")}[key]
        ref_mean = sum((1.0 / 4) * ref[k].astype(np.float32).astype(np.float64) for k in range(4))
        assert np.abs(mean - ref_mean).max() / np.abs(ref_mean).max() < 1e-5
    ids = {r["key"]: r["value"] for r in au.read_dir(out + "/partitionIds")}
    assert ids == {"1.0#0": 0, "1.0#1": 1, "1.0#2": 2, "1.0#3": 3, "10.0#0": 4, "10.0#1": 5, "10.0#2": 6, "10.0#3": 7}


def _prepared_fixture(host, tmp_path, binary=False):
    npz = np.load(os.path.join(GOLDEN, "sample_data.npz"))
    names = [str(n) for n in npz["feature_names"]]
    recs = au.fixture_records(npz, with_key=lambda i: i // 250)
    au.write_avro(str(tmp_path / "in" / "part-0.avro"), au.pig_schema_with_key(), recs, block=500)
    out = str(tmp_path / "out")
    cfg = _cfg(str(tmp_path / "p.job"), input_paths=str(tmp_path / "in"), output_path=out + "/tmp-data", map_key="pkey", num_blocks=4,
               binary_feature="true" if binary else "false")
    assert host.mlease_job_run(b"RegressionPrepare", cfg.encode()) == 0, host.mlease_job_last_error().decode()
    return names, out


LAMBDA_MAP_SCHEMA = {"type": "record", "name": "LambdaMap", "fields": [{"name": "name", "type": "string"}, {"name": "term", "type": "string"},
                                                                       {"name": "value", "type": "float"}]}


def _lambda_map(tmp_path, names):
    listed = {names[3]: 0.05, names[17]: 25.0, names[100]: 3.0, "not-a-feature": 7.0}
    au.write_avro(str(tmp_path / "lmap" / "part-0.avro"), LAMBDA_MAP_SCHEMA, [{"name": k, "term": "", "value": v} for k, v in listed.items()])
    lm = np.zeros(len(names), np.float32)
    for k, v in listed.items():
        if k in names:
            lm[names.index(k)] = v
    return str(tmp_path / "lmap"), lm


def test_admm_train_job_with_initialize_boost_rate_and_lambda_map(host, tmp_path, fixture_data):
    """Job-level initialize.boost.rate (jobs/RegressionAdmmTrain.java:236-266): <out>/initialModel holds the NaiveTrain fits
    (lambda.map passed on to them, :248), z starts at their mean, iteration 1 runs on rho * boost; lambda.map file ingest
    (ReadLambdaMapConsumer) feeds the z-update weights (:382-386).  Final model vs the oracle in exact mode."""
    names, out = _prepared_fixture(host, tmp_path)
    lmdir, lm = _lambda_map(tmp_path, names)
    cfg = _cfg(str(tmp_path / "t.job"), output_base_path=out, num_blocks=4, num_iters=8, regularizer=2, epsilon=0, initialize_boost_rate=2.5,
               lambda_map=lmdir)
    open(cfg, "a").write("lambda=1,10\n")
    assert host.mlease_job_run(b"RegressionAdmmTrain", cfg.encode()) == 0, host.mlease_job_last_error().decode()
    prs = [0, 250, 500, 750, 1000]
    init = {r["key"]: _model_vec(r["model"], names) for r in au.read_dir(out + "/initialModel")}
    assert len(init) == 8
    for lam, key in ((1.0, "1.0"), (10.0, "10.0")):
        ref, _, _ = orc.naive_train(fixture_data, prs, lam, lambda_map=lm, mode="exact", nthreads=4)
        for k in range(4):
            assert np.abs(init["%s#%d" % (key, k)] - ref[k]).max() / np.abs(ref[k]).max() < 1e-5
    ref = orc.admm_run(fixture_data, prs, [1.0, 10.0], niters=8, mode="exact", nthreads=8, epsilon=0.0, initialize_boost_rate=2.5, lambda_map=lm)
    plain = orc.admm_run(fixture_data, prs, [1.0, 10.0], niters=8, mode="exact", nthreads=8, epsilon=0.0)
    fm = {r["key"]: _model_vec(r["model"], names) for r in au.read_dir(out + "/final-model")}
    for li, key in enumerate(("1.0", "10.0")):
        zr = ref["z_hist"][-1, li]
        assert np.abs(fm[key] - zr).max() / np.abs(zr).max() < 1e-5
        assert np.abs(plain["z_hist"][-1, li] - zr).max() / np.abs(zr).max() > 1e-3      # boost + map really change the model
    # iter-1/init-value is the mean initial model (the reducers' start), not the empty map
    iv = {r["key"]: _model_vec(r["model"], names) for r in au.read_dir(out + "/iter-1/init-value")}
    z0 = sum((1.0 / 4) * init["1.0#%d" % k].astype(np.float64) for k in range(4))
    assert np.abs(iv["1.0"] - z0).max() <= 1e-6 * np.abs(z0).max()


def test_admm_train_job_l1_regularizer(host, tmp_path, fixture_data):
    names, out = _prepared_fixture(host, tmp_path)
    cfg = _cfg(str(tmp_path / "t.job"), output_base_path=out, num_blocks=4, num_iters=6, regularizer=1, epsilon=0)
    open(cfg, "a").write("lambda=1,10\n")
    assert host.mlease_job_run(b"RegressionAdmmTrain", cfg.encode()) == 0, host.mlease_job_last_error().decode()
    ref = orc.admm_run(fixture_data, [0, 250, 500, 750, 1000], [1.0, 10.0], niters=6, mode="exact", nthreads=8, epsilon=0.0, regularizer=1)
    fm = {r["key"]: _model_vec(r["model"], names) for r in au.read_dir(out + "/final-model")}
    for li, key in enumerate(("1.0", "10.0")):
        zr = ref["z_hist"][-1, li]
        assert np.abs(fm[key] - zr).max() / np.abs(zr).max() < 1e-5


def test_naive_train_job_lambda_map_and_binary_feature(host, tmp_path, fixture_data):
    """NaiveTrain job on per-key SPARSE datasets (one CSR upload for all lambdas), lambda.map -> per-feature prior variance
    (jobs/RegressionNaiveTrain.java:318-343), binary.feature -> every listed feature counts as 1."""
    names, out = _prepared_fixture(host, tmp_path, binary=True)
    lmdir, lm = _lambda_map(tmp_path, names)
    cfg = _cfg(str(tmp_path / "n.job"), output_base_path=out, compute_model_mean="false", remove_tmp_dir="false", binary_feature="true", lambda_map=lmdir,
               prior_mean=0.25)
    open(cfg, "a").write("lambda=0.5,5\n")
    assert host.mlease_job_run(b"NaiveTrain", cfg.encode()) == 0, host.mlease_job_last_error().decode()
    models = {r["key"]: _model_vec(r["model"], names) for r in au.read_dir(out + "/models")}
    assert len(models) == 8
    d = fixture_data
    ones = orc.Csr(d.rowptr, d.colidx, np.ones_like(d.val), d.response, d.weight, d.offset, d.n_features)
    for lam, key in ((0.5, "0.5"), (5.0, "5.0")):
        ref, _, _ = orc.naive_train(ones, [0, 250, 500, 750, 1000], lam, lambda_map=lm, prior_mean=0.25, mode="exact", nthreads=4)
        for k in range(4):
            m = models["%s#%d" % (key, k)]
            assert np.abs(m - ref[k]).max() / np.abs(ref[k]).max() < 1e-5, (key, k)


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpus() < 2, reason="needs 2 GPUs (gpu.devices=0,1)")
def test_admm_train_job_on_two_gpus_matches_one(host, tmp_path, fixture_data, frozen):
    """gpu.devices=0,1: the C++ job layer drives both GPUs through mlease_world (worker thread per GPU + NCCL all-reduce inside
    the library); same final model as the frozen oracle run and as the single-GPU job."""
    names, out = _prepared_fixture(host, tmp_path)
    res = {}
    for tag, devs in (("one", "0"), ("two", "0,1")):
        o = out + "_" + tag
        cfg = _cfg(str(tmp_path / (tag + ".job")), input_paths=out + "/tmp-data", output_base_path=o, num_blocks=4, num_iters=20, regularizer=2, epsilon=0,
                   gpu_devices=devs)
        open(cfg, "a").write("lambda=1,10,100\n")
        assert host.mlease_job_run(b"RegressionAdmmTrain", cfg.encode()) == 0, host.mlease_job_last_error().decode()
        res[tag] = {r["key"]: _model_vec(r["model"], names) for r in au.read_dir(o + "/final-model")}
    for li, key in enumerate(("1.0", "10.0", "100.0")):
        ref = frozen["exact_z_hist"][-1, li]
        assert np.abs(res["two"][key] - ref).max() / np.abs(ref).max() < 1e-5
        assert np.abs(res["two"][key] - res["one"][key]).max() <= 2e-6 * np.abs(ref).max()
