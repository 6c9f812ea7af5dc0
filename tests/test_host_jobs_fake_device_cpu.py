"""CPU test of the job layer's ORCHESTRATION: the C++ jobs linked against a test double of the device library
(tests/fake_device/fake_mlease_b200.c: canned numbers, no computation) so that RegressionAdmmTrain / RegressionTest /
RegressionTestLoglik / RegressionNaiveTrain run end to end without a GPU.  What is checked is the host side: the whole output
tree (iter-i/{u,init-value,model}, final-model, best-model, sample-test-loglik, test/lambda-*, _loglik, models, partitionIds)
is the same whether the avro work goes through the block-parallel plan-walker code or the generic Value-tree code, and the
output layout is the reference's.  The numbers in the files mean nothing; the GPU tests (test_gpu_jobs.py) check those."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import avro_util as au  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
HOST = os.path.join(ROOT, "ml-ease_b200", "host")


@pytest.fixture(scope="module")
def fake_host(tmp_path_factory):
    d = tmp_path_factory.mktemp("fakehost")
    so = str(d / "libmlease_host_fake.so")
    san = os.environ.get("MLEASE_TEST_SANITIZE", "")   # e.g. "address,undefined" or "thread" (run pytest with the matching runtime LD_PRELOADed)
    extra = ["-g", "-fsanitize=" + san, "-fno-omit-frame-pointer"] if san else []
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared"] + extra + ["-o", so, os.path.join(HOST, "avro_io.cpp"), os.path.join(HOST, "regression_jobs.cpp"),
                           "-x", "c", os.path.join(ROOT, "tests", "fake_device", "fake_mlease_b200.c"), "-lz", "-pthread", "-lm"])
    h = C.CDLL(so)
    h.mlease_job_last_error.restype = C.c_char_p
    return h


def _cfg(path, **kv):
    with open(path, "w") as f:
        for k, v in kv.items():
            f.write("%s=%s\n" % (k.replace("_", "."), v))
    return path


def _tree(root):
    """relative path -> decoded records of every avro file under root."""
    out = {}
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith(".avro"):
                p = os.path.join(dp, f)
                out[os.path.relpath(p, root)] = au.read_avro(p)[:2]
    return out


def _run(h, job, cfg):
    rc = h.mlease_job_run(job.encode(), cfg.encode())
    assert rc == 0, h.mlease_job_last_error().decode()


def test_regression_chain_and_naive_train_output_trees_do_not_depend_on_the_avro_path(fake_host, tmp_path, monkeypatch):
    npz = np.load(os.path.join(GOLDEN, "sample_data.npz"))
    recs = au.fixture_records(npz, with_key=lambda i: i // 250)
    au.write_avro(str(tmp_path / "in" / "part-0.avro"), au.pig_schema_with_key(), recs[:700], codec="deflate", block=64)
    au.write_avro(str(tmp_path / "in" / "part-1.avro"), au.pig_schema_with_key(), recs[700:], block=500)
    trees = {}
    for mode in ("fast", "generic"):
        monkeypatch.setenv("MLEASE_HOST_GENERIC_INGEST", "1" if mode == "generic" else "0")
        out = str(tmp_path / ("out_" + mode))
        cfg = _cfg(str(tmp_path / (mode + "_r.job")), input_paths=str(tmp_path / "in"), output_base_path=out, test_path=str(tmp_path / "in"),
                   map_key="pkey", num_blocks=4, num_iters=5, regularizer=2, initialize_boost_rate=2.0)
        open(cfg, "a").write("lambda=1,10\n")
        _run(fake_host, "Regression", cfg)
        nout = str(tmp_path / ("nout_" + mode))
        _run(fake_host, "RegressionPrepare", _cfg(str(tmp_path / (mode + "_p.job")), input_paths=str(tmp_path / "in"), output_path=nout + "/tmp-data",
                                                  num_blocks=4, num_click_replicates=2, random_seed=3))
        ncfg = _cfg(str(tmp_path / (mode + "_n.job")), output_base_path=nout, num_blocks=4, heavy_per_item_train="true", remove_tmp_dir="false")
        open(ncfg, "a").write("lambda=10,1\n")
        _run(fake_host, "NaiveTrain", ncfg)
        trees[mode] = (_tree(out), _tree(nout))
    for a, b in zip(trees["fast"], trees["generic"]):
        assert sorted(a) == sorted(b)
        for k in a:
            assert a[k] == b[k], k
    chain, naive = trees["fast"]
    # the reference's output layout: per-iteration files, final / best model, test outputs per lambda as typed
    for rel in ("tmp-data/part-00000.avro", "lambda-rho/part-r-00000.avro", "initialModel/part-r-00000.avro", "iter-1/u/part-r-00000.avro",
                "iter-1/init-value/part-r-00000.avro", "iter-5/model/part-r-00000.avro", "final-model/part-r-00000.avro",
                "sample-test-loglik/iteration-5.avro", "sample-test-loglik/iteration-0.avro", "sample-test-loglik/write-test-00000.avro", "test/lambda-1/part-r-00000.avro", "test/lambda-10/part-r-00001.avro",
                "test/lambda-1/_loglik/part-r-00000.avro"):
        assert rel in chain, (rel, sorted(chain)[:40])
    assert chain["sample-test-loglik/write-test-00000.avro"][1] == [] and [r["iter"] for r in chain["sample-test-loglik/iteration-0.avro"][1]] == [0, 0]
    assert not any(k.startswith("best-model/best-iteration-0") for k in chain)
    assert len(chain["iter-5/model/part-r-00000.avro"][1]) == 8 and len(chain["iter-2/u/part-r-00000.avro"][1]) == 8 and chain["iter-1/u/part-r-00000.avro"][1] == []
    m0 = chain["iter-5/model/part-r-00000.avro"][1][0]
    assert m0["key"] == "1.0#0" and m0["model"][0]["name"] == "(INTERCEPT)" and len(m0["model"]) == 201 and len(m0["uplusx"]) == 201
    assert [f["name"] for f in chain["test/lambda-1/part-r-00000.avro"][0]["fields"]][-1] == "pred"
    assert len(naive["models/part-r-00000.avro"][1]) == 8 and "partitionIds/part-r-00000.avro" in naive and "final-model/part-r-00000.avro" in naive


def test_write_iteration_files_false_keeps_the_final_outputs(fake_host, tmp_path):
    npz = np.load(os.path.join(GOLDEN, "sample_data.npz"))
    recs = au.fixture_records(npz, with_key=lambda i: i // 250)
    au.write_avro(str(tmp_path / "in" / "part-0.avro"), au.pig_schema_with_key(), recs, block=300)
    trees = {}
    for flag in ("true", "false"):
        out = str(tmp_path / ("out_" + flag))
        cfg = _cfg(str(tmp_path / (flag + ".job")), input_paths=str(tmp_path / "in"), output_base_path=out, test_path=str(tmp_path / "in"), map_key="pkey",
                   num_blocks=4, num_iters=3, regularizer=2, write_iteration_files=flag)
        open(cfg, "a").write("lambda=1\n")
        _run(fake_host, "Regression", cfg)
        trees[flag] = _tree(out)
    assert any(k.startswith("iter-") for k in trees["true"]) and not any(k.startswith("iter-") for k in trees["false"])
    rest = {k: v for k, v in trees["true"].items() if not k.startswith("iter-")}
    assert rest == trees["false"] and "final-model/part-r-00000.avro" in rest and "test/lambda-1/_loglik/part-r-00000.avro" in rest


def test_naive_train_intercept_key_other_than_the_dataset_intercept(fake_host, tmp_path):
    """jobs/RegressionNaiveTrain.java:340-343 files the intercept's variance 100000 under `intercept.key`; the dataset's intercept
    is always "(INTERCEPT)", so another name moves that entry onto the feature of that name (the test double adds lambda_map[j]
    to coefficient j, which makes the entry visible)."""
    npz = np.load(os.path.join(GOLDEN, "sample_data.npz"))
    names = [str(n) for n in npz["feature_names"]]
    recs = au.fixture_records(npz, with_key=lambda i: i // 500)
    au.write_avro(str(tmp_path / "in" / "part-0.avro"), au.pig_schema_with_key(), recs, block=300)
    models = {}
    for tag, extra in (("default", {}), ("moved", {"intercept_key": names[5]})):
        out = str(tmp_path / ("out_" + tag))
        _run(fake_host, "RegressionPrepare", _cfg(str(tmp_path / (tag + "_p.job")), input_paths=str(tmp_path / "in"), output_path=out + "/tmp-data", map_key="pkey", num_blocks=2))
        cfg = _cfg(str(tmp_path / (tag + "_n.job")), output_base_path=out, compute_model_mean="false", remove_tmp_dir="false", **extra)
        open(cfg, "a").write("lambda=1\n")
        _run(fake_host, "NaiveTrain", cfg)
        models[tag] = {r["key"]: {f["name"]: f["value"] for f in r["model"]} for r in au.read_dir(out + "/models")}
    for key in models["default"]:
        d, m = models["default"][key], models["moved"][key]
        diff = {n: m[n] - d[n] for n in d if m[n] != d[n]}
        assert set(diff) == {names[5]} and abs(diff[names[5]] - 1e-5) < 2e-6, diff


def test_per_iteration_test_loglik_ignores_num_click_replicates_like_the_reference(fake_host, tmp_path):
    """updateLogLikBestModel is handed the train job's num.click.replicates but calls testloglik(conf, z, testPath, 1, ignoreValue)
    (jobs/RegressionAdmmTrain.java:817): the per-iteration value is computed with n = 1 whatever the job says, i.e. the intercept term
    -log(n - 1 + n exp(-b)) (models/LinearModel.java:241-244) is b itself."""
    npz = np.load(os.path.join(GOLDEN, "sample_data.npz"))
    recs = au.fixture_records(npz, with_key=lambda i: i // 500)
    au.write_avro(str(tmp_path / "in" / "part-0.avro"), au.pig_schema_with_key(), recs, block=300)
    lls = {}
    for reps in (1, 3):
        out = str(tmp_path / ("out%d" % reps))
        _run(fake_host, "RegressionPrepare", _cfg(str(tmp_path / "p.job"), input_paths=str(tmp_path / "in"), output_path=out + "/tmp-data", map_key="pkey", num_blocks=2))
        cfg = _cfg(str(tmp_path / "t.job"), output_base_path=out, num_blocks=2, num_iters=2, regularizer=2, test_path=str(tmp_path / "in"), num_click_replicates=reps,
                   write_iteration_files="false")
        open(cfg, "a").write("lambda=1\n")
        _run(fake_host, "RegressionAdmmTrain", cfg)
        lls[reps] = au.read_avro(out + "/sample-test-loglik/iteration-2.avro")[1][0]["testLoglik"]
        zfin = {f["name"]: f["value"] for f in au.read_dir(out + "/final-model")[0]["model"]}
    assert np.isfinite(lls[1]) and lls[1] == lls[3]
    # n = 1 against a direct evaluation with the (float) final model: agreement to float precision of the model
    names = [str(n) for n in npz["feature_names"]]
    beta = np.array([zfin[n] for n in names]); b = zfin["(INTERCEPT)"]
    xb = b + np.array([sum(beta[c] * v for c, v in zip(npz["colidx"][npz["rowptr"][i]:npz["rowptr"][i + 1]], npz["val"][npz["rowptr"][i]:npz["rowptr"][i + 1]])) for i in range(len(recs))])
    y = npz["response"]
    ref = np.mean(np.where(y == 1, -np.log1p(np.exp(-xb)), -np.log1p(np.exp(xb))))
    assert abs(lls[1] - ref) < 1e-4 * abs(ref)


def test_remove_tmp_dir_leaves_only_the_results(fake_host, tmp_path):
    """remove.tmp.dir=true (jobs/RegressionAdmmTrain.java:474-478,503-520): iteration directories, initialModel and tmp-data are gone at
    the end; lambda-rho, final-model, best-model and the sample log-likelihoods stay."""
    npz = np.load(os.path.join(GOLDEN, "sample_data.npz"))
    recs = au.fixture_records(npz, with_key=lambda i: i // 500)
    au.write_avro(str(tmp_path / "in" / "part-0.avro"), au.pig_schema_with_key(), recs, block=300)
    out = str(tmp_path / "out")
    _run(fake_host, "RegressionPrepare", _cfg(str(tmp_path / "p.job"), input_paths=str(tmp_path / "in"), output_path=out + "/tmp-data", map_key="pkey", num_blocks=2))
    cfg = _cfg(str(tmp_path / "t.job"), output_base_path=out, num_blocks=2, num_iters=4, regularizer=2, test_path=str(tmp_path / "in"), remove_tmp_dir="true",
               initialize_boost_rate=1.5)
    open(cfg, "a").write("lambda=1,10\n")
    _run(fake_host, "RegressionAdmmTrain", cfg)
    left = sorted(os.listdir(out))
    assert left == ["best-model", "final-model", "lambda-rho", "sample-test-loglik"], left


def test_naive_train_models_list_the_features_of_their_key_only(fake_host, tmp_path):
    """A NaiveTrain reducer's dataset holds the features its key's rows list, and so does its model (llf/LibLinear.java:343-350; no
    prior-mean map in NaiveTrain): not every feature of the job.  The mean model still covers all of them."""
    schema = {"type": "record", "name": "r", "fields": [
        {"name": "features", "type": {"type": "array", "items": {"type": "record", "name": "f", "fields": [
            {"name": "name", "type": "string"}, {"name": "term", "type": "string"}, {"name": "value", "type": "float"}]}}},
        {"name": "response", "type": "int"}, {"name": "pkey", "type": "int"}]}
    def rec(key, feats, y):
        return {"features": [{"name": n, "term": t, "value": 1.0} for n, t in feats], "response": y, "pkey": key}
    recs = [rec(0, [("a", ""), ("b", "x")], 1), rec(0, [("a", "")], 0), rec(1, [("b", "x"), ("c", "")], 1), rec(1, [("c", "")], 0)]
    au.write_avro(str(tmp_path / "in" / "p.avro"), schema, recs)
    out = str(tmp_path / "out")
    _run(fake_host, "RegressionPrepare", _cfg(str(tmp_path / "p.job"), input_paths=str(tmp_path / "in"), output_path=out + "/tmp-data", map_key="pkey", num_blocks=2))
    cfg = _cfg(str(tmp_path / "n.job"), output_base_path=out, num_blocks=2, remove_tmp_dir="false")
    open(cfg, "a").write("lambda=1\n")
    _run(fake_host, "NaiveTrain", cfg)
    models = {r["key"]: [(f["name"], f["term"]) for f in r["model"]] for r in au.read_dir(out + "/models")}
    assert models == {"1.0#0": [("(INTERCEPT)", ""), ("a", ""), ("b", "x")], "1.0#1": [("(INTERCEPT)", ""), ("b", "x"), ("c", "")]}
    mean = au.read_dir(out + "/final-model")
    assert len(mean) == 1 and [(f["name"], f["term"]) for f in mean[0]["model"]] == [("(INTERCEPT)", ""), ("a", ""), ("b", "x"), ("c", "")]
