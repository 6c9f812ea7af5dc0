#!/usr/bin/env python
"""bench.py -- ADMM iterations/s of the B200 hot path on BASELINE.json's metric config.

Workload (config.workload): BASELINE configs[1] = 8 partitions x 1M rows x 1k dense features, lambda = 1
(synthetic).  The 8 partitions are sharded over the N ranks (p % N), so per-GPU work shrinks as N
grows: scaling = "strong".  One "step" = one ADMM iteration (x-update of every partition + the consensus
all-reduce + z/u update).  The timed region is a complete job of K iterations FROM THE COLD STATE z = u = 0
(the reference's num.iters loop, jobs/RegressionAdmmTrain.java:281), after W warm-up iterations of a throw-away
job; inputs are 4 GB per partition, far larger than the 126 MB L2, so no flush is needed between iterations.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference ...                      (CPU arm: the oracle port, rank 0 only)

Prints ONE JSON line (rank 0).  `value` = iterations/s with inputs resident in HBM; `e2e` = the same job through
the public API from pinned HOST buffers (upload + K iterations + model read-back in the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "ml-ease_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--partitions", type=int, default=8)
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--features", type=int, default=1000)
    ap.add_argument("--lambda_", type=float, default=1.0)
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows per partition of the bounded CPU sample (0 = auto)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--hessian-policy", type=int, default=0)
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            j = json.load(f)
        return dict(hbm=float(j["hbm_gbs"]), tf_burst=float(j["bf16_tflops"]), tf_sust=float(j["bf16_tflops_sustained"]), src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


# ------------------------------------------------------------------------------------------------ synthetic data
def true_beta(D, seed=999):
    rng = np.random.default_rng(seed)
    return (rng.normal(size=D) / np.sqrt(D)).astype(np.float32)


def gen_partition_torch(p, n, D, beta, device):
    """Partition p: x ~ N(0,1) fp32, y ~ Bernoulli(sigmoid(x.beta* - 1)); seed 1000+p."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(1000 + p)
    X = torch.randn(n, D, generator=g, device=device, dtype=torch.float32)
    b = torch.as_tensor(beta, device=device)
    s = X @ b - 1.0
    y = (torch.rand(n, generator=g, device=device) < torch.sigmoid(s)).to(torch.int32)
    return X, y


# ------------------------------------------------------------------------------------------------ clocks sampler
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.lines, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._rd, daemon=True).start()
        except Exception:
            self.proc = None

    def _rd(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:   # timed region shorter than one sampling period: take one reading now (GPU still warm)
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=10).stdout.strip()
                f = [x.strip() for x in o.split(",")]
                sm.append(float(f[1])); mx = max(mx, float(f[2]))
                for nm, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
def cpu_arm(args, steps, rows_cpu=None):
    """The reference's CPU path for this metric: the faithful oracle (TRON, reference tolerance schedule), one
    single-threaded solve per (partition, lambda) like a Hadoop reducer, min(P, cores) solves in parallel.  Times
    x-update + z/u update only (no Hadoop launch / shuffle / per-iteration avro re-ingest: flatters the reference).
    Bounded sample: `rows_cpu` rows per partition, iterations/s extrapolated linearly in rows to args.rows."""
    from oracle import oracle as orc
    import torch
    cores = os.cpu_count() or 1
    P, D = args.partitions, args.features
    threads = min(P, cores)
    beta = true_beta(D)
    n = rows_cpu or args.cpu_rows or 10000
    Xs, ys = [], []
    for p in range(P):
        X, y = gen_partition_torch(p, n, D, beta, "cpu")
        Xs.append(X.numpy()); ys.append(y.numpy())
    data = orc.Csr.from_dense(np.vstack(Xs), np.concatenate(ys))
    prs = np.arange(P + 1, dtype=np.int64) * n
    t0 = time.perf_counter()
    r = orc.admm_run(data, prs, [args.lambda_], niters=steps, epsilon=0.0, mode="faithful", nthreads=threads)
    dt = time.perf_counter() - t0
    its = r["iters_done"]
    scale = n / float(args.rows)
    val = its / dt * scale
    return dict(value=val, unit="ADMM iterations/s", cores=threads, kind="port",
                sample="%d partitions x %d rows x %d features (%.1f%% of rows), %d iterations in %.1f s, %d sparse passes; "
                       "extrapolated linearly in rows to %d rows/partition; host has %d cores" % (P, n, D, 100 * scale, its, dt, r["passes"], args.rows, cores),
                seconds=dt, iters=its, passes=int(r["passes"]))


_REAL_STDOUT = None


def isolate_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, etc.) may write to fd 1 from any
    rank, so fd 1 is pointed at stderr for the whole process and the JSON line is written to the saved descriptor."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, line)
    else:
        sys.stdout.write(line.decode()); sys.stdout.flush()


def main():
    args = parse()
    isolate_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K, W = args.steps, max(args.warmup, 0)
    cfg = {"workload": "8 partitions x 1M x 1k dense, lambda=1 (BASELINE configs[1]); partitions sharded p%N over ranks"
           if (args.partitions, args.rows, args.features) == (8, 1_000_000, 1000) else
           "%d partitions x %d x %d dense, lambda=%g" % (args.partitions, args.rows, args.features, args.lambda_),
           "partitions": args.partitions, "rows_per_partition": args.rows, "features": args.features, "lambda": args.lambda_,
           "num_iters": K, "timed_region": "cold-start job of K iterations (z=u=0)", "l2": "inputs_larger_than_L2 (4 GB/partition)",
           "parallelism": "partitions p%%N over %d rank(s), one NCCL all-reduce of [L][D'] fp64 per iteration" % world}
    base = {"metric": "ADMM iterations/sec", "unit": "ADMM iterations/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 data / f64 reductions / bf16 Gram operands",
            "data": "synthetic", "config": cfg}

    if args.impl == "reference":
        if rank != 0:
            return
        cb = cpu_arm(args, K)
        out = dict(base)
        out.update({"impl": "reference", "value": cb["value"], "ms_per_step": 1000.0 / cb["value"], "n_gpus": args.gpus,
                    "samples_per_s": cb["value"] * args.partitions * args.rows, "gpu_launches": 0,
                    "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                    "e2e": {"value": cb["value"], "unit": "ADMM iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        emit(out)
        return

    import torch
    import torch.distributed as dist
    import mlease_b200 as mb
    from mlease_b200.distributed import CudaAdmmBackend, admm_loop, shard_partitions

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = "cuda:%d" % local_rank
    if world > 1:
        # keep stdout to the single JSON line: NCCL's version banner / debug lines go to a file
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nccl_debug_%h_%p.log")
        dist.init_process_group("nccl", device_id=torch.device(dev))
    P, n, D, L = args.partitions, args.rows, args.features, 1
    my_parts = shard_partitions(P, world, rank)
    beta = true_beta(D)
    stream = torch.cuda.current_stream().cuda_stream

    def allreduce(buf):
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_session():
        return mb.AdmmSession(P, D, [args.lambda_], device=local_rank, stream=stream, epsilon=0.0, hessian_policy=args.hessian_policy)

    # ---------------- device-resident leg ("value") ----------------
    sess = make_session()
    host_parts = {}
    for p in my_parts:
        X, y = gen_partition_torch(p, n, D, beta, dev)
        sess.add_partition_dense(p, X, y)                      # device pointers: D2D copy into the padded layout
        if not args.no_e2e:
            hx = torch.empty((n, D), dtype=torch.float32, pin_memory=True)
            hy = torch.empty((n,), dtype=torch.int32, pin_memory=True)
            hx.copy_(X); hy.copy_(y)
            host_parts[p] = (hx, hy)
        del X, y
    torch.cuda.synchronize()
    be = CudaAdmmBackend(sess)
    ar = allreduce if world > 1 else None
    if W > 0:
        admm_loop(be, W, ar)                                   # warm-up: a throw-away job of W iterations
    barrier()
    sess.profile(2)
    st0 = sess.stats()
    clk = Clocks(local_rank); clk.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    done, hist = admm_loop(be, K, ar)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = clk.stop()
    st1 = sess.stats()
    prof = sess.profile(0)
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    z_final = sess.z(0)
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    sess.close(); del be, sess
    torch.cuda.empty_cache()

    # ---------------- end-to-end leg (host buffers, public API) ----------------
    e2e = None
    if not args.no_e2e:
        barrier()
        t0 = time.perf_counter()
        s2 = make_session()
        h2d = 0
        for p in my_parts:
            hx, hy = host_parts[p]
            s2.add_partition_dense(p, hx, hy)                  # pinned host -> device inside the timed region
            h2d += hx.numel() * 4 + hy.numel() * 4
        be2 = CudaAdmmBackend(s2)
        done2, _ = admm_loop(be2, K, ar)
        model = s2.final_model(0)                              # device -> host read of the job's result
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tdt = torch.tensor([dt], dtype=torch.float64, device=dev)
        th = torch.tensor([float(h2d)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tdt, op=dist.ReduceOp.MAX); dist.all_reduce(th, op=dist.ReduceOp.SUM)
        e2e = {"value": done2 / float(tdt.item()), "unit": "ADMM iterations/s", "h2d_bytes_per_step": float(th.item()) / done2,
               "d2h_bytes_per_step": (model.nbytes + 8 * done2) * world / done2, "seconds": float(tdt.item()),
               "note": "upload once (the reference re-ingests every iteration), K iterations, model read-back"}
        s2.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    out = dict(base)
    traffic = None   # DRAM bytes of one steady-state K1 launch from the committed ncu --set full capture (same command, N=1)
    try:
        if world == 1 and (P, n, D) == (8, 1_000_000, 1000):
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_k1_traffic.json")))["traffic_bytes_per_launch"]
    except Exception:
        traffic = None
    val = done / (ms / 1000.0)
    k1_ms, k1_n = prof["ms"]["k1"], prof["launches"]["k1"]
    gr_ms, gr_n = prof["ms"]["gram"], prof["launches"]["gram"]
    k1_total_bytes = prof["k1_bytes"]
    k1_gbs = (k1_total_bytes / 1e9) / (k1_ms / 1e3) if k1_ms > 0 else None
    roof = {"kernel": "k1_dense_kernel (fused score+reweight+gradient, one pass over X)", "bound": "hbm",
            "achieved": k1_gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": (k1_gbs / pk["hbm"]) if k1_gbs else None, "traffic": traffic,
            "peak_source": pk["src"] + " hbm_gbs (copy)", "launches": k1_n, "avg_launch_ms": k1_ms / max(k1_n, 1),
            "algorithmic_bytes_per_launch": k1_total_bytes / max(k1_n, 1),
            "emit_bytes_not_counted": prof["k1_emit_bytes"], "share_of_step": k1_ms / ms}
    gram_tf = (prof["gram_flops"] / 1e12) / (gr_ms / 1e3) if gr_ms > 0 else None
    roof_gram = {"kernel": "gram_tcgen05_kernel", "bound": "tensor", "achieved": gram_tf, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                 "frac": (gram_tf / pk["tf_sust"]) if gram_tf else None, "launches": gr_n, "avg_launch_ms": gr_ms / max(gr_n, 1),
                 "flops": "n*D'*(D'+1) per build (lower triangle)", "peak_source": pk["src"] + " bf16 sustained", "share_of_step": gr_ms / ms}
    out.update({"value": val, "ms_per_step": ms / done, "samples_per_s": val * P * n, "iters_done": done,
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "roofline_gram": roof_gram,
                "kernel_ms": prof["ms"], "kernel_launch_counts": prof["launches"],
                "solver": {"k1_passes": st1["k1_passes"] - st0["k1_passes"], "gram_builds": st1["gram_builds"] - st0["gram_builds"],
                           "newton_steps": st1["newton_steps"] - st0["newton_steps"], "rejected": st1["rejected_steps"] - st0["rejected_steps"],
                           "not_converged": st1["not_converged"], "last_maxdiff": hist[-1] if hist else None},
                "z_checksum": float(np.abs(z_final).sum())})
    if e2e:
        out["e2e"] = e2e
    if world == 1 and not args.no_cpu:
        cb = cpu_arm(args, min(K, 20))
        out["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
