#!/usr/bin/env python
"""Informational probe of BASELINE config 2/3 shape (sparse, 1% nnz, D=10k, multi-lambda) on ONE GPU with the partitions
that GPU would own at N=8 (default 1 partition x 3 lambdas).  Not the bench line (that is config 1 dense, bench.py)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ml-ease_b200"))
import numpy as np
import torch
import mlease_b200 as mb

P = int(os.environ.get("P", 1)); n = int(os.environ.get("N", 1_000_000)); D = int(os.environ.get("D", 10_000)); nnz = int(os.environ.get("NNZ", 100))
lambdas = [float(x) for x in os.environ.get("LAMBDAS", "0.1,1,10").split(",")]
iters = int(os.environ.get("ITERS", 10))
dev = "cuda:0"
g = torch.Generator(device=dev); g.manual_seed(3)
beta = torch.randn(D, generator=g, device=dev) / nnz ** 0.5
parts = []
for p in range(P):
    start = torch.randint(0, D, (n, 1), generator=g, device=dev)
    stride = torch.randint(1, D // nnz, (n, 1), generator=g, device=dev)
    cols = (start + stride * torch.arange(nnz, device=dev)[None, :]) % D          # distinct columns per row
    cols, _ = torch.sort(cols, dim=1)
    vals = torch.randn(n, nnz, generator=g, device=dev)
    s = (vals * beta[cols]).sum(1) - 1.0
    y = (torch.rand(n, generator=g, device=dev) < torch.sigmoid(s)).to(torch.int32)
    rowptr = torch.arange(n + 1, device=dev, dtype=torch.int64) * nnz
    parts.append((rowptr, cols.to(torch.int32).reshape(-1).contiguous(), vals.reshape(-1).contiguous(), y))
torch.cuda.synchronize()
t0 = time.perf_counter()
sess = mb.AdmmSession(P, D, lambdas, epsilon=0.0, stream=torch.cuda.current_stream().cuda_stream)
for p, (rp, ci, v, y) in enumerate(parts):
    sess.add_partition_csr(p, rp, ci, v, y)
sess.begin()
torch.cuda.synchronize()
print("setup %.2f s, mem %.1f GB" % (time.perf_counter() - t0, torch.cuda.mem_get_info()[0] / 1e9), flush=True)
sess.profile(2)
for it in range(iters):
    t1 = time.perf_counter()
    md, stop = sess.iterate()
    torch.cuda.synchronize()
    print("iter %d: %.3f s  maxdiff %.3e  stats %s" % (it + 1, time.perf_counter() - t1, md, {k: v for k, v in sess.stats().items() if k in ("k1_passes", "gram_builds", "newton_steps", "last_iter_slots", "not_converged")}), flush=True)
print(json.dumps(sess.profile(0)))
