#!/usr/bin/env python
"""Throughput probe of BASELINE config 4 (NaiveTrain: K keys x n rows x D dense features, independent fits) at reduced K."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ml-ease_b200"))
import numpy as np
import torch
import mlease_b200 as mb

K, n, D = int(sys.argv[1]) if len(sys.argv) > 1 else 2048, 1000, 256
g = torch.Generator(device="cuda"); g.manual_seed(5)
X = torch.randn(K * n, D, generator=g, device="cuda")
beta = torch.randn(D, generator=g, device="cuda") / D ** 0.5
y = (torch.rand(K * n, generator=g, device="cuda") < torch.sigmoid(X @ beta - 0.5)).to(torch.int32)
krs = np.arange(K + 1, dtype=np.int64) * n
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    m, sk = mb.naive_train_dense(X, krs, y, 1.0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("naive_train_dense: %d keys x %d x %d in %.3f s -> %.0f fits/s (%.2f GB of X)" % (K, n, D, dt, K / dt, K * n * D * 4 / 1e9))
# spot check 3 keys against the oracle
from oracle import oracle as orc
Xh, yh = X[:3 * n].cpu().numpy(), y[:3 * n].cpu().numpy()
ref, _, _ = orc.naive_train(orc.Csr.from_dense(Xh, yh), [0, n, 2 * n, 3 * n], 1.0, mode="exact")
print("max rel err vs oracle (3 keys):", float(np.abs(m[:3] - ref).max() / np.abs(ref).max()))
