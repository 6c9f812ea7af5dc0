"""world_size-2 gloo tests (CPU) of the N>1 control flow: partition sharding + the single all-reduce +
redundant consensus.  The per-partition solve is played by the ORACLE here (test double, tests only); the
product backend is mlease_b200.distributed.CudaAdmmBackend."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    """Restates one rank's share of an ADMM iteration with oracle solves (float rounding points as the reference)."""

    def __init__(self, data, part_rowstart, my_parts, P, lam, rho=1.0):
        from oracle import oracle as orc
        self.orc, self.data, self.prs, self.parts, self.P = orc, data, part_rowstart, my_parts, P
        self.lam, self.rho = np.float32(lam), np.float32(rho)
        self.Dt = data.n_features + 1
        self.exchange = torch.zeros(self.Dt, dtype=torch.float64)

    def _sub(self, p):
        d, r0, r1 = self.data, self.prs[p], self.prs[p + 1]
        sl = slice(d.rowptr[r0], d.rowptr[r1])
        return self.orc.Csr(d.rowptr[r0:r1 + 1] - d.rowptr[r0], d.colidx[sl], d.val[sl], d.response[r0:r1], d.weight[r0:r1],
                            d.offset[r0:r1], d.n_features)

    def begin(self):
        self.z = np.zeros(self.Dt)
        self.u = {p: np.zeros(self.Dt, np.float32) for p in self.parts}
        self.uplusx = {}

    def local_step(self, exchange):
        zf = self.z.astype(np.float32).astype(np.float64)
        s = np.zeros(self.Dt)
        for p in self.parts:
            u = self.u[p].astype(np.float64)
            x, _ = self.orc.liblinear_train(self._sub(p), zf, zf - u, np.full(self.Dt, 1.0 / float(self.rho)), 1e-14, 100000)
            self.uplusx[p] = (u + x).astype(np.float32)
            s += x.astype(np.float32).astype(np.float64) + u
        exchange.copy_(torch.from_numpy(s))

    def consensus(self, exchange):
        S = exchange.numpy().copy()
        pr = np.float32(self.P) * self.rho
        w = float(pr / (self.lam + pr))
        zn = w * (S / self.P)
        zn[-1] = S[-1] / self.P
        md = float(np.abs(zn - self.z).max())
        self.z = zn
        for p in self.parts:
            self.u[p] = (self.uplusx[p].astype(np.float64) - zn).astype(np.float32)
        return md, False


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "ml-ease_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mlease_b200.distributed import admm_loop, shard_partitions
    from oracle import oracle as orc
    d = np.load(os.path.join(ROOT, "tests", "golden", "sample_data.npz"))
    data = orc.Csr(d["rowptr"], d["colidx"], d["val"], d["response"], d["weight"], d["offset"], 200)
    P = 4
    prs = np.linspace(0, 1000, P + 1).astype(np.int64)
    be = OracleBackend(data, prs, shard_partitions(P, world, rank), P, 10.0)
    done, hist = admm_loop(be, 4, (lambda b: dist.all_reduce(b)) if world > 1 else None)
    q.put((rank, done, be.z, hist))
    dist.barrier()
    dist.destroy_process_group()


def _run(world, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    out = [q.get(timeout=240) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
    return sorted(out, key=lambda t: t[0])


def test_shard_partitions():
    from mlease_b200.distributed import shard_partitions
    assert shard_partitions(8, 1, 0) == list(range(8))
    assert shard_partitions(8, 2, 1) == [1, 3, 5, 7]
    assert sorted(sum((shard_partitions(64, 8, r) for r in range(8)), [])) == list(range(64))
    with pytest.raises(ValueError):
        shard_partitions(8, 2, 2)


def test_two_ranks_equal_one_rank_and_oracle(frozen):
    one = _run(1, 29611)
    two = _run(2, 29613)
    z1 = one[0][2]
    for r in range(2):
        assert two[r][1] == 4
        # the all-reduce changes only the summation order of the exchange vector
        np.testing.assert_allclose(two[r][2], z1, rtol=0, atol=1e-12)
    np.testing.assert_array_equal(two[0][2], two[1][2])   # every rank holds the same z (redundant consensus)
    ref = frozen["exact_z_hist"][3, 1]                     # lambda = 10 is index 1 of the frozen run, 4 iterations
    assert np.abs(z1 - ref).max() / np.abs(ref).max() < 1e-6
