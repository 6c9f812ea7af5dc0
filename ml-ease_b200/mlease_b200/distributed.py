"""One process per GPU: partition sharding and the per-iteration all-reduce (torch.distributed is plumbing only).

The path shards exactly where ADMM does: partitions are independent in the x-update (AdmmReducer.reduce,
jobs/RegressionAdmmTrain.java:642-718) and meet in ONE exchange per iteration, the mean of x+u the driver
computes from the reducer outputs (:362-364).  Each rank owns the partitions p with p % world_size == rank,
adds sum_p float(x_p)+u_p over its local partitions (K4 admm_pack), the ranks all-reduce that [L][D'] double
buffer (NCCL over NVLink on GPUs; gloo in the CPU tests of this control flow), and every rank applies the
z-prox redundantly, so no broadcast and no second collective (convergence scalar included) is needed.
"""
from __future__ import annotations


def shard_partitions(num_blocks: int, world_size: int, rank: int):
    """Partition ids owned by `rank` (p -> GPU p mod G)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return [p for p in range(num_blocks) if p % world_size == rank]


def admm_loop(backend, num_iters, allreduce=None):
    """Drives begin / local_step / all-reduce / consensus.

    backend: object with begin(), local_step(exchange), consensus(exchange) -> (maxdiff, stop) and an
    `exchange` attribute (tensor/array the all-reduce acts on in place).  Returns (iters_done, maxdiff history).
    Every rank takes the same stop decision because consensus() is a pure function of the reduced buffer."""
    backend.begin()
    hist = []
    done = 0
    for i in range(1, num_iters + 1):
        backend.local_step(backend.exchange)
        if allreduce is not None:
            allreduce(backend.exchange)
        maxdiff, stop = backend.consensus(backend.exchange)
        hist.append(maxdiff)
        done = i
        if stop:
            break
    return done, hist


class CudaAdmmBackend:
    """AdmmSession + a torch CUDA exchange buffer on the current stream."""

    def __init__(self, session):
        import torch
        self.session = session
        self.exchange = torch.zeros(session.L * session.Dt, dtype=torch.float64, device="cuda:%d" % session.device)

    def begin(self):
        self.session.begin()

    def local_step(self, exchange):
        self.session.local_step(exchange.data_ptr())

    def consensus(self, exchange):
        return self.session.consensus(exchange.data_ptr())


def run_distributed(session, num_iters, group=None):
    """Multi-process ADMM: all ranks call this with their own session (local partitions already added)."""
    import torch.distributed as dist
    be = CudaAdmmBackend(session)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        return admm_loop(be, num_iters, lambda buf: dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group))
    return admm_loop(be, num_iters, None)
