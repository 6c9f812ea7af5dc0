"""One process per GPU: partition sharding and the per-iteration all-reduce.

Product path (`make_comm` + `AdmmSession.set_comm` + `session.run`): the all-reduce is the library's own NCCL call inside the
C loop (csrc/comm.cu); torch.distributed only ships the 128-byte NCCL id between the processes.  `admm_loop` below is the
same control flow written out in Python over an abstract backend -- it is what the gloo CPU tests drive (and a fallback
driver for callers that want to own the collective through the `mlease_allreduce_fn` callback).

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


def make_comm(device, group=None):
    """The library's NCCL communicator for this rank of an initialised torch.distributed job (None for a single process):
    rank 0 draws the id, torch.distributed broadcasts the 128 bytes, every rank joins with its own GPU."""
    import torch.distributed as dist
    from .admm import Comm
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return Comm(box[0], rank, world, device)


def run_distributed(session, num_iters, group=None, comm=None):
    """Multi-process ADMM: all ranks call this with their own session (local partitions already added).  The loop runs in C
    (mlease_admm_run) with one ncclAllReduce per iteration on the session's stream.  Returns (iters_done, [last maxdiff])."""
    own = comm is None
    if own:
        comm = make_comm(session.device, group)
    session.set_comm(comm)
    try:
        done = session.run(num_iters)
        return done, [session.stats()["last_maxdiff"]]
    finally:
        session.set_comm(None)
        if own and comm is not None:
            comm.close()
