"""Rank program of test_two_gpus_world_and_process_per_gpu_match_the_oracle: one process per GPU, partitions p % N == rank,
the whole ADMM loop in C (mlease_admm_run) with the library's NCCL all-reduce; torch.distributed only ships the NCCL id."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ml-ease_b200")):
    sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    import mlease_b200 as mb
    from mlease_b200.distributed import run_distributed, shard_partitions
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")                       # plumbing only: the data path's collective is the library's own NCCL
    g = np.load(os.path.join(ROOT, "tests", "golden", "sample_data.npz"))
    prs = np.load(os.path.join(ROOT, "tests", "golden", "oracle_frozen.npz"))["part_rowstart"]
    P = len(prs) - 1
    with mb.AdmmSession(P, len(g["feature_names"]), [1.0, 10.0, 100.0], device=local, epsilon=0.0) as s:
        for p in shard_partitions(P, world, rank):
            r0, r1 = prs[p], prs[p + 1]
            sl = slice(g["rowptr"][r0], g["rowptr"][r1])
            s.add_partition_csr(p, g["rowptr"][r0:r1 + 1] - g["rowptr"][r0], g["colidx"][sl], g["val"][sl], g["response"][r0:r1], g["weight"][r0:r1],
                                g["offset"][r0:r1])
        done, _ = run_distributed(s, 20)
        assert done == 20
        z = np.stack([s.z(l) for l in range(3)])
    zs = [None] * world
    dist.all_gather_object(zs, z)
    assert all(np.array_equal(zs[0], zz) for zz in zs), "ranks disagree on z"
    if rank == 0:
        np.save(sys.argv[1], z)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
