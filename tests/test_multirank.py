"""world_size-2 gloo test (CPU) of the independent-job partition used for multi-GPU runs: every
job runs on exactly one rank, no data-path collective, timings reduce with MAX over ranks."""
from __future__ import annotations

import os
import socket

import torch.multiprocessing as mp

from style_transfer_visualizer_b200 import jobs


def test_partition_covers_every_job_once() -> None:
    for n_jobs in (0, 1, 7, 64):
        for world in (1, 2, 3, 8):
            parts = [jobs.partition_jobs(n_jobs, world, r) for r in range(world)]
            flat = [j for p in parts for j in p]
            assert flat == list(range(n_jobs))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert [len(jobs.partition_jobs(64, 8, r)) for r in range(8)] == [8] * 8


def _worker(rank: int, world: int, port: int, out) -> None:  # noqa: ANN001
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    info = jobs.init_distributed(backend="gloo")
    mine = jobs.partition_jobs(5, info.world_size, info.rank)
    jobs.barrier()
    slowest = jobs.max_over_ranks(1.0 + info.rank)
    total = jobs.sum_over_ranks(float(len(mine)))
    gathered = jobs.gather_objects({"rank": info.rank, "jobs": mine})
    out.put((rank, mine, slowest, total, gathered))
    import torch.distributed as dist

    dist.destroy_process_group()


def test_two_rank_gloo_job_partition() -> None:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, jobs0, slow0, tot0, gath0), (r1, jobs1, slow1, tot1, gath1) = results
    assert (r0, r1) == (0, 1)
    assert jobs0 == [0, 1, 2] and jobs1 == [3, 4]
    assert slow0 == slow1 == 2.0 and tot0 == tot1 == 5.0
    assert [g["jobs"] for g in gath0] == [[0, 1, 2], [3, 4]] and gath1 == []
