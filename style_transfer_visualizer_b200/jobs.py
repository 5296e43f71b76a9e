"""Independent-job data parallelism (BASELINE.json configs[3]: N content/style pairs across
1/2/4/8 GPUs).  Each pair is its own optimisation with no shared state except the read-only VGG
weights, so the jobs are statically partitioned over ranks -- one process per GPU -- and there is
NO collective on the data path; ``torch.distributed`` is used only to rendezvous, to barrier the
timed region and to gather per-rank timings / summaries."""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch
import torch.distributed as dist


def partition_jobs(n_jobs: int, world_size: int, rank: int) -> list[int]:
    """Contiguous static partition; the first ``n_jobs % world_size`` ranks get one extra job."""
    if world_size < 1 or not 0 <= rank < world_size:
        msg = f"bad rank {rank} for world size {world_size}"
        raise ValueError(msg)
    base, extra = divmod(n_jobs, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


@dataclass
class RankInfo:
    rank: int
    world_size: int
    local_rank: int


def init_distributed(backend: str | None = None) -> RankInfo:
    """Join the process group described by RANK / WORLD_SIZE / MASTER_* (torchrun); a plain
    single-process launch needs no group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        be = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kwargs = {}
        if be == "nccl":
            torch.cuda.set_device(local)
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=be, rank=rank, world_size=world, **kwargs)
    return RankInfo(rank=rank, world_size=world, local_rank=local)


def shutdown() -> None:
    if dist.is_initialized():
        dist.destroy_process_group()


def barrier() -> None:
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device: torch.device | None = None) -> float:
    """MAX-reduce a host scalar (timings are reported as the slowest rank's)."""
    if not dist.is_initialized():
        return value
    t = torch.tensor([value], dtype=torch.float64,
                     device=device if dist.get_backend() == "nccl" else None)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device: torch.device | None = None) -> float:
    if not dist.is_initialized():
        return value
    t = torch.tensor([value], dtype=torch.float64,
                     device=device if dist.get_backend() == "nccl" else None)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_objects(obj: object) -> list[object]:
    """Rank 0 receives every rank's (small, picklable) summary; other ranks get []."""
    if not dist.is_initialized():
        return [obj]
    out: list[object] | None = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(obj, out, dst=0)
    return out or []
