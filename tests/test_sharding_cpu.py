"""CPU tests of the row-band sharding host logic: band planning and the halo exchange protocol
(world size 2 and 3, gloo).  The kernels themselves need GPUs (tools/sharded_check.py)."""
from __future__ import annotations

import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from style_transfer_visualizer_b200.sharded import exchange_rows, plan_bands


def test_plan_bands_alignment_and_coverage() -> None:
    assert plan_bands(2160, 8) == [(0, 272), (272, 544), (544, 816), (816, 1088), (1088, 1360),
                                   (1360, 1632), (1632, 1904), (1904, 2160)]
    for h in (64, 150, 1080, 2160, 2161):
        for world in (1, 2, 3, 4):
            bands = plan_bands(h, world)
            assert bands[0][0] == 0 and bands[-1][1] == h
            assert all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
            assert all(y0 % 16 == 0 for y0, _ in bands)
            assert all(y1 > y0 for y0, y1 in bands)
    with pytest.raises(ValueError, match="too few"):
        plan_bands(32, 4)


def _worker(rank: int, world: int, port: int, out) -> None:  # noqa: ANN001
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    buf = torch.full((4 + 2, 3, 2), float(rank + 1))       # own rows 1..4 hold rank+1
    buf[1] += 0.25                                          # first own row
    buf[4] += 0.5                                           # last own row
    buf[0] = -7.0                                           # stale halos
    buf[5] = -7.0
    exchange_rows(buf, rank, world)
    out.put((rank, buf[0, 0, 0].item(), buf[5, 0, 0].item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_gloo(world: int) -> None:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict((r, (top, bot)) for r, top, bot in (q.get(timeout=120) for _ in range(world)))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        top, bot = got[r]
        assert top == (0.0 if r == 0 else (r + 0.5))              # upper neighbour's LAST own row
        assert bot == (0.0 if r == world - 1 else (r + 2 + 0.25))  # lower neighbour's FIRST own row
