"""Multi-GPU host logic on CPU (gloo, world_size 2): the night is dealt out to ranks
with no data-path collective -- every file goes to exactly one
rank -- and the only reduction is the max-over-ranks of the elapsed time that bench.py
reports."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, files, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = files[rank::world]                       # cli.main's partition
    # what the ranks would exchange at the end: nothing but counts and the slowest time
    counts = [None] * world
    dist.all_gather_object(counts, mine)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)   # this rank's elapsed time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    units = torch.tensor([float(len(mine))], dtype=torch.float64)
    dist.all_reduce(units, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put((counts, float(t.item()), float(units.item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_files_are_dealt_out_once_and_time_is_max_over_ranks():
    files = [f"night/f{k:03d}.fits" for k in range(11)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, files, q)) for r in range(2)]
    for p in procs:
        p.start()
    counts, tmax, units = q.get(timeout=90)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert sorted(counts[0] + counts[1]) == files and not set(counts[0]) & set(counts[1])
    assert abs(len(counts[0]) - len(counts[1])) <= 1
    assert tmax == 2.0 and units == len(files)          # whole-job units / slowest rank
