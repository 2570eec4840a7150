"""Multi-GPU host logic on CPU (gloo, world_size 2): the three partitions of
gppd_b200.sharding -- files -> ranks (what cli.main uses), windows -> ranks, the 8
(telescope, side) groups of ONE call -> ranks (gppd_options.group_mask) -- and the
host-side gathers that put the ranks' pieces back together.  No data-path collective
(SURVEY.md section 8e): the only reduction is the max-over-ranks of the elapsed time that
bench.py reports."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _full_result(n, nwin):
    """A stand-in for the result of one unsharded call: every element encodes where it is."""
    out = (np.arange(n)[:, None] * 100.0 + np.arange(40)[None, :]).astype(np.complex128)
    par = np.arange(nwin * 32 * 6, dtype=np.float64).reshape(nwin * 32, 6)
    chi = np.arange(nwin * 32, dtype=np.float64) + 0.5
    return out, par, chi


def _worker(rank, world, port, files, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    from gppd_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.partition_files(files, rank, world)          # the package's partition
    mask = sharding.partition_groups(rank, world)
    n, wrows = 1030, 100
    nwin = (n + wrows - 1) // wrows
    wlo, whi = sharding.partition_windows(nwin, rank, world)
    # what each rank would have computed: only ITS groups of the group-sharded call and
    # only ITS windows of the window-sharded call are valid, the rest is poison
    out, par, chi = _full_result(n, 1)
    ch = sharding.mask_channels(mask)
    o = np.full_like(out, np.nan)
    o[:, ch] = out[:, ch]
    p = np.full_like(par, np.nan)
    c = np.full_like(chi, np.nan)
    fits = [d for d in ch if d < 32]
    p[fits], c[fits] = par[fits], chi[fits]
    outw, parw, chiw = _full_result(n, nwin)
    rows = slice(wlo * wrows, min(whi * wrows, n))
    piece_w = (outw[rows], parw[wlo * 32:whi * 32], chiw[wlo * 32:whi * 32])
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(files=mine, mask=mask, wrange=(wlo, whi),
                                          piece_g=(o, p, c), piece_w=piece_w))
    t = torch.tensor([1.0 + rank], dtype=torch.float64)   # this rank's elapsed time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    units = torch.tensor([float(len(mine))], dtype=torch.float64)
    dist.all_reduce(units, op=dist.ReduceOp.SUM)
    if rank == 0:
        masks = [g["mask"] for g in gathered]
        mg = sharding.gather_groups([g["piece_g"] for g in gathered], masks)
        mw = sharding.gather_windows([g["piece_w"] for g in gathered], [g["wrange"] for g in gathered],
                                     n, wrows)
        ok_g = all(np.array_equal(a, b) for a, b in zip(mg, _full_result(n, 1)))
        ok_w = all(np.array_equal(a, b) for a, b in zip(mw, _full_result(n, nwin)))
        q.put(([g["files"] for g in gathered], masks, [g["wrange"] for g in gathered],
               float(t.item()), float(units.item()), ok_g, ok_w))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_partitions_and_gathers_world2():
    files = [f"night/f{k:03d}.fits" for k in range(11)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, files, q)) for r in range(2)]
    for p in procs:
        p.start()
    counts, masks, wranges, tmax, units, ok_g, ok_w = q.get(timeout=90)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert sorted(counts[0] + counts[1]) == files and not set(counts[0]) & set(counts[1])
    assert abs(len(counts[0]) - len(counts[1])) <= 1
    assert masks == [0x0f, 0xf0]                        # FT groups on rank 0, SC groups on rank 1
    assert wranges == [(0, 6), (6, 11)]
    assert ok_g and ok_w                                # the gathers rebuild the unsharded result
    assert tmax == 2.0 and units == len(files)          # whole-job units / slowest rank


def test_partition_rules():
    from gppd_b200 import sharding
    for world in (1, 2, 4, 8):
        masks = [sharding.partition_groups(r, world) for r in range(world)]
        assert sum(masks) == 0xff and all(bin(m).count("1") == 8 // world for m in masks)
    assert sharding.mask_channels(0x03) == [0, 1, 2, 3, 4, 5, 6, 7, 32, 33]
    with pytest.raises(ValueError):
        sharding.partition_groups(0, 3)
    with pytest.raises(ValueError):
        sharding.gather_groups([_full_result(4, 1)] * 2, [0x0f, 0x1f])
    for nwin, world in ((11, 2), (5, 8), (64, 4), (1, 2)):
        r = [sharding.partition_windows(nwin, k, world) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == nwin and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    files = list(range(10))
    assert sorted(sum((sharding.partition_files(files, r, 3) for r in range(3)), [])) == files
