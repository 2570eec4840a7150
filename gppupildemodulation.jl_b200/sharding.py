"""How the demodulateall path is split over the GPUs of one box (one process per GPU).

The path has three kinds of independent units and NO exchange step,
so every split is a pure partition followed by a host-side gather:

  files    the reference processes them one after the other
           (src/GPPupilDemodulation.jl:356)            -> ``partition_files``
  windows  ``--window``: every window is its own demodulateall call
           (src/GPPupilDemodulation.jl:204-205)        -> ``partition_windows``
  groups   inside one call the 8 (telescope, side) groups are independent, each
           reads its 4 diode channels + its FC channel (src/Modulation.jl:387-390)
                                                       -> ``partition_groups`` and
           ``gppd_options.group_mask``: one global fit over a very long exposure is
           split 8 groups <-> 1/2/4/8 GPUs

NumPy only; the ranks never talk to each other on the data path.
"""
from __future__ import annotations

import numpy as np

NGROUP = 8


def partition_files(files, rank: int, world: int) -> list:
    """The files rank ``rank`` of ``world`` processes: every world-th one, so that a
    night sorted by time is spread evenly."""
    world = max(1, int(world))
    if not 0 <= rank < world:
        raise ValueError("rank outside 0..world-1")
    return list(files)[rank::world]


def partition_groups(rank: int, world: int) -> int:
    """``gppd_options.group_mask`` of rank ``rank``: contiguous runs of the 8
    (telescope, side) groups, bit g = group g (channels 4g..4g+3 and FC channel 32+g).
    ``world`` must divide 8 (1, 2, 4 or 8 GPUs); more ranks than groups get mask 0
    (nothing to do)."""
    world = max(1, int(world))
    if not 0 <= rank < world:
        raise ValueError("rank outside 0..world-1")
    if world > NGROUP:
        return (1 << rank) if rank < NGROUP else 0
    if NGROUP % world:
        raise ValueError("the number of ranks must divide the 8 diode groups")
    per = NGROUP // world
    return ((1 << per) - 1) << (rank * per)


def mask_groups(mask: int) -> list:
    return [g for g in range(NGROUP) if (mask >> g) & 1]


def mask_channels(mask: int) -> list:
    """0-based channel numbers (of 40) a mask reads and writes."""
    ch = []
    for g in mask_groups(mask):
        ch += [4 * g, 4 * g + 1, 4 * g + 2, 4 * g + 3]
    return ch + [32 + g for g in mask_groups(mask)]


def partition_windows(nwin: int, rank: int, world: int) -> tuple:
    """Contiguous window range [lo, hi) of rank ``rank`` (sizes differ by at most one)."""
    world = max(1, int(world))
    if not 0 <= rank < world:
        raise ValueError("rank outside 0..world-1")
    base, extra = divmod(int(nwin), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_groups(parts, masks, nwin: int = 1):
    """Host-side gather of a call that was split by groups: ``parts`` = one
    (output (N, 40), params (nwin*32, 6), chi2 (nwin*32,)) triple per rank, ``masks`` =
    their group masks.  Returns the merged triple; every group must come from exactly
    one rank."""
    seen = 0
    for m in masks:
        if seen & m:
            raise ValueError("a group was processed by two ranks")
        seen |= m
    if seen != (1 << NGROUP) - 1:
        raise ValueError("groups missing from the gather")
    out = np.empty_like(parts[0][0])
    par = np.empty_like(parts[0][1])
    chi = np.empty_like(parts[0][2])
    for (o, p, c), m in zip(parts, masks):
        ch = mask_channels(m)
        out[:, ch] = o[:, ch]
        fits = np.array([w * 32 + d for w in range(nwin) for d in ch if d < 32], dtype=np.int64)
        par[fits] = p[fits]
        chi[fits] = c[fits]
    return out, par, chi


def gather_windows(parts, ranges, n: int, wrows: int):
    """Host-side gather of a call that was split by windows: ``parts`` = one
    (output rows of its windows, params, chi2) triple per rank, ``ranges`` their
    [lo, hi) window ranges in order."""
    out = np.concatenate([p[0] for p in parts], axis=0)
    par = np.concatenate([p[1] for p in parts], axis=0)
    chi = np.concatenate([p[2] for p in parts], axis=0)
    if out.shape[0] != n or ranges[0][0] != 0 or any(a[1] != b[0] for a, b in zip(ranges, ranges[1:])):
        raise ValueError("window ranges do not tile the exposure")
    return out, par, chi
