"""Command-line front end with the reference's flags (``main``,
src/GPPupilDemodulation.jl:257-426) and a night-level scheduler on top of the C ABI.

    python -m gppd_b200.cli [-s SUFFIX] [-o] [-f] [-r] [-v] [-k] [-c CENTER]
                            [-w WINDOW] [-d DIR] INPUT...

Same gating as the reference: a file is processed when its primary header has
``ESO INS PMC1 MODULATE = T`` and ``ESO INS MET MODE`` is not ``OFF``; ``FAINT`` mode
builds the laser on/off timers from the ``ESO INS ANLO3`` keywords unless
``--nofaint``.  The output is a copy of the input file with the ``METROLOGY`` table
and its header replaced, written to ``DIR/<name><SUFFIX>.fits``.

What differs is the execution: files are not processed one after the other but
pipelined over the handle's slots -- while one file's records are on the GPU the next
one is being read and uploaded and the previous one written -- and with several GPUs
(``RANK`` / ``WORLD_SIZE`` from torchrun, or ``--rank/--world``) every rank takes every
WORLD_SIZE-th file; there is no communication between ranks.  The METROLOGY records
travel as raw FITS bytes (``gppd_submit_fits_rows``): byte swapping, de-interleaving,
the fit, the demodulation and the re-packing all happen on the device.
"""
from __future__ import annotations

import argparse
import ctypes as C
import logging
import os
import sys
import time

import numpy as np

from . import _lib, fits
from .sharding import partition_files
from .api import (_options, buildfaintparameters, demodulation_keys, read_stefan_file,
                  window_columns)

SUFFIXES = (".fits", ".fits.gz", "fits.Z")      # src/GPPupilDemodulation.jl:14
log = logging.getLogger("GPPupilDemodulation")


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog="GPPupilDemodulation",
                                description="Simple tool to demodulate Gravity metrology table.")
    p.add_argument("--version", action="version", version="0.1")
    p.add_argument("--suffix", "-s", default="",
                   help="Store the demodulated metrology in the INPUT.SUFFIX.fits file")
    p.add_argument("--onlyhigh", "-o", action="store_true",
                   help="Demodulate metrology using parameters estimated only on HIGH and NORMAL")
    p.add_argument("--nofaint", "-f", action="store_true",
                   help="Do no use the faint mode state to demodulate")
    p.add_argument("--recursive", "-r", action="store_true", help="Recursively explore entire directories.")
    p.add_argument("--verbose", "-v", action="store_true", help="Verbose mode")
    p.add_argument("--keepraw", "-k", action="store_true", help="keep raw")
    p.add_argument("--center", "-c", default="stefan",
                   help="center voltages: stefan (default) | empirical | uncentered | fit")
    p.add_argument("--window", "-w", type=float, default=0.0,
                   help="Compute demodulation on non overlapping window of WINDOW second")
    p.add_argument("--dir", "-d", default=os.getcwd(), help="output folder")
    p.add_argument("--rank", type=int, default=int(os.environ.get("RANK", "0")), help=argparse.SUPPRESS)
    p.add_argument("--world", type=int, default=int(os.environ.get("WORLD_SIZE", "1")), help=argparse.SUPPRESS)
    p.add_argument("--device", type=int, default=int(os.environ.get("LOCAL_RANK", "0")), help=argparse.SUPPRESS)
    p.add_argument("INPUT", nargs="*", default=["."],
                   help="List of all TARGET to process. In conjunction with -r TARGET can contain directories.")
    return p


class _Job:
    """One file in flight on one pipeline slot."""

    def __init__(self, filename, outname, hdus, imet, faintparam, mjd, layout, window, keepraw):
        self.filename, self.outname, self.hdus, self.imet = filename, outname, hdus, imet
        self.faintparam, self.mjd, self.window, self.keepraw = faintparam, mjd, window, keepraw
        self.row_bytes, self.n, self.cols = layout
        self.tstart = time.time()


def _plan_file(filename, args, folder):
    """Header gating of one file (reference :358-392).  Returns a _Job or None."""
    if not os.path.isfile(filename) or not filename.endswith(SUFFIXES):
        return None
    hdus = fits.read_fits(filename)
    prim = hdus[0].header
    if "ESO INS PMC1 MODULATE" not in prim:
        log.info("no ESO INS PMC1 MODULATE keyword in %s", filename)
        return None
    if not prim["ESO INS PMC1 MODULATE"]:
        log.info("ESO INS PMC1 MODULATE set to false in %s", filename)
        return None
    log.info("Processing %s", filename)
    metmod = prim.get("ESO INS MET MODE", "ON")
    log.info("%s use %s metrology mode", filename, metmod)
    if metmod == "OFF":
        return None
    faintparam = None
    if metmod == "FAINT":
        if not args.nofaint:
            faintparam = buildfaintparameters(prim)
        else:
            log.info("FAINT mode deactivated")
    mjd = float(prim["MJD-OBS"])
    imet = next((i for i, h in enumerate(hdus) if h.is_bintable and h.name == "METROLOGY"), None)
    if imet is None:
        raise KeyError(f"{filename}: no METROLOGY table")
    layout = fits.bintable_layout(hdus[imet].header)
    for col, width in (("TIME", 4), ("VOLT", 320)):
        if col not in layout[2] or fits.tform_bytes(layout[2][col][1]) != width:
            raise ValueError(f"{filename}: METROLOGY column {col} missing or of unexpected width")
    fname = os.path.basename(filename).split(".fits")[0]
    outname = os.path.join(folder, fname + args.suffix + ".fits")
    return _Job(filename, outname, hdus, imet, faintparam, mjd, layout,
                args.window if args.window != 0.0 else None, args.keepraw)


class NightScheduler:
    """Files -> pipeline slots of one handle (one GPU)."""

    def __init__(self, handle, offsets, onlyhigh, method="auto"):
        self.h, self.L = handle, _lib.lib()
        self.empirical = offsets is True       # --center empirical: centres fitted on the device
        self.offsets = (None if (offsets is None or self.empirical)
                        else np.ascontiguousarray(offsets, dtype=np.complex128))
        self.onlyhigh, self.method = onlyhigh, method
        self.nslots = handle.num_slots
        self.busy = [None] * self.nslots       # (job, buffers) per slot
        self.pinned = [dict() for _ in range(self.nslots)]

    def _pinned(self, slot, key, nbytes):
        """A pinned staging buffer of the slot, grown on demand."""
        cur = self.pinned[slot].get(key)
        if cur is None or cur[1] < nbytes:
            if cur is not None:
                _lib.check(self.L.gppd_free_pinned(self.h.raw, cur[0]))
            p = C.c_void_p()
            cap = int(nbytes * 1.25) + 4096
            _lib.check(self.L.gppd_alloc_pinned(self.h.raw, cap, C.byref(p)))
            cur = (p, cap)
            self.pinned[slot][key] = cur
        return np.ctypeslib.as_array(C.cast(cur[0], C.POINTER(C.c_uint8)), shape=(cur[1],))[:nbytes]

    def submit(self, slot, job: _Job):
        n, rb = job.n, job.row_bytes
        rb_out = rb + (256 if job.keepraw else 0)
        src = np.frombuffer(job.hdus[job.imet].data, dtype=np.uint8, count=n * rb)
        rows = self._pinned(slot, "rows", n * rb)
        rows[:] = src
        rows_out = self._pinned(slot, "rows_out", n * rb_out)
        # window arithmetic needs TIME[0:2]; the library reads them from the records
        wrows, nwin = n, 1
        if job.window is not None:
            toff = job.cols["TIME"][0]
            t01 = np.array([int.from_bytes(src[k * rb + toff:k * rb + toff + 4].tobytes(), "big", signed=True)
                            for k in range(2)], dtype=np.int32)
            w, k = C.c_int64(0), C.c_int64(0)
            _lib.check(self.L.gppd_table_windows(2, _lib.ptr(t01, _lib._i32p), job.mjd, float(job.window),
                                                 C.byref(w), C.byref(k)))
            wrows = w.value
            nwin = (n + wrows - 1) // wrows
        params = np.empty((nwin * 32, 6))
        chi2 = np.empty(nwin * 32)
        state = np.empty(n, dtype=np.int8) if job.faintparam is not None else None
        o = _options(onlyhigh=self.onlyhigh, keepraw=job.keepraw, method=self.method,
                     empirical=self.empirical)
        fp = job.faintparam
        t1 = fp.timer1 if fp is not None else None
        t2 = fp.timer2 if fp is not None else None
        off = None if self.offsets is None else self.offsets.view(np.float64)
        _lib.check(self.L.gppd_submit_fits_rows(
            self.h.raw, slot, n, rows.ctypes.data_as(C.c_void_p), rb, job.cols["TIME"][0],
            job.cols["VOLT"][0], job.mjd, _lib.ptr(off), _lib.ptr(t1), 0 if t1 is None else t1.size,
            _lib.ptr(t2), 0 if t2 is None else t2.size, float(job.window or 0.0), C.byref(o),
            rows_out.ctypes.data_as(C.c_void_p), _lib.ptr(params), _lib.ptr(chi2), None,
            _lib.ptr(state, _lib._i8p)))
        self.busy[slot] = (job, dict(rows_out=rows_out, params=params, chi2=chi2, state=state,
                                     wrows=wrows, nwin=nwin, rb_out=rb_out, keep=(o, off, t1, t2)))

    def finish(self, slot):
        """Wait for the slot's file, take its records out of the slot's pinned buffer and hand
        the assembling and writing of the output file to the writer threads (reference
        :406-412); returns a future of the output name."""
        job, b = self.busy[slot]
        self.busy[slot] = None
        _lib.check(self.L.gppd_wait(self.h.raw, slot))
        b = dict(b, rows_out=b["rows_out"].copy())     # the slot's staging buffer is reused
        return self.writers.submit(self._write, job, b)

    def _write(self, job, b):
        n = job.n
        fitoffsets = self.offsets is None and not self.empirical
        records = b["rows_out"].reshape(n, b["rb_out"])
        keys, newcols = [], []
        if job.window is None:
            keys += list(demodulation_keys(b["params"], fitoffsets).items())
        else:
            cols = window_columns(b["params"], n, b["wrows"], b["nwin"], fitoffsets)
            extra = []
            for name in ("ABSA", "ARGA", "B", "PHI", "X0", "Y0"):
                if name in cols:
                    extra.append(cols[name].astype(">f4").view(np.uint8).reshape(n, 128))
                    newcols.append((name, "32E", None))
            if b["state"] is not None:      # Int8 column: FITS 'B' with TZERO = -128
                extra.append((b["state"].astype(np.int16) + 128).astype(np.uint8).reshape(n, 1))
                newcols.append(("STATE", "1B", None))
            records = np.concatenate([records] + extra, axis=1)
        keys.append(("PROCSOFT", "GPPupilDemodulation.jl"))
        met = job.hdus[job.imet]
        new = fits.replace_bintable(met, records,
                                    tform_changes={"VOLT": "144E"} if job.keepraw else None,
                                    new_columns=newcols, new_keys=keys)
        if b["state"] is not None and job.window is not None:
            i = int(new.header["TFIELDS"])
            new.cards += [fits.format_card(f"TSCAL{i}", 1), fits.format_card(f"TZERO{i}", -128)]
        hdus = list(job.hdus)
        hdus[job.imet] = new
        fits.write_fits(job.outname, hdus)
        log.info("%s processed in %.3f s", job.filename, time.time() - job.tstart)
        log.info(" %s written", job.outname)
        return job.outname

    def run(self, jobs):
        """Pipeline the jobs over the slots -- reader threads plan (read and gate) the next
        files, the GPU works on up to `nslots` files, writer threads assemble and write the
        finished ones; returns the output names in input order."""
        from concurrent.futures import ThreadPoolExecutor
        done, slot = [], 0
        with ThreadPoolExecutor(max_workers=3) as self.writers:
            for job in jobs:
                if self.busy[slot] is not None:
                    done.append(self.finish(slot))
                self.submit(slot, job)
                slot = (slot + 1) % self.nslots
            for k in range(self.nslots):
                s = (slot + k) % self.nslots
                if self.busy[s] is not None:
                    done.append(self.finish(s))
            return [f.result() for f in done]


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    logging.basicConfig(level=logging.INFO if args.verbose else logging.WARNING,
                        format="[ Info: %(message)s", stream=sys.stderr)
    files = []
    for arg in args.INPUT:              # reference :324-331
        if os.path.isdir(arg) and args.recursive:
            for root, _, names in os.walk(arg):
                files += [os.path.join(root, f) for f in sorted(names)]
        else:
            files.append(arg)
    folder = args.dir
    if not folder.startswith("/"):
        folder = os.path.join(os.getcwd(), folder)
    os.makedirs(folder, exist_ok=True)   # (the reference creates relative folders only)
    if args.center == "stefan":          # reference :346-354
        offsets = read_stefan_file()
    elif args.center == "uncentered":
        offsets = np.zeros(40, dtype=np.complex128)
    elif args.center == "empirical":     # offsets = true: compute_offsets, :105-125
        offsets = True
    elif args.center == "fit":
        offsets = None
    else:
        raise SystemExit(f"unknown --center {args.center}")
    mine = partition_files(files, args.rank, args.world)   # no communication between ranks
    handle = _lib.Handle(args.device)
    sched = NightScheduler(handle, offsets, args.onlyhigh)

    def jobs():
        # read and gate a few files ahead of the GPU (file reads release the interpreter lock)
        from collections import deque
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=2) as readers:
            ahead = deque()
            it = iter(mine)
            for f in it:
                ahead.append(readers.submit(_plan_file, f, args, folder))
                if len(ahead) < 4:
                    continue
                j = ahead.popleft().result()
                if j is not None:
                    yield j
            while ahead:
                j = ahead.popleft().result()
                if j is not None:
                    yield j

    sched.run(jobs())
    return 0


if __name__ == "__main__":
    sys.exit(main())
