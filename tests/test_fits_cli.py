"""The callers on either side of the hot path:
FITS ingest / egress of raw table records and the command line of the reference
(src/GPPupilDemodulation.jl:257-426, src/FitsUtils.jl:95-156)."""
import os

import numpy as np
import pytest


def _night(gp, ora, d, n=3000):
    """bright.fits, sub/faint.fits, notmod.fits (MODULATE = F), off.fits (MET MODE OFF), junk.txt"""
    os.makedirs(os.path.join(d, "sub"), exist_ok=True)
    from conftest import make_case
    tabs = {}
    tb = make_case(gp.synthetic, n, k=41)
    gp.synthetic.make_fits(os.path.join(d, "bright.fits"), tb)
    tabs["bright"] = tb
    tf = make_case(gp.synthetic, n, k=42, faint=True, ora=ora)
    gp.synthetic.make_fits(os.path.join(d, "sub", "faint.fits"), tf, tf["header"])
    tabs["faint"] = tf
    gp.synthetic.make_fits(os.path.join(d, "notmod.fits"), tb, modulate=False)
    gp.synthetic.make_fits(os.path.join(d, "off.fits"), tb, met_mode="OFF")
    open(os.path.join(d, "junk.txt"), "w").write("not a fits file")
    return tabs


def test_fits_roundtrip_and_cards(gp, tmp_path):
    from gppd_b200 import fits
    tab = gp.synthetic.make_table(257, k=2)
    p = gp.synthetic.make_fits(str(tmp_path / "a.fits"), tab, gp.synthetic.faint_header(tab["mjd"]))
    hdus = fits.read_fits(p)
    assert [h.name for h in hdus] == ["", "IMAGING_DATA_ACQ", "OPDC", "METROLOGY"]
    assert os.path.getsize(p) % 2880 == 0
    prim = hdus[0].header
    assert prim["ESO INS PMC1 MODULATE"] is True and prim["ESO INS MET MODE"] == "FAINT"
    assert prim["MJD-OBS"] == tab["mjd"] and prim["ESO INS ANLO3 REPEAT1"] == 60
    rb, n, cols = fits.bintable_layout(hdus[3].header)
    assert (rb, n) == (332, 257) and cols["TIME"][:2] == (0, "1J") and cols["VOLT"][:2] == (4, "80E")
    rec = np.frombuffer(hdus[3].data, np.uint8, n * rb).reshape(n, rb)
    assert np.array_equal(rec[:, 4:324].copy().view(">f4"), tab["volt"])
    fits.write_fits(str(tmp_path / "b.fits"), hdus)
    assert open(p, "rb").read() == open(tmp_path / "b.fits", "rb").read()      # byte-exact pass-through
    # cards: HIERARCH for long keywords, exact float round trip, quotes
    for key, val in [("DEMODULATION SIN PHASE FT T1 D1", -3.0141592653589793), ("PROCSOFT", "it's"),
                     ("NAXIS1", 588), ("ESO INS PMC1 MODULATE", False), ("X", 1e-300)]:
        card = fits.format_card(key, val)
        assert len(card) == 80 and fits.parse_card(card) == (key, val)
    # replacing a table: widths, counts and appended columns are book-kept
    new = fits.replace_bintable(hdus[3], np.zeros((n, 588 + 128), np.uint8), {"VOLT": "144E"},
                                [("B", "32E", None)], [("PROCSOFT", "x")])
    rb2, n2, cols2 = fits.bintable_layout(new.header)
    assert (rb2, n2) == (716, n) and cols2["VOLT"][1] == "144E" and cols2["B"][:2] == (588, "32E")
    assert cols2["POWER_LASER"][0] == 4 + 576 and new.header["PROCSOFT"] == "x"


def test_cli_gating_without_gpu(gp, ora, tmp_path):
    """Header gating of main() (:358-392) needs no GPU."""
    from gppd_b200 import cli
    d = str(tmp_path)
    _night(gp, ora, d, n=400)
    args = cli.build_parser().parse_args(["-r", "-s", "_x", "-w", "0.5", "-k", d])
    assert (args.suffix, args.window, args.keepraw, args.recursive, args.center) == ("_x", 0.5, True, True, "stefan")
    plan = {f: cli._plan_file(os.path.join(d, f), args, d)
            for f in ("bright.fits", "sub/faint.fits", "notmod.fits", "off.fits", "junk.txt")}
    assert plan["notmod.fits"] is None and plan["off.fits"] is None and plan["junk.txt"] is None
    jb, jf = plan["bright.fits"], plan["sub/faint.fits"]
    assert jb.faintparam is None and jf.faintparam is not None and jf.faintparam.timer1.size == 2
    assert jb.outname == os.path.join(d, "bright_x.fits") and (jb.row_bytes, jb.n) == (332, 400)
    args2 = cli.build_parser().parse_args(["-f", d])
    assert cli._plan_file(os.path.join(d, "sub/faint.fits"), args2, d).faintparam is None   # --nofaint


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["whole", "keepraw", "window", "fit", "empirical"])
def test_cli_night_end_to_end(gp, ora, tmp_path, mode):
    from gppd_b200 import cli, fits
    d, out = str(tmp_path / "night"), str(tmp_path / "out")
    tabs = _night(gp, ora, d)
    argv = ["-r", "-d", out, "-s", "_demod"]
    argv += {"whole": [], "keepraw": ["-k"], "window": ["-w", "2.0"], "fit": ["-c", "fit"],
             "empirical": ["-c", "empirical"]}[mode]
    assert cli.main(argv + [d]) == 0
    assert sorted(os.listdir(out)) == ["bright_demod.fits", "faint_demod.fits"]
    off = {"fit": False, "empirical": True}.get(mode, gp.synthetic.stefan_centres())
    for name, sub in (("bright", ""), ("faint", "sub")):
        tab = tabs[name]
        src = fits.read_fits(os.path.join(d, sub, name + ".fits"))
        dst = fits.read_fits(os.path.join(out, name + "_demod.fits"))
        assert [h.name for h in dst] == [h.name for h in src]
        for a, b in zip(src[:3], dst[:3]):                      # every other HDU: byte-exact copy
            assert a.cards == b.cards and a.data == b.data
        fs = gp.FaintStates(tab["faintstates"].timer1, tab["faintstates"].timer2, 1.0, 2.0) if name == "faint" else None
        tg, hg = gp.processmetrology({"TIME": tab["time_us"], "VOLT": tab["volt"]}, tab["mjd"],
                                     window=2.0 if mode == "window" else None, faintparam=fs,
                                     keepraw=mode == "keepraw", offsets=off)
        rb, n, cols = fits.bintable_layout(dst[3].header)
        rec = np.frombuffer(dst[3].data, np.uint8, n * rb).reshape(n, rb)
        nv = 144 if mode == "keepraw" else 80
        assert cols["VOLT"][1] == f"{nv}E" and rb == 332 + 4 * (nv - 80) + (0 if mode != "window" else rb - 332)
        volt = rec[:, 4:4 + 4 * nv].copy().view(">f4").astype(np.float32)
        assert np.array_equal(volt, tg["VOLT"])                 # same kernels, same bits as the array path
        # untouched columns keep their bytes
        srec = np.frombuffer(src[3].data, np.uint8, n * 332).reshape(n, 332)
        assert np.array_equal(rec[:, :4], srec[:, :4])
        po = cols["POWER_LASER"][0]
        assert np.array_equal(rec[:, po:po + 8], srec[:, 324:332])
        assert dst[3].header["PROCSOFT"] == "GPPupilDemodulation.jl"
        if mode == "window":
            for k in ("ABSA", "ARGA", "B", "PHI"):
                o = cols[k][0]
                assert np.array_equal(rec[:, o:o + 128].copy().view(">f4").astype(np.float32), tg[k])
            if name == "faint":
                o = cols["STATE"][0]
                assert np.array_equal(rec[:, o].astype(np.int16) - 128, tg["STATE"].astype(np.int16))
        else:
            for k, v in hg.items():
                assert dst[3].header[k] == v, k


def _lzw_compress(data: bytes, maxbits: int = 16, clear_when_full: bool = False) -> bytes:
    """compress(1)-format encoder (block mode), for the .Z tests only."""
    body = bytearray()
    st = dict(acc=0, nacc=0, bits=9, ingroup=0)

    def emit(code):
        st["acc"] |= code << st["nacc"]
        st["nacc"] += st["bits"]
        while st["nacc"] >= 8:
            body.append(st["acc"] & 0xFF)
            st["acc"] >>= 8
            st["nacc"] -= 8
        st["ingroup"] = (st["ingroup"] + 1) % 8

    def pad():
        while st["ingroup"]:
            emit(0)

    table, free = {}, 257
    w = data[:1]
    for i in range(1, len(data)):
        c = data[i:i + 1]
        if w + c in table:
            w = w + c
            continue
        emit(table[w] if len(w) > 1 else w[0])
        full = (1 << maxbits) if st["bits"] == maxbits else (1 << st["bits"]) - 1
        if free > full and st["bits"] < maxbits:
            pad()
            st["bits"] += 1
        if free < (1 << maxbits):
            table[w + c] = free
            free += 1
        elif clear_when_full:
            emit(256)
            pad()
            st["bits"] = 9
            table.clear()
            free = 257
        w = c
    if data:
        emit(table[w] if len(w) > 1 else w[0])
    while st["nacc"] > 0:
        body.append(st["acc"] & 0xFF)
        st["acc"] >>= 8
        st["nacc"] -= 8
    return bytes([0x1F, 0x9D, 0x80 | maxbits]) + bytes(body)


def test_fits_Z_files(gp, tmp_path):
    """`fits.Z` is one of the reference's suffixes (src/GPPupilDemodulation.jl:14; CFITSIO
    decodes compress'ed files).  The decoder is held to streams that gzip's own `-d` accepts."""
    import shutil
    import subprocess
    from gppd_b200 import fits
    rng = np.random.default_rng(1)
    samples = [b"", b"x", b"TOBEORNOTTOBEORTOBEORNOT" * 2000,
               rng.integers(0, 256, 150000, dtype=np.uint8).tobytes(),      # fills the 16-bit table
               rng.integers(0, 4, 300000, dtype=np.uint8).tobytes()]
    for data in samples:
        for maxbits, clear in ((16, False), (12, True), (10, False)):
            z = _lzw_compress(data, maxbits, clear)
            assert fits.unlzw(z) == data
            if shutil.which("gzip"):
                p = tmp_path / "t.Z"
                p.write_bytes(z)
                r = subprocess.run(["gzip", "-dc", str(p)], capture_output=True)
                assert r.returncode == 0 and r.stdout == data       # the test encoder writes valid streams
    with pytest.raises(ValueError):
        fits.unlzw(b"\x1f\x8b\x08")
    # a whole file
    tab = gp.synthetic.make_table(300, k=5)
    plain = gp.synthetic.make_fits(str(tmp_path / "a.fits"), tab)
    (tmp_path / "a.fits.Z").write_bytes(_lzw_compress(open(plain, "rb").read()))
    a, b = fits.read_fits(plain), fits.read_fits(str(tmp_path / "a.fits.Z"))
    assert len(a) == len(b) and all(x.cards == y.cards and x.data == y.data for x, y in zip(a, b))


@pytest.mark.gpu
def test_cli_reads_compressed_inputs(gp, ora, tmp_path):
    """The same table as .fits, .fits.gz and .fits.Z gives the same output file."""
    import gzip
    from gppd_b200 import cli, fits
    d, out = tmp_path / "in", tmp_path / "out"
    d.mkdir()
    tab = gp.synthetic.make_table(2500, k=6)
    plain = gp.synthetic.make_fits(str(d / "p.fits"), tab)
    raw = open(plain, "rb").read()
    with gzip.open(str(d / "g.fits.gz"), "wb") as fh:
        fh.write(raw)
    (d / "z.fits.Z").write_bytes(_lzw_compress(raw))
    assert cli.main(["-d", str(out), str(d / "p.fits"), str(d / "g.fits.gz"), str(d / "z.fits.Z")]) == 0
    assert sorted(os.listdir(out)) == ["g.fits", "p.fits", "z.fits"]
    ref = open(out / "p.fits", "rb").read()
    assert open(out / "g.fits", "rb").read() == ref and open(out / "z.fits", "rb").read() == ref


@pytest.mark.gpu
def test_native_file_path_many_files_and_errors(gp, ora, tmp_path):
    """The native ingest (gppd_file_*): more files than pipeline slots, of different lengths and
    modes, give the same bytes as the Python record path (--no-native); a missing input file is
    reported by the wait / drain, not lost on the I/O threads."""
    import ctypes as C
    from gppd_b200 import cli, _lib
    d, out1, out2 = tmp_path / "in", tmp_path / "native", tmp_path / "python"
    d.mkdir()
    from conftest import make_case
    for k in range(19):                     # 19 files over 8 slots, ragged sizes, every third FAINT
        faint = k % 3 == 2
        tab = make_case(gp.synthetic, 700 + 137 * k, k=60 + k, faint=faint, ora=ora)
        gp.synthetic.make_fits(str(d / ("f%02d.fits" % k)), tab, tab["header"])
    for mode in ([], ["-w", "0.6", "-c", "fit"], ["-k"]):
        for o in (out1, out2):
            if o.exists():
                for f in os.listdir(o):
                    os.remove(o / f)
        assert cli.main(mode + ["-d", str(out1)] + [str(d / f) for f in sorted(os.listdir(d))]) == 0
        assert cli.main(mode + ["--no-native", "-d", str(out2)] + [str(d / f) for f in sorted(os.listdir(d))]) == 0
        names = sorted(os.listdir(out1))
        assert len(names) == 19 and names == sorted(os.listdir(out2))
        for f in names:
            assert open(out1 / f, "rb").read() == open(out2 / f, "rb").read(), (mode, f)
    # error reporting of the asynchronous parts
    L, h = _lib.lib(), gp.default_handle()
    o = gp.api._options()
    _lib.check(L.gppd_file_submit(h.raw, 3, os.fsencode(str(d / "missing.fits")), 2880, 100, 332, 0, 4, 59949.0,
                                  None, None, 0, None, 0, 0.0, C.byref(o)))
    par, chi = np.empty((32, 6)), np.empty(32)
    rc = L.gppd_file_wait(h.raw, 3, _lib.ptr(par), _lib.ptr(chi), None, None)
    assert rc == 6 and b"missing.fits" in L.gppd_last_error()          # GPPD_ERR_IO
    assert L.gppd_file_drain(h.raw) == 6                                # ... and once more by the drain
    assert L.gppd_file_drain(h.raw) == 0                                # which clears it
    assert L.gppd_file_submit(h.raw, 99, b"x", 0, 100, 332, 0, 4, 0.0, None, None, 0, None, 0, 0.0, C.byref(o)) == 1
